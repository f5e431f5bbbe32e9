"""Static attribution of a kernel's SASS to engine stages, from `nvdisasm --print-line-info-inline` output.
  cuobjdump -xelf all x.o; nvdisasm --print-line-info-inline x.sm_100a.cubin > x.dis
  python tools/sass_by_stage.py x.dis <kernel-substring> mujoco-template_b200/csrc/b2_engine.cuh"""
import collections
import re
import sys

dis, kern, engine = sys.argv[1:4]
# function start lines in the engine source
funcs = []
for n, line in enumerate(open(engine), 1):
    m = re.match(r"\s+(?:template <[^>]*>\s*)?(?:static )?B2_(?:STAGE|DEV) [\w:<>\* ]+?\b(\w+)\(", line)
    if m:
        funcs.append((n, m.group(1)))
def fn_of(line):
    name = "?"
    for n, f in funcs:
        if n <= line:
            name = f
        else:
            break
    return name
WRAPPERS = {"forward", "forward_position", "forward_rest", "step", "constrained_acceleration", "euler_wrap"}
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l)
stage = "kernel"
by = collections.defaultdict(collections.Counter)
chain = []
fresh = False
for l in lines[start + 1:]:
    if l.startswith(".text."):
        break
    m = re.search(r'//## File ".*?/(\w+\.cuh?)", line (\d+)', l)
    if m:
        if not fresh:
            chain, fresh = [], True
        chain.append((m.group(1), int(m.group(2))))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
    if m:
        if fresh:
            names = [fn_of(ln) for f, ln in chain if f == "b2_engine.cuh"]
            names = [x for x in names if x not in WRAPPERS]
            stage = names[-1] if names else ("kernel:" + chain[-1][0] if chain else "kernel")
            fresh = False
        op = m.group(1)
        cls = "fp64" if op in ("DFMA", "DADD", "DMUL", "DSETP") else ("mem" if op in ("LDL", "STL", "LDG", "STG") else "other")
        by[stage][cls] += 1
tot = collections.Counter()
for c in by.values():
    tot.update(c)
print("%-28s %7s %7s %7s" % ("stage", "fp64", "mem", "other"))
for s, c in sorted(by.items(), key=lambda kv: -sum(kv[1].values())):
    print("%-28s %7d %7d %7d" % (s, c["fp64"], c["mem"], c["other"]))
print("%-28s %7d %7d %7d" % ("TOTAL", tot["fp64"], tot["mem"], tot["other"]))
