"""Dynamic attribution of a kernel's executed instructions / stall samples to source functions.
  ncu -i rep --page source --csv > src.csv          (SASS rows with executed counts, same binary as below)
  cuobjdump -xelf all x.o; nvdisasm --print-line-info-inline x.sm_100a.cubin > x.dis
  python tools/ncu_by_function.py src.csv x.dis <kernel-substring> <source-file> [depth]
Instructions are matched by their offset from the start of the kernel."""
import collections
import csv
import re
import sys

src_csv, dis, kern, srcfile = sys.argv[1:5]
depth = int(sys.argv[5]) if len(sys.argv) > 5 else 1
base = srcfile.split("/")[-1]
funcs = []
for n, line in enumerate(open(srcfile), 1):
    m = re.match(r"\s*(?:template <[^>]*>\s*)?(?:static )?(?:B2_(?:STAGE|DEV)|__device__(?: __forceinline__)?|__global__)[\w:<>\*& ,]*?\b(\w+)\(", line)
    if m and m.group(1) not in ("if", "for", "while"):
        funcs.append((n, m.group(1)))
def fn_of(line):
    name = "?"
    for n, f in funcs:
        if n <= line:
            name = f
        else:
            break
    return name
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l)
off2fn = {}
chain, fresh, cur = [], False, "?"
for l in lines[start + 1:]:
    if l.startswith(".text."):
        break
    m = re.search(r'//## File ".*?/([\w.]+)", line (\d+)', l)
    if m:
        if not fresh:
            chain, fresh = [], True
        chain.append((m.group(1), int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+", l)
    if m:
        if fresh:
            names = [fn_of(ln) for f, ln in chain if f == base]
            cur = "/".join(names[-depth:][::-1]) if names else "other:" + (chain[0][0] if chain else "?")
            fresh = False
        off2fn[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
h = rows[1]
ia, ie, isamp = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
ith = h.index("Thread Instructions Executed")
first = None
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in rows[2:]:
    if len(r) <= ith or not r[ie].isdigit():
        continue
    a = int(r[ia], 16)
    if first is None:
        first = a
    fn = off2fn.get(a - first, "unmapped")
    agg[fn][0] += int(r[ie]); agg[fn][1] += int(r[isamp]); agg[fn][2] += int(r[ith])
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("%-44s %12s %6s %7s %6s" % ("function", "warp_instr", "%", "stall%", "lanes"))
for fn, (n, s, th) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if n:
        print("%-44s %12d %5.1f%% %6.1f%% %6.1f" % (fn[:44], n, 100 * n / ti, 100 * s / ts, th / n))
print("total warp instr", ti, "samples", ts)
