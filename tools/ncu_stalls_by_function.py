"""Stall-reason samples per source function of one kernel (companion of ncu_by_function.py).
  python tools/ncu_stalls_by_function.py src.csv x.dis <kernel-substring> <source-file> [depth]"""
import collections
import csv
import re
import sys

src_csv, dis, kern, srcfile = sys.argv[1:5]
depth = int(sys.argv[5]) if len(sys.argv) > 5 else 2
base = srcfile.split("/")[-1]
funcs = []
for n, line in enumerate(open(srcfile), 1):
    m = re.match(r"\s*(?:template <[^>]*>\s*)?(?:static )?(?:B2_(?:STAGE|DEV)|__device__(?: __forceinline__)?(?: __noinline__)?|__global__)[\w:<>\*& ,]*?\b(\w+)\(", line)
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
off2 = {}
chain, fresh, cur = [], False, None
ninstr = 0
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
            names = [fn_of(ln) if f == base else "other:" + f for f, ln in chain]
            names = [n for i, n in enumerate(names) if i == 0 or n != names[i - 1]]
            cur = "/".join(reversed(names[:depth][::1])) if names else "?"
            cur = "/".join(list(reversed(names))[-depth:])
            fresh = False
        off2[int(m.group(1), 16)] = cur
        ninstr += 1
rows = list(csv.reader(open(src_csv)))
h = rows[1]
ia, ie = h.index("Address"), h.index("Instructions Executed")
reasons = ["stall_no_inst", "stall_wait", "stall_long_sb", "stall_short_sb", "stall_selected", "stall_branch_resolving", "stall_math", "stall_barrier"]
ir = [h.index(r) for r in reasons]
agg = collections.defaultdict(lambda: [0] * (len(reasons) + 2))
static = collections.Counter(off2.values())
first = None
for r in rows[2:]:
    if len(r) <= max(ir) or not r[ie].isdigit():
        continue
    a = int(r[ia], 16)
    if first is None:
        first = a
    c = off2.get(a - first, "?")
    agg[c][0] += int(r[ie])
    for k, i in enumerate(ir):
        agg[c][2 + k] += int(r[i] or 0)
tot = [sum(v[k] for v in agg.values()) for k in range(len(reasons) + 2)]
print("static SASS instructions:", ninstr)
print("%-44s %7s %9s " % ("function", "static", "instr%") + " ".join("%9s" % r.replace("stall_", "")[:9] for r in reasons))
for c, v in sorted(agg.items(), key=lambda x: -sum(x[1][2:])):
    if sum(v[2:]) < 0.004 * sum(tot[2:]):
        continue
    print("%-44s %7d %8.1f%% " % (c[-44:], static[c], 100 * v[0] / tot[0]) + " ".join("%8.1f%%" % (100 * x / sum(tot[2:])) for x in v[2:]))
print("%-44s %7d %8.1f%% " % ("total", ninstr, 100) + " ".join("%8.1f%%" % (100 * x / sum(tot[2:])) for x in tot[2:]))
