"""Line-level view (instructions, stall samples, active lanes) of one source range of a kernel; see ncu_by_function.py.
  python tools/ncu_by_line.py src.csv x.dis <kernel-substring> <source-file> <first-line> <last-line>"""
import collections
import csv
import re
import sys

src_csv, dis, kern, srcfile, lo, hi = sys.argv[1:7]
lo, hi = int(lo), int(hi)
base = srcfile.split("/")[-1]
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l)
off2 = {}
chain, fresh, cur = [], False, None
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
            hit = [ln for f, ln in chain if f == base and lo <= ln <= hi]
            cur = hit[0] if hit else None
            fresh = False
        off2[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
h = rows[1]
ia, ie, isamp, ith = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
first = None
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in rows[2:]:
    if len(r) <= ith or not r[ie].isdigit():
        continue
    a = int(r[ia], 16)
    if first is None:
        first = a
    c = off2.get(a - first)
    if c:
        agg[c][0] += int(r[ie]); agg[c][1] += int(r[isamp]); agg[c][2] += int(r[ith])
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
text = open(srcfile).read().split("\n")
for ln, (n, s, th) in sorted(agg.items()):
    if n > ti * 0.01 or s > ts * 0.01:
        print("%5d %11d %5.1f%% samples %6d %5.1f%% lanes %4.1f  %s" % (ln, n, 100 * n / ti, s, 100 * s / max(ts, 1), th / max(n, 1), text[ln - 1].strip()[:80]))
print("range total: instr", ti, "samples", ts)
