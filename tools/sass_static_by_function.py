"""Static SASS instruction count of one kernel per innermost source function (code-size view; the warp engine is bound by
instruction fetch).   nvdisasm --print-line-info-inline x.cubin > x.dis
  python tools/sass_static_by_function.py x.dis <kernel-substring> <source-file> [top-lines]"""
import collections
import re
import sys

dis, kern, srcfile = sys.argv[1:4]
ntop = int(sys.argv[4]) if len(sys.argv) > 4 else 0
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
chain, fresh, cur = [], False, None
cnt, byline = collections.Counter(), collections.Counter()
for l in lines[start + 1:]:
    if l.startswith(".text."):
        break
    m = re.search(r'//## File ".*?/([\w.]+)", line (\d+)', l)
    if m:
        if not fresh:
            chain, fresh = [], True
        chain.append((m.group(1), int(m.group(2))))
        continue
    if re.match(r"\s+/\*[0-9a-f]+\*/\s+", l):
        if fresh:
            eng = [ln for f, ln in chain if f == base]
            cur = (fn_of(eng[0]) if eng else "other:" + chain[0][0], eng[0] if eng else 0)
            fresh = False
        cnt[cur[0]] += 1
        byline[cur] += 1
print("total", sum(cnt.values()))
for k, v in cnt.most_common(30):
    print("%6d %s" % (v, k))
if ntop:
    src = open(srcfile).read().split("\n")
    print("--- top lines")
    for k, v in byline.most_common(ntop):
        print("%6d %s:%d  %s" % (v, k[0], k[1], src[k[1] - 1].strip()[:90] if k[1] else ""))
