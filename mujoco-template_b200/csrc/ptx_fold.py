"""Constant folding of exact zeros and ones in the FP64 / FP32 arithmetic of a PTX file.

The model-specialised kernels (generated/spec_*.cu) unroll the physics of one model completely; a large share of their
operands are then literal constants -- a slide joint's motion axis (0, 0, 0, 1, 0, 0), the identity rotation of a body
that cannot rotate, zero offsets.  The compiler front end may not drop `x * 0.0` or `x + 0.0` (IEEE: NaN, infinities and
the sign of zero), so about a fifth of the FP64 instructions of such a kernel multiply by a literal zero.  This pass does
what -ffast-math's no-NaNs / no-signed-zeros folding would do, on the PTX between cicc and ptxas:

    mul d, a, 0        -> mov d, 0            fma d, a, 0, c     -> mov d, c
    mul d, a, 1        -> mov d, a            fma d, a, b, 0     -> mul d, a, b
    add d, a, 0        -> mov d, a            fma d, a, 1, c     -> add d, a, c
    sub d, a, 0        -> mov d, a            sub d, 0, b        -> neg d, b
    neg d, 0           -> mov d, 0            mov d, <constant register> -> mov d, constant

and propagates the constants it creates (registers with exactly one, unpredicated definition) to a fixed point; ptxas
removes the copies and the dead code.  The results differ from the unfolded kernel only where a NaN or an infinity would
have met a zero (a diverged env: flagged from its state, not from such products) and in the sign of zeros.

    python ptx_fold.py in.ptx out.ptx        (prints the fold counts)
"""
import re
import sys

ZERO = {"f64": ("0d0000000000000000", "0d8000000000000000"), "f32": ("0f00000000", "0f80000000")}
ONE = {"f64": "0d3FF0000000000000", "f32": "0f3F800000"}
INS = re.compile(r"^(\s*)(@!?%p\d+\s+)?(mov|mul|add|sub|fma|neg)((?:\.rn|\.ftz)*)\.(f64|f32)\s+(%\w+),\s*([^;]+);(.*)$")
DEF = re.compile(r"^\s*(?:@!?%p\d+\s+)?[\w.:]+\s+(\{[^}]*\}|%\w+)")


def fold_function(lines):
    """lines: the instruction lines of one function body (modified in place); returns the number of rewrites"""
    ndef = {}
    for l in lines:
        m = DEF.match(l)
        if not m:
            continue
        for r in re.findall(r"%\w+", m.group(1)):
            ndef[r] = ndef.get(r, 0) + 1
    const = {}
    total = 0
    changed = True
    while changed:
        changed = False
        for idx, l in enumerate(lines):
            m = INS.match(l)
            if not m:
                continue
            ind, pred, op, mods, ty, dst, srcs, tail = m.groups()
            ops = [s.strip() for s in srcs.split(",")]
            val = [const.get(o, o) for o in ops]
            zero = [v in ZERO[ty] for v in val]
            one = [v == ONE[ty] for v in val]
            z0 = ZERO[ty][0]
            new = None
            if op == "mov":
                if ops[0] in const:
                    new = ("mov", [const[ops[0]]])
            elif op == "mul":
                if zero[0] or zero[1]:
                    new = ("mov", [z0])
                elif one[0]:
                    new = ("mov", [ops[1]])
                elif one[1]:
                    new = ("mov", [ops[0]])
            elif op == "add":
                if zero[0]:
                    new = ("mov", [ops[1]])
                elif zero[1]:
                    new = ("mov", [ops[0]])
            elif op == "sub":
                if zero[1]:
                    new = ("mov", [ops[0]])
                elif zero[0]:
                    new = ("neg", [ops[1]])
            elif op == "neg":
                if zero[0]:
                    new = ("mov", [z0])
            elif op == "fma":
                if zero[0] or zero[1]:
                    new = ("mov", [ops[2]])
                elif zero[2]:
                    new = ("mul", [ops[0], ops[1]])
                elif one[0]:
                    new = ("add", [ops[1], ops[2]])
                elif one[1]:
                    new = ("add", [ops[0], ops[2]])
            if new is not None:
                nop, nsrc = new
                nsrc = [const.get(s, s) for s in nsrc]
                nm = mods if nop in ("mul", "add") else (".ftz" if nop == "neg" and ".ftz" in mods else "")
                lines[idx] = "%s%s%s%s.%s \t%s, %s;%s" % (ind, pred or "", nop, nm, ty, dst, ", ".join(nsrc), tail)
                total += 1
                changed = True
                op, ops = nop, nsrc
            # a register with one unpredicated definition that is a literal is a constant from here on
            if op == "mov" and not pred and ndef.get(dst, 0) == 1 and dst not in const and re.match(r"^0[df][0-9A-Fa-f]+$", ops[0]):
                const[dst] = ops[0]
                changed = True
    return total


def fold(text):
    out, body, depth, total = [], None, 0, 0
    for line in text.split("\n"):
        if body is None:
            out.append(line)
            if line.startswith("{"):
                body = []
            continue
        if line.startswith("}"):
            total += fold_function(body)
            out.extend(body)
            out.append(line)
            body = None
        else:
            body.append(line)
    return "\n".join(out), total


if __name__ == "__main__":
    src = open(sys.argv[1]).read()
    res, n = fold(src)
    open(sys.argv[2], "w").write(res)
    print("ptx_fold: %d instructions rewritten" % n)
