"""csrc/ptx_fold.py: the zero / one folding pass that runs on the specialised kernels' PTX (host logic, no GPU)."""
import importlib.util
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mujoco-template_b200", "csrc")


def _mod():
    spec = importlib.util.spec_from_file_location("ptx_fold", os.path.join(CSRC, "ptx_fold.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _fold(body):
    text = ".entry k()\n{\n" + "\n".join("\t" + l for l in body) + "\n}\n"
    out, n = _mod().fold(text)
    return [l.strip() for l in out.split("\n")[2:-2]], n


def _norm(line):
    return " ".join(line.replace("\t", " ").split())


def test_rewrites_of_literal_zero_and_one():
    out, n = _fold([
        "mul.f64 %fd1, %fd9, 0d0000000000000000;",
        "fma.rn.f64 %fd2, %fd9, 0d0000000000000000, %fd8;",
        "fma.rn.f64 %fd3, %fd9, %fd8, 0d0000000000000000;",
        "fma.rn.f64 %fd4, %fd9, 0d3FF0000000000000, %fd8;",
        "add.f64 %fd5, %fd9, 0d8000000000000000;",
        "sub.f64 %fd6, 0d0000000000000000, %fd9;",
        "mul.f64 %fd7, 0d3FF0000000000000, %fd9;",
        "mul.f64 %fd10, %fd9, 0d4000000000000000;",
    ])
    assert n == 7
    assert [_norm(l) for l in out] == [
        "mov.f64 %fd1, 0d0000000000000000;",
        "mov.f64 %fd2, %fd8;",
        "mul.rn.f64 %fd3, %fd9, %fd8;",
        "add.rn.f64 %fd4, %fd9, %fd8;",
        "mov.f64 %fd5, %fd9;",
        "neg.f64 %fd6, %fd9;",
        "mov.f64 %fd7, %fd9;",
        "mul.f64 %fd10, %fd9, 0d4000000000000000;",
    ]


def test_constants_propagate_to_a_fixed_point_through_single_definitions():
    out, n = _fold([
        "mov.f64 %fd1, 0d0000000000000000;",
        "mul.f64 %fd2, %fd1, %fd9;",          # 0 * x -> 0
        "fma.rn.f64 %fd3, %fd2, %fd8, %fd7;",  # (that 0) * y + z -> z
        "add.f64 %fd4, %fd3, %fd2;",          # z + 0 -> z
    ])
    assert [_norm(l) for l in out][1:] == ["mov.f64 %fd2, 0d0000000000000000;", "mov.f64 %fd3, %fd7;", "mov.f64 %fd4, %fd3;"]
    assert n == 3


def test_registers_with_several_or_predicated_definitions_are_not_constants():
    body = [
        "mov.f64 %fd1, 0d0000000000000000;",
        "@%p1 mov.f64 %fd1, %fd9;",            # a second (predicated) definition: %fd1 is not a constant
        "mul.f64 %fd2, %fd1, %fd8;",
        "@%p2 mov.f64 %fd3, 0d0000000000000000;",
        "mul.f64 %fd4, %fd3, %fd8;",
        "ld.global.f64 %fd5, [%rd1];",
        "mul.f64 %fd6, %fd5, %fd8;",
    ]
    out, n = _fold(body)
    assert n == 0 and [_norm(l) for l in out] == [_norm(l) for l in body]


def test_f32_and_function_scope():
    m = _mod()
    text = (".entry a()\n{\n\tmov.f32 %f1, 0f00000000;\n\tmul.f32 %f2, %f1, %f9;\n}\n"
            ".entry b()\n{\n\tmul.f32 %f2, %f1, %f9;\n}\n")     # %f1 of b() is a different register
    out, n = m.fold(text)
    assert n == 1
    assert "mov.f32 \t%f2, 0f00000000;" in out.split(".entry b()")[0]
    assert _norm(out.split(".entry b()")[1].split("\n")[2]) == "mul.f32 %f2, %f1, %f9;"


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc")
def test_fold_build_produces_an_object_and_falls_back_cleanly(tmp_path):
    src = tmp_path / "t.cu"
    src.write_text("__global__ void k(const double* x, double* y) { double a[3] = {0.0, 1.0, 0.0}; double s = 0;\n"
                   "  for (int i = 0; i < 3; i++) s += a[i] * x[i]; y[0] = s; }\n")
    out = tmp_path / "t.o"
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-Xptxas", "-v"]
    r = subprocess.run(["python3", os.path.join(CSRC, "fold_build.py"), "nvcc", str(src), str(out)] + flags, capture_output=True, text=True)
    assert r.returncode == 0 and out.exists(), r.stderr
    assert "ptx_fold:" in r.stderr
    sass = subprocess.run(["cuobjdump", "-sass", str(out)], capture_output=True, text=True).stdout
    assert "DFMA" not in sass and "DMUL" not in sass  # y = x[1]: nothing left to multiply
    out.unlink()
    r = subprocess.run(["python3", os.path.join(CSRC, "fold_build.py"), "nvcc", str(src), str(out)] + flags, capture_output=True, text=True,
                       env=dict(os.environ, B2_PTX_FOLD="0"))
    assert r.returncode == 0 and out.exists() and "ptx_fold:" not in r.stderr
