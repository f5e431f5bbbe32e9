"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/b2mj.h declares.
No compute entry point is called here (that needs a GPU)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, load_model


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b2mj.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from mujoco_template import _capi

    lib = _capi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/b2mj.h but not exported"
    assert sorted(_capi.EXPORTED_SYMBOLS) == declared
    assert b"b2mj" in lib.b2_version()


def test_model_create_is_host_only_and_picks_size_class_and_specialisation():
    from mujoco_template import _capi
    import torch

    for name, cls in (("pendulum", 0), ("cartpole", 0), ("drone", 1), ("humanoid", 2)):
        nm = _capi.NativeModel(load_model(name).blob)
        assert nm.handle
    with pytest.raises(_capi.ConfigError):
        _capi.NativeModel(b"not a model blob at all")
    if not torch.cuda.is_available():
        nm = _capi.NativeModel(load_model("cartpole").blob)
        with pytest.raises(_capi.TemplateError, match="no CUDA device"):
            _capi.NativeBatch(nm, 4)


def test_product_fails_loudly_without_gpu():
    import torch

    import mujoco_template as mt

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mt.TemplateError, match="no CPU fallback"):
        mt.BatchedEnv(load_model("cartpole"), 8)
    with pytest.raises(mt.TemplateError):
        mt.Env(mt.ModelHandle(load_model("pendulum")))


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may import, load or link it."""
    pkg = os.path.join(ROOT, "mujoco-template_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|liborc|oracle/|orc_[a-z_]+\(", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not pat.search(text), f"{os.path.join(dirpath, f)} references the oracle"


def test_generated_headers_are_current():
    from mujoco_template import _layout, _specialize

    assert open(os.path.join(ROOT, "include", "b2_model_layout.h")).read() == _layout.emit_c_header()
    for name in ("pendulum", "cartpole", "drone"):
        path = os.path.join(ROOT, "mujoco-template_b200", "csrc", "generated", f"spec_{name}.cu")
        assert open(path).read() == _specialize.emit_spec(load_model(name)._c, name)


def test_jit_specialize_builds_and_registers(tmp_path, monkeypatch):
    """Run-time specialisation of a user model: nvcc cross-compiles the generated translation unit for sm_100a and the
    object's registrar resolves b2::register_spec against the loaded libb2mj.so (no GPU needed for either)."""
    import shutil
    import warnings

    import pytest

    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        pytest.skip("nvcc not available")
    from mujoco_template import _mj as mj
    from mujoco_template._specialize import jit_specialize
    from test_mjcf_compiler import ARM_XML

    monkeypatch.setenv("B2_SPEC_CACHE", str(tmp_path))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = mj.MjModel.from_xml_string(ARM_XML.replace('timestep="0.004"', 'timestep="0.0045"'))
    so = jit_specialize(model, "cpu_side_check")
    assert so is not None and os.path.exists(so) and os.path.getsize(so) > 100_000
    assert jit_specialize(model, "cpu_side_check") == so  # cached
    big = mj.MjModel.from_compiled(os.path.join(os.path.dirname(__file__), "golden", "models", "humanoid.b2m"))
    assert jit_specialize(big) is None  # nv = 27: stays on the generic / warp engine


def test_model_create_rejects_truncated_and_corrupt_blobs():
    """b2_model_create validates the whole blob before anything dereferences it (offset table inside the blob, array
    lengths equal to what the dimensions imply, index arrays in range): truncated or corrupted input gives B2_ERR_BLOB
    (ConfigError), never an out-of-bounds read."""
    import struct

    import numpy as np

    from mujoco_template import _capi, _layout

    for name in ("cartpole", "drone", "humanoid"):
        blob = bytes(load_model(name).blob)
        assert _capi.NativeModel(blob).handle
        # every truncation point in the header / table region, and a spread of points in the data region
        cuts = list(range(0, 1300, 7)) + list(np.linspace(1300, len(blob) - 1, 60).astype(int))
        for n in cuts:
            with pytest.raises(_capi.ConfigError):
                _capi.NativeModel(blob[:n])
        head = (5 + len(_layout.ISCALARS)) * 4
        head += (8 - head % 8) % 8
        table = head + 8 * sum(w for _, w in _layout.DSCALARS)
        # negative and absurd dimensions
        for i in range(len(_layout.ISCALARS)):
            for bad in (-1, 1 << 30):
                b = bytearray(blob)
                struct.pack_into("<i", b, 20 + 4 * i, bad)
                if _layout.ISCALARS[i] in ("integrator", "iterations", "ls_iterations", "has_fluid", "has_dofdamping", "disableflags",
                                           "nmocap_unused") and bad > 0:
                    continue
                with pytest.raises(_capi.ConfigError):
                    _capi.NativeModel(bytes(b))
        # a dimension that disagrees with the recorded array lengths
        b = bytearray(blob)
        struct.pack_into("<i", b, 20 + 4 * _layout.ISCALARS.index("nbody"), struct.unpack_from("<i", blob, 20 + 4 * 3)[0] + 1)
        with pytest.raises(_capi.ConfigError):
            _capi.NativeModel(bytes(b))
        # offsets pointing outside the blob, negative, or misaligned; wrong type tag
        for k in range(0, len(_layout.ARRAYS), 5):
            for field, bad in ((2, len(blob) + 8), (2, -8), (2, 4), (0, 7)):
                b = bytearray(blob)
                n = struct.unpack_from("<i", blob, table + 12 * k + 4)[0]
                if n == 0 and field == 2 and bad in (len(blob) + 8,):
                    pass  # zero-length arrays must still have an in-range offset
                struct.pack_into("<i", b, table + 12 * k + 4 * field, bad)
                with pytest.raises(_capi.ConfigError):
                    _capi.NativeModel(bytes(b))
        # out-of-range ids in the index arrays the kernels dereference
        names = [a[0] for a in _layout.ARRAYS]
        for arr in ("body_parentid", "jnt_bodyid", "jnt_qposadr", "jnt_dofadr", "dof_bodyid", "dof_parentid", "geom_bodyid", "pair_geom1",
                    "pair_geom2", "actuator_trnid"):
            k = names.index(arr)
            n, off = struct.unpack_from("<ii", blob, table + 12 * k + 4)
            if n == 0:
                continue
            for bad in (10_000, -5):
                b = bytearray(blob)
                struct.pack_into("<i", b, off + 4 * (n - 1), bad)
                with pytest.raises(_capi.ConfigError):
                    _capi.NativeModel(bytes(b))
    rng = np.random.default_rng(0)
    blob = bytes(load_model("humanoid").blob)
    for _ in range(300):  # random byte flips in the header / table: either rejected or still a consistent model -- never a crash
        b = bytearray(blob)
        for pos in rng.integers(0, 1300, 3):
            b[pos] = int(rng.integers(0, 256))
        try:
            _capi.NativeModel(bytes(b))
        except _capi.ConfigError:
            pass
