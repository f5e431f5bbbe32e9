"""Binary layout of a compiled model ("blob") shared by the CUDA library and the CPU oracle.

The blob is the only thing that crosses the C-ABI at model-creation time
(``b2_model_create``).  It plays the role of MuJoCo's ``mjModel`` for the hot path:
the reference owns an ``mj.MjModel`` (reference ``mujoco_template/model.py:14-25``);
we own a flat, versioned, position-independent byte string.

The field list below is the single source of truth; ``emit_c_header()`` generates
``include/b2_model_layout.h`` from it (a test checks the committed header is current).
"""

from __future__ import annotations

import struct

import numpy as np

MAGIC = 0x4A4D3242  # "B2MJ"
VERSION = 4

# integer scalars, in blob order
ISCALARS = (
    "nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite", "ntendon", "nwrap", "nkey",
    "npair", "integrator", "iterations", "ls_iterations", "has_fluid", "has_dofdamping",
    "disableflags", "nmocap_unused", "nsensor", "nsensordata",
)

# double scalars / small fixed vectors, in blob order: (name, width)
DSCALARS = (
    ("timestep", 1), ("gravity", 3), ("wind", 3), ("density", 1), ("viscosity", 1),
    ("tolerance", 1), ("ls_tolerance", 1), ("impratio", 1), ("meaninertia", 1),
)

# arrays: (name, dtype 'i'|'d', width per item, count key)
ARRAYS = (
    # bodies
    ("body_parentid", "i", 1, "nbody"), ("body_rootid", "i", 1, "nbody"),
    ("body_weldid", "i", 1, "nbody"), ("body_jntnum", "i", 1, "nbody"),
    ("body_jntadr", "i", 1, "nbody"), ("body_dofnum", "i", 1, "nbody"),
    ("body_dofadr", "i", 1, "nbody"),
    ("body_pos", "d", 3, "nbody"), ("body_quat", "d", 4, "nbody"),
    ("body_ipos", "d", 3, "nbody"), ("body_iquat", "d", 4, "nbody"),
    ("body_mass", "d", 1, "nbody"), ("body_subtreemass", "d", 1, "nbody"),
    ("body_inertia", "d", 3, "nbody"), ("body_invweight0", "d", 2, "nbody"),
    # joints
    ("jnt_type", "i", 1, "njnt"), ("jnt_qposadr", "i", 1, "njnt"),
    ("jnt_dofadr", "i", 1, "njnt"), ("jnt_bodyid", "i", 1, "njnt"),
    ("jnt_limited", "i", 1, "njnt"),
    ("jnt_pos", "d", 3, "njnt"), ("jnt_axis", "d", 3, "njnt"),
    ("jnt_stiffness", "d", 1, "njnt"), ("jnt_range", "d", 2, "njnt"),
    ("jnt_margin", "d", 1, "njnt"), ("jnt_solref", "d", 2, "njnt"),
    ("jnt_solimp", "d", 5, "njnt"),
    ("qpos0", "d", 1, "nq"), ("qpos_spring", "d", 1, "nq"),
    # dofs
    ("dof_bodyid", "i", 1, "nv"), ("dof_jntid", "i", 1, "nv"), ("dof_parentid", "i", 1, "nv"),
    ("dof_armature", "d", 1, "nv"), ("dof_damping", "d", 1, "nv"),
    ("dof_invweight0", "d", 1, "nv"),
    # geoms
    ("geom_type", "i", 1, "ngeom"), ("geom_bodyid", "i", 1, "ngeom"),
    ("geom_size", "d", 3, "ngeom"), ("geom_rbound", "d", 1, "ngeom"),
    ("geom_pos", "d", 3, "ngeom"), ("geom_quat", "d", 4, "ngeom"),
    # sites
    ("site_bodyid", "i", 1, "nsite"), ("site_pos", "d", 3, "nsite"),
    ("site_quat", "d", 4, "nsite"),
    # fixed tendons
    ("tendon_adr", "i", 1, "ntendon"), ("tendon_num", "i", 1, "ntendon"),
    ("tendon_limited", "i", 1, "ntendon"),
    ("tendon_range", "d", 2, "ntendon"), ("tendon_margin", "d", 1, "ntendon"),
    ("tendon_solref", "d", 2, "ntendon"), ("tendon_solimp", "d", 5, "ntendon"),
    ("tendon_invweight0", "d", 1, "ntendon"), ("tendon_stiffness", "d", 1, "ntendon"),
    ("tendon_damping", "d", 1, "ntendon"), ("tendon_lengthspring", "d", 2, "ntendon"),
    ("wrap_jntid", "i", 1, "nwrap"), ("wrap_coef", "d", 1, "nwrap"),
    # actuators
    ("actuator_trntype", "i", 1, "nu"), ("actuator_trnid", "i", 1, "nu"),
    ("actuator_ctrllimited", "i", 1, "nu"), ("actuator_forcelimited", "i", 1, "nu"),
    ("actuator_disabled", "i", 1, "nu"),
    ("actuator_gear", "d", 6, "nu"), ("actuator_ctrlrange", "d", 2, "nu"),
    ("actuator_forcerange", "d", 2, "nu"), ("actuator_gainprm", "d", 1, "nu"),
    ("actuator_biasprm", "d", 3, "nu"),
    # keyframes
    ("key_time", "d", 1, "nkey"), ("key_qpos", "d", "nq", "nkey"),
    ("key_qvel", "d", "nv", "nkey"), ("key_ctrl", "d", "nu", "nkey"),
    # statically filtered collision candidates with pre-mixed contact parameters
    ("pair_geom1", "i", 1, "npair"), ("pair_geom2", "i", 1, "npair"),
    ("pair_dim", "i", 1, "npair"),
    ("pair_margin", "d", 1, "npair"), ("pair_gap", "d", 1, "npair"),
    ("pair_friction", "d", 5, "npair"), ("pair_solref", "d", 2, "npair"),
    ("pair_solimp", "d", 5, "npair"),
    # sensors (type codes: mjcf.SENSOR_TYPES; objtype follows mjtObj: 1 body, 2 xbody, 3 joint, 5 geom, 6 site)
    ("sensor_type", "i", 1, "nsensor"), ("sensor_objtype", "i", 1, "nsensor"),
    ("sensor_objid", "i", 1, "nsensor"), ("sensor_adr", "i", 1, "nsensor"),
    ("sensor_dim", "i", 1, "nsensor"), ("sensor_cutoff", "d", 1, "nsensor"),
)


def _width(model: dict, w) -> int:
    return int(model[w]) if isinstance(w, str) else int(w)


def pack(model: dict) -> bytes:
    """Serialise a compiled-model dict (see ``mjcf.compile_model``) to the blob format."""
    head = [MAGIC, VERSION, len(ISCALARS), sum(w for _, w in DSCALARS), len(ARRAYS)]
    isc = [int(model.get(k, 0)) for k in ISCALARS]
    dsc: list[float] = []
    for name, w in DSCALARS:
        v = np.atleast_1d(np.asarray(model[name], dtype=np.float64)).ravel()
        if v.size != w:
            raise ValueError(f"{name}: expected {w} values, got {v.size}")
        dsc.extend(float(x) for x in v)
    pre = struct.pack(f"<{len(head)}i", *head) + struct.pack(f"<{len(isc)}i", *isc)
    if len(pre) % 8:
        pre += b"\0" * (8 - len(pre) % 8)
    pre += struct.pack(f"<{len(dsc)}d", *dsc)
    table_off = len(pre)
    table_bytes = len(ARRAYS) * 3 * 4
    data_off = table_off + table_bytes
    if data_off % 8:
        data_off += 8 - data_off % 8
    table: list[int] = []
    chunks: list[bytes] = []
    cursor = data_off
    for name, dt, w, cnt in ARRAYS:
        n = int(model[cnt]) * _width(model, w)
        arr = np.asarray(model[name], dtype=np.int32 if dt == "i" else np.float64).ravel()
        if arr.size != n:
            raise ValueError(f"{name}: expected {n} values, got {arr.size}")
        raw = arr.tobytes()
        if len(raw) % 8:
            raw += b"\0" * (8 - len(raw) % 8)
        table.extend([0 if dt == "i" else 1, n, cursor])
        chunks.append(raw)
        cursor += len(raw)
    blob = pre + struct.pack(f"<{len(table)}i", *table)
    blob += b"\0" * (data_off - len(blob))
    blob += b"".join(chunks)
    assert len(blob) == cursor
    return blob


def emit_c_header() -> str:
    """Generate include/b2_model_layout.h (plain C, usable from gcc and nvcc)."""
    L: list[str] = []
    A = L.append
    A("/* GENERATED by mujoco_template/_layout.py (emit_c_header) -- do not edit by hand.")
    A(" * Flat compiled-model blob: the data the hot path needs from an MJCF model.")
    A(" * Replaces the reference's mj.MjModel ownership (reference mujoco_template/model.py:14-25). */")
    A("#ifndef B2_MODEL_LAYOUT_H")
    A("#define B2_MODEL_LAYOUT_H")
    A("#include <stddef.h>")
    A("#include <stdint.h>")
    A("#include <string.h>")
    A(f"#define B2M_MAGIC 0x{MAGIC:08X}")
    A(f"#define B2M_VERSION {VERSION}")
    A("typedef struct b2m_view {")
    for k in ISCALARS:
        A(f"  int {k};")
    for name, w in DSCALARS:
        A(f"  double {name}{'[%d]' % w if w > 1 else ''};")
    for name, dt, w, cnt in ARRAYS:
        A(f"  const {'int' if dt == 'i' else 'double'}* {name};  /* {cnt} x {w} */")
    A("} b2m_view;")
    A("")
    A("/* returns 0 on success, nonzero on malformed blob */")
    A("static inline int b2m_view_init(b2m_view* v, const void* blob, size_t nbytes) {")
    A("  const unsigned char* base = (const unsigned char*)blob;")
    A("  const int32_t* h = (const int32_t*)blob;")
    A("  if (nbytes < 20 || h[0] != (int32_t)B2M_MAGIC) return 1;")
    A("  if (h[1] != B2M_VERSION) return 2;")
    A(f"  if (h[2] != {len(ISCALARS)} || h[3] != {sum(w for _, w in DSCALARS)} || h[4] != {len(ARRAYS)}) return 3;")
    A("  const int32_t* is = h + 5;")
    for i, k in enumerate(ISCALARS):
        A(f"  v->{k} = is[{i}];")
    A(f"  size_t off = (5 + {len(ISCALARS)}) * 4; if (off % 8) off += 8 - off % 8;")
    A("  const double* ds = (const double*)(base + off);")
    j = 0
    for name, w in DSCALARS:
        if w == 1:
            A(f"  v->{name} = ds[{j}];")
        else:
            for t in range(w):
                A(f"  v->{name}[{t}] = ds[{j + t}];")
        j += w
    A(f"  off += {j} * 8;")
    A("  const int32_t* tab = (const int32_t*)(base + off);")
    for i, (name, dt, w, cnt) in enumerate(ARRAYS):
        ctype = "int" if dt == "i" else "double"
        A(f"  if ((size_t)tab[{3 * i + 2}] > nbytes) return 4;")
        A(f"  v->{name} = (const {ctype}*)(base + tab[{3 * i + 2}]);")
    A("  return 0;")
    A("}")
    A("#endif")
    return "\n".join(L) + "\n"
