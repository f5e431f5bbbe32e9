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
    A("/* returns 0 on success, nonzero on a malformed blob.  Everything is checked against nbytes before it is read:")
    A(" * header and offset table inside the blob, dimensions non-negative and bounded, every array's recorded type and")
    A(" * length equal to what the dimensions imply, its bytes inside the blob, and every index array within range of the")
    A(" * arrays it indexes (codes: 1 magic, 2 version, 3 field counts, 4 truncated, 5 dimension, 6 table entry, 7 index). */")
    A("static inline int b2m_view_init(b2m_view* v, const void* blob, size_t nbytes) {")
    A("  const unsigned char* base = (const unsigned char*)blob;")
    A("  const int32_t* h = (const int32_t*)blob;")
    A("  if (nbytes < 20 || h[0] != (int32_t)B2M_MAGIC) return 1;")
    A("  if (h[1] != B2M_VERSION) return 2;")
    A(f"  if (h[2] != {len(ISCALARS)} || h[3] != {sum(w for _, w in DSCALARS)} || h[4] != {len(ARRAYS)}) return 3;")
    nd = sum(w for _, w in DSCALARS)
    off0 = (5 + len(ISCALARS)) * 4
    off0 += (8 - off0 % 8) % 8
    table_end = off0 + nd * 8 + len(ARRAYS) * 12
    A(f"  if (nbytes < {table_end}) return 4;  /* scalars + offset table */")
    A("  const int32_t* is = h + 5;")
    A(f"  for (int i = 0; i < {len(ISCALARS)}; i++) if (is[i] < 0 || is[i] > (1 << 20)) return 5;")
    for i, k in enumerate(ISCALARS):
        A(f"  v->{k} = is[{i}];")
    A(f"  size_t off = {off0};")
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
    A("#define B2M_ARRAY(i, field, ctype, isdouble, count)                                                     \\")
    A("  do {                                                                                                \\")
    A("    const long long n_ = (long long)(count);                                                          \\")
    A("    const int32_t o_ = tab[3 * (i) + 2];                                                              \\")
    A("    if (tab[3 * (i)] != (isdouble) || (long long)tab[3 * (i) + 1] != n_ || o_ < 0 || (o_ & 7)) return 6; \\")
    A("    if ((unsigned long long)o_ + (unsigned long long)n_ * sizeof(ctype) > (unsigned long long)nbytes) return 4; \\")
    A("    v->field = (const ctype*)(base + o_);                                                              \\")
    A("  } while (0)")
    for i, (name, dt, w, cnt) in enumerate(ARRAYS):
        ctype = "int" if dt == "i" else "double"
        width = f"v->{w}" if isinstance(w, str) else str(w)
        A(f"  B2M_ARRAY({i}, {name}, {ctype}, {0 if dt == 'i' else 1}, (long long)v->{cnt} * {width});")
    A("#undef B2M_ARRAY")
    A("  /* index arrays: every id the kernels dereference must lie inside the array it indexes */")
    A("#define B2M_RANGE(arr, n, lo, hi) for (int i_ = 0; i_ < (n); i_++) if (v->arr[i_] < (lo) || v->arr[i_] >= (hi)) return 7")
    A("  if (v->nbody < 1 || v->nsensordata > (1 << 16)) return 5;")
    A("  for (int i = 0; i < v->nbody; i++) {")
    A("    if (v->body_parentid[i] < 0 || v->body_parentid[i] > (i ? i - 1 : 0)) return 7;")
    A("    if (v->body_jntnum[i] < 0 || v->body_dofnum[i] < 0) return 7;")
    A("    if (v->body_jntnum[i] && (v->body_jntadr[i] < 0 || v->body_jntadr[i] + v->body_jntnum[i] > v->njnt)) return 7;")
    A("    if (v->body_dofnum[i] && (v->body_dofadr[i] < 0 || v->body_dofadr[i] + v->body_dofnum[i] > v->nv)) return 7;")
    A("  }")
    A("  B2M_RANGE(body_rootid, v->nbody, 0, v->nbody);")
    A("  B2M_RANGE(body_weldid, v->nbody, 0, v->nbody);")
    A("  B2M_RANGE(jnt_type, v->njnt, 0, 4);")
    A("  B2M_RANGE(jnt_bodyid, v->njnt, 0, v->nbody);")
    A("  for (int j = 0; j < v->njnt; j++) {")
    A("    const int t = v->jnt_type[j], wq = t == 0 ? 7 : (t == 1 ? 4 : 1), wv = t == 0 ? 6 : (t == 1 ? 3 : 1);")
    A("    if (v->jnt_qposadr[j] < 0 || v->jnt_qposadr[j] + wq > v->nq || v->jnt_dofadr[j] < 0 || v->jnt_dofadr[j] + wv > v->nv) return 7;")
    A("  }")
    A("  B2M_RANGE(dof_bodyid, v->nv, 0, v->nbody);")
    A("  B2M_RANGE(dof_jntid, v->nv, 0, v->njnt);")
    A("  for (int i = 0; i < v->nv; i++) if (v->dof_parentid[i] < -1 || v->dof_parentid[i] >= i) return 7;")
    A("  B2M_RANGE(geom_bodyid, v->ngeom, 0, v->nbody);")
    A("  B2M_RANGE(geom_type, v->ngeom, 0, 8);")
    A("  B2M_RANGE(site_bodyid, v->nsite, 0, v->nbody);")
    A("  for (int t = 0; t < v->ntendon; t++)")
    A("    if (v->tendon_adr[t] < 0 || v->tendon_num[t] < 0 || v->tendon_adr[t] + v->tendon_num[t] > v->nwrap) return 7;")
    A("  B2M_RANGE(wrap_jntid, v->nwrap, 0, v->njnt);")
    A("  for (int a = 0; a < v->nu; a++) {")
    A("    const int t = v->actuator_trntype[a], id = v->actuator_trnid[a];")
    A("    if (t == 0 ? (id < 0 || id >= v->njnt) : (t == 4 ? (id < 0 || id >= v->nsite) : 1)) return 7;")
    A("  }")
    A("  B2M_RANGE(pair_geom1, v->npair, 0, v->ngeom);")
    A("  B2M_RANGE(pair_geom2, v->npair, 0, v->ngeom);")
    A("  B2M_RANGE(pair_dim, v->npair, 1, 7);")
    A("  for (int s = 0; s < v->nsensor; s++) {")
    A("    const int ot = v->sensor_objtype[s], id = v->sensor_objid[s];")
    A("    const int lim = ot == 1 || ot == 2 ? v->nbody : (ot == 3 ? v->njnt : (ot == 5 ? v->ngeom : (ot == 6 ? v->nsite : 0)));")
    A("    if (id < 0 || id >= lim) return 7;")
    A("    if (v->sensor_dim[s] < 0 || v->sensor_adr[s] < 0 || v->sensor_adr[s] + v->sensor_dim[s] > v->nsensordata) return 7;")
    A("  }")
    A("#undef B2M_RANGE")
    A("  return 0;")
    A("}")
    A("#endif")
    return "\n".join(L) + "\n"
