"""ctypes binding of the C ABI in include/b200sim.h.

This is the reference-side stub a maintainer would add where `mjlab.sim.sim`
imports `mujoco_warp` (src/mjlab/sim/sim.py:5,139-149,184-190): field tables that
mirror the header enums, the two ctypes structs, derived integer tables for the
device kernels, and a loader that FAILS LOUDLY when the CUDA library is missing
(there is no CPU fallback on the product path).
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libb200sim.so"

# (enum name, ModelDesc attribute, row length as a function of the model)
MODEL_FLOAT_FIELDS = [
  ("OPT_TIMESTEP", "opt_timestep", lambda m: 1),
  ("OPT_GRAVITY", "opt_gravity", lambda m: 3),
  ("OPT_TOLERANCE", "opt_tolerance", lambda m: 1),
  ("OPT_LS_TOLERANCE", "opt_ls_tolerance", lambda m: 1),
  ("OPT_IMPRATIO", "opt_impratio", lambda m: 1),
  ("STAT_MEANINERTIA", "stat_meaninertia", lambda m: 1),
  ("QPOS0", "qpos0", lambda m: m.nq),
  ("QPOS_SPRING", "qpos_spring", lambda m: m.nq),
  ("BODY_POS", "body_pos", lambda m: m.nbody * 3),
  ("BODY_QUAT", "body_quat", lambda m: m.nbody * 4),
  ("BODY_IPOS", "body_ipos", lambda m: m.nbody * 3),
  ("BODY_IQUAT", "body_iquat", lambda m: m.nbody * 4),
  ("BODY_MASS", "body_mass", lambda m: m.nbody),
  ("BODY_SUBTREEMASS", "body_subtreemass", lambda m: m.nbody),
  ("BODY_INERTIA", "body_inertia", lambda m: m.nbody * 3),
  ("BODY_INVWEIGHT0", "body_invweight0", lambda m: m.nbody * 2),
  ("JNT_POS", "jnt_pos", lambda m: m.njnt * 3),
  ("JNT_AXIS", "jnt_axis", lambda m: m.njnt * 3),
  ("JNT_RANGE", "jnt_range", lambda m: m.njnt * 2),
  ("JNT_SOLREF", "jnt_solref", lambda m: m.njnt * 2),
  ("JNT_SOLIMP", "jnt_solimp", lambda m: m.njnt * 5),
  ("JNT_MARGIN", "jnt_margin", lambda m: m.njnt),
  ("JNT_STIFFNESS", "jnt_stiffness", lambda m: m.njnt),
  ("DOF_ARMATURE", "dof_armature", lambda m: m.nv),
  ("DOF_DAMPING", "dof_damping", lambda m: m.nv),
  ("DOF_INVWEIGHT0", "dof_invweight0", lambda m: m.nv),
  ("GEOM_POS", "geom_pos", lambda m: m.ngeom * 3),
  ("GEOM_QUAT", "geom_quat", lambda m: m.ngeom * 4),
  ("GEOM_SIZE", "geom_size", lambda m: m.ngeom * 3),
  ("GEOM_RBOUND", "geom_rbound", lambda m: m.ngeom),
  ("GEOM_FRICTION", "geom_friction", lambda m: m.ngeom * 3),
  ("GEOM_SOLREF", "geom_solref", lambda m: m.ngeom * 2),
  ("GEOM_SOLIMP", "geom_solimp", lambda m: m.ngeom * 5),
  ("GEOM_SOLMIX", "geom_solmix", lambda m: m.ngeom),
  ("GEOM_MARGIN", "geom_margin", lambda m: m.ngeom),
  ("GEOM_GAP", "geom_gap", lambda m: m.ngeom),
  ("SITE_POS", "site_pos", lambda m: m.nsite * 3),
  ("SITE_QUAT", "site_quat", lambda m: m.nsite * 4),
  ("ACTUATOR_GAINPRM", "actuator_gainprm", lambda m: m.nu * 10),
  ("ACTUATOR_BIASPRM", "actuator_biasprm", lambda m: m.nu * 10),
  ("ACTUATOR_GEAR", "actuator_gear", lambda m: m.nu * 6),
  ("ACTUATOR_FORCERANGE", "actuator_forcerange", lambda m: m.nu * 2),
  ("ACTUATOR_CTRLRANGE", "actuator_ctrlrange", lambda m: m.nu * 2),
  ("SENSOR_CUTOFF", "sensor_cutoff", lambda m: m.nsensor),
  ("SGEOM_XPOS", "sgeom_xpos", lambda m: m.ngeom * 3),
  ("SGEOM_XMAT", "sgeom_xmat", lambda m: m.ngeom * 9),
  ("FIELD_AABB", "field_aabb", lambda m: m.nfield * 6),
  ("FIELD_GRID", "field_grid", lambda m: 4),
  ("DOF_FRICTIONLOSS", "dof_frictionloss", lambda m: m.nv),
  ("DOF_SOLREF", "dof_solref", lambda m: m.nv * 2),
  ("DOF_SOLIMP", "dof_solimp", lambda m: m.nv * 5),
]

MODEL_INT_FIELDS = [
  ("BODY_PARENTID", "body_parentid"), ("BODY_ROOTID", "body_rootid"), ("BODY_MOCAPID", "body_mocapid"),
  ("BODY_JNTNUM", "body_jntnum"), ("BODY_JNTADR", "body_jntadr"), ("BODY_DOFNUM", "body_dofnum"),
  ("BODY_DOFADR", "body_dofadr"), ("BODY_LASTDOF", "body_lastdof"), ("BODY_DOFMASK", "body_dofmask"),
  ("LEVEL_ADR", "level_adr"), ("LEVEL_BODY", "level_body"), ("CHILD_ADR", "child_adr"), ("CHILD_BODY", "child_body"),
  ("JNT_TYPE", "jnt_type"), ("JNT_QPOSADR", "jnt_qposadr"), ("JNT_DOFADR", "jnt_dofadr"), ("JNT_BODYID", "jnt_bodyid"),
  ("JNT_LIMITED", "jnt_limited"), ("DOF_BODYID", "dof_bodyid"), ("DOF_JNTID", "dof_jntid"), ("DOF_PARENTID", "dof_parentid"),
  ("GEOM_TYPE", "geom_type"), ("GEOM_BODYID", "geom_bodyid"), ("GEOM_CONDIM", "geom_condim"),
  ("GEOM_PRIORITY", "geom_priority"), ("SITE_BODYID", "site_bodyid"), ("PAIR_GEOM", "pair_geom"),
  ("ACTUATOR_TRNID", "actuator_trnid"), ("ACTUATOR_BIASTYPE", "actuator_biastype"),
  ("ACTUATOR_FORCELIMITED", "actuator_forcelimited"), ("ACTUATOR_CTRLLIMITED", "actuator_ctrllimited"),
  ("SENSOR_TYPE", "sensor_type"), ("SENSOR_OBJTYPE", "sensor_objtype"), ("SENSOR_OBJID", "sensor_objid"),
  ("SENSOR_REFTYPE", "sensor_reftype"), ("SENSOR_REFID", "sensor_refid"), ("SENSOR_INTPRM", "sensor_intprm"),
  ("SENSOR_DIM", "sensor_dim"), ("SENSOR_ADR", "sensor_adr"), ("DOF_ACTID", "dof_actid"),
  ("BODY_ANCMASK", "body_ancmask"), ("DYN_GEOM", "dyn_geom"), ("FIELD_GEOM", "field_geom"), ("FIELD_CELL0", "field_cell0"),
  ("FIELD_CELL_ADR", "field_cell_adr"), ("FIELD_CELL_ITEM", "field_cell_item"), ("FQ_GEOM", "fq_geom"),
]

# (enum name, attribute on Data, per-world shape as a function of (model, nconmax, njmax))
DATA_FLOAT_FIELDS = [
  ("QPOS", "qpos", lambda m, c, j: (m.nq,)),
  ("QVEL", "qvel", lambda m, c, j: (m.nv,)),
  ("CTRL", "ctrl", lambda m, c, j: (m.nu,)),
  ("QACC", "qacc", lambda m, c, j: (m.nv,)),
  ("QACC_WARMSTART", "qacc_warmstart", lambda m, c, j: (m.nv,)),
  ("QFRC_APPLIED", "qfrc_applied", lambda m, c, j: (m.nv,)),
  ("XFRC_APPLIED", "xfrc_applied", lambda m, c, j: (m.nbody, 6)),
  ("MOCAP_POS", "mocap_pos", lambda m, c, j: (m.nmocap, 3)),
  ("MOCAP_QUAT", "mocap_quat", lambda m, c, j: (m.nmocap, 4)),
  ("TIME", "time", lambda m, c, j: ()),
  ("XPOS", "xpos", lambda m, c, j: (m.nbody, 3)),
  ("XQUAT", "xquat", lambda m, c, j: (m.nbody, 4)),
  ("XMAT", "xmat", lambda m, c, j: (m.nbody, 3, 3)),
  ("XIPOS", "xipos", lambda m, c, j: (m.nbody, 3)),
  ("XIMAT", "ximat", lambda m, c, j: (m.nbody, 3, 3)),
  ("XANCHOR", "xanchor", lambda m, c, j: (m.njnt, 3)),
  ("XAXIS", "xaxis", lambda m, c, j: (m.njnt, 3)),
  ("GEOM_XPOS", "geom_xpos", lambda m, c, j: (m.ngeom, 3)),
  ("GEOM_XMAT", "geom_xmat", lambda m, c, j: (m.ngeom, 3, 3)),
  ("SITE_XPOS", "site_xpos", lambda m, c, j: (m.nsite, 3)),
  ("SITE_XMAT", "site_xmat", lambda m, c, j: (m.nsite, 3, 3)),
  ("SUBTREE_COM", "subtree_com", lambda m, c, j: (m.nbody, 3)),
  ("CINERT", "cinert", lambda m, c, j: (m.nbody, 10)),
  ("CDOF", "cdof", lambda m, c, j: (m.nv, 6)),
  ("CDOF_DOT", "cdof_dot", lambda m, c, j: (m.nv, 6)),
  ("CVEL", "cvel", lambda m, c, j: (m.nbody, 6)),
  ("CACC", "cacc", lambda m, c, j: (m.nbody, 6)),
  ("QM", "qM", lambda m, c, j: (m.nv, leading_dim(m.nv))),  # padded rows; the API view slices [..., :nv]
  ("QFRC_BIAS", "qfrc_bias", lambda m, c, j: (m.nv,)),
  ("QFRC_PASSIVE", "qfrc_passive", lambda m, c, j: (m.nv,)),
  ("QFRC_ACTUATOR", "qfrc_actuator", lambda m, c, j: (m.nv,)),
  ("QFRC_SMOOTH", "qfrc_smooth", lambda m, c, j: (m.nv,)),
  ("QACC_SMOOTH", "qacc_smooth", lambda m, c, j: (m.nv,)),
  ("QFRC_CONSTRAINT", "qfrc_constraint", lambda m, c, j: (m.nv,)),
  ("ACTUATOR_FORCE", "actuator_force", lambda m, c, j: (m.nu,)),
  ("ACTUATOR_LENGTH", "actuator_length", lambda m, c, j: (m.nu,)),
  ("ACTUATOR_VELOCITY", "actuator_velocity", lambda m, c, j: (m.nu,)),
  ("SENSORDATA", "sensordata", lambda m, c, j: (m.nsensordata,)),
  ("SUBTREE_LINVEL", "subtree_linvel", lambda m, c, j: (m.nbody, 3)),
  ("SUBTREE_ANGMOM", "subtree_angmom", lambda m, c, j: (m.nbody, 3)),
  ("CONTACT_DIST", "contact_dist", lambda m, c, j: (c,)),
  ("CONTACT_POS", "contact_pos", lambda m, c, j: (c, 3)),
  ("CONTACT_FRAME", "contact_frame", lambda m, c, j: (c, 9)),
  ("CONTACT_INCLUDEMARGIN", "contact_includemargin", lambda m, c, j: (c,)),
  ("CONTACT_FRICTION", "contact_friction", lambda m, c, j: (c, 5)),
  ("CONTACT_SOLREF", "contact_solref", lambda m, c, j: (c, 2)),
  ("CONTACT_SOLIMP", "contact_solimp", lambda m, c, j: (c, 5)),
  ("EFC_J", "efc_J", lambda m, c, j: (j, m.nv)),
  ("EFC_D", "efc_D", lambda m, c, j: (j,)),
  ("EFC_AREF", "efc_aref", lambda m, c, j: (j,)),
  ("EFC_POS", "efc_pos", lambda m, c, j: (j,)),
  ("EFC_MARGIN", "efc_margin", lambda m, c, j: (j,)),
  ("EFC_FORCE", "efc_force", lambda m, c, j: (j,)),
  ("MEMO_QPOS", "memo_qpos", lambda m, c, j: (m.nq,)),
  ("MEMO_QVEL", "memo_qvel", lambda m, c, j: (m.nv,)),
]

DATA_INT_FIELDS = [
  ("NCON", "ncon", lambda m, c, j: ()),
  ("NEFC", "nefc", lambda m, c, j: ()),
  ("NL", "nl", lambda m, c, j: ()),
  ("SOLVER_NITER", "solver_niter", lambda m, c, j: ()),
  ("CONTACT_GEOM", "contact_geom", lambda m, c, j: (c, 2)),
  ("CONTACT_DIM", "contact_dim", lambda m, c, j: (c,)),
  ("CONTACT_EFC_ADDRESS", "contact_efc_address", lambda m, c, j: (c,)),
  ("NCON_DROPPED", "ncon_dropped", lambda m, c, j: ()),
  ("NEFC_DROPPED", "nefc_dropped", lambda m, c, j: ()),
  ("SOLVE_ORDER", "solve_order", lambda m, c, j: ()),
  ("SOLVE_NEXT", "solve_next", lambda m, c, j: ()),
  ("SOLVE_RING", "solve_ring", lambda m, c, j: ()),
  ("SOLVE_CTR", "solve_ctr", lambda m, c, j: (4,)),
  ("SMOOTH_MEMO", "smooth_memo", lambda m, c, j: ()),
  ("NF", "nf", lambda m, c, j: ()),
]

def leading_dim(n: int) -> int:
  """Row stride of dense nv x nv matrices in the engine (b2_ld in csrc/b2_common.cuh)."""
  ld = (n + 3) & ~3
  if (ld & 7) != 4:
    ld += 4
  return ld


NMF, NMI, NDF, NDI = len(MODEL_FLOAT_FIELDS), len(MODEL_INT_FIELDS), len(DATA_FLOAT_FIELDS), len(DATA_INT_FIELDS)


class B2Model(C.Structure):
  _fields_ = (
    [(n, C.c_int) for n in ("nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite", "nsensor", "nsensordata", "npair",
                            "nmocap", "nlevel", "integrator", "cone", "solver", "iterations", "ls_iterations",
                            "disableflags", "has_accelerometer", "has_subtreevel", "nlimited", "maxchain", "nfriction",
                            "nfield", "nfq", "ndyn", "field_nx", "field_ny")]
    + [("mf", C.c_void_p * NMF), ("mf_stride", C.c_int * NMF), ("mi", C.c_void_p * NMI)]
  )


class B2Data(C.Structure):
  _fields_ = [("nworld", C.c_int), ("nconmax", C.c_int), ("njmax", C.c_int), ("flags", C.c_int), ("df", C.c_void_p * NDF), ("di", C.c_void_p * NDI)]


def derived_int_tables(m) -> dict[str, np.ndarray]:
  """Tree-level / child CSR tables and ancestor-dof masks the kernels walk."""
  nb = m.nbody
  level = np.zeros(nb, np.int64)
  for i in range(1, nb):
    level[i] = level[m.body_parentid[i]] + 1
  nlevel = int(level.max()) + 1
  order = np.argsort(level, kind="stable").astype(np.int32)
  level_adr = np.zeros(nlevel + 1, np.int32)
  for lv in level:
    level_adr[lv + 1] += 1
  level_adr = np.cumsum(level_adr).astype(np.int32)
  child_adr = np.zeros(nb + 1, np.int32)
  children: list[list[int]] = [[] for _ in range(nb)]
  for i in range(1, nb):
    children[m.body_parentid[i]].append(i)
  flat: list[int] = []
  for i in range(nb):
    child_adr[i] = len(flat)
    flat.extend(children[i])
  child_adr[nb] = len(flat)
  lastdof = -np.ones(nb, np.int32)
  mask = np.zeros(nb, np.uint64)
  for i in range(1, nb):
    lastdof[i] = lastdof[m.body_parentid[i]]
    mask[i] = mask[m.body_parentid[i]]
    if m.body_dofnum[i] > 0:
      lastdof[i] = m.body_dofadr[i] + m.body_dofnum[i] - 1
      for k in range(m.body_dofnum[i]):
        mask[i] |= np.uint64(1) << np.uint64(m.body_dofadr[i] + k)
  dof_actid = -np.ones(max(m.nv, 1), np.int32)
  for a in range(m.nu):
    dof = int(m.jnt_dofadr[m.actuator_trnid[a, 0]])
    dof_actid[dof] = a if dof_actid[dof] == -1 else -2
  anc = np.zeros(nb, np.uint64)
  if nb <= 64:
    for i in range(nb):
      anc[i] = (anc[m.body_parentid[i]] if i > 0 else np.uint64(0)) | (np.uint64(1) << np.uint64(i))
  ancmask = np.zeros((nb, 2), np.uint32)
  ancmask[:, 0] = (anc & np.uint64(0xFFFFFFFF)).astype(np.uint32)
  ancmask[:, 1] = (anc >> np.uint64(32)).astype(np.uint32)
  dofmask = np.zeros((nb, 2), np.uint32)
  dofmask[:, 0] = (mask & np.uint64(0xFFFFFFFF)).astype(np.uint32)
  dofmask[:, 1] = (mask >> np.uint64(32)).astype(np.uint32)
  return dict(
    nlevel=nlevel, level_adr=level_adr, level_body=order, child_adr=child_adr,
    child_body=np.array(flat if flat else [0], np.int32), body_lastdof=lastdof, body_dofmask=dofmask.view(np.int32),
    dof_actid=dof_actid, body_ancmask=ancmask.view(np.int32),
  )


SUPPORTED_PAIRS = {(0, 2), (0, 3), (0, 6), (2, 2), (2, 3), (3, 3), (2, 6), (3, 6), (6, 6)}  # (plane|sphere|capsule|box) subset


def check_supported(m) -> None:
  """Setup-time validation (raises ValueError like the reference's setup path, sim.py:231-233)."""
  if m.nv > 64:
    raise ValueError(f"nv={m.nv}: the engine supports at most 64 dofs per world")
  if np.any(m.jnt_type == 1):
    raise ValueError("ball joints are outside the supported subset")
  for g1, g2 in m.pair_geom:
    t = (int(m.geom_type[g1]), int(m.geom_type[g2]))
    if t not in SUPPORTED_PAIRS:
      raise ValueError(f"collision between geom types {t} ({m.geom_names[g1]}, {m.geom_names[g2]}) is not supported")
  if np.any(~np.isin(np.asarray(m.geom_condim), (1, 3))):
    raise ValueError("contact dimensions 4 / 6 (torsional / rolling friction rows) are outside the supported subset: condim must be 1 or 3")
  if m.opt_cone not in (0, 1):
    raise ValueError(f"unknown friction cone {m.opt_cone} (0 pyramidal, 1 elliptic)")
  if m.opt_solver not in (1, 2):
    raise ValueError("supported solvers: Newton, CG (PGS is not implemented)")
  if m.opt_integrator not in (0, 3):
    raise ValueError("supported integrators: euler, implicitfast")
  if m.nu and np.any(m.actuator_trnid[:, 0] < 0):
    raise ValueError("only joint transmissions are supported")
  fl = np.asarray(getattr(m, "dof_frictionloss", 0.0))
  if np.any(fl < 0):
    raise ValueError("dof_frictionloss must be >= 0")


# model fields that exist on the compiled model and have physical meaning, but that no kernel reads: making them
# per-world / randomising them would silently do nothing, so `expand_model_fields` and `randomize_field` refuse them
# (ADVICE r01).  Empty since the friction-loss rows were built (dof_frictionloss is consumed by k_solve).
UNCONSUMED_PHYSICS_FIELDS: frozenset = frozenset()


def count_friction_dofs(m) -> int:
  """Dofs that get a friction-loss constraint row (mj_instantiateFriction: dof_frictionloss > 0)."""
  return int(np.sum(np.asarray(getattr(m, "dof_frictionloss", 0.0)) > 0))


def model_scalars(m, tables) -> dict[str, int]:
  st = np.asarray(m.sensor_type) if m.nsensor else np.zeros(0, int)
  return dict(
    nq=m.nq, nv=m.nv, nu=m.nu, nbody=m.nbody, njnt=m.njnt, ngeom=m.ngeom, nsite=m.nsite, nsensor=m.nsensor,
    nsensordata=m.nsensordata, npair=m.npair, nmocap=m.nmocap, nlevel=tables["nlevel"], integrator=int(m.opt_integrator),
    cone=int(m.opt_cone), solver=int(m.opt_solver), iterations=int(m.opt_iterations),
    ls_iterations=int(m.opt_ls_iterations), disableflags=int(m.opt_disableflags),
    has_accelerometer=int(np.any(st == 1)), has_subtreevel=int(np.any((st == 35) | (st == 36))),
    nlimited=int(np.sum((np.asarray(m.jnt_limited) != 0) & (np.asarray(m.jnt_type) >= 2))) if m.njnt else 0,
    maxchain=_max_chain(m, tables), nfriction=count_friction_dofs(m),
    nfield=int(m.nfield), nfq=int(m.nfq), ndyn=int(m.ndyn), field_nx=int(m.field_nx), field_ny=int(m.field_ny),
  )


def _max_chain(m, tables) -> int:
  mask = tables["body_dofmask"].view(np.uint32).astype(np.uint64)
  full = mask[:, 0] | (mask[:, 1] << np.uint64(32))
  best = 1
  for g1, g2 in np.asarray(m.pair_geom).reshape(-1, 2):
    u = int(full[m.geom_bodyid[g1]] | full[m.geom_bodyid[g2]])
    best = max(best, bin(u).count("1"))
  for g in np.asarray(m.fq_geom).reshape(-1):   # field geoms are static: the chain is the query geom's own
    best = max(best, bin(int(full[m.geom_bodyid[g]])).count("1"))
  return best


def model_float_array(m, attr: str) -> np.ndarray:
  v = getattr(m, attr)
  return np.atleast_1d(np.asarray(v, np.float64)).reshape(-1)


def model_int_array(m, attr: str, tables) -> np.ndarray:
  v = tables[attr] if attr in tables else getattr(m, attr)
  a = np.ascontiguousarray(np.asarray(v).reshape(-1)).astype(np.int32, copy=False)
  return a if a.size else np.zeros(1, np.int32)


def init_static_numpy(m, geom_xpos: np.ndarray, geom_xmat: np.ndarray) -> None:
  """Host-side twin of b2_init_static for the numpy-backed test / baseline builds of the device code."""
  static = np.ones(m.ngeom, bool)
  static[np.asarray(m.dyn_geom, np.int64)] = False
  geom_xpos[:, static] = np.asarray(m.sgeom_xpos)[static]
  geom_xmat[:, static] = np.asarray(m.sgeom_xmat).reshape(m.ngeom, 3, 3)[static].reshape((-1,) + geom_xmat.shape[2:])


_lib = None


def load_library() -> C.CDLL:
  """Loads libb200sim.so; raises (no fallback) if it has not been built."""
  global _lib
  if _lib is not None:
    return _lib
  path = Path(os.environ.get("B200SIM_LIB", LIB_PATH))
  if not path.exists():
    raise RuntimeError(
      f"{path} not found: the CUDA extension is required (no CPU fallback). "
      "Build it with `python -c 'import __graft_entry__ as g; g.build()'`."
    )
  lib = C.CDLL(str(path))
  lib.b2_abi_version.restype = C.c_int
  lib.b2_last_error.restype = C.c_char_p
  lib.b2_sizeof_model.restype = C.c_int
  lib.b2_sizeof_data.restype = C.c_int
  lib.b2_launch_count.restype = C.c_longlong
  for name in ("b2_check",):
    getattr(lib, name).argtypes = [C.POINTER(B2Model), C.POINTER(B2Data)]
    getattr(lib, name).restype = C.c_int
  for name in ("b2_step", "b2_forward"):
    getattr(lib, name).argtypes = [C.POINTER(B2Model), C.POINTER(B2Data), C.c_void_p]
    getattr(lib, name).restype = C.c_int
  lib.b2_substeps.argtypes = [C.POINTER(B2Model), C.POINTER(B2Data), C.c_int, C.c_void_p]
  lib.b2_substeps.restype = C.c_int
  lib.b2_substeps_after_forward.argtypes = [C.POINTER(B2Model), C.POINTER(B2Data), C.c_int, C.c_void_p]
  lib.b2_substeps_after_forward.restype = C.c_int
  lib.b2_stage.argtypes = [C.POINTER(B2Model), C.POINTER(B2Data), C.c_int, C.c_int, C.c_void_p]
  lib.b2_stage.restype = C.c_int
  lib.b2_reset.argtypes = [C.POINTER(B2Model), C.POINTER(B2Data), C.c_void_p, C.c_void_p]
  lib.b2_reset.restype = C.c_int
  lib.b2_repeat_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
  lib.b2_repeat_rows.restype = C.c_int
  lib.b2_init_static.argtypes = [C.POINTER(B2Model), C.POINTER(B2Data), C.c_void_p]
  lib.b2_init_static.restype = C.c_int
  if lib.b2_sizeof_model() != C.sizeof(B2Model) or lib.b2_sizeof_data() != C.sizeof(B2Data):
    raise RuntimeError("libb200sim.so ABI mismatch with mjlab_b200/engine.py (rebuild the extension)")
  _lib = lib
  return lib


def check_call(lib, rc: int) -> None:
  if rc != 0:
    raise RuntimeError(f"b200sim: {lib.b2_last_error().decode()}")
