"""Box-side helper (needs `mujoco`, which this image does not have): dump a compiled mujoco.MjModel into the
array container `mjlab_b200.mjcf.ModelDesc` uses, so models compiled by MuJoCo itself can be fed to the
engine and cross-checked against `mjlab_b200.mjcf`'s own compiler.  The array copy is UNTESTED here (no mujoco wheel); the
collision tables (static-geom poses, geom field, candidate pairs) come from the same `collision_tables` the compiler uses.

    python tools/export_mjmodel.py scene.xml out.npz        # on a box with MuJoCo
"""
import sys

import numpy as np


def export(mj_model, path):
  import mujoco  # noqa: F401  (only available where MuJoCo is installed)
  from mjlab_b200.mjcf import ModelDesc, collision_tables

  m = ModelDesc()
  ints = ["nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite", "nsensor", "nsensordata", "nmocap", "nkey"]
  for k in ints:
    setattr(m, k, int(getattr(mj_model, k)))
  arrays = [
    "body_parentid", "body_rootid", "body_weldid", "body_mocapid", "body_jntnum", "body_jntadr", "body_dofnum",
    "body_dofadr", "body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_inertia", "body_subtreemass",
    "body_invweight0", "jnt_type", "jnt_qposadr", "jnt_dofadr", "jnt_bodyid", "jnt_pos", "jnt_axis", "jnt_range",
    "jnt_limited", "jnt_solref", "jnt_solimp", "jnt_margin", "jnt_stiffness", "dof_bodyid", "dof_jntid", "dof_parentid",
    "dof_armature", "dof_damping", "dof_frictionloss", "dof_invweight0", "qpos0", "qpos_spring", "geom_type",
    "geom_bodyid", "geom_pos", "geom_quat", "geom_size", "geom_contype", "geom_conaffinity", "geom_condim",
    "geom_priority", "geom_friction", "geom_solref", "geom_solimp", "geom_solmix", "geom_margin", "geom_gap",
    "geom_rbound", "geom_group", "geom_rgba", "site_bodyid", "site_pos", "site_quat", "actuator_trnid",
    "actuator_gaintype", "actuator_biastype", "actuator_gainprm", "actuator_biasprm", "actuator_gear",
    "actuator_forcelimited", "actuator_forcerange", "actuator_ctrllimited", "actuator_ctrlrange", "sensor_type",
    "sensor_objtype", "sensor_objid", "sensor_reftype", "sensor_refid", "sensor_intprm", "sensor_dim", "sensor_adr",
    "sensor_cutoff", "key_qpos", "key_qvel", "key_ctrl",
  ]
  for k in arrays:
    setattr(m, k, np.array(getattr(mj_model, k)))
  for kind, n in (("body", m.nbody), ("jnt", m.njnt), ("geom", m.ngeom), ("site", m.nsite), ("actuator", m.nu), ("sensor", m.nsensor)):
    obj = {"jnt": "joint"}.get(kind, kind)
    setattr(m, f"{kind}_names", [getattr(mj_model, obj)(i).name for i in range(n)])
  o = mj_model.opt
  m.opt_timestep, m.opt_gravity = float(o.timestep), np.array(o.gravity)
  m.opt_integrator, m.opt_cone, m.opt_solver = int(o.integrator), int(o.cone), int(o.solver)
  m.opt_iterations, m.opt_ls_iterations = int(o.iterations), int(o.ls_iterations)
  m.opt_tolerance, m.opt_ls_tolerance, m.opt_impratio = float(o.tolerance), float(o.ls_tolerance), float(o.impratio)
  m.opt_disableflags = int(o.disableflags)
  m.stat_meaninertia = float(mj_model.stat.meaninertia)
  # <exclude> body pairs: MuJoCo stores them as signatures (body1 << 16) + body2
  sig = np.asarray(getattr(mj_model, "exclude_signature", np.zeros(0, np.int64))).astype(np.int64)
  excl = {(int(min(x >> 16, x & 0xFFFF)), int(max(x >> 16, x & 0xFFFF))) for x in sig}
  m.exclude_pairs = np.array(sorted(excl), np.int32).reshape(-1, 2)
  collision_tables(m, excl)
  m.save(path)
  return m


if __name__ == "__main__":
  if len(sys.argv) != 3:
    print(__doc__)
    sys.exit(0)
  import mujoco
  from pathlib import Path
  sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
  export(mujoco.MjModel.from_xml_path(sys.argv[1]), sys.argv[2])
