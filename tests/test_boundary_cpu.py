"""Drop-in boundary pieces that need no GPU: the MjModel -> ModelDesc conversion behind `Simulation(..., mujoco.MjModel, ...)`
(sim/sim.py:122-149), MuJoCo's `model.opt` object surface on `ModelDesc`, the read-only `WarpBridge`
(reference tests/test_sim_data.py:31-83), `SceneEntityCfg.resolve` (managers/scene_entity_config.py:76-200), and the
setup-time rejection of parsed-but-unimplemented physics."""
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

import mjlab_b200
from mjlab_b200 import engine as E
from mjlab_b200.managers import SceneEntityCfg
from mjlab_b200.mjcf import ModelDesc
from mjlab_b200.sim import WarpBridge
from tests.fake_mujoco import FakeMjModel


@pytest.mark.parametrize("name", ["g1_flat", "go1_flat", "g1_rough"])
def test_from_mjmodel_reproduces_every_array_and_derived_table(name):
  m = mjlab_b200.load_model(name)
  m2 = ModelDesc.from_mjmodel(FakeMjModel(m))
  for k, v in m.__dict__.items():
    if k in ("qM0", "key_names", "opt_ccd_iterations") or k.startswith("terrain_"):   # not part of the MjModel surface (terrain_*: TerrainImporter metadata)
      continue
    assert hasattr(m2, k), k
    w = getattr(m2, k)
    if isinstance(v, np.ndarray):
      np.testing.assert_array_equal(np.asarray(w).reshape(v.shape), v, err_msg=k)
    else:
      assert w == v, k
  E.check_supported(m2)
  t1, t2 = E.derived_int_tables(m), E.derived_int_tables(m2)
  for k in t1:
    np.testing.assert_array_equal(t1[k], t2[k], err_msg=k)


def test_modeldesc_has_mujocos_opt_surface():
  m = mjlab_b200.load_model("go1_flat")
  assert m.opt.timestep == m.opt_timestep and tuple(m.opt.gravity) == tuple(m.opt_gravity)
  m.opt.timestep = 0.01
  m.opt.gravity = (0, 0, 7.5)
  assert m.opt_timestep == 0.01 and tuple(m.opt_gravity) == (0.0, 0.0, 7.5)
  assert m.opt.ccd_iterations == 50
  with pytest.raises(AttributeError):
    m.opt.no_such_option = 1
  assert m.body(1).name == m.body_names[1] and m.geom(m.geom_names[3]).id == 3 and m.stat.meaninertia == m.stat_meaninertia


def test_bridge_is_read_only_and_shares_memory():
  """tests/test_sim_data.py:31-83 of the reference, on the bridge class itself."""
  t = torch.tensor([[1.0, 2.0], [3.0, 4.0]])
  b = WarpBridge({"arr": t, "val": 1.0})
  assert b.arr is b.arr and b.val == 1.0                       # same object on every access (cached wrapper in the reference)
  ptr = b.arr.data_ptr()
  b.arr[:] = torch.zeros(2, 2)
  assert b.arr.data_ptr() == ptr and torch.all(t == 0)         # slice assignment keeps the address, writes through
  b.arr[0, 1] = 99.0
  assert t[0, 1] == 99.0
  assert torch.allclose(b.arr * 2, t * 2) and torch.sum(b.arr).item() == 99.0
  with pytest.raises(AttributeError, match="Cannot set attribute 'arr' on WarpBridge"):
    b.arr = torch.zeros(2, 2)
  with pytest.raises(AttributeError, match="Use in-place operations instead"):
    b.val = 42.0
  with pytest.raises(AttributeError):
    b.missing


def test_scene_entity_cfg_resolve():
  names = ["hip", "knee", "ankle"]

  def find(keys, preserve_order=False):
    import re
    keys = [keys] if isinstance(keys, str) else list(keys)
    ids = [i for i, n in enumerate(names) if any(re.fullmatch(k, n) for k in keys)]
    return ids, [names[i] for i in ids]

  ent = NS(joint_names=names, num_joints=3, find_joints=find, body_names=["a", "b"], num_bodies=2,
           find_bodies=lambda k, preserve_order=False: ([0], ["a"]))
  scene = {"robot": ent}
  c = SceneEntityCfg("robot", joint_names=(".*",))
  c.resolve(scene)
  assert c.joint_ids == [0, 1, 2] and c.joint_names == [".*"]            # regex: ids resolved, not collapsed (names != all names)
  c = SceneEntityCfg("robot", joint_names=("hip", "knee", "ankle"))
  c.resolve(scene)
  assert c.joint_ids == slice(None)                                      # everything, in order -> slice(None)
  c = SceneEntityCfg("robot", joint_ids=[2, 0])
  c.resolve(scene)
  assert c.joint_names == ["ankle", "hip"]
  c = SceneEntityCfg("robot", joint_names=("knee",), joint_ids=[2])
  with pytest.raises(ValueError, match="Inconsistent joint names and indices"):
    c.resolve(scene)
  c = SceneEntityCfg("robot")
  c.resolve(scene)
  assert c.joint_ids == slice(None) and c.body_ids == slice(None) and c.joint_names is None
  with pytest.raises(KeyError):
    SceneEntityCfg("nobody", joint_names="hip").resolve(scene)


def test_frictionloss_is_consumed_not_rejected():
  """Round-1 advice: parsed-but-unused physics must not be silent.  Friction loss was first rejected at setup; since the
  friction-loss rows were built it is consumed (tests/test_frictionloss_cpu.py): positive values size the solver for the
  rows, negative ones are an error, and no model field is left in the 'unconsumed' set."""
  m = mjlab_b200.load_model("go1_flat")
  E.check_supported(m)
  tables = E.derived_int_tables(m)
  assert E.model_scalars(m, tables)["nfriction"] == 0
  m.dof_frictionloss = np.asarray(m.dof_frictionloss, float).copy()
  m.dof_frictionloss[7] = 0.1
  E.check_supported(m)
  assert E.model_scalars(m, tables)["nfriction"] == 1
  m.dof_frictionloss[7] = -0.1
  with pytest.raises(ValueError, match="frictionloss"):
    E.check_supported(m)
  assert not E.UNCONSUMED_PHYSICS_FIELDS
  assert "dof_frictionloss" in [a for _, a, _ in E.MODEL_FLOAT_FIELDS]