"""Contact sensors beyond what the shipped tasks declare (reduce none / mindist / maxforce with several slots; found, force, dist,
pos, normal, tangent fields; geom, body and subtree matching; with and without a secondary): the kernel source's sensor stage
(single-lane host build of b2_solve.cuh `post_world`) against the oracle's (oracle/mjstep.py `sensors`), which restates MuJoCo's
mjSENS_CONTACT semantics that the reference's ContactSensor relies on (src/mjlab/sensor/contact_sensor.py:1-120; the cases mirror
the reference's tests/test_contact_sensor.py:147-760: fields, reduce modes with multiple contacts, body / subtree modes, several
slots, no contacts)."""
import numpy as np
import pytest

from mjlab_b200 import mjcf
from oracle import mjstep as O
from tests.hostemu.harness import HostSim

SCENE = """
<mujoco>
  <option timestep="0.005"/>
  <worldbody>
    <geom name="floor" type="plane" size="5 5 0.1"/>
    <body name="slab" pos="0 0 0.1"><freejoint/><geom name="slab_g" type="box" size="0.5 0.4 0.1" mass="4"/>
      <body name="post" pos="0.3 0.2 0.2"><geom name="post_g" type="capsule" size="0.04 0.08" mass="0.2"/></body>
    </body>
    <body name="cube" pos="0.1 0.05 0.3"><freejoint/><geom name="cube_g" type="box" size="0.1 0.1 0.1" mass="1"/></body>
    <body name="ball" pos="-0.25 0.1 0.27"><freejoint/><geom name="ball_g" type="sphere" size="0.07" mass="0.3"/></body>
    <body name="bar" pos="-0.1 -0.15 0.26" euler="0 0 30"><freejoint/><geom name="bar_g" type="capsule" size="0.05 0.2" euler="0 90 0" mass="0.5"/></body>
    <body name="far" pos="2 2 0.5"><freejoint/><geom name="far_g" type="sphere" size="0.05" mass="0.1"/></body>
  </worldbody>
</mujoco>
"""
BIT = {"found": 1, "force": 2, "dist": 8, "pos": 16, "normal": 32, "tangent": 64}
RED = {"none": 0, "mindist": 1, "maxforce": 2, "netforce": 3}
OBJ = {"geom": mjcf.OBJ_GEOM, "body": mjcf.OBJ_BODY, "subtree": mjcf.OBJ_XBODY}
# (name, primary mode, primary, secondary mode, secondary, fields, reduce, slots)
SENSORS = [
  ("slab_any", "geom", "slab_g", None, None, ("found", "force", "dist", "pos", "normal", "tangent"), "none", 3),
  ("slab_floor", "geom", "slab_g", "geom", "floor", ("found", "force"), "netforce", 1),
  ("slab_body", "body", "slab", None, None, ("found", "force"), "maxforce", 2),
  ("slab_tree", "subtree", "slab", None, None, ("found", "dist", "normal"), "mindist", 4),
  ("cube_on_slab", "body", "cube", "body", "slab", ("found", "force", "pos"), "maxforce", 4),
  ("slab_under_cube", "body", "slab", "body", "cube", ("found", "force", "normal"), "mindist", 2),      # same contacts, sign flipped
  ("ball", "geom", "ball_g", None, None, ("found", "force", "normal"), "netforce", 1),
  ("far", "body", "far", None, None, ("found", "force", "dist"), "maxforce", 2),                        # never touches: zeros
]


def build():
  spec = mjcf.Spec.from_string(SCENE)
  dims = {}
  for name, pm, prim, sm, sec, fields, red, num in SENSORS:
    for f in fields:    # one MuJoCo sensor per (sensor, field), as the reference's ContactSensor declares them
      kw = dict(name=f"{name}_{f}", type=mjcf.SENS_CONTACT, objtype=OBJ[pm], objname=prim, intprm=(BIT[f], RED[red], num))
      if sm is not None:
        kw["reftype"], kw["refname"] = OBJ[sm], sec
      spec.add_sensor(**kw)
      dims[f"{name}_{f}"] = (1 if f in ("found", "dist") else 3) * (1 if red == "netforce" else num)
  return spec.compile(), dims


def test_kernel_sensor_stage_matches_oracle_on_all_contact_sensor_modes():
  """float64 host build: the sensor LOGIC (matching, ranking, slots, signs, fields) digit for digit.  A float32 comparison of this
  scene is not meaningful per element - four redundant corner contacts share a load to within percents only, and near-ties in
  the mindist / maxforce ranking order differently - the float32 path is the same source and is covered by the task sensors."""
  real = "double"
  m, dims = build()
  assert m.nsensor == sum(len(s[5]) for s in SENSORS) and m.nsensordata == sum(dims.values())
  nw, nstep = 3, 40
  rng = np.random.default_rng(2)
  hs = HostSim(m, nw, nconmax=64, njmax=300, real=real)
  ds = []
  for w in range(nw):
    d = O.Data(m)
    d.qpos[:] = m.qpos0
    d.qpos[:21:7] += rng.normal(size=3) * 0.01 * w
    d.qvel[:] = rng.normal(size=m.nv) * 0.1
    hs.qpos[w], hs.qvel[w] = d.qpos, d.qvel
    ds.append(d)
  seen_multi = seen_sign = 0
  adr = {n: int(a) for n, a in zip(m.sensor_names, m.sensor_adr)}
  for s in range(nstep):
    hs.step()
    for w, d in enumerate(ds):
      O.step(m, d, 64, 300)
      assert hs.ncon[w] == len(d.contacts)
      np.testing.assert_allclose(hs.sensordata[w], d.sensordata, err_msg=f"step {s} world {w}", rtol=1e-8, atol=1e-9)
      sd = d.sensordata
      seen_multi += sd[adr["slab_any_found"]] >= 3                  # several matches: the slot ranking is exercised
      f1, f2 = sd[adr["cube_on_slab_force"]: adr["cube_on_slab_force"] + 3], sd[adr["slab_under_cube_force"]: adr["slab_under_cube_force"] + 3]
      seen_sign += bool(np.abs(f1).sum() > 0 and np.abs(f2).sum() > 0)
      assert not np.any(sd[adr["far_found"]: adr["far_found"] + 2]) and not np.any(sd[adr["far_force"]: adr["far_force"] + 6])
  assert seen_multi > 20 and seen_sign > 20
