"""CPU: rough terrain (BASELINE config 4) - the terrain generator against the reference's own output, the box narrowphase
against closed forms, and the kernel source (single-lane host build) against the numpy oracle on the terrain grid."""
import numpy as np
import pytest

import mjlab_b200
from mjlab_b200 import terrains as T
from oracle import mjstep as O
from tests import util
from tests.hostemu.harness import HostSim

GOLD = np.load("tests/golden/terrain_rough.npz")


@pytest.mark.parametrize("play", [False, True])
def test_terrain_generator_matches_the_reference(play):
  """Every box (centre, half extents) and every spawn origin of ROUGH_TERRAINS_CFG, seed 0, equals what the reference's
  TerrainGenerator.compile produced (tools/make_terrain_golden.py)."""
  t = T.generate(T.rough_terrains_cfg(play=play, seed=0))
  k = "play" if play else "train"
  assert t.boxes.shape == GOLD[k + "_boxes"].shape
  np.testing.assert_allclose(t.boxes, GOLD[k + "_boxes"], rtol=0, atol=1e-12)
  np.testing.assert_allclose(t.origins, GOLD[k + "_origins"], rtol=0, atol=1e-12)


def test_env_placement_and_level_update_match_the_reference():
  assert np.array_equal(T.terrain_types(64, 20), GOLD["imp_types"])
  ids, lv = GOLD["cur_env_ids"], GOLD["cur_levels0"]
  dist = np.linalg.norm(GOLD["cur_root_xy"][ids] - GOLD["cur_origins0"][ids, :2], axis=1)
  up = dist > 8.0 / 2
  down = (dist < np.linalg.norm(GOLD["cur_command"][ids, :2], axis=1) * 20.0 * 0.5) & ~up
  new = T.update_levels(lv[ids], up, down, 10, -1)
  wrapped = GOLD["cur_wrapped"]
  assert wrapped.sum() >= 4 and (new[wrapped] == -1).all()       # the random restart draw itself is engine-specific
  assert np.array_equal(new[~wrapped], GOLD["cur_levels1"][ids][~wrapped])
  untouched = np.setdiff1d(np.arange(64), ids)
  assert np.array_equal(GOLD["cur_levels1"][untouched], lv[untouched])


def test_rough_model_field_tables():
  m = mjlab_b200.load_model("g1_rough")
  t = T.generate(T.rough_terrains_cfg(seed=0))
  assert m.nfield == len(t.boxes) == 3564 and m.nfq == 33 and m.ndyn == m.ngeom - m.nfield
  np.testing.assert_allclose(m.sgeom_xpos[m.field_geom], t.boxes[:, :3], atol=1e-12)
  np.testing.assert_allclose(m.field_aabb, t.boxes, atol=1e-12)       # axis-aligned boxes: AABB == the box
  np.testing.assert_allclose(m.terrain_origins, t.origins, atol=1e-12)
  # every box is listed in exactly the cells its AABB overlaps
  x0, y0, cell, _ = m.field_grid
  rng = np.random.default_rng(0)
  for fi in rng.choice(m.nfield, 50, replace=False):
    a = m.field_aabb[fi]
    cx = np.arange(int(np.floor((a[0] - a[3] - x0) / cell)), int(np.floor((a[0] + a[3] - x0) / cell)) + 1)
    cy = np.arange(int(np.floor((a[1] - a[4] - y0) / cell)), int(np.floor((a[1] + a[4] - y0) / cell)) + 1)
    cx, cy = cx[(cx >= 0) & (cx < m.field_nx)], cy[(cy >= 0) & (cy < m.field_ny)]
    listed = [(c // m.field_ny, c % m.field_ny) for c in range(m.field_nx * m.field_ny)
              if fi in m.field_cell_item[m.field_cell_adr[c]:m.field_cell_adr[c + 1]]] if fi < 3 else None
    if listed is not None:
      assert sorted(listed) == sorted((int(x), int(y)) for x in cx for y in cy)
    assert tuple(m.field_cell0[fi]) == (cx[0], cy[0])
  # the moving geoms that query the field: all 33 collision geoms of the G1 (spheres and capsules)
  assert set(m.geom_type[m.fq_geom]) <= {2, 3}


# ---- narrowphase closed forms ---------------------------------------------------------------------------------------
def test_sphere_box_closed_forms():
  bp, bR, bs = np.array([1.0, 2.0, 0.5]), np.eye(3), np.array([2.0, 1.0, 0.5])
  # above the top face: contact straight down, depth = gap - radius
  dist, pos, n = O._sphere_box(np.array([1.5, 2.2, 1.08]), 0.1, bp, bR, bs, 0.0)
  assert np.isclose(dist, -0.02) and np.allclose(n, [0, 0, -1]) and np.allclose(pos, [1.5, 2.2, 0.99])
  # off an edge: normal along the shortest segment to the edge
  dist, pos, n = O._sphere_box(np.array([3.05, 2.0, 1.05]), 0.1, bp, bR, bs, 0.0)
  assert np.isclose(dist, np.hypot(0.05, 0.05) - 0.1) and np.allclose(n, -np.array([1, 0, 1]) / np.sqrt(2))
  # separated
  assert O._sphere_box(np.array([1.0, 2.0, 1.2]), 0.1, bp, bR, bs, 0.0) is None
  # centre inside the box: leaves through the nearest face (+z here), normal pointing into the box
  dist, pos, n = O._sphere_box(np.array([1.0, 2.0, 0.95]), 0.1, bp, bR, bs, 0.0)
  assert np.isclose(dist, -0.05 - 0.1) and np.allclose(n, [0, 0, -1])
  # rotated box == rotated problem
  q = np.array([0.9, 0.1, 0.3, 0.2]); q /= np.linalg.norm(q)
  R = O.quat_to_mat(q)
  sp = np.array([1.5, 2.2, 1.08])
  d2, p2, n2 = O._sphere_box(bp + R @ (sp - bp), 0.1, bp, R, bs, 0.0)
  assert np.isclose(d2, -0.02) and np.allclose(n2, R @ [0, 0, -1]) and np.allclose(p2, bp + R @ (np.array([1.5, 2.2, 0.99]) - bp))


def test_capsule_box_configurations():
  bp, bR, bs = np.zeros(3), np.eye(3), np.array([1.0, 1.0, 0.5])
  r, h = 0.05, 0.2
  # lying flat on the top face: two contacts at the segment ends, equal depth
  out = O._capsule_box(np.array([0.1, 0.2, 0.54]), np.array([1.0, 0, 0]), r, h, bp, bR, bs, 0.0)
  assert len(out) == 2 and np.allclose([o[0] for o in out], -0.01)
  assert np.allclose(sorted(o[1][0] for o in out), [-0.1, 0.3]) and all(np.allclose(o[2], [0, 0, -1]) for o in out)
  # half over the edge x = 1: contacts at the inner end and above the edge
  out = O._capsule_box(np.array([0.95, 0.0, 0.54]), np.array([1.0, 0, 0]), r, h, bp, bR, bs, 0.0)
  assert len(out) == 2 and np.allclose(sorted(o[1][0] for o in out), [0.75, 1.0], atol=1e-6)
  # tilted, one end down: deepest end first; the other end is in the air
  ax = np.array([np.cos(0.5), 0, np.sin(0.5)])
  out = O._capsule_box(np.array([0.0, 0.0, 0.54 + h * np.sin(0.5)]), ax, r, h, bp, bR, bs, 0.0)
  assert len(out) == 1 and np.isclose(out[0][0], -0.01) and np.isclose(out[0][1][0], -h * np.cos(0.5))
  # standing upright above the face: one contact under the lower cap
  out = O._capsule_box(np.array([0.3, 0.3, 0.5 + h + 0.04]), np.array([0, 0, 1.0]), r, h, bp, bR, bs, 0.0)
  assert len(out) == 1 and np.isclose(out[0][0], -0.01) and np.allclose(out[0][2], [0, 0, -1])
  # across a corner edge (the stair nose), segment perpendicular to the edge direction
  out = O._capsule_box(np.array([1.02, 0.0, 0.52]), np.array([1.0, 0, -1.0]) / np.sqrt(2), r, h, bp, bR, bs, 0.0)
  assert len(out) >= 1 and np.isclose(out[0][0], np.hypot(0.02, 0.02) - r)
  # separated
  assert O._capsule_box(np.array([0.0, 0.0, 0.8]), np.array([1.0, 0, 0]), r, h, bp, bR, bs, 0.0) == []


# ---- kernel source on the terrain grid vs the oracle --------------------------------------------------------------------
def rough_states(m, nworld, seed):
  """Robots dropped onto stairs tiles (pyramids and inverted pyramids, several levels), shifted off the platform so
  that the feet meet step edges, with jittered joints."""
  rng = np.random.default_rng(seed)
  qpos, qvel, ctrl, xfrc = util.random_states(m, nworld, seed=seed, drop=0.0, jitter=0.08, vel=0.3)
  R, C = m.terrain_origins.shape[:2]
  for w in range(nworld):
    r, c = rng.integers(2, R), rng.integers(8, C) if w % 4 else rng.integers(0, 8)
    o = m.terrain_origins[r, c]
    off = rng.uniform(-2.6, 2.6, 2) if w % 2 else rng.uniform(-0.5, 0.5, 2)
    qpos[w, :2] = o[:2] + off
    qpos[w, 2] += o[2] - 0.02 * (w % 3)
    # stairs descend away from the platform: lower the robot by the steps it has moved past
  return qpos, qvel, ctrl, xfrc


def test_kernel_source_on_host_matches_oracle_on_the_terrain_grid():
  m = mjlab_b200.load_model("g1_rough")
  nw, nstep = 6, 3
  qpos, qvel, ctrl, xfrc = rough_states(m, nw, seed=11)
  hs = HostSim(m, nw, nconmax=64, njmax=300, real="double")
  hs.qpos[:], hs.qvel[:], hs.ctrl[:] = qpos, qvel, ctrl
  ds = []
  for w in range(nw):
    d = O.Data(m)
    d.qpos[:], d.qvel[:], d.ctrl[:] = qpos[w], qvel[w], ctrl[w]
    ds.append(d)
  seen_box_contacts = 0
  for s in range(nstep):
    hs.step()
    for w, d in enumerate(ds):
      O.step(m, d, 64, 300)
      n = len(d.contacts)
      assert hs.ncon[w] == n, (s, w)
      assert hs.contact_geom[w, :n].tolist() == [list(c.geom) for c in d.contacts]
      if n:
        np.testing.assert_allclose(hs.contact_dist[w, :n], [c.dist for c in d.contacts], rtol=0, atol=1e-7)
        np.testing.assert_allclose(hs.contact_pos[w, :n], [c.pos for c in d.contacts], rtol=0, atol=1e-7)
        np.testing.assert_allclose(hs.contact_frame[w, :n].reshape(n, 3, 3), [c.frame for c in d.contacts], rtol=0, atol=1e-6)
      seen_box_contacts += sum(1 for c in d.contacts if c.geom[1] in set(m.field_geom.tolist()))
      assert hs.nefc[w] == d.nefc
      assert np.abs(hs.qpos[w] - d.qpos).max() < 1e-8
      assert np.abs(hs.qvel[w] - d.qvel).max() < 1e-6 * max(1.0, np.abs(d.qvel).max())
  assert seen_box_contacts >= 20           # the case really exercises the terrain boxes
  # static geoms keep the constant pose; moving ones were updated
  np.testing.assert_allclose(hs.geom_xpos[0, m.field_geom], m.sgeom_xpos[m.field_geom], atol=0)
  np.testing.assert_allclose(hs.geom_xpos[0], ds[0].geom_xpos, atol=1e-9)


def test_float_host_build_on_the_terrain_grid_stays_close():
  m = mjlab_b200.load_model("g1_rough")
  nw = 4
  qpos, qvel, ctrl, xfrc = rough_states(m, nw, seed=12)
  a, b = HostSim(m, nw, 64, 300, real="double"), HostSim(m, nw, 64, 300, real="float")
  for h in (a, b):
    h.qpos[:], h.qvel[:], h.ctrl[:] = qpos, qvel, ctrl
    for _ in range(4):
      h.step()
  assert a.ncon.tolist() == b.ncon.tolist()
  assert np.abs(a.qpos - b.qpos).max() < 5e-5


def test_rough_task_configs():
  """Train: curriculum on the 10 x 20 grid; play: 5 x 5 random grid with the terrain-randomisation reset event as well."""
  from mjlab_b200 import envcfg
  from mjlab_b200.env import VelocityEnvCfg
  train = envcfg.velocity_cfg(mjlab_b200.load_model("g1_rough"), "g1", play=False)
  assert train.terrain_mode == 1 and train.terrain_origins.shape == (10, 20, 3) and train.terrain_half_size == 4.0
  play = envcfg.velocity_cfg(mjlab_b200.load_model("g1_rough_play"), "g1", play=True)
  assert play.terrain_mode == 3 and play.terrain_origins.shape == (5, 5, 3) and not play.push_enabled
  flat = envcfg.velocity_cfg(mjlab_b200.load_model("g1_flat"), "g1")
  assert flat.terrain_mode == 0 and flat.terrain_rows == 0
  c = VelocityEnvCfg.from_task("Mjlab-Velocity-Rough-Unitree-G1", num_envs=8)
  assert (c.robot, c.terrain) == ("g1", "rough")
  with pytest.raises(ValueError):
    VelocityEnvCfg.from_task("Mjlab-Velocity-Rough-Unitree-H1")
