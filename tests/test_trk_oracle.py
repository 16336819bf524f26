"""The tracking-task oracle (oracle/trackstep.py) against golden vectors produced by THE REFERENCE'S OWN MotionCommand,
reward / observation / termination functions (tools/make_trk_golden.py)."""
import numpy as np

import mjlab_b200
from mjlab_b200 import envcfg
from oracle import trackstep as OT
from tests.envutil import snapshot_from

TOL = dict(rtol=3e-5, atol=3e-6)


def setup(g):
  m = mjlab_b200.load_model("g1_flat")
  mo = {k[len("motion_"):]: g[k].astype(np.float32) for k in g.files if k.startswith("motion_") and k != "motion_fps"}
  cfg = envcfg.tracking_cfg(m, mo["joint_pos"].shape[0])
  cfg.enable_corruption = False
  cfg.push_enabled = False
  cfg.sampling_mode = "start"
  return m, mo, cfg


def test_oracle_matches_reference_functions():
  g = np.load(snapshot_from("mdp_tracking_g1"))
  m, mo, cfg = setup(g)
  assert tuple(g["body_names"]) == envcfg.G1_TRACKED_BODIES and tuple(g["ee_names"]) == envcfg.G1_EE_BODIES
  np.testing.assert_allclose(g["reward_weights"], cfg.reward_weights)
  T, B = g["actions"].shape[:2]
  s = OT.make_state(cfg, B, g["origins"])
  s.encoder_bias[:] = g["enc_bias"]
  s.time_steps[:] = g["time0"] + 1
  s.rel_pos[:], s.rel_quat[:] = g["rel0_pos"], g["rel0_quat"]
  nterm = 0
  for t in range(T):
    d = snapshot_from.load(g, t, m)
    OT.apply_action(cfg, s, d, g["actions"][t])
    OT.post(cfg, m, mo, s, d, u=np.zeros((B, OT.A.NDRAW), np.float32), do_reset=False)
    np.testing.assert_array_equal(s.term_dones[:, 1], g["ref_term_anchor_pos"][t])
    np.testing.assert_array_equal(s.term_dones[:, 2], g["ref_term_anchor_ori"][t])
    np.testing.assert_array_equal(s.term_dones[:, 3], g["ref_term_ee"][t])
    nterm += int(s.term_dones[:, 1:].sum())
    np.testing.assert_allclose(s.last_terms, g["ref_rew_terms"][t], **TOL)
    np.testing.assert_allclose(s.reward, g["ref_reward"][t], **TOL)
    rb = OT.robot_bodies(cfg, d)
    OT.update_metrics(cfg, s, d, OT.motion_frame(cfg, mo, s), rb)
    np.testing.assert_allclose(s.metrics, g["ref_metrics"][t], rtol=1e-4, atol=1e-5)
    fr = OT.update_command(cfg, mo, s, d, rb, None)
    np.testing.assert_array_equal(s.time_steps, g["ref_time_steps"][t])
    np.testing.assert_allclose(s.rel_pos, g["ref_rel_pos"][t], **TOL)
    np.testing.assert_allclose(s.rel_quat, g["ref_rel_quat"][t], **TOL)
    OT.observations(cfg, s, d, fr, rb)
    np.testing.assert_allclose(s.obs_policy, g["ref_obs_policy"][t], **TOL)
    np.testing.assert_allclose(s.obs_critic, g["ref_obs_critic"][t], **TOL)
  assert nterm > 0
  # every motion termination is True somewhere in the reference's own output
  assert g["ref_term_anchor_pos"].any() and g["ref_term_anchor_ori"].any() and g["ref_term_ee"].any()


def test_adaptive_sampling_probabilities():
  g = np.load(snapshot_from("mdp_tracking_g1"))
  m, mo, cfg = setup(g)
  cfg.sampling_mode, cfg.adaptive_kernel_size, cfg.bin_count = "adaptive", 3, 6
  s = OT.make_state(cfg, 4)
  s.bin_failed_count[:] = [0, 2, 0, 0, 1, 0]
  p = OT.sampling_probabilities(cfg, s)
  assert abs(p.sum() - 1) < 1e-6 and p.argmax() == 1 and p[0] > p[2] > 0   # non-causal kernel: bin 0 sees the failures ahead of it


def test_3d_termination_variants_match_reference_functions():
  """bad_anchor_pos / bad_motion_body_pos (tasks/tracking/mdp/terminations.py:18-25,55-70: 3-D distance instead of the height
  difference), pinned by the reference's functions at a threshold where they differ from the z-only variants on these states."""
  g = np.load(snapshot_from("mdp_tracking_g1"))
  m, mo, cfg = setup(g)
  thr = 0.05
  assert (g["ref_term_anchor_pos_3d"] != g["ref_term_anchor_pos_z_at_thr3d"]).any() and (g["ref_term_ee_3d"] != g["ref_term_ee_z_at_thr3d"]).any()
  for k in ("ref_term_anchor_pos_3d", "ref_term_ee_3d"):
    assert g[k].any() and not g[k].all()
  T, B = g["actions"].shape[:2]
  for full in (3, 0):
    cfg.anchor_pos_threshold = cfg.ee_pos_threshold = thr
    cfg.term_full3d = full
    s = OT.make_state(cfg, B, g["origins"])
    s.encoder_bias[:] = g["enc_bias"]
    s.time_steps[:] = g["time0"] + 1
    s.rel_pos[:], s.rel_quat[:] = g["rel0_pos"], g["rel0_quat"]
    for t in range(T):
      d = snapshot_from.load(g, t, m)
      OT.apply_action(cfg, s, d, g["actions"][t])
      OT.post(cfg, m, mo, s, d, u=np.zeros((B, OT.A.NDRAW), np.float32), do_reset=False)
      np.testing.assert_array_equal(s.term_dones[:, 1], g["ref_term_anchor_pos_3d" if full else "ref_term_anchor_pos_z_at_thr3d"][t])
      np.testing.assert_array_equal(s.term_dones[:, 3], g["ref_term_ee_3d" if full else "ref_term_ee_z_at_thr3d"][t])
      rb = OT.robot_bodies(cfg, d)
      OT.update_metrics(cfg, s, d, OT.motion_frame(cfg, mo, s), rb)
      OT.update_command(cfg, mo, s, d, rb, None)
