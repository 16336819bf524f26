"""The env-step oracle (oracle/envstep.py) against golden vectors produced by THE REFERENCE'S OWN FUNCTIONS
(tools/make_mdp_golden.py): rewards, terminations, observations, air-time state, swing-height state, heading control."""
import numpy as np
import pytest

import mjlab_b200
from mjlab_b200 import envabi as A
from mjlab_b200 import envcfg
from oracle import envstep as OE
from tests.envutil import init_state_from_golden, load_golden, snapshot

TOL = dict(rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("fixture", ["g1", "go1", "g1_fallen", "go1_fallen"])
def test_oracle_matches_reference_functions(fixture):
  robot = fixture.split("_")[0]
  m = mjlab_b200.load_model(f"{robot}_flat")
  cfg = envcfg.velocity_cfg(m, robot)
  g = load_golden(fixture)
  T, B = g["actions"].shape[:2]
  s = OE.make_state(cfg, B)
  init_state_from_golden(s, g)
  nterm_checked = 0
  for t in range(T):
    d = snapshot(g, t, m)
    OE.apply_action(cfg, s, d, g["actions"][t])
    for k in range(4):
      OE.update_air_time(cfg, s, g["sub_sensordata"][t, k].astype(np.float32), g["sub_time"][t, k])
    np.testing.assert_allclose(s.air_state.transpose(1, 0, 2), g["ref_air"][t], **TOL)
    OE.post(cfg, m, s, d, u=np.zeros((B, A.NDRAW), np.float32), do_reset=False)
    np.testing.assert_array_equal(s.term_dones[:, 0], g["ref_time_out"][t])
    np.testing.assert_array_equal(s.term_dones[:, 1], g["ref_fell_over"][t])
    np.testing.assert_array_equal(s.term_dones[:, 2], g["ref_illegal"][t])
    nterm_checked += int(s.term_dones.sum())
    np.testing.assert_allclose(s.last_terms, g["ref_rew_terms"][t], **TOL)
    np.testing.assert_allclose(s.reward, g["ref_reward"][t], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(s.peak_heights, g["ref_peak"][t], **TOL)
    R = OE.root_kinematics(cfg, d)
    np.testing.assert_allclose(R.lin_b, g["ref_root_lin_vel_b"][t], **TOL)
    np.testing.assert_allclose(R.ang_b, g["ref_root_ang_vel_b"][t], **TOL)
    np.testing.assert_allclose(R.pg, g["ref_projected_gravity"][t], **TOL)
    OE.update_command(cfg, s, R.quat)
    np.testing.assert_allclose(OE.heading_w(R.quat), g["ref_heading_w"][t], **TOL)
    np.testing.assert_allclose(s.heading_error, g["ref_heading_error"][t], rtol=2e-5, atol=5e-6)
    np.testing.assert_allclose(s.command, g["ref_cmd_after"][t], rtol=2e-5, atol=5e-6)
    cfg_nonoise = cfg
    OE.observations(cfg_nonoise, s, d, R, u=None)
    np.testing.assert_allclose(s.obs_policy, g["ref_obs_policy"][t], **TOL)
    np.testing.assert_allclose(s.obs_critic, g["ref_obs_critic"][t], **TOL)
  # generic terminations: state-only flags, compared on every snapshot
  cfg2 = envcfg.velocity_cfg(m, robot)
  cfg2.min_root_height, cfg2.nan_termination = float(g["min_root_height"]), True
  for t in range(T):
    d = snapshot(g, t, m)
    flags = OE.terminations(cfg2, s, d, OE.root_kinematics(cfg2, d))
    np.testing.assert_array_equal(flags[3], g["ref_root_low"][t])
    np.testing.assert_array_equal(flags[4], g["ref_nan"][t])
  assert g["ref_root_low"].any() and not g["ref_root_low"].all()
  assert nterm_checked > 0  # the fixture exercises at least one termination flag
  if fixture.endswith("fallen"):   # ... and these make the reference's own bad_orientation / illegal_contact return True
    assert g["ref_fell_over"].any() and not g["ref_fell_over"].all()
    assert robot != "go1" or (g["ref_illegal"].any() and not g["ref_illegal"].all())


def test_philox_known_answer():
  # Random123 known-answer vectors for philox4x32-10 (kat_vectors): zero and all-ones inputs
  o = OE.philox4x32(0, 0, 0, 0, 0, 0)
  assert [hex(int(x)) for x in o.reshape(-1)] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
  o = OE.philox4x32(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)
  assert [hex(int(x)) for x in o.reshape(-1)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


def test_draws_are_uniform_and_reproducible():
  u = OE.draws(1234, 512, 7)
  assert u.shape == (512, A.NDRAW) and u.dtype == np.float32
  assert 0 <= u.min() and u.max() < 1
  assert abs(u.mean() - 0.5) < 0.01
  np.testing.assert_array_equal(u, OE.draws(1234, 512, 7))
  assert not np.array_equal(u, OE.draws(1234, 512, 8))
  assert not np.array_equal(u, OE.draws(1235, 512, 7))
