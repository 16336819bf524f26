"""CPU: the oracle and the single-lane host build of the kernel source against the committed golden vectors."""
import numpy as np
import pytest

import mjlab_b200
from tests import util
from tests.hostemu.harness import HostSim

GOLD = {n: np.load(f"tests/golden/physics_{n}.npz") for n in ("go1_flat", "g1_flat")}


@pytest.mark.parametrize("name", ["go1_flat", "g1_flat"])
def test_oracle_reproduces_golden(name):
  g = GOLD[name]
  m = mjlab_b200.load_model(name)
  nw = 3  # a prefix of the worlds keeps the CPU suite short
  ds = util.oracle_rollout(m, g["qpos0"][:nw], g["qvel0"][:nw], g["ctrl"][:nw], g["xfrc"][:nw], int(g["nstep"]),
                           int(g["nconmax"]), int(g["njmax"]))
  for w, d in enumerate(ds):
    np.testing.assert_allclose(d.qpos, g["qpos"][w], rtol=0, atol=1e-12)
    np.testing.assert_allclose(d.qvel, g["qvel"][w], rtol=0, atol=1e-10)
    assert len(d.contacts) == g["ncon"][w] and d.nefc == g["nefc"][w]
    assert util.contact_pairs(d) == [tuple(p) for p in g["contact_geom"][w][: len(d.contacts)]]


@pytest.mark.parametrize("name", ["go1_flat", "g1_flat"])
@pytest.mark.parametrize("real,tol_q,tol_v", [("double", 1e-10, 1e-8), ("float", 2e-5, 2e-3)])
def test_kernel_source_on_host_matches_golden(name, real, tol_q, tol_v):
  """Same .cuh sources as the CUDA library, compiled single-lane for the host (test harness only)."""
  g = GOLD[name]
  m = mjlab_b200.load_model(name)
  nw = g["qpos0"].shape[0]
  hs = HostSim(m, nw, nconmax=int(g["nconmax"]), njmax=int(g["njmax"]), real=real)
  hs.qpos[:], hs.qvel[:], hs.ctrl[:], hs.xfrc_applied[:] = g["qpos0"], g["qvel0"], g["ctrl"], g["xfrc"]
  for _ in range(int(g["nstep"])):
    hs.step()
  # bit-exact: contact counts, pair lists, row counts
  assert hs.ncon.tolist() == g["ncon"].tolist()
  assert hs.nefc.tolist() == g["nefc"].tolist()
  for w in range(nw):
    n = int(g["ncon"][w])
    assert hs.contact_geom[w, :n].tolist() == g["contact_geom"][w, :n].tolist()
  assert np.abs(hs.qpos - g["qpos"]).max() < tol_q
  assert np.abs(hs.qvel - g["qvel"]).max() / np.abs(g["qvel"]).max() < tol_v
  assert np.abs(hs.time - g["time"]).max() < 1e-6


def test_stagewise_host_vs_oracle():
  """Every stage output of one forward pass (float64 host build) against the oracle, to 1e-9."""
  from oracle import mjstep as O
  m = mjlab_b200.load_model("g1_flat")
  qpos, qvel, ctrl, xfrc = util.random_states(m, 3, seed=5, xfrc_world=1)
  hs = HostSim(m, 3, nconmax=64, njmax=300, real="double")
  hs.qpos[:], hs.qvel[:], hs.ctrl[:], hs.xfrc_applied[:] = qpos, qvel, ctrl, xfrc
  hs.forward()
  ds = util.oracle_rollout(m, qpos, qvel, ctrl, xfrc, 1, 64, 300, integrate_last=False)
  for w, d in enumerate(ds):
    n = d.nefc
    pairs = [("xpos", d.xpos), ("xquat", d.xquat), ("xipos", d.xipos), ("geom_xpos", d.geom_xpos),
             ("site_xpos", d.site_xpos), ("subtree_com", d.subtree_com), ("cdof", d.cdof), ("cvel", d.cvel),
             ("qfrc_bias", d.qfrc_bias), ("qfrc_passive", d.qfrc_passive), ("qfrc_actuator", d.qfrc_actuator),
             ("actuator_force", d.actuator_force), ("qfrc_smooth", d.qfrc_smooth), ("qacc_smooth", d.qacc_smooth),
             ("qacc", d.qacc), ("qfrc_constraint", d.qfrc_constraint), ("sensordata", d.sensordata),
             ("subtree_angmom", d.subtree_angmom), ("subtree_linvel", d.subtree_linvel)]
    for fld, want in pairs:
      got = getattr(hs, fld)[w]
      assert np.abs(np.asarray(got).reshape(np.shape(want)) - want).max() <= 1e-9 * max(1.0, np.abs(want).max()), (w, fld)
    assert np.abs(hs.qM[w][:, : m.nv] - d.qM).max() < 1e-12
    if n:
      assert np.abs(hs.efc_J[w, :n] - d.efc_J).max() < 1e-12
      assert np.abs(hs.efc_D[w, :n] - d.efc_D).max() < 1e-9 * np.abs(d.efc_D).max()
      assert np.abs(hs.efc_aref[w, :n] - d.efc_aref).max() < 1e-9 * max(1.0, np.abs(d.efc_aref).max())
      assert np.abs(hs.efc_force[w, :n] - d.efc_force).max() < 1e-8 * max(1.0, np.abs(d.efc_force).max())
    assert hs.solver_niter[w] == d.solver_niter
    # state must be untouched by forward()
    np.testing.assert_array_equal(hs.qvel[w], qvel[w])
    assert hs.time[w] == 0.0


def test_nconmax_and_njmax_overflow_match_oracle():
  """Contact / row overflow is truncated deterministically, identically in oracle and kernel source."""
  from oracle import mjstep as O
  m = mjlab_b200.load_model("g1_flat")
  qpos, qvel, ctrl, xfrc = util.random_states(m, 2, seed=6, drop=0.06)
  for nconmax, njmax in ((5, 300), (35, 14)):
    hs = HostSim(m, 2, nconmax=nconmax, njmax=njmax, real="double")
    hs.qpos[:], hs.qvel[:], hs.ctrl[:] = qpos, qvel, ctrl
    hs.forward()
    ds = util.oracle_rollout(m, qpos, qvel, ctrl, xfrc, 1, nconmax, njmax, integrate_last=False)
    for w, d in enumerate(ds):
      assert hs.ncon[w] == len(d.contacts) <= nconmax
      assert hs.nefc[w] == d.nefc <= njmax
      assert np.abs(hs.qacc[w] - d.qacc).max() < 1e-8 * max(1.0, np.abs(d.qacc).max())
    assert hs.ncon_dropped.sum() > 0 or hs.nefc_dropped.sum() > 0
