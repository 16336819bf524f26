"""GPU parity of the elliptic-cone solver instantiations (k_solve_t<*, false, true>) against the float64 oracle, through the C ABI
(`MujocoCfg(cone="elliptic")`, /root/reference/src/mjlab/sim/sim.py:51-56).  Row counts / contact addresses bit-exact; D, aref,
forces, accelerations and sensor data at north_star's 1e-4 relative; the velocity env runs on the elliptic cone end to end."""
import copy

import numpy as np
import pytest
import torch

from tests import util
from tests.test_gpu_parity import load_state, relerr

pytestmark = pytest.mark.gpu


def _sim(m, nw, solver="newton", iterations=20, impratio=2.0):
  from mjlab_b200.sim import MujocoCfg, Simulation, SimulationCfg
  cfg = SimulationCfg(nconmax=35, njmax=300, export_efc_J=True,
                      mujoco=MujocoCfg(timestep=0.005, iterations=iterations, ls_iterations=50, cone="elliptic", solver=solver, impratio=impratio))
  return Simulation(nw, cfg, m, "cuda:0")


@pytest.mark.parametrize("name,solver,fl", [("go1_flat", "newton", 0.0), ("g1_flat", "newton", 0.0), ("go1_flat", "cg", 0.0),
                                            ("go1_flat", "newton", 0.3)])
def test_forward_on_the_elliptic_cone(name, solver, fl):
  import mjlab_b200
  m = copy.deepcopy(mjlab_b200.load_model(name))
  if fl > 0:       # friction-loss rows next to the cone contacts: k_solve_t<false, true, true>
    m.dof_frictionloss = np.concatenate((np.zeros(6), np.linspace(0.05, fl, m.nv - 6)))
  nw = 8
  qpos, qvel, ctrl, xfrc = util.random_states(m, nw, seed=11)
  sim = _sim(m, nw, solver, 100 if solver == "cg" else 20)
  assert sim.mj_model.opt_cone == 1
  load_state(sim, qpos, qvel, ctrl, xfrc)
  sim.forward()
  torch.cuda.synchronize()
  ds = util.oracle_rollout(sim.mj_model, qpos, qvel, ctrl, xfrc, 1, 35, 300, integrate_last=False)
  D = sim.data
  ncone = 0
  tol = 1e-4 if solver == "newton" else 2e-3     # CG stops on the cost-improvement test: looser in float32
  for w, d in enumerate(ds):
    assert int(D.nefc[w]) == d.nefc and int(D.ncon[w]) == len(d.contacts) and int(D.nf[w]) == d.nf == (m.nv - 6 if fl > 0 else 0)
    np.testing.assert_array_equal(D.contact_efc_address[w, : len(d.contacts)].cpu().numpy(), [c.efc_address for c in d.contacts])
    n = d.nefc
    ncone += len(d.efc_cones)
    if n:
      assert relerr(D.efc_J[w, :n].cpu().numpy(), d.efc_J) < 1e-5
      assert relerr(D.efc_D[w, :n].cpu().numpy(), d.efc_D) < 5e-4
      assert relerr(D.efc_aref[w, :n].cpu().numpy(), d.efc_aref) < 1e-4
      f = D.efc_force[w, :n].cpu().numpy()
      assert relerr(f, d.efc_force) < 2 * tol, (w, relerr(f, d.efc_force))
      for a, dim, mu, fri in d.efc_cones:           # the device forces lie in their cones too
        assert f[a] >= 0 and np.linalg.norm(f[a + 1:a + dim] / fri) <= f[a] * (1 + 1e-4) + 1e-4
    assert relerr(D.qacc[w].cpu().numpy(), d.qacc) < tol, (w, relerr(D.qacc[w].cpu().numpy(), d.qacc))
    assert relerr(D.qfrc_constraint[w].cpu().numpy(), d.qfrc_constraint) < 2 * tol
    assert relerr(D.sensordata[w].cpu().numpy(), d.sensordata) < 2 * tol
  assert ncone >= 10


def test_step_trajectory_on_the_elliptic_cone():
  import mjlab_b200
  m = copy.deepcopy(mjlab_b200.load_model("go1_flat"))
  nw, nstep = 6, 5
  qpos, qvel, ctrl, xfrc = util.random_states(m, nw, seed=2)
  sim = _sim(m, nw, impratio=1.0)
  load_state(sim, qpos, qvel, ctrl, xfrc)
  for _ in range(nstep):
    sim.step()
  torch.cuda.synchronize()
  ds = util.oracle_rollout(sim.mj_model, qpos, qvel, ctrl, xfrc, nstep, 35, 300)
  for w, d in enumerate(ds):
    assert int(sim.data.ncon[w]) == len(d.contacts)
    assert relerr(sim.data.qpos[w].cpu().numpy(), d.qpos) < 1e-4 * nstep
    assert relerr(sim.data.qvel[w].cpu().numpy(), d.qvel) < 1e-4 * nstep


def test_elliptic_and_pyramidal_agree_where_no_contact_slides_and_differ_in_row_count():
  """A standing Go1 (keyframe pose, zero velocity): all contacts stick, so both cones give (nearly) the same accelerations, with
  3 rows per frictional contact instead of 4."""
  import mjlab_b200
  from mjlab_b200.sim import MujocoCfg, Simulation, SimulationCfg
  m = mjlab_b200.load_model("go1_flat")
  res = {}
  for cone in ("pyramidal", "elliptic"):
    sim = Simulation(4, SimulationCfg(nconmax=35, njmax=300, mujoco=MujocoCfg(timestep=0.005, iterations=50, ls_iterations=50, cone=cone)),
                     copy.deepcopy(m), "cuda:0")
    sim.data.qpos[:, 2] -= 0.045
    for _ in range(200):
      sim.step()
    torch.cuda.synchronize()
    res[cone] = (sim.data.nefc.cpu().numpy().copy(), sim.data.ncon.cpu().numpy().copy(), sim.data.nl.cpu().numpy().copy(), sim.data.qpos.cpu().numpy().copy())
  (ne_p, nc_p, nl_p, qp), (ne_e, nc_e, nl_e, qe) = res["pyramidal"], res["elliptic"]
  assert nc_p.min() >= 4 and np.array_equal(nc_p, nc_e)
  np.testing.assert_array_equal(ne_p - nl_p, 4 * nc_p)
  np.testing.assert_array_equal(ne_e - nl_e, 3 * nc_e)
  assert np.isfinite(qe).all() and np.abs(qe[:, 2] - qp[:, 2]).max() < 2e-3


@pytest.mark.parametrize("solver", ["newton", "cg"])
def test_velocity_env_runs_on_the_elliptic_cone(solver):
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  from mjlab_b200.sim import MujocoCfg, SimulationCfg
  sim = SimulationCfg(nconmax=35, njmax=300, mujoco=MujocoCfg(timestep=0.005, iterations=10, ls_iterations=20, cone="elliptic", solver=solver))
  env = ManagerBasedRlEnv(VelocityEnvCfg(robot="go1", num_envs=64, seed=3, sim=sim), "cuda:0")
  assert env.sim.mj_model.opt_cone == 1 and env.sim.mj_model.opt_solver == {"newton": 2, "cg": 1}[solver]
  obs, _ = env.reset()
  tot = 0.0
  for _ in range(30):
    obs, rew, term, trunc, _ = env.step(torch.zeros(64, env.vcfg.nj, device="cuda:0"))
    tot += float(rew.sum())
  assert torch.isfinite(obs["policy"]).all() and np.isfinite(tot)
  f = env.scene["feet_ground_contact"].data.force
  assert torch.isfinite(f).all() and float(f.abs().max()) > 1.0
