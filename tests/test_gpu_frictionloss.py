"""GPU parity of the friction-loss constraint rows (k_solve_t<*, true>) against the float64 oracle, through the C ABI.

Row counts / order bit-exact; forces and accelerations at north_star's 1e-4 relative (fp32 device arithmetic vs fp64).
Also: per-world friction loss through `expand_model_fields` + direct writes to `sim.model.dof_frictionloss` on a model whose
compiled friction loss is ZERO (the solver is re-planned for nv friction rows and switches kernel instantiation), and the
reference's `frictionloss` actuator parameter reaching the joints of the compiled robot.
"""
import copy

import numpy as np
import pytest
import torch

from tests import util
from tests.test_gpu_parity import load_state, make_sim, relerr

pytestmark = pytest.mark.gpu


def _go1_with_friction(solver=2):
  import mjlab_b200
  m = copy.deepcopy(mjlab_b200.load_model("go1_flat"))
  fl = np.zeros(m.nv)
  fl[6:] = np.linspace(0.05, 0.6, m.nv - 6)
  fl[8] = 0.0
  m.dof_frictionloss = fl
  m.opt_solver = solver
  return m


@pytest.mark.parametrize("solver", [2, 1])
def test_forward_with_friction_rows(solver):
  from mjlab_b200.sim import MujocoCfg, Simulation, SimulationCfg
  m = _go1_with_friction(solver)
  nw = 8
  qpos, qvel, ctrl, xfrc = util.random_states(m, nw, seed=11)
  qvel[:, 6:] *= np.array([0.0, 0.02, 1.0, 5.0, 0.2, 1.0, 0.5, 2.0])[:, None]
  cfg = SimulationCfg(nconmax=35, njmax=300, export_efc_J=True,
                      mujoco=MujocoCfg(timestep=0.005, iterations=100, ls_iterations=50, solver={2: "newton", 1: "cg"}[solver]))
  sim = Simulation(nw, cfg, m, "cuda:0")
  load_state(sim, qpos, qvel, ctrl, xfrc)
  sim.forward()
  torch.cuda.synchronize()
  ds = util.oracle_rollout(m, qpos, qvel, ctrl, xfrc, 1, 35, 300, integrate_last=False)
  D = sim.data
  nfr = m.nv - 7
  for w, d in enumerate(ds):
    assert int(D.nf[w]) == d.nf == nfr
    assert int(D.nl[w]) == int(np.sum(d.efc_type == 0))
    assert int(D.nefc[w]) == d.nefc
    n = d.nefc
    assert relerr(D.efc_J[w, :n].cpu().numpy(), d.efc_J) < 1e-5
    assert relerr(D.efc_D[w, :n].cpu().numpy(), d.efc_D) < 5e-4
    assert relerr(D.efc_aref[w, :n].cpu().numpy(), d.efc_aref) < 1e-4
    # friction forces: bounded by the row's frictionloss exactly, equal to the oracle's at 1e-4 of the largest force
    ff = D.efc_force[w, :nfr].cpu().numpy()
    assert np.all(np.abs(ff) <= d.efc_frictionloss[:nfr].astype(np.float32))
    tol = 1e-4 if solver == 2 else 2e-3    # CG stops on the cost-improvement test: looser in float32
    assert relerr(D.efc_force[w, :n].cpu().numpy(), d.efc_force) < 2 * tol, (w, relerr(D.efc_force[w, :n].cpu().numpy(), d.efc_force))
    assert relerr(D.qacc[w].cpu().numpy(), d.qacc) < tol, (w, relerr(D.qacc[w].cpu().numpy(), d.qacc))
    assert relerr(D.qfrc_constraint[w].cpu().numpy(), d.qfrc_constraint) < 2 * tol


def test_step_trajectory_with_friction_rows():
  m = _go1_with_friction()
  nw, nstep = 6, 5
  qpos, qvel, ctrl, xfrc = util.random_states(m, nw, seed=2)
  sim = make_sim(m, nw)
  load_state(sim, qpos, qvel, ctrl, xfrc)
  for _ in range(nstep):
    sim.step()
  torch.cuda.synchronize()
  ds = util.oracle_rollout(m, qpos, qvel, ctrl, xfrc, nstep, 35, 300)
  for w, d in enumerate(ds):
    assert int(sim.data.ncon[w]) == len(d.contacts)
    assert relerr(sim.data.qpos[w].cpu().numpy(), d.qpos) < 1e-4 * nstep
    assert relerr(sim.data.qvel[w].cpu().numpy(), d.qvel) < 1e-4 * nstep


def test_friction_free_model_is_bit_identical_to_before_and_per_world_frictionloss(go1):
  """A model with frictionloss = 0 runs the friction-free kernel (nf = 0 everywhere).  After expand_model_fields the field is
  per world: worlds left at 0 keep bit-identical results, worlds given a friction loss match the oracle run with that value."""
  nw = 6
  qpos, qvel, ctrl, xfrc = util.random_states(go1, nw, seed=4)
  base = make_sim(go1, nw)
  load_state(base, qpos, qvel, ctrl, xfrc)
  base.forward()
  torch.cuda.synchronize()
  assert int(base.data.nf.max()) == 0
  sim = make_sim(go1, nw)
  sim.expand_model_fields(("dof_frictionloss",))
  fl = sim.model.dof_frictionloss
  assert fl.shape == (nw, go1.nv) and fl.stride(0) == go1.nv
  fl[1, 6:] = 0.3
  fl[4, 9] = 0.15
  load_state(sim, qpos, qvel, ctrl, xfrc)
  sim.forward()
  torch.cuda.synchronize()
  assert sim.data.nf.cpu().tolist() == [0, go1.nv - 6, 0, 0, 1, 0]
  for w in (0, 2, 3, 5):
    assert torch.equal(sim.data.qacc[w], base.data.qacc[w])
    assert torch.equal(sim.data.efc_force[w], base.data.efc_force[w])
  for w in (1, 4):
    m = copy.deepcopy(go1)
    m.dof_frictionloss = fl[w].cpu().numpy().astype(float)
    d = util.oracle_rollout(m, qpos[w : w + 1], qvel[w : w + 1], ctrl[w : w + 1], xfrc[w : w + 1], 1, 35, 300, integrate_last=False)[0]
    assert int(sim.data.nefc[w]) == d.nefc
    assert relerr(sim.data.qacc[w].cpu().numpy(), d.qacc) < 1e-4
    assert relerr(sim.data.efc_force[w, : d.nefc].cpu().numpy(), d.efc_force) < 2e-4
  # graphs were re-captured with the new plan: stepping works and stays finite
  for _ in range(3):
    sim.step()
  torch.cuda.synchronize()
  assert bool(torch.isfinite(sim.data.qpos).all())


def test_randomize_field_dof_frictionloss_in_env():
  """`randomize_field("dof_frictionloss")` (envs/mdp/events.py:267) on the velocity env: per-world values land in the model,
  the env steps with friction rows, and the joints of a world with a large friction loss move less under the same actions."""
  from mjlab_b200 import events
  from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg
  torch.manual_seed(0)
  env = ManagerBasedRlEnv(VelocityEnvCfg(robot="go1", num_envs=64, seed=3), "cuda:0")
  env.sim.expand_model_fields(("dof_frictionloss",))
  env.reset()
  events.randomize_field(env, None, "dof_frictionloss", (0.2, 0.6), operation="abs", joint_names=(".*",))
  fl = env.sim.model.dof_frictionloss
  assert float(fl[:, 6:].min()) >= 0.2 and float(fl[:, 6:].max()) <= 0.6 and float(fl[:, :6].abs().max()) == 0.0
  assert float(fl[:, 6:].std()) > 0.05
  fl[:32, 6:] = 0.0
  fl[32:, 6:] = 3.0
  act = torch.zeros(64, env.action_manager.total_action_dim, device="cuda:0")
  for _ in range(5):
    a = torch.rand_like(act) * 2 - 1
    a[32:] = a[:32]
    env.step(a)
  torch.cuda.synchronize()
  nf = env.sim.data.nf
  assert int(nf[:32].max()) == 0 and int(nf[32:].min()) == env.sim.model.dof_frictionloss.shape[1] - 6
  assert bool(torch.isfinite(env.sim.data.qpos).all())
  jv = env.sim.data.qvel[:, 6:].abs().mean(dim=1)
  assert float(jv[32:].mean()) < 0.7 * float(jv[:32].mean())
