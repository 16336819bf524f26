"""CG solver (mjSOL_CG: Polak-Ribiere directions preconditioned with M, the solver the reference's tests/test_sim.py:54 selects):
the kernel source (host build) against the oracle iteration by iteration, and CG against Newton - both minimise the same convex
cost, so with enough iterations they reach the same accelerations and forces."""
import copy

import numpy as np

import mjlab_b200
from oracle import mjstep as O
from tests import util
from tests.hostemu.harness import HostSim


def test_cg_kernel_source_matches_oracle_and_newton():
  m0 = mjlab_b200.load_model("go1_flat")
  nw = 6
  qpos, qvel, ctrl, xfrc = util.random_states(m0, nw, seed=5)
  res = {}
  for solver, iters in ((1, 100), (2, 100)):
    m = copy.deepcopy(m0)
    m.opt_solver, m.opt_iterations, m.opt_ls_iterations = solver, iters, 50
    hs = HostSim(m, nw, 35, 300, real="double")
    hs.qpos[:], hs.qvel[:], hs.ctrl[:] = qpos, qvel, ctrl
    hs.forward()
    ds = util.oracle_rollout(m, qpos, qvel, ctrl, xfrc, 1, 35, 300, integrate_last=False)
    for w, d in enumerate(ds):
      assert hs.nefc[w] == d.nefc and hs.solver_niter[w] == d.solver_niter, (solver, w, hs.solver_niter[w], d.solver_niter)
      np.testing.assert_allclose(hs.qacc[w], d.qacc, rtol=1e-9, atol=1e-9)
      if d.nefc:
        np.testing.assert_allclose(hs.efc_force[w, : d.nefc], d.efc_force, rtol=1e-8, atol=1e-8)
    res[solver] = (hs.qacc.copy(), hs.solver_niter.copy(), hs.nefc.copy())
  cgq, cgn, nefc = res[1]
  nq, nn, _ = res[2]
  assert nefc.max() > 0 and cgn[nefc > 0].min() >= 1
  assert cgn.mean() > nn.mean()                          # first-order directions need more iterations than Newton's ...
  np.testing.assert_allclose(cgq, nq, rtol=1e-3, atol=2e-2)   # ... and end at the same minimum (|qacc| up to 500; CG stops on the cost-improvement test)
