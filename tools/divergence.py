"""N-step trajectory divergence of the CUDA physics step (float32) from the float64 CPU build of the same algorithm
(oracle/cport, which reproduces the numpy oracle to 1e-9): the statistic BASELINE.json's north_star asks for.

  python tools/divergence.py [robot] [nworld] > gpurun_out/divergence.json       # on the GPU box
  python tools/divergence.py g1_rough 128                                        # robots spread over the terrain tiles

Seeded states around the init keyframe (tests/util.random_states: some worlds standing, some falling), constant
controls, `nconmax=35, njmax=300, dt=0.005`.  Reported at N = 1, 2, 5, 10, 20, 50, 100, 200 steps: per-world max-norm
error of qpos and qvel (median / p90 / max over worlds) and the fraction of worlds whose contact COUNT history is still
bit-identical.  Contacts make the dynamics non-smooth, so once a contact is made a step earlier or later in one of the
two arithmetic widths the trajectories separate; before that the error grows roughly linearly at ~1e-6 per step.
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mjlab_b200  # noqa: E402
from mjlab_b200.sim import MujocoCfg, Simulation, SimulationCfg  # noqa: E402
from oracle.cport import CpuSim  # noqa: E402
from tests.util import random_states  # noqa: E402

robot = sys.argv[1] if len(sys.argv) > 1 else "g1"
nworld = int(sys.argv[2]) if len(sys.argv) > 2 else 128
m = mjlab_b200.load_model(robot if robot.endswith("_rough") else f"{robot}_flat")
sim = Simulation(nworld, SimulationCfg(nconmax=35, njmax=300, mujoco=MujocoCfg(timestep=0.005, iterations=10, ls_iterations=20)), m, "cuda:0")
qpos, qvel, ctrl, _ = random_states(m, nworld, seed=123, vel=0.2)
if robot.endswith("_rough"):   # one robot per sub-terrain origin (rows = difficulty), shifted onto the steps
  rng = np.random.default_rng(7)
  R, C = m.terrain_origins.shape[:2]
  for w in range(nworld):
    o = m.terrain_origins[rng.integers(0, R), rng.integers(0, C)]
    qpos[w, :2] += o[:2] + rng.uniform(-2.5, 2.5, 2)
    qpos[w, 2] += o[2] + 0.05
cpu = CpuSim(m, nworld, 35, 300)
cpu.qpos[:], cpu.qvel[:], cpu.ctrl[:] = qpos, qvel, ctrl
f32 = lambda a: torch.as_tensor(a, dtype=torch.float32, device="cuda:0")  # noqa: E731
sim.data.qpos[:], sim.data.qvel[:], sim.data.ctrl[:] = f32(qpos), f32(qvel), f32(ctrl)
marks = (1, 2, 5, 10, 20, 50, 100, 200)
same = np.ones(nworld, bool)
out = []
for n in range(1, marks[-1] + 1):
  sim.step()
  cpu.step(1, 16)
  same &= sim.data.ncon.cpu().numpy().reshape(-1) == cpu.ncon.reshape(-1)
  if n in marks:
    eq = np.abs(sim.data.qpos.cpu().numpy() - cpu.qpos).max(1)
    ev = np.abs(sim.data.qvel.cpu().numpy() - cpu.qvel).max(1)
    q = lambda a: [float(np.median(a)), float(np.quantile(a, 0.9)), float(a.max())]  # noqa: E731
    out.append({"nstep": n, "qpos_err_median_p90_max": q(eq), "qvel_err_median_p90_max": q(ev), "qpos_err_median_while_contacts_identical": float(np.median(eq[same])) if same.any() else None,
                "worlds_with_identical_contact_count_history": float(same.mean())})
print(json.dumps({"robot": robot, "nworld": nworld, "dtype_gpu": "f32", "dtype_ref": "f64", "trajectory": out}, indent=1))
