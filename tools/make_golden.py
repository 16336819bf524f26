"""Generate the committed golden fixtures under tests/golden/ from the float64 oracle.

    python tools/make_golden.py

The reference itself cannot run here or on the GPU box (mujoco / mujoco_warp / warp are not
installable), so these vectors pin the ORACLE (regression) and give the GPU tests a fixed target
that does not need the oracle at run time.  Seeds, sizes and step counts are recorded in the file.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import mjlab_b200  # noqa: E402
from tests import util  # noqa: E402

NWORLD, NSTEP, NCONMAX, NJMAX = 6, 4, 35, 300

if __name__ == "__main__":
  out = ROOT / "tests" / "golden"
  out.mkdir(exist_ok=True)
  for name, seed in (("go1_flat", 11), ("g1_flat", 12)):
    m = mjlab_b200.load_model(name)
    qpos, qvel, ctrl, xfrc = util.random_states(m, NWORLD, seed=seed, xfrc_world=2)
    ds = util.oracle_rollout(m, qpos, qvel, ctrl, xfrc, NSTEP, NCONMAX, NJMAX)
    np.savez_compressed(
      out / f"physics_{name}.npz",
      seed=seed, nstep=NSTEP, nconmax=NCONMAX, njmax=NJMAX,
      qpos0=qpos, qvel0=qvel, ctrl=ctrl, xfrc=xfrc,
      qpos=np.stack([d.qpos for d in ds]), qvel=np.stack([d.qvel for d in ds]),
      qacc=np.stack([d.qacc for d in ds]), sensordata=np.stack([d.sensordata for d in ds]),
      time=np.array([d.time for d in ds]),
      ncon=np.array([len(d.contacts) for d in ds]), nefc=np.array([d.nefc for d in ds]),
      contact_geom=np.stack([np.pad(np.array(util.contact_pairs(d), int).reshape(-1, 2), ((0, NCONMAX - len(d.contacts)), (0, 0)), constant_values=-1) for d in ds]),
      xpos=np.stack([d.xpos for d in ds]), subtree_com=np.stack([d.subtree_com for d in ds]),
    )
    print(name, "ncon", [len(d.contacts) for d in ds], "nefc", [d.nefc for d in ds])
