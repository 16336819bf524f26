"""Shared helpers for parity tests: seeded states and oracle roll-outs."""
import numpy as np

from oracle import mjstep as O


def random_states(m, nworld, seed=0, drop=0.03, jitter=0.1, vel=0.5, ctrl_noise=0.2, xfrc_world=None):
  """Seeded initial states around the init keyframe (some worlds in contact, some in the air)."""
  rng = np.random.default_rng(seed)
  qpos = np.tile(m.key_qpos[0], (nworld, 1))
  qpos[:, 2] -= drop * (np.arange(nworld) % 4) - 0.03
  qpos[:, 7:] += rng.normal(size=(nworld, m.nq - 7)) * jitter
  q = qpos[:, 3:7] + rng.normal(size=(nworld, 4)) * 0.05
  qpos[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
  qvel = rng.normal(size=(nworld, m.nv)) * vel
  ctrl = np.tile(m.key_ctrl[0], (nworld, 1)) + rng.normal(size=(nworld, m.nu)) * ctrl_noise
  xfrc = np.zeros((nworld, m.nbody, 6))
  if xfrc_world is not None:
    xfrc[xfrc_world, 5] = rng.normal(size=6)
  return qpos, qvel, ctrl, xfrc


def oracle_rollout(m, qpos, qvel, ctrl, xfrc, nstep, nconmax, njmax, integrate_last=True):
  out = []
  for w in range(qpos.shape[0]):
    d = O.Data(m)
    d.qpos[:], d.qvel[:], d.ctrl[:] = qpos[w], qvel[w], ctrl[w]
    d.xfrc_applied[:] = xfrc[w]
    for s in range(nstep):
      if s == nstep - 1 and not integrate_last:
        O.forward(m, d, nconmax, njmax)
      else:
        O.step(m, d, nconmax, njmax)
    out.append(d)
  return out


def contact_pairs(d):
  return [tuple(c.geom) for c in d.contacts]
