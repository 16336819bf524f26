"""Reference-motion files of the tracking task.

On-disk format = what the reference's converter writes (src/mjlab/scripts/csv_to_npz.py:206-308) and its `MotionLoader`
reads (src/mjlab/tasks/tracking/mdp/commands.py:32-68): `fps[1]`, `joint_pos[T,nj]`, `joint_vel[T,nj]`,
`body_pos_w[T,nb,3]`, `body_quat_w[T,nb,4]` (w,x,y,z), `body_lin_vel_w[T,nb,3]`, `body_ang_vel_w[T,nb,3]` for ALL robot
bodies; the command then selects the tracked subset.

`synthetic_motion` produces the file of SURVEY.md section 8(d), config 5 (there is no network for real motion data): the
KNEES_BENT pose plus 0.1 rad * sin(2 pi 0.5 t) on every joint, root fixed at its keyframe height, body poses and
velocities obtained the way the reference's converter gets them - by writing each frame's state into the simulator and
running the forward pass (here: one world per frame, one launch).
"""

from __future__ import annotations

import math
from pathlib import Path

import numpy as np
import torch

KEYS = ("joint_pos", "joint_vel", "body_pos_w", "body_quat_w", "body_lin_vel_w", "body_ang_vel_w")


class MotionLoader:
  def __init__(self, motion, body_indexes: torch.Tensor, device: str = "cuda:0"):
    data = np.load(motion) if isinstance(motion, (str, Path)) else motion
    t = lambda k: torch.as_tensor(np.asarray(data[k]), dtype=torch.float32, device=device).contiguous()  # noqa: E731
    self.fps = float(np.asarray(data["fps"]).reshape(-1)[0]) if "fps" in data else 50.0
    self.joint_pos, self.joint_vel = t("joint_pos"), t("joint_vel")
    self._body_pos_w, self._body_quat_w = t("body_pos_w"), t("body_quat_w")
    self._body_lin_vel_w, self._body_ang_vel_w = t("body_lin_vel_w"), t("body_ang_vel_w")
    self._body_indexes = body_indexes
    self.time_step_total = self.joint_pos.shape[0]

  body_pos_w = property(lambda self: self._body_pos_w[:, self._body_indexes])
  body_quat_w = property(lambda self: self._body_quat_w[:, self._body_indexes])
  body_lin_vel_w = property(lambda self: self._body_lin_vel_w[:, self._body_indexes])
  body_ang_vel_w = property(lambda self: self._body_ang_vel_w[:, self._body_indexes])


def synthetic_motion(model, T: int = 500, fps: float = 50.0, amplitude: float = 0.1, freq_hz: float = 0.5, device: str = "cuda:0") -> dict:
  """dict with the npz keys above for the compiled scene `model` (robot entity 'robot')."""
  from .entity import Entity
  from .sim import MujocoCfg, Simulation, SimulationCfg
  import copy
  sim = Simulation(T, SimulationCfg(nconmax=35, njmax=250, mujoco=MujocoCfg(timestep=0.005, iterations=10, ls_iterations=20)),
                   copy.deepcopy(model), device)
  robot = Entity("robot", sim)
  t = torch.arange(T, device=device, dtype=torch.float32) / fps
  w = 2 * math.pi * freq_hz
  jp = robot.data.default_joint_pos + amplitude * torch.sin(w * t)[:, None]
  jv = (amplitude * w * torch.cos(w * t))[:, None].expand_as(jp).contiguous()
  lo, hi = robot.data.soft_joint_pos_limits[..., 0], robot.data.soft_joint_pos_limits[..., 1]
  inside = (jp >= lo) & (jp <= hi)
  jp = torch.clamp(jp, lo, hi)
  jv = torch.where(inside, jv, torch.zeros_like(jv))
  robot.write_root_state_to_sim(robot.data.default_root_state.clone())
  robot.write_joint_state_to_sim(jp, jv)
  sim.forward()
  torch.cuda.synchronize()
  d = robot.data
  out = {"fps": np.array([fps]), "joint_pos": d.joint_pos, "joint_vel": d.joint_vel, "body_pos_w": d.body_link_pos_w,
         "body_quat_w": d.body_link_quat_w, "body_lin_vel_w": d.body_link_lin_vel_w, "body_ang_vel_w": d.body_link_ang_vel_w}
  return {k: (v.cpu().numpy().astype(np.float32) if isinstance(v, torch.Tensor) else v) for k, v in out.items()}


def save_motion(path, motion: dict) -> None:
  np.savez(path, **motion)
