"""Policy export: the deployment metadata of `rl/exporter_utils.py:23-62` and the policy
network itself (actor + observation normaliser, `tasks/velocity/rl/exporter.py:11-22`).

`export_policy_as_onnx` needs the `onnx` package like the reference does (it raises a clear error when the package is
missing - this image has none); `export_policy_as_jit` writes the same network as TorchScript plus the metadata as a JSON
side-car, which needs nothing beyond torch.
"""

from __future__ import annotations

import json
import os

import numpy as np
import torch
from torch import nn


def list_to_csv_str(arr, *, decimals: int = 3, delimiter: str = ",") -> str:
  """rl/exporter_utils.py:11-20."""
  fmt = f"{{:.{decimals}f}}"
  return delimiter.join(fmt.format(x) if isinstance(x, (int, float)) else str(x) for x in arr)


def get_base_metadata(env, run_path: str) -> dict:
  """rl/exporter_utils.py:23-62: joint names / gains / default pose in natural joint order, command and observation term
  names, action scale."""
  robot = env.scene["robot"]
  m, cfg = env.sim.mj_model, env.vcfg
  ctrl = np.asarray(cfg.ctrl_ids)
  return {
    "run_path": run_path,
    "joint_names": list(robot.joint_names),
    "joint_stiffness": np.asarray(m.actuator_gainprm)[ctrl, 0].tolist(),
    "joint_damping": (-np.asarray(m.actuator_biasprm)[ctrl, 2]).tolist(),
    "default_joint_pos": np.asarray(cfg.default_joint_pos, float).tolist(),
    "command_names": list(env.command_manager.active_terms),
    "observation_names": list(env.observation_manager.active_terms["policy"]),
    "action_scale": np.asarray(cfg.action_scale, float).tolist(),
  }


class _PolicyModule(nn.Module):
  """obs -> actions = actor(normalizer(obs)): what `_OnnxPolicyExporter` traces (the network's internal padding of the
  observation / action widths stays inside)."""

  def __init__(self, actor_critic):
    super().__init__()
    import copy
    self.ac = copy.deepcopy(actor_critic)

  def forward(self, obs: torch.Tensor) -> torch.Tensor:
    return self.ac.act_inference(obs)


def export_policy_as_jit(actor_critic, env, path: str, run_path: str = "", filename: str = "policy.pt") -> str:
  os.makedirs(path, exist_ok=True)
  mod = _PolicyModule(actor_critic).to("cpu").eval()
  traced = torch.jit.trace(mod, torch.zeros(1, actor_critic.n_actor_obs))
  out = os.path.join(path, filename)
  traced.save(out)
  meta = get_base_metadata(env, run_path)
  with open(os.path.splitext(out)[0] + ".json", "w") as f:
    json.dump({k: (list_to_csv_str(v) if isinstance(v, list) else str(v)) for k, v in meta.items()}, f, indent=1)
  return out


def export_policy_as_onnx(actor_critic, env, path: str, run_path: str = "", filename: str = "policy.onnx") -> str:
  """export_velocity_policy_as_onnx + attach_onnx_metadata (tasks/velocity/rl/exporter.py:11-36)."""
  try:
    import onnx
  except ImportError as e:
    raise RuntimeError("ONNX export needs the `onnx` package (use export_policy_as_jit without it)") from e
  os.makedirs(path, exist_ok=True)
  mod = _PolicyModule(actor_critic).to("cpu").eval()
  out = os.path.join(path, filename)
  torch.onnx.export(mod, torch.zeros(1, actor_critic.n_actor_obs), out, input_names=["obs"], output_names=["actions"], dynamo=False)
  model = onnx.load(out)
  for k, v in get_base_metadata(env, run_path).items():
    entry = onnx.StringStringEntryProto()
    entry.key, entry.value = k, (list_to_csv_str(v) if isinstance(v, list) else str(v))
    model.metadata_props.append(entry)
  onnx.save(model, out)
  return out
