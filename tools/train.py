"""Train a task with the PPO learner on the B200 engine (the reference's `train` CLI, src/mjlab/scripts/train.py:46-171,
minus logging / checkpoint plumbing).

  python tools/train.py --task velocity --robot g1 --num-envs 4096 --iterations 300
  python tools/train.py --task tracking --iterations 300                       # synthetic clip unless --motion-file is given
  python tools/train.py --task rough --num-envs 8192 --iterations 300          # G1 on the ROUGH_TERRAINS_CFG box grid
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/train.py ...   # --gpu-ids path
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mjlab_b200.env import ManagerBasedRlEnv, VelocityEnvCfg  # noqa: E402
from mjlab_b200.ppo import OnPolicyRunner, PpoCfg  # noqa: E402
from mjlab_b200.vecenv import RslRlVecEnvWrapper  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--task", default="velocity", choices=["velocity", "tracking", "rough"])
ap.add_argument("--motion-file", default=None)
ap.add_argument("--robot", default="g1")
ap.add_argument("--num-envs", type=int, default=4096)
ap.add_argument("--iterations", type=int, default=300)
ap.add_argument("--forward-mode", default="reference")
ap.add_argument("--log-every", type=int, default=10)
ap.add_argument("--out", default="")
ap.add_argument("--matmul-precision", default="high", choices=["highest", "high"])
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
if world > 1:
  dist.init_process_group("nccl", device_id=torch.device(dev))
if a.task == "tracking":
  from mjlab_b200.tracking_env import TrackingEnvCfg, TrackingRlEnv
  env = TrackingRlEnv(TrackingEnvCfg(num_envs=a.num_envs, seed=42 + rank, forward_mode=a.forward_mode, motion_file=a.motion_file), dev)
else:
  env = ManagerBasedRlEnv(VelocityEnvCfg(robot=a.robot, terrain="rough" if a.task == "rough" else "flat", num_envs=a.num_envs,
                                         seed=42 + rank, forward_mode=a.forward_mode), dev)
runner = OnPolicyRunner(RslRlVecEnvWrapper(env), PpoCfg(matmul_precision=a.matmul_precision))
# the first iterations capture the roll-out and update graphs (and run them eagerly once): time the rest
warm = min(5, max(a.iterations - 1, 0))
hist = runner.learn(warm, log_every=a.log_every)
torch.cuda.synchronize()
t0 = time.perf_counter()
hist = runner.learn(a.iterations - warm, log_every=a.log_every)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
# split of one iteration: roll-out (env steps + policy inference) vs PPO update
torch.cuda.synchronize(); t1 = time.perf_counter()
ro = runner.collect()
torch.cuda.synchronize(); t2 = time.perf_counter()
runner.update(ro)
torch.cuda.synchronize(); t3 = time.perf_counter()
if rank == 0:
  sps = (a.iterations - warm) * 24 * a.num_envs * world / dt
  summary = {"task": a.task, "robot": a.robot, "num_envs_per_gpu": a.num_envs, "n_gpus": world, "iterations": a.iterations, "timed_iterations": a.iterations - warm, "wall_s": dt, "train_env_steps_per_s": sps,
             "iteration_split_ms": {"rollout_24_steps": 1e3 * (t2 - t1), "ppo_update": 1e3 * (t3 - t2)},
             "first": hist[0], "last": hist[-1]}
  print(json.dumps(summary))
  if a.out:
    Path(a.out).write_text(json.dumps({"summary": summary, "history": hist}))
if world > 1:
  dist.destroy_process_group()
