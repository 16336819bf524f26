"""Multi-GPU is 'replicas only' (SURVEY.md 8e): worlds shard across ranks with no data-path collective.
The only cross-rank logic is seeding (seed + rank, src/mjlab/scripts/train.py:59) and the max-over-ranks
timing reduction of bench.py; both are exercised here with world_size 2 on gloo."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]

WORKER = r'''
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["B2_ROOT"])
from tests import util
import mjlab_b200
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
m = mjlab_b200.load_model("go1_flat")
# every rank owns an independent batch: same model, seed + rank
qpos, qvel, ctrl, _ = util.random_states(m, 4, seed=42 + rank)
sig = torch.tensor([float(np.abs(qpos).sum()), float(np.abs(qvel).sum())], dtype=torch.float64)
gathered = [torch.zeros_like(sig) for _ in range(world)]
dist.all_gather(gathered, sig)
assert not torch.equal(gathered[0], gathered[1]), "ranks must not simulate identical batches"
# max-over-ranks timing, whole-job throughput
t = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() == 15.0
value = 100 * 4 * world / (t.item() * 1e-3)
if rank == 0:
  print(json.dumps({"value": value}))
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_replicas(tmp_path):
  script = tmp_path / "worker.py"
  script.write_text(WORKER)
  env = dict(os.environ, B2_ROOT=str(ROOT))
  out = subprocess.run(
    [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
     "--master-port", "29731", str(script)], env=env, capture_output=True, text=True, timeout=240)
  assert out.returncode == 0, out.stderr[-2000:]
  assert '"value": 53333.3' in out.stdout
