"""Dump the REFERENCE'S OWN task configuration objects (src/mjlab/tasks/registry.py `load_env_cfg`, built by
tasks/velocity/velocity_env_cfg.py + config/{g1,go1}/env_cfgs.py and tasks/tracking/...) to JSON:

    python tools/make_cfg_golden.py          # needs /root/reference (build container only)

tools/refstub.py stubs the native imports; the config modules themselves are plain dataclasses.  Output:
tests/golden/task_cfg/<task>[-play].json.  tests/test_task_cfg.py feeds them to `mjlab_b200.envcfg.from_manager_cfg` (the path a
user's `ManagerBasedRlEnvCfg` takes) and requires the result to equal this repo's built-in task tables - so every weight,
threshold, noise range, reset / push / command range and curriculum stage of the kernels' configuration is pinned by the
reference's config objects, not by a transcription.
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import refstub  # noqa: E402

refstub.install()

import mjlab.tasks  # noqa: E402,F401  (registers the tasks)
from mjlab.tasks.registry import load_env_cfg  # noqa: E402

from mjlab_b200.envcfg import cfg_to_dict  # noqa: E402

TASKS = ["Mjlab-Velocity-Flat-Unitree-G1", "Mjlab-Velocity-Flat-Unitree-Go1", "Mjlab-Velocity-Rough-Unitree-G1",
         "Mjlab-Velocity-Rough-Unitree-Go1", "Mjlab-Tracking-Flat-Unitree-G1", "Mjlab-Tracking-Flat-Unitree-G1-No-State-Estimation"]

if __name__ == "__main__":
  out = ROOT / "tests" / "golden" / "task_cfg"
  out.mkdir(parents=True, exist_ok=True)
  for task in TASKS:
    for play in (False, True):
      cfg = load_env_cfg(task, play=play)
      d = cfg_to_dict(cfg)
      d["__task__"], d["__play__"] = task, play
      path = out / f"{task}{'-play' if play else ''}.json"
      path.write_text(json.dumps(d, indent=1))   # key order kept: the order of observation terms IS the observation layout
      print(path.name, len(path.read_text()) // 1024, "KB", "rewards:", len(d["rewards"]), "obs:", {g: len(v["terms"]) for g, v in d["observations"].items()})
