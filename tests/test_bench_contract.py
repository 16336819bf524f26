"""bench.py's reference (CPU) arm runs without a GPU: check the JSON contract of the line it prints."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_one_contract_line():
  out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
  assert out.returncode == 0, out.stderr[-2000:]
  lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
  assert len(lines) == 1
  j = json.loads(lines[0])
  assert j["impl"] == "reference" and j["metric"] == "env-steps/s" and j["unit"] == "env-steps/s"
  assert j["n_gpus"] == 1 and j["steps"] == 1 and j["warmup"] == 1 and j["higher_is_better"] is True
  assert j["scaling"] == "weak" and j["vs_baseline"] is None and j["data"] == "synthetic" and j["dtype"] == "f64"
  assert j["value"] > 0 and j["ms_per_step"] > 0 and "workload" in j["config"] and "model" not in j["config"]
  cb = j["cpu_baseline"]
  assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
  e = j["e2e"]
  assert e["value"] == j["value"] and e["unit"] == j["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


import pytest  # noqa: E402


@pytest.mark.gpu
def test_own_arm_prints_the_full_contract_line():
  out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--gpus", "1", "--steps", "6", "--warmup", "3"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
  assert out.returncode == 0, out.stderr[-2000:]
  lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
  assert len(lines) == 1
  j = json.loads(lines[0])
  for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
            "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
    assert k in j, k
  assert j["metric"] == "env-steps/s" and j["steps"] == 6 and j["warmup"] == 3 and j["dtype"] == "f32" and j["vs_baseline"] is None
  assert j["value"] > 1e5 and abs(j["value"] - 4096 / (j["ms_per_step"] * 1e-3)) < 1e-3 * j["value"]
  assert j["gpu_launches"] == 6 * 29                      # action, 4 x (smooth, order, solve, post, airtime), post + gate, forward (4), obs, log
  e = j["e2e"]
  assert e["value"] > 1e5 and e["h2d_bytes_per_step"] == 4096 * 29 * 4 and e["d2h_bytes_per_step"] > 0
  r = j["roofline"]
  assert r["bound"] == "hbm" and r["kernel"] == "k_solve" and 0 < r["frac"] < 1 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
  c = j["cpu_baseline"]
  assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
  assert set(j["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
  assert "workload" in j["config"] and "model" not in j["config"]


@pytest.mark.gpu
@pytest.mark.parametrize("flags,metric_min", [
  (["--mode", "physics", "--model", "go1_flat", "--steps", "4", "--warmup", "3", "--no-roofline-pass", "--prewarm-s", "0.2"], 1e5),
  (["--task", "train", "--steps", "1", "--warmup", "3", "--prewarm-s", "0.2"], 1e4),
  (["--task", "tracking", "--steps", "4", "--warmup", "3", "--no-roofline-pass", "--prewarm-s", "0.2"], 1e5),
])
def test_other_bench_modes_print_one_line(flags, metric_min):
  """BASELINE configs 2 (physics-only stepping), 3 (PPO train iteration) and 5 (tracking) stay driver-runnable."""
  out = subprocess.run([sys.executable, str(ROOT / "bench.py"), *flags], capture_output=True, text=True, timeout=900, cwd=ROOT)
  assert out.returncode == 0, out.stderr[-2000:]
  lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
  assert len(lines) == 1
  j = json.loads(lines[0])
  assert j["value"] > metric_min and j["n_gpus"] == 1 and j["gpu_launches"] > 0 and "workload" in j["config"]
  assert j["e2e"]["value"] > 0 and set(j["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
