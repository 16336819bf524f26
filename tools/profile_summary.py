"""Condense an ncu report (read here, without a GPU) into a text summary for profiles/.

  python tools/profile_summary.py gpurun_out/prof.ncu-rep k_solve profiles/r01_k_solve.txt "note"
"""
import csv
import io
import subprocess
import sys

KEYS = [
  "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
  "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem",
  "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
  "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
  "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio",
  "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
  "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
  "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main():
  rep, kernel, out, note = sys.argv[1], sys.argv[2], sys.argv[3], (sys.argv[4] if len(sys.argv) > 4 else "")
  raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(raw)))
  h, units, v = rows[0], rows[1], rows[2]
  lines = [f"# ncu summary: {kernel}  ({rep.split('/')[-1]})", f"# {note}", ""]
  for i, n in enumerate(h):
    if n in KEYS:
      lines.append(f"{n:70s} {v[i]:>16s} {units[i]}")
  lines.append("")
  lines.append("stall reasons (warps stalled per issue-active cycle):")
  for i, n in enumerate(h):
    if "issue_stalled" in n and n.endswith("per_issue_active.ratio"):
      try:
        x = float(v[i])
      except ValueError:
        continue
      if x >= 0.05:
        lines.append(f"  {n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {x:8.3f}")
  src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
  tmp = out + ".sass.csv"
  open(tmp, "w").write(src)
  try:
    top = subprocess.run([sys.executable, "tools/ncu_lines.py", tmp, "mjlab_b200/libb200sim.so", kernel, "25"],
                         capture_output=True, text=True).stdout
    lines += ["", "hottest CUDA source lines (stall samples / executed warp instructions; line numbers are those of the",
              "build that was profiled, attribution is only exact while libb200sim.so is unchanged):", top]
  finally:
    import os
    os.remove(tmp)
  open(out, "w").write("\n".join(lines) + "\n")
  print("\n".join(lines[:40]))


if __name__ == "__main__":
  main()
