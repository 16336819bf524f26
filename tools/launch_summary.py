"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file x.csv ...`) per kernel:

  python tools/launch_summary.py gpurun_out/launches.csv "note" > profiles/rNN_launches.txt
"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ik, iv, im = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
  if r[im] != "gpu__time_duration.sum":
    continue
  v = float(r[iv].replace(",", ""))
  iu = h.index("Metric Unit")
  v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(r[iu], 1e-3)
  tot[r[ik]] += v
  cnt[r[ik]] += 1
allt = sum(tot.values())
print(f"# ncu launch list summary (gpu__time_duration.sum, --clock-control none), {sum(cnt.values())} launches: {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("# Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's live CUDA-event split, not absolutes.")
print(f"{'kernel':60s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda t: -t[1]):
  print(f"{k[:60]:60s} {cnt[k]:8d} {v:12.1f} {v / cnt[k]:10.2f} {100 * v / allt:6.1f}%")
