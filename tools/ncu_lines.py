"""Attribute an ncu SASS-level source page to CUDA source lines.

  ncu -i prof.ncu-rep --page source --csv > sass.csv
  python tools/ncu_lines.py sass.csv mjlab_b200/libb200sim.so k_solve [top]

Joins the per-instruction rows of the ncu export with `nvdisasm -g` line info of
the same cubin (instruction order is identical) and prints the hottest source
lines by stall samples and by executed instructions.
"""
import csv
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path


# template instantiations known under a plain name (b2_api.cu: k_solve = k_solve_t<false>, k_solve_cg = k_solve_t<true>)
MANGLED = {"k_solve": "_Z9k_solve_tILb0EE", "k_solve_cg": "_Z9k_solve_tILb1EE"}


def line_table(so: str, kernel: str):
  with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", str(Path(so).resolve())], cwd=td, check=True, capture_output=True)
    cubin = next(Path(td).glob("*.cubin"))
    txt = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True, check=True).stdout
  out, cur, inside = [], ("?", 0), False
  for ln in txt.splitlines():
    if ln.startswith("\t.section"):
      inside = (".text." in ln) and (MANGLED.get(kernel, f"_Z{len(kernel)}{kernel}") in ln)   # exact (Itanium-mangled) function name
      continue
    if not inside:
      continue
    mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if mm:
      cur = (Path(mm.group(1)).name, int(mm.group(2)))
      continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
      out.append(cur)
  return out


def region_of(key):
  """Map (file, line) to a named region using `// @region name` markers in the source file."""
  f, l = key
  marks = _marks(f)
  name = f
  for ml, mn in marks:
    if ml <= l:
      name = f"{f}:{mn}"
    else:
      break
  return name


_MARKS = {}


def _marks(f):
  if f not in _MARKS:
    cands = list(Path(".").rglob(f))
    out = []
    if cands:
      for i, ln in enumerate(cands[0].read_text().splitlines(), 1):
        mm = re.search(r"//\s*@region\s+(\S+)", ln)
        if mm:
          out.append((i, mm.group(1)))
    _MARKS[f] = out
  return _MARKS[f]


def main():
  sass_csv, so, kernel = sys.argv[1:4]
  top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
  rows = list(csv.reader(open(sass_csv)))
  hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
  hdr = rows[hdr_i]
  body = rows[hdr_i + 1:]
  lines = line_table(so, kernel)
  if len(lines) != len(body):
    print(f"warning: {len(lines)} disassembled instructions vs {len(body)} profiled rows", file=sys.stderr)
  ci = {n: hdr.index(n) for n in ("# Samples", "Instructions Executed", "Source")}
  reason = sys.argv[5] if len(sys.argv) > 5 else None  # e.g. stall_no_inst: rank by that stall reason
  if reason:
    ci["# Samples"] = hdr.index(reason)
  samples, insts = defaultdict(int), defaultdict(int)
  for (f, l), r in zip(lines, body):
    samples[(f, l)] += int(r[ci["# Samples"]] or 0)
    insts[(f, l)] += int(r[ci["Instructions Executed"]] or 0)
  ts, ti = sum(samples.values()), sum(insts.values())
  src_cache = {}

  def text(f, l):
    if f not in src_cache:
      cands = list(Path(".").rglob(f))
      src_cache[f] = cands[0].read_text().splitlines() if cands else []
    s = src_cache[f]
    return s[l - 1].strip()[:110] if 0 < l <= len(s) else ""

  print(f"total samples {ts}, warp instructions {ti}")
  print("--- by stall samples")
  for k, v in sorted(samples.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{100*v/ts:5.1f}% smp {100*insts[k]/ti:5.1f}% ins  {k[0]}:{k[1]:<4d} {text(*k)}")
  print("--- by executed warp instructions")
  for k, v in sorted(insts.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{100*samples[k]/ts:5.1f}% smp {100*v/ti:5.1f}% ins  {k[0]}:{k[1]:<4d} {text(*k)}")
  # coarse regions: consecutive source-line ranges of the solver file (see tools/ncu_lines.py REGIONS)
  regions = defaultdict(lambda: [0, 0])
  for k in insts:
    name = region_of(k)
    regions[name][0] += samples[k]
    regions[name][1] += insts[k]
  print("--- by region")
  for name, (sv, iv) in sorted(regions.items(), key=lambda kv: -kv[1][1]):
    print(f"{100*sv/ts:5.1f}% smp {100*iv/ti:5.1f}% ins  {name}")


if __name__ == "__main__":
  main()
