#!/bin/bash
# Attempt to install the reference's physics stack (MuJoCo C, Warp, mujoco_warp) on the GPU box from the default
# package index and the offline wheelhouse, and log the outcome. The box has no network; this
# records that fact. NOT attempted: the pinned mujoco_warp git revision (uv.lock:1378-1380) - installing from that
# GitHub repository runs its build code and needs explicit human authorisation, which this run does not have.
# Output: gpurun_out/install_attempt.log
out=gpurun_out/install_attempt.log
mkdir -p gpurun_out
{
  echo "== date: $(date -u)"; echo "== host: $(hostname)"; nvidia-smi -L
  echo "== python -c 'import mujoco / warp / mujoco_warp / rsl_rl'"
  for m in mujoco warp mujoco_warp rsl_rl; do python -c "import $m; print('$m', getattr($m,'__version__','?'))" 2>&1 | tail -1; done
  echo "== find wheels / trees on the box"
  find / \( -iname '*mujoco*' -o -iname 'warp_lang*' -o -iname 'warp-lang*' \) -not -path '/proc/*' -not -path "$PWD/*" -not -path '/root/repo/*' 2>/dev/null | head -20
  echo "== pip install from the default index (expected: no network)"
  timeout 120 python -m pip install --disable-pip-version-check --retries 0 --timeout 10 mujoco mujoco-warp warp-lang 2>&1 | tail -15
  echo "== pip install from the offline wheelhouse"
  timeout 120 python -m pip install --disable-pip-version-check --no-index --find-links /opt/wheelhouse mujoco mujoco-warp warp-lang 2>&1 | tail -8
  echo "== pinned versions from the default index (uv.lock:1334-1335,3362-3363)"
  timeout 60 python -m pip install --disable-pip-version-check --retries 0 --timeout 10 'mujoco==3.4.1.dev843752024' 'warp-lang==1.11.0.dev20251211' 2>&1 | tail -6
  echo "== pinned mujoco_warp git revision: NOT ATTEMPTED (needs human authorisation to run an external repo's build code)"
  echo "== connectivity"
  timeout 10 python - <<'PY'
import socket
for h in ("pypi.org",):
    try:
        print(h, socket.gethostbyname(h))
    except Exception as e:
        print(h, "DNS failure:", e)
PY
  echo "== done"
} > $out 2>&1
cat $out
