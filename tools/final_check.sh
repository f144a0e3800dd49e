#!/bin/bash
# GPU box, one B200: A/B of the opt-in programmatic dependent launch of the one-pass CG kernels (SM_PDL) on the
# mid-size lattices it is meant for, then the whole GPU suite under both settings -- first the one the A/B favours
# (same solution bit for bit and >= 3 % fewer microseconds per iteration at 1024^2), so that a call that runs out of
# time has at least validated that one.  Logs under gpurun_out/.
set -u
LIMIT=${1:-280}          # seconds this script may take in all (the caller's gpurun limit minus a margin)
mkdir -p gpurun_out
timeout 60 python tools/cg_mid.py 1024 "SM_PDL=0;SM_PDL=1;SM_PDL=0,SM_GRAPHS=0;SM_PDL=1,SM_GRAPHS=0;SM_PDL=0;SM_PDL=1" \
    > gpurun_out/r02_pdl_ab_1024.txt 2>&1
echo "ab 1024 rc $?"
timeout 40 python tools/cg_mid.py 2048 "SM_PDL=0;SM_PDL=1" 0.0 > gpurun_out/r02_pdl_ab_2048.txt 2>&1
echo "ab 2048 rc $?"
cat gpurun_out/r02_pdl_ab_1024.txt gpurun_out/r02_pdl_ab_2048.txt
first=$(python - <<'PY'
import json
rows = []
try:
    for line in open("gpurun_out/r02_pdl_ab_1024.txt"):
        if line.startswith("{"):
            rows.append(json.loads(line))
    g = [r for r in rows if "SM_GRAPHS" not in r["env"]]
    off = min(r["us_per_it"] for r in g if r["env"]["SM_PDL"] == "0")
    on = min(r["us_per_it"] for r in g if r["env"]["SM_PDL"] == "1")
    same = all(r["relerr_vs_first"] == 0.0 and r["ok"] == rows[0]["ok"] and r["its"] == rows[0]["its"] for r in rows)
    print(1 if (len(rows) == 6 and same and on <= 0.97 * off) else 0)
except Exception:
    print(0)
PY
)
other=$((1 - first))
echo "first suite: SM_PDL=$first"
SM_PDL=$first timeout 170 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_pdl$first.log 2>&1
echo "suite SM_PDL=$first rc $?"; tail -2 gpurun_out/r02_pytest_gpu_pdl$first.log
left=$((LIMIT - SECONDS - 5))
echo "second suite: SM_PDL=$other, $left s left"
if [ "$left" -lt 60 ]; then echo "second suite skipped (no time left)"; exit 0; fi
SM_PDL=$other timeout $left python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_pdl$other.log 2>&1
echo "suite SM_PDL=$other rc $?"; tail -2 gpurun_out/r02_pytest_gpu_pdl$other.log
