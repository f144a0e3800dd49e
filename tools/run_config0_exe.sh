#!/bin/bash
# BASELINE configs[0] end to end through the two EXECUTABLES with the same stdin:
#   64x64, beta=2, m0=0, 10 MD steps, trajectory length 1, ranks 1x1, 80 thermalisation + 20 measurements (100 trajectories)
#   reference: oracle/_ref/SM_64x64 (the unmodified src/main.cpp over the mini-MPI shim, 1 core)
#   this repo: host/bin/SM_64x64 (same prompts, trajectories on the GPU)
set -u
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
PARAMS="1\n1\n0\n10\n1\n2\n80\n20\n0\n0\n"
export HOSTNAME="${HOSTNAME:-gpubox}"
for who in b200 reference; do
  d=$(mktemp -d); cd "$d"
  if [ $who = b200 ]; then exe="$ROOT/host/bin/SM_64x64"; else exe="$ROOT/oracle/_ref/SM_64x64"; fi
  [ -x "$exe" ] || { echo "{\"impl\": \"$who\", \"unavailable\": \"$exe missing\"}"; continue; }
  t0=$(date +%s.%N)
  printf "$PARAMS" | "$exe" > out.txt 2> err.txt
  rc=$?
  t1=$(date +%s.%N)
  ep=$(grep "Average plaquette" out.txt | sed 's/.*: Ep = \([^ ]*\) dEp = \(.*\)/\1/')
  dep=$(grep "Average plaquette" out.txt | sed 's/.*: Ep = \([^ ]*\) dEp = \(.*\)/\2/')
  acc=$(grep "Acceptance rate" out.txt | sed 's/.*: //')
  ex=$(grep "Execution time" out.txt | sed 's/.*= \([^ ]*\) s/\1/')
  echo "{\"impl\": \"$who\", \"rc\": $rc, \"wall_s\": $(python3 -c "print(round($t1 - $t0, 3))"), \"execution_time_s\": ${ex:-null}, \"Ep\": ${ep:-null}, \"dEp\": ${dep:-null}, \"acceptance\": ${acc:-null}, \"trajectories\": 100, \"simdata_lines\": $(cat 2D_U1_64x64_m00_SimData.txt 2>/dev/null | wc -l)}"
  cd /; rm -rf "$d"
done
