#!/bin/bash
# GPU box: ncu capture of the two kernels of a CG iteration at 1024^2 + exported pages (see tools/ncu_cg_1024.py)
set -u
mkdir -p gpurun_out
python tools/ncu_cg_1024.py > gpurun_out/r02_cg_1024_plain.log 2>&1 &&
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"k_dd_tma|k_cg_resid" -s 12 -c 4 \
    -o gpurun_out/r02_cg_1024 -f python tools/ncu_cg_1024.py > gpurun_out/r02_cg_1024_ncu.log 2>&1
echo "ncu rc $?"
if [ -f gpurun_out/r02_cg_1024.ncu-rep ]; then
  ncu -i gpurun_out/r02_cg_1024.ncu-rep --page raw --csv > gpurun_out/r02_cg_1024_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02_cg_1024.ncu-rep --page details > gpurun_out/r02_cg_1024_details.txt 2>/dev/null
fi
tail -3 gpurun_out/r02_cg_1024_plain.log gpurun_out/r02_cg_1024_ncu.log
