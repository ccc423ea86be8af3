#!/bin/bash
# ncu --set full of the two kernels of a steady-state C5 step (eager launches so that ncu sees kernel nodes one by one).
# usage: bash profiles/c5_ncu.sh [tag]   -> gpurun_out/c5_<tag>_{bulk,rows}.ncu-rep + text summaries
TAG=${1:-r2}
export SBC_NO_GRAPHS=1
CMD="python profiles/bench_c5.py --steps 30 --warmup 30"
$CMD > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:update_bulk_kernel -s 40 -c 1 -f -o gpurun_out/c5_${TAG}_bulk $CMD > gpurun_out/c5_${TAG}_ncu_bulk.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_cells_lanes_kernel -s 40 -c 1 -f -o gpurun_out/c5_${TAG}_rows $CMD > gpurun_out/c5_${TAG}_ncu_rows.log 2>&1
for k in bulk rows; do
  python profiles/summarize.py report gpurun_out/c5_${TAG}_${k}.ncu-rep > gpurun_out/c5_${TAG}_${k}.txt
  ncu -i gpurun_out/c5_${TAG}_${k}.ncu-rep --page source --csv > gpurun_out/c5_${TAG}_${k}_source.csv 2>/dev/null
done
cat gpurun_out/c5_${TAG}_bulk.txt gpurun_out/c5_${TAG}_rows.txt
