#!/bin/bash
# Round-1 profile capture (run under gpurun): launch list of the default bench command, ncu --set full of
# the ensemble kernel and of the dense all-pairs kernel. Summaries are produced here by profiles/summarize.py.
set -x
CMD="python bench.py --steps 3 --warmup 3 --no-cpu"
$CMD > gpurun_out/plain_r1.json 2> gpurun_out/plain_r1.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $CMD > /dev/null 2>&1
CMD2="python bench.py --steps 3 --warmup 3 --no-cpu --tanks 4096 --pair-n 16384"
$CMD2 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:swarm_warp_kernel -s 3 -c 1 -o gpurun_out/prof_r1_ensemble $CMD2 > gpurun_out/ncu_r1_ens.log 2>&1
CMD3="python profiles/pair_bench.py 65536 2 open"
$CMD3 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pair_dense_kernel -s 1 -c 1 -o gpurun_out/prof_r1_pair $CMD3 > gpurun_out/ncu_r1_pair.log 2>&1
ls -la gpurun_out/*.ncu-rep
