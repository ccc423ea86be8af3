#!/bin/bash
# C5 on one GPU: graph-replayed vs eager step time, then the ncu launch list of eager steps (per-kernel durations,
# cold-cache and serialised: compare shares) -> gpurun_out/c5_*.  usage: bash profiles/c5_probe.sh [tag]
TAG=${1:-r2}
python profiles/bench_c5.py --steps 400 --warmup 60 > gpurun_out/c5_${TAG}_graph.json 2> gpurun_out/c5_${TAG}_graph.err
SBC_NO_GRAPHS=1 python profiles/bench_c5.py --steps 400 --warmup 60 > gpurun_out/c5_${TAG}_eager.json 2>> gpurun_out/c5_${TAG}_graph.err
cat gpurun_out/c5_${TAG}_graph.json gpurun_out/c5_${TAG}_eager.json
SBC_NO_GRAPHS=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c5_${TAG}_launches.csv \
    python profiles/bench_c5.py --steps 60 --warmup 30 > /dev/null 2> gpurun_out/c5_${TAG}_ncu.err
python profiles/summarize.py launches gpurun_out/c5_${TAG}_launches.csv | head -20
