#!/bin/bash
# Round-2 evidence capture (run under gpurun, ONE GPU). Every ncu pass follows a plain run of the same command that
# exited 0. Raw artefacts land in gpurun_out/; profiles/finish_r2.sh (run in the build container) turns them into the
# committed summaries profiles/r2_*.txt that bench.py's `recorded_profile` keys read back.
#   usage: bash profiles/capture_r2.sh [tests] [bench] [ens] [pair] [c5]      (no argument = all)
WHAT=${*:-tests bench ens pair c5}
has() { [[ " $WHAT " == *" $1 "* ]]; }
O=gpurun_out
if has tests; then
  timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/r2_pytest_gpu.log; tail -3 $O/r2_pytest_gpu.log
fi
if has bench; then
  timeout 600 python bench.py > $O/r2_bench.json 2> $O/r2_bench.err; echo "bench rc $?"
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_ref.json 2> $O/r2_bench_ref.err; echo "ref rc $?"
  CMD="python bench.py --steps 3 --warmup 3 --no-cpu"
  $CMD > /dev/null 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_bench.csv $CMD > /dev/null 2>&1
fi
if has ens; then   # the headline kernel AT THE BENCH SIZE (16384 tanks x 32 agents x 1000 steps per launch)
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-c5 --pair-n 0"
  $CMD > /dev/null 2> $O/r2_ens_plain.err &&
  ncu --set full --clock-control none --import-source on -k regex:swarm_warp_kernel -s 3 -c 1 -f -o $O/r2_ensemble $CMD > $O/r2_ncu_ens.log 2>&1
fi
if has pair; then
  for c in "65536 1 open" "16384 1 open" "4096 16 open" "4096 16 allnear" "1024 64 open" "65536 1 periodic"; do
    set -- $c
    CMD="python profiles/pair_case.py $1 $2 $3"
    $CMD > $O/r2_pair_$3_n$1_r$2.json 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:pair_dense -s 2 -c 1 -f -o $O/r2_pair_$3_n$1_r$2 $CMD > $O/r2_ncu_pair.log 2>&1
  done
fi
if has c5; then
  bash profiles/c5_probe.sh r2f > /dev/null 2>&1
  bash profiles/c5_ncu.sh r2f > /dev/null 2>&1
fi
ls -la $O/*.ncu-rep 2>/dev/null | tail -20
