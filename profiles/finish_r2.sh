#!/bin/bash
# Turn the artefacts profiles/capture_r2.sh left in gpurun_out/ into the committed summaries profiles/r2_*.txt.
O=gpurun_out; P=profiles
if [ -f $O/r2_ensemble.ncu-rep ]; then
  { echo "# ncu --set full --clock-control none of the headline kernel at the bench size (profiles/capture_r2.sh)";
    echo "# config: tanks=16384 agents=32 chunk=1000"; python $P/summarize.py report $O/r2_ensemble.ncu-rep; } > $P/r2_ncu_ensemble_kernel.txt
fi
for f in $O/r2_pair_*_n*_r*.ncu-rep; do
  [ -f "$f" ] || continue
  b=$(basename $f .ncu-rep); j=$O/$b.json
  tag=$(python -c "import json,sys; print(json.loads(open('$j').read().strip().splitlines()[-1])['tag'])")
  n=$(echo $b | sed 's/.*_n\([0-9]*\)_r.*/\1/'); r=$(echo $b | sed 's/.*_r\([0-9]*\)$/\1/')
  { echo "# ncu --set full --clock-control none of pair_dense_kernel, case $tag (profiles/pair_case.py)";
    echo "# config: n=$n replicates=$r"; python $P/summarize.py report $f pair_dense; } > $P/r2_ncu_pair_dense_$tag.txt
done
[ -f $O/r2_launches_bench.csv ] && python $P/summarize.py launches $O/r2_launches_bench.csv > $P/r2_launches_bench.txt
if [ -f $O/c5_r2f_launches.csv ]; then
  { echo "# C5 (N = 1,048,576, L = 10,240) on one GPU: graph-replayed and eager step time, then the eager launch list"; echo "# config: n_agents=1048576 box=10240"; cat $O/c5_r2f_graph.json $O/c5_r2f_eager.json;
    python $P/summarize.py launches $O/c5_r2f_launches.csv; echo; echo "# ncu --set full of the two kernels of a steady-state step";
    cat $O/c5_r2f_bulk.txt $O/c5_r2f_rows.txt; } > $P/r2_c5_step_breakdown.txt
fi
ls -la $P/r2_*
