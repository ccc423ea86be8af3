#!/bin/bash
# Strong scaling of BASELINE config C2: EXACTLY 16,384 tanks x N=32 sharded over 1/2/4/8 GPUs (bench.py's default run is
# weak: 16,384 tanks PER GPU). One JSON line per GPU count into gpurun_out/r2_strong_n<N>.json. usage: bash profiles/strong_c2.sh
NG=$(python -c "import torch; print(torch.cuda.device_count())")
for n in 1 2 4 8; do
  [ $n -le $NG ] || continue
  t=$((16384 / n))
  ARGS="bench.py --gpus $n --tanks $t --steps 10 --warmup 3 --no-cpu --no-c5 --no-domain --pair-n 0"
  if [ $n -eq 1 ]; then python $ARGS > gpurun_out/r2_strong_n$n.json 2> gpurun_out/r2_strong_n$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800 + n)) $ARGS > gpurun_out/r2_strong_n$n.json 2> gpurun_out/r2_strong_n$n.err; fi
  python - <<P
import json
for l in open('gpurun_out/r2_strong_n$n.json'):
    if l.startswith('{'):
        d = json.loads(l); print('strong n=$n tanks/gpu=$t value=%.4g e2e=%.4g ms/step=%.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
P
done
