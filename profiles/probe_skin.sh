#!/bin/bash
# skin sweep of the decomposed swarm (2 GPUs) and of the undecomposed one (GPU 0) -> gpurun_out/p4_*
O=gpurun_out
for f in 0.2 0.3 0.4 0.5; do
  echo "skin $f 2gpu: $(SBC_SKIN_FRAC=$f python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/domain_timeline.py 2>/dev/null | grep '"rank": 0' | cut -c1-330)"
  echo "skin $f 1gpu: $(SBC_SKIN_FRAC=$f python profiles/c5_timeline.py 2>/dev/null | tail -1 | cut -c1-200)"
done > $O/p4_skin.txt
