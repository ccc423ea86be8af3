#!/bin/bash
# final-state check on two GPUs: multi-GPU parity tests, timeline of the decomposed swarm, timeline of the undecomposed one
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_domain.py -m gpu -q -x 2>&1 | tail -3 > $O/p8_tests.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/domain_timeline.py > $O/p8_domain_timeline.txt 2>/dev/null
python profiles/c5_timeline.py > $O/p8_c5_timeline.txt 2>&1
