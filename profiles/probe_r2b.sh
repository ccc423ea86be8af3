#!/bin/bash
# whole cell-list cycles per graph launch: parity tests, C5 timeline, C4 sweep (same script as profiles/r2_runs/r2_c4_sweep.jsonl)
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_pair.py tests/test_gpu_step.py tests/test_gpu_domain.py tests/test_gpu_predator.py -m gpu -q -x 2>&1 | tail -3 > $O/p13_tests.txt
python profiles/c5_timeline.py 2>&1 | tail -1 | cut -c1-420 > $O/p13_c5.txt
python profiles/bench_configs.py --set c4sweep --cpu-seconds 0.5 > $O/p13_c4sweep.jsonl 2> $O/p13_c4sweep.err
python profiles/time_c4.py > $O/p13_c4.txt 2>&1
