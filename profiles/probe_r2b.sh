#!/bin/bash
# tuning probe of the C5 step (one GPU): cell-list parity tests, then the in-situ timeline with the rows kernel's phase marks
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_pair.py tests/test_gpu_step.py tests/test_gpu_properties.py tests/test_gpu_domain.py -m gpu -q -x 2>&1 | tail -3 > $O/p5_tests.txt
python profiles/c5_timeline.py > $O/p5_timeline.txt 2>&1
python profiles/bench_c5.py --steps 400 --warmup 60 2>/dev/null | tail -1 | cut -c1-420 > $O/p5_bench_c5.txt
