#!/bin/bash
# Tuning probe of the C5 step (one GPU) as used in round 2: cell-list parity tests, the in-situ timeline with the rows
# kernel's phase marks, the C4 sweep. A/B runs of two builds: SBC_LIB_PATH=<other libsbc.so>; knobs: SBC_MAX_BATCH (steps
# per graph launch), SBC_ROWS_BLOCKS_PER_SM, SBC_SKIN_FRAC; compile-time: -DSBC_ROWS_NB=, -DSBC_ROWS_MINB= (celllist.cu).
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_pair.py tests/test_gpu_step.py tests/test_gpu_domain.py tests/test_gpu_predator.py -m gpu -q -x 2>&1 | tail -3 > $O/probe_tests.txt
python profiles/c5_timeline.py 2>&1 | tail -1 > $O/probe_c5_timeline.json
python profiles/bench_configs.py --set c4sweep --cpu-seconds 0.5 > $O/probe_c4sweep.jsonl 2> $O/probe_c4sweep.err
python profiles/time_c4.py > $O/probe_c4.txt 2>&1
