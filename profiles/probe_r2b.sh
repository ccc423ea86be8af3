#!/bin/bash
# A/B of builds of the rows kernel (register budget): in-situ timeline and C5 step time
O=gpurun_out
for v in "" sde_burst_coast_b200/lib/variants/libsbc_rows6.so sde_burst_coast_b200/lib/variants/libsbc_rows5.so; do
  echo "lib=$v: $(SBC_LIB_PATH=$v python profiles/c5_timeline.py 2>&1 | tail -1)"
  echo "lib=$v: $(SBC_LIB_PATH=$v python profiles/bench_c5.py --steps 400 --warmup 60 2>/dev/null | tail -1 | python -c "import json,sys; print(json.loads(sys.stdin.read())['us_per_step'])")"
done > $O/p7_rows_regs.txt
