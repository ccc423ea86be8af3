#!/usr/bin/env python
"""One case of bench.py's pair-kernel leg on its own (same seeded state), for the ncu captures of profiles/capture_r2.sh.
usage: python profiles/pair_case.py N REPLICATES open|periodic|sparse|allnear"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench
from sde_burst_coast_b200.swarm import Swarm

n, reps, kind = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
want = {'open': 'open, dense', 'periodic': 'periodic, dense', 'sparse': 'periodic, sparse (C4 density)', 'allnear': 'open, all near'}[kind]
case = [c for c in bench.pair_cases(n) if c[0] == want and c[1] == n and c[2] == reps][0]
sp, st, L, over = bench.pair_case_state(*case)
g = Swarm(sp, n, n_replicates=reps)
g.set_state(**st)
ms = g.bench_pair_pass(1, iters=4)
g.close()
print(json.dumps({'tag': bench.pair_case_tag(case[0], n, reps, case[5], case[6]), 'n': n, 'replicates': reps, 'box': L, 'ms_per_launch': ms}))
