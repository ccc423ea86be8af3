#!/usr/bin/env python
"""Drive only the dense all-pairs pass (for ncu captures and quick timing).
usage: python profiles/pair_bench.py [N] [iters] [open|periodic|both] [resort]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from helpers import make_params, f32
from sde_burst_coast_b200.swarm import Swarm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
which = sys.argv[3] if len(sys.argv) > 3 else 'both'
resort = 2 if (len(sys.argv) > 4 and sys.argv[4] == 'resort') else 0
L = 100.0 * (n / 16384.0) ** 0.5     # keep the density of the N=16384 / L=100 probe state
out = {}
for name, mode, over, flop in (('open', 'burst_coast', dict(BC=-1), 13),
                               ('periodic', 'natPred', dict(BC=0, size=L, Npred=0), 21)):
    if which not in ('both', name):
        continue
    sp, _, _ = make_params(mode, n, **over)
    rng = np.random.default_rng(20261019)
    st = dict(x=f32(L * rng.random(n)), y=f32(L * rng.random(n)), phi=f32(2 * np.pi * rng.random(n)), vproj=np.ones(n))
    g = Swarm(sp, n)
    g.set_state(**st)
    ms = g.bench_pair_pass(1 | resort, iters=iters)
    pairs = n * (n - 1)
    out[name] = dict(n=n, L=L, ms=ms, pairs_per_s=pairs / (ms * 1e-3), tflops=pairs * flop / (ms * 1e-3) / 1e12,
                     frac_of_74_45=pairs * flop / (ms * 1e-3) / 74.45e12, amb=g.stats()['ambiguous_pairs'])
    g.close()
print(json.dumps(out))
