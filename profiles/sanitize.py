"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck):
   compute-sanitizer --tool memcheck python profiles/sanitize.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from helpers import make_params, random_state, base_dict
from sde_burst_coast_b200 import params as PP
from sde_burst_coast_b200.swarm import Swarm
rng = np.random.default_rng(0)
# ensemble kernels (warp: 8/16/32 lanes per tank; block) + observables + state IO
for N, R in ((5, 7), (13, 5), (30, 4), (100, 3), (1000, 1)):
    sp, _, _ = make_params('burst_coast', N)
    g = Swarm(sp, N, n_replicates=R, seed=1)
    sts = [random_state(N, sp, rng) for _ in range(R)]
    g.set_state(**{k: np.stack([s[k] for s in sts]) for k in sts[0]})
    g.step(0, 150); g.observables(); g.get_particles(); g.get_state(); g.close()
# dense pass (sorted / unsorted, periodic / open), burst-gated rows, nn lists
for mode, over, N in (('natPred', dict(BC=0, size=100.0, Npred=0, alg_range=10.0), 700), ('burst_coast', dict(BC=-1), 300),
                      ('natPred', dict(BC=0, size=45.0, Npred=0), 200)):
    sp, _, _ = make_params(mode, N, **over)
    g = Swarm(sp, N)
    g.set_state(**random_state(N, sp, rng, spread=60.0 if sp.BC == -1 else None))
    g.pair_interactions(1); g.pair_interactions(0); g.close()
# multi-kernel step with predators, graphs, cell list
sp, npred, _ = make_params('mulFisher', 300, BC=0, size=100.0, Npred=50, pred_time=0.01, time=5.0)
g = Swarm(sp, 300, npred, seed=3); g.force_multikernel(True)
g.set_state(**random_state(300, sp, rng, spread=40.0)); g.step(0, 60); g.get_predators(); g.counts(); g.close()
d = base_dict('natPred', 9000, BC=0, size=300.0, Npred=0)
sp, _ = PP.make_sbc_params(d, path=PP.PATH_CELLLIST)
g = Swarm(sp, 9000, seed=4); g.set_state(**random_state(9000, sp, rng)); g.step(0, 12); g.close()
# domain decomposition, loop-back
import torch
from sde_burst_coast_b200.domain import SlabSwarm, step_loopback, gather_global
d = base_dict('natPred', 4000, BC=0, size=400.0, Npred=0)
sp, _ = PP.make_sbc_params(d, path=PP.PATH_CELLLIST)
st = random_state(4000, sp, rng)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
slabs = [SlabSwarm(sp, 4000, r, 3, seed=5, stream=ts.cuda_stream) for r in range(3)]
for s in slabs: s.scatter_initial(st)
step_loopback(slabs, 0, 10); gather_global(slabs, 4000)
for s in slabs: s.close()
print('sanitize run complete')
