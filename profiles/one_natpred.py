"""The reference's default run (RunSingle.py: natPred, N = 30, one chasing predator, kill rule 3, periodic box
L = 100) as an ensemble of R swarms on the in-warp predator kernel (ncu target):
python profiles/one_natpred.py [R] [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'profiles'))
import numpy as np
from helpers import base_dict
from sde_burst_coast_b200 import params as PP
from sde_burst_coast_b200.swarm import Swarm
from bench_configs import state
R = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
N = 30
d = base_dict('natPred', N, BC=0, size=100.0, Npred=1, pred_move=1, pred_kill=3, pred_time=0.02, time=12.0)
sp, npred = PP.make_sbc_params(d)
g = Swarm(sp, N, npred, n_replicates=R, seed=1)
st = state(N, sp.sizeL, sp.BC, 7)
g.set_state(**{k: np.tile(v, (R, 1)) for k, v in st.items()})
g.step(0, 50); g.synchronize()
t0 = time.perf_counter(); g.step(50, steps); g.synchronize(); dt = time.perf_counter() - t0
print('natPred N=30 x%d: %.2f us/step, %.3g agent-steps/s, dead %d' % (R, 1e6 * dt / steps, N * R * steps / dt, int(g.counts()[1].sum())))
g.close()
