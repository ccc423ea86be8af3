"""Run one multi-kernel configuration for a few steps (ncu launch lists): python profiles/one_config.py c3|c4 [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'profiles'))
from helpers import base_dict
from sde_burst_coast_b200 import params as PP
from sde_burst_coast_b200.swarm import Swarm
from bench_configs import state
which = sys.argv[1]; steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
if which == 'c3':
    d = base_dict('natPred', 1024, BC=0, size=584.0, Npred=1, pred_move=1, pred_kill=3, pred_time=0.02, time=12.0, pred_speed0=20)
else:
    d = base_dict('mulFisher', 16384, BC=0, size=1280.0, Npred=640, pred_kill=1, pred_time=0.02, time=12.0)
sp, npred = PP.make_sbc_params(d)
N = int(d['N'])
g = Swarm(sp, N, npred, seed=1)
g.set_state(**state(N, sp.sizeL, sp.BC, 7))
g.step(0, 40 + steps); g.synchronize(); print(g.stats()); g.close()
