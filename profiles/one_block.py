"""One swarm of N agents on the one-block-per-swarm / cluster kernel (ncu target):
python profiles/one_block.py [N] [steps] [pred_kill (-1 = no predator)]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'profiles'))
from helpers import base_dict
from sde_burst_coast_b200 import params as PP
from sde_burst_coast_b200.swarm import Swarm
from bench_configs import state
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
kill = int(sys.argv[3]) if len(sys.argv) > 3 else -1
if kill < 0:
    d = base_dict('natPred', N, BC=0, size=584.0, Npred=0)
else:
    d = base_dict('natPred', N, BC=0, size=584.0, Npred=1, pred_move=1, pred_kill=kill, pred_time=0.02, time=12.0, pred_speed0=20)
sp, npred = PP.make_sbc_params(d)
g = Swarm(sp, N, npred, seed=1)
g.set_state(**state(N, sp.sizeL, sp.BC, 7))
g.step(0, 50); g.synchronize()
g.step(50, steps); g.synchronize(); print(g.stats()); g.close()
