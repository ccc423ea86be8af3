"""Mid-size single swarms on the cooperative grid kernel vs the multi-kernel path (GPU only):
python profiles/time_grid.py"""
import sys, time; sys.path.insert(0, "profiles"); sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
from helpers import base_dict
from sde_burst_coast_b200 import params as PP
from sde_burst_coast_b200.swarm import Swarm
from bench_configs import state

def t(name, d, steps=2000):
    sp, npred = PP.make_sbc_params(d)
    N = int(d['N'])
    for hook, tag in ((0, 'auto'), (1, 'multi-kernel')):
        ts = torch.cuda.Stream()
        g = Swarm(sp, N, npred, seed=1, stream=ts.cuda_stream)
        g.force_multikernel(hook)
        g.set_state(**state(N, sp.sizeL, sp.BC, 7))
        g.step(0, 50); g.synchronize()
        t0 = time.perf_counter(); g.step(50, steps); g.synchronize(); dt = time.perf_counter() - t0
        st = g.stats()
        print('%-36s [%-12s] %8.2f us/step  block launches %d' % (name, tag, 1e6 * dt / steps, st['block_launches'])); g.close()

for N in (3072, 4096, 8192):
    t("no predator N=%d" % N, base_dict("natPred", N, BC=0, size=584.0 * (N / 1024) ** 0.5, Npred=0))
t("mulFisher N=4096 L=640", base_dict("mulFisher", 4096, BC=0, size=640.0, Npred=320, pred_kill=1, pred_time=0.02, time=12.0))
t("C4 mulFisher N=16384 L=1280", base_dict("mulFisher", 16384, BC=0, size=1280.0, Npred=640, pred_kill=1, pred_time=0.02, time=12.0))
