"""GPU-only timing of the multi-kernel configs (C4 mulFisher N = 16,384; natPred N = 4,096 with one predator):
python profiles/time_c4.py"""
import sys, time; sys.path.insert(0, "profiles"); sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
from helpers import base_dict
from sde_burst_coast_b200 import params as PP
from sde_burst_coast_b200.swarm import Swarm
from bench_configs import state

def t(name, d, steps=2000):
    sp, npred = PP.make_sbc_params(d)
    N = int(d['N'])
    ts = torch.cuda.Stream()
    g = Swarm(sp, N, npred, seed=1, stream=ts.cuda_stream)
    g.set_state(**state(N, sp.sizeL, sp.BC, 7))
    g.step(0, 50); g.synchronize()
    s0 = g.stats()
    t0 = time.perf_counter(); g.step(50, steps); g.synchronize(); dt = time.perf_counter() - t0
    s1 = g.stats()
    n = sum(s1[k] - s0[k] for k in ('pair_launches', 'update_launches', 'block_launches', 'other_launches')) / steps
    print('%-40s %8.2f us/step  %.2f kernels/step  alive %d' % (name, 1e6 * dt / steps, n, int(g.counts()[0][0]))); g.close()

t("C4 mulFisher N=16384 L=1280", base_dict("mulFisher", 16384, BC=0, size=1280.0, Npred=640, pred_kill=1, pred_time=0.02, time=12.0))
t("C4 mulFisher N=16384 L=100", base_dict("mulFisher", 16384, BC=0, size=100.0, Npred=50, pred_kill=1, pred_time=0.02, time=12.0))
t("natPred N=4096, one predator", base_dict("natPred", 4096, BC=0, size=1168.0, Npred=1, pred_move=1, pred_kill=3, pred_time=0.02, time=12.0, pred_speed0=20))
t("no predator N=8192", base_dict("natPred", 8192, BC=0, size=1652.0, Npred=0))
