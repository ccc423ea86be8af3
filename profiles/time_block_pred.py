import sys, time; sys.path.insert(0,"profiles"); sys.path.insert(0,"tests"); sys.path.insert(0,".")
import numpy as np, torch
from helpers import base_dict
from sde_burst_coast_b200 import params as PP
from sde_burst_coast_b200.swarm import Swarm
from bench_configs import state
def t(name, N, npred, kill, move, R=1, steps=4000, mk=False):
    d = base_dict("natPred", N, BC=0, size=584.0, Npred=npred, pred_move=move, pred_kill=kill, pred_time=0.02, time=12.0, pred_speed0=20)
    sp, npred = PP.make_sbc_params(d)
    g = Swarm(sp, N, npred, n_replicates=R, seed=1)
    st = state(N, sp.sizeL, sp.BC, 7)
    g.set_state(**{k: np.tile(v, (R, 1)) for k, v in st.items()})
    g.force_multikernel(mk)
    g.step(0, 50); g.synchronize()
    t0 = time.perf_counter(); g.step(50, steps); g.synchronize(); dt = time.perf_counter() - t0
    print("%-40s %8.2f us/step  %.3g agent-steps/s" % (name, 1e6 * dt / steps, N * R * steps / dt)); g.close()
t("N=1024 no predator (auto)", 1024, 0, 0, 0)
t("N=1024 pred kill0 chase (auto)", 1024, 1, 0, 1)
t("N=1024 pred kill1 chase (auto)", 1024, 1, 1, 1)
t("N=1024 pred kill3 chase (auto)", 1024, 1, 3, 1)
t("N=1024 pred kill1 walk (auto)", 1024, 1, 1, 0)
t("N=256 pred kill3 chase (auto)", 256, 1, 3, 1)
t("N=30 pred kill3 chase x4096 (auto)", 30, 1, 3, 1, R=4096, steps=2000)
t("N=30 no pred x4096 (warp kernel)", 30, 0, 0, 0, R=4096, steps=2000)
t("N=30 pred kill3 chase x4096 (block, forced)", 30, 1, 3, 1, R=4096, steps=2000, mk=2)
t("N=30 pred kill3 chase x16384 (auto)", 30, 1, 3, 1, R=16384, steps=2000)
t("N=30 pred kill1 walk x16384 (auto)", 30, 1, 1, 0, R=16384, steps=2000)
t("N=30 no pred x16384 (warp kernel)", 30, 0, 0, 0, R=16384, steps=2000)
t("N=1024 pred kill3 chase (multikernel)", 1024, 1, 3, 1, mk=True, steps=2000)
t("N=1024 no predator (block, forced)", 1024, 0, 0, 0, mk=2)
t("N=2048 pred kill3 chase (cluster)", 2048, 1, 3, 1)
t("N=512 pred kill3 chase (cluster)", 512, 1, 3, 1)
t("N=1024 net 50 nodes kill1 (cluster)", 1024, 50, 1, 0)
