import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, threading
import test_gpu_domain as T
N, L, world = 6000, 400.0, 2
kill = int(sys.argv[1]) if len(sys.argv) > 1 else 3
sp, st = T._pred_setup(N, L, 34, 'natPred', 1, kill, 1, pred_speed0=40.0, kill_rate=400.0, N_confu=500)
slabs = T._pred_slabs(sp, st, N, world, 23, 1)
period = min(sl.local_rebuild_period() for sl in slabs)
for sl in slabs: sl.set_rebuild_period(period)
for sl in slabs: sl.swarm._check(sl.swarm.lib.sbc_domain_post(sl.swarm.h))
for sl in slabs: sl.swarm._check(sl.swarm.lib.sbc_domain_complete(sl.swarm.h))
for s in range(int(os.environ.get("DBG_K", "70"))):
    def run(sl):
        try:
            sl.step_local(s); (sl.swarm.synchronize() if os.environ.get("DBG_SYNC") else None); print('step', s, 'rank', sl.rank, 'ok', "", flush=True)
        except Exception as e:
            print('step', s, 'rank', sl.rank, 'ERR', e, flush=True)
    th = [threading.Thread(target=run, args=(sl,)) for sl in slabs]
    [t.start() for t in th]; [t.join() for t in th]
    for sl in slabs: sl.swarm._check(sl.swarm.lib.sbc_domain_post(sl.swarm.h))
    for sl in slabs: sl.swarm._check(sl.swarm.lib.sbc_domain_complete(sl.swarm.h))
