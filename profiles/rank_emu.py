import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from helpers import make_params
from bench import tank_ensemble_state, SEED
from sde_burst_coast_b200.swarm import Swarm
sp, _, _ = make_params('burst_coast', 32)
R = 16384
for rank in (1, 2, 3):
    g = Swarm(sp, 32, n_replicates=R, seed=SEED, replicate_offset=rank * R)
    g.set_state(**tank_ensemble_state(R, 32, sp, SEED, first_replicate=rank * R))
    g.step(0, 12000); g.synchronize(); print('rank-emu', rank, 'ok', g.stats()['agent_steps']); g.close()
