"""Batch driver for ensembles of independent swarms (SURVEY f-4): the reference launches one `swarmdyn` process per
replicate from its Python scripts (SwarmDynByPy.py / RunSingle.py); here one call runs all replicates of a parameter
dictionary on the device - initial conditions (sbc_init_state), the step loop (sbc_step up to each output step, the
reference's schedule `s % step_output == 0 and s >= trans_time/dt`, swarmdyn.cpp:67-78) and the per-output-step
observables (sbc_out_swarm / sbc_out_swarm_fishnet) - and returns what the reference writes to /swarm, /swarm_fishNet
and, on request, /part. Replicates can be sharded over GPUs by giving every process its own `replicate_offset`:
a replicate's trajectory depends only on (seed, global replicate id)."""
import numpy as np

from . import params as P
from .swarm import Swarm


def run_ensemble(dic, replicates, seed=0, replicate_offset=0, device=0, neighbour_mode=P.NEIGHBOUR_METRIC,
                 want_particles=False, no_pairs=False):
    """dic: parameter dictionary with the reference's key names (get_base_params / selection_line_parameter).
    Returns a dict: 'steps' (output step indices), 'swarm' [n_out, R, 16], 'swarm_fishNet' [n_out_pred, R, 4] or None,
    'part' [n_out, R, N, 7] and 'alive' [n_out, R, N] when want_particles."""
    sp, npred = P.make_sbc_params(dic, neighbour_mode=neighbour_mode)
    N, R = int(dic['N']), int(replicates)
    dt = float(dic['dt'])
    sim_steps = int(sp.sim_time / dt)
    step_output = max(1, int(float(dic['output']) / dt))
    s_trans, s_pred = int(sp.trans_time / dt), sp.pred_time / dt
    g = Swarm(sp, N, npred, n_replicates=R, seed=seed, replicate_offset=replicate_offset, device=device)
    try:
        g.init_state(int(dic.get('IC', 0)))
        out = {'steps': [], 'swarm': [], 'swarm_fishNet': [], 'part': [], 'alive': []}
        s = 0
        while s < sim_steps:
            s_out = ((max(s, s_trans) + step_output - 1) // step_output) * step_output
            last = s_out >= sim_steps
            if last:
                s_out = sim_steps - 1
            g.step(s, s_out - s + 1)
            s = s_out + 1
            if last:
                break
            out['steps'].append(s_out)
            out['swarm'].append(g.out_swarm(no_pairs=no_pairs))
            if npred > 0 and s_out >= s_pred:
                out['swarm_fishNet'].append(g.out_swarm_fishnet())
            if want_particles:
                part, alive = g.get_particles()
                out['part'].append(part.reshape(R, N, 7))
                out['alive'].append(alive.reshape(R, N))
        res = {'steps': np.asarray(out['steps']), 'swarm': np.asarray(out['swarm']).reshape(-1, R, 16),
               'swarm_fishNet': np.asarray(out['swarm_fishNet']).reshape(-1, R, 4) if out['swarm_fishNet'] else None}
        if want_particles:
            res['part'], res['alive'] = np.asarray(out['part']), np.asarray(out['alive'])
        res['stats'] = g.stats()
        return res
    finally:
        g.close()
