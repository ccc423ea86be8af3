#!/usr/bin/env python
"""BASELINE.json configs C1, C3, C4 on one GPU next to the CPU oracle:
agent-steps/s and directed pair evaluations/s. (C2 is bench.py, C5 is profiles/bench_c5.py.)
usage: python profiles/bench_configs.py [--cpu-seconds 8]   -> one JSON line per config"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from helpers import base_dict, f32
from sde_burst_coast_b200 import params as PP


def state(N, L, BC, seed):
    rng = np.random.default_rng(seed)
    if BC == 0:
        x, y = L * rng.random(N), L * rng.random(N)
    else:
        r, th = 0.8 * L * np.sqrt(rng.random(N)), 2 * np.pi * rng.random(N)
        x, y = r * np.cos(th), r * np.sin(th)
    return dict(x=f32(x), y=f32(y), phi=f32(2 * np.pi * rng.random(N) - np.pi), vproj=np.ones(N))


def run(name, d, npred_expected, steps, cpu_steps, neighbour=PP.NEIGHBOUR_METRIC, voronoi_cpu=False, R=1):
    import torch
    from sde_burst_coast_b200.swarm import Swarm
    from oracle import pyorc
    sp, npred = PP.make_sbc_params(d, neighbour_mode=neighbour)
    N = int(d['N'])
    st = state(N, sp.sizeL, sp.BC, 7)
    ts = torch.cuda.Stream()
    g = Swarm(sp, N, npred, n_replicates=R, seed=1, stream=ts.cuda_stream)
    g.set_state(**({k: np.tile(v, (R, 1)) for k, v in st.items()} if R > 1 else st))
    with torch.cuda.stream(ts):
        g.step(0, 50)
        torch.cuda.synchronize()
        s0 = g.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ts)
        g.step(50, steps)
        e1.record(ts)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    s1 = g.stats()
    alive = int(g.counts()[0][0])
    g.close()
    kind = 'reference' if pyorc.have_reference() else 'port'
    w = pyorc.World(sp, N, npred, seed=1, kind=kind)
    w.set_state(**st)
    w.step(0, 2)
    t = w.timed_steps(2, cpu_steps)
    w.close()
    out = {'config': name, 'n_agents': N, 'replicates': R, 'n_pred': npred, 'box': sp.sizeL, 'BC': sp.BC,
           'neighbour_rule': 'voronoi' if neighbour else 'metric', 'gpu_steps': steps,
           'gpu_agent_steps_per_s': N * R * steps / (ms * 1e-3), 'gpu_us_per_step': 1e3 * ms / steps,
           'gpu_pairs_or_candidates_per_s': (s1['pairs'] - s0['pairs']) / (ms * 1e-3),
           'gpu_launches_per_step': sum(s1[k] - s0[k] for k in ('pair_launches', 'update_launches', 'block_launches', 'other_launches')) / steps,
           'alive_at_end': alive,
           'cpu_kind': kind, 'cpu_cores': 1, 'cpu_steps': cpu_steps, 'cpu_agent_steps_per_s': N * cpu_steps / t,
           'cpu_us_per_step': 1e6 * t / cpu_steps}
    out['speedup_vs_1_core'] = out['gpu_agent_steps_per_s'] / out['cpu_agent_steps_per_s']
    print(json.dumps(out), flush=True)


def main():
    # C1: single tank, N = 8 and 30, metric and Voronoi rule
    for N in (8, 30):
        d = base_dict('burst_coast', N)
        run('C1 tank N=%d (metric)' % N, d, 0, 40000, 40000)
        run('C1 tank N=%d (voronoi)' % N, d, 0, 40000, 40000, neighbour=PP.NEIGHBOUR_VORONOI)
    # the reference's default run (RunSingle.py: natPred, N = 30, L = 100, chasing predator, kill rule 3): one swarm
    # and an ensemble of 16,384 replicates in one launch
    d = base_dict('natPred', 30, BC=0, size=100.0, Npred=1, pred_move=1, pred_kill=3, pred_time=0.02, time=12.0)
    run('natPred default N=30', d, 1, 40000, 20000)
    run('natPred default N=30 x 16384 replicates', d, 1, 2000, 20000, R=16384)
    # C3: natPred N = 1024, sparse (L = 584) and dense (L = 100)
    for L in (584.0, 100.0):
        d = base_dict('natPred', 1024, BC=0, size=L, Npred=1, pred_move=1, pred_kill=3, pred_time=0.02, time=12.0, pred_speed0=20)
        run('C3 natPred N=1024 L=%g' % L, d, 1, 4000, 300)
    # C4: mulFisher N = 16384, sparse (L = 1280, 640 net nodes) and dense (L = 100, 50 nodes)
    for L in (1280.0, 100.0):
        d = base_dict('mulFisher', 16384, BC=0, size=L, Npred=int(L / 2), pred_kill=1, pred_time=0.02, time=12.0)
        run('C4 mulFisher N=16384 L=%g' % L, d, int(L / 2), 1000, 3)


if __name__ == '__main__':
    main()
