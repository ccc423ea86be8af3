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


def c2_runs(cpu_seconds):
    """16,384 independent tanks, N in {8, 16, 32, 64} (10,000 steps each), a mixed ensemble
    (N = 8 * 2^(r mod 4): four handles of 4,096 tanks on four streams at once) and one 40,000-step run; the reference's
    own translation units on ONE core beside each (a bounded sample of single tanks, throughput per agent-step)."""
    import torch
    import bench
    from helpers import make_params
    from sde_burst_coast_b200.swarm import Swarm
    from oracle import pyorc
    kind = 'reference' if pyorc.have_reference() else 'port'
    R = 16384

    def cpu_rate(sp, N):
        rate, secs, n = bench.cpu_sample(sp, N, max(1, 8 // max(1, N // 8)), int(2000 * cpu_seconds), kind, 1, bench.SEED)
        return rate

    def gpu_run(N, tanks, steps, stream, first=0):
        sp, _, _ = make_params('burst_coast', N)
        g = Swarm(sp, N, n_replicates=tanks, seed=bench.SEED, replicate_offset=first, stream=stream.cuda_stream)
        g.set_state(**bench.tank_ensemble_state(tanks, N, sp, bench.SEED, first_replicate=first))
        return sp, g

    for N, steps in ((8, 10000), (16, 10000), (32, 10000), (64, 10000), (32, 40000)):
        ts = torch.cuda.Stream()
        sp, g = gpu_run(N, R, steps, ts)
        with torch.cuda.stream(ts):
            g.step(0, 1000)
            torch.cuda.synchronize()
            s0 = g.stats()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ts)
            g.step(1000, steps)
            e1.record(ts)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        s1 = g.stats()
        g.close()
        cpu = cpu_rate(sp, N)
        out = {'config': 'C2 ensemble %d tanks x N=%d, %d steps' % (R, N, steps), 'n_agents': N, 'replicates': R, 'gpu_steps': steps,
               'gpu_ms': ms, 'gpu_agent_steps_per_s': N * R * steps / (ms * 1e-3),
               'gpu_pairs_per_s': (s1['pairs'] - s0['pairs']) / (ms * 1e-3), 'gpu_launches': int(s1['block_launches'] - s0['block_launches']),
               'kernel': 'swarm_warp_kernel (%d lanes per tank)' % max(8, N) if N <= 32 else 'swarm_block_kernel (one block per tank)',
               'cpu_kind': kind, 'cpu_cores': 1, 'cpu_agent_steps_per_s': cpu}
        out['speedup_vs_1_core'] = out['gpu_agent_steps_per_s'] / cpu
        print(json.dumps(out), flush=True)
    # mixed ensemble: replicate r holds N = 8 * 2^(r mod 4) agents -> four homogeneous handles of 4096 tanks, run at once
    steps = 10000
    parts = []
    for k, N in enumerate((8, 16, 32, 64)):
        ts = torch.cuda.Stream()
        sp, g = gpu_run(N, R // 4, steps, ts, first=k * (R // 4))
        parts.append((N, ts, g))
    for N, ts, g in parts:
        g.step(0, 1000)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for N, ts, g in parts:
        g.step(1000, steps)
    torch.cuda.synchronize()
    secs = time.perf_counter() - t0
    agents = sum(N * (R // 4) for N, _, _ in parts)
    for _, _, g in parts:
        g.close()
    print(json.dumps({'config': 'C2 mixed ensemble: %d tanks, N = 8 * 2^(r mod 4), %d steps, four handles on four streams' % (R, steps),
                      'agents_total': agents, 'gpu_steps': steps, 'gpu_ms': secs * 1e3, 'timing': 'wall clock around the four launches, synchronised',
                      'gpu_agent_steps_per_s': agents * steps / secs}), flush=True)


def c4_sweep():
    """mulFisher N = 16,384, L = 1280, 640 net nodes, prob_social in {0, .25, .5, .71, .74, 1}."""
    for ps in (0.0, 0.25, 0.5, 0.71, 0.74, 1.0):
        d = base_dict('mulFisher', 16384, BC=0, size=1280.0, Npred=640, pred_kill=1, pred_time=0.02, time=12.0, prob_social=ps)
        run('C4 mulFisher N=16384 L=1280 prob_social=%g' % ps, d, 640, 1000, 2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--set', default='base', choices=['base', 'c2', 'c4sweep'])
    ap.add_argument('--cpu-seconds', type=float, default=4.0)
    args = ap.parse_args()
    if args.set == 'c2':
        return c2_runs(args.cpu_seconds)
    if args.set == 'c4sweep':
        return c4_sweep()
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
