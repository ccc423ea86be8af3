#!/usr/bin/env python
"""In-situ timeline of the multi-kernel step of config C5 (one GPU): sbc_debug_trace records, per step, when the first
block of each kernel started and the last one ended (%globaltimer). Prints the median offsets of the rows kernel and the
bulk update inside a step and the step period - what a profiler's timeline view would show for the graph-replayed run.
usage: python profiles/c5_timeline.py [--agents N] [--steps K]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--agents', type=int, default=1 << 20)
    ap.add_argument('--steps', type=int, default=400)
    ap.add_argument('--warmup', type=int, default=60)
    args = ap.parse_args()
    import torch
    from helpers import base_dict, f32
    from sde_burst_coast_b200 import params as PP
    from sde_burst_coast_b200.swarm import Swarm
    N = args.agents
    L = float(np.sqrt(N / 0.01))
    sp, _ = PP.make_sbc_params(base_dict('natPred', N, BC=0, size=L, Npred=0), path=PP.PATH_CELLLIST)
    rng = np.random.default_rng(20261019)
    st = dict(x=f32(L * rng.random(N)), y=f32(L * rng.random(N)), phi=f32(2 * np.pi * rng.random(N) - np.pi), vproj=np.ones(N))
    ts = torch.cuda.Stream()
    g = Swarm(sp, N, seed=20261019, stream=ts.cuda_stream)
    g.set_state(**st)
    g.step(0, args.warmup)
    g.trace(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ts)
    g.step(args.warmup, args.steps)
    e1.record(ts)
    ts.synchronize()
    tr = g.trace(False, read=True).astype(np.float64)
    g.close()
    steps = np.arange(args.warmup, args.warmup + args.steps) % 1024
    rows, bulk = tr[steps, 0], tr[steps, 1]
    ok = (rows[:, 1] > 0) & (bulk[:, 1] > 0)
    t0 = np.minimum(rows[:, 0], bulk[:, 0])
    t1 = np.maximum(rows[:, 1], bulk[:, 1])
    period = np.diff(t0)
    out = {'n_agents': N, 'steps': int(ok.sum()), 'us_per_step_events': 1e3 * e0.elapsed_time(e1) / args.steps,
           'step_period_us': float(np.median(period)) / 1e3,
           'rows_start_us': float(np.median(rows[:, 0] - t0)) / 1e3, 'rows_end_us': float(np.median(rows[:, 1] - t0)) / 1e3,
           'bulk_start_us': float(np.median(bulk[:, 0] - t0)) / 1e3, 'bulk_end_us': float(np.median(bulk[:, 1] - t0)) / 1e3,
           'rows_last_block_start_us': float(np.median(tr[steps, 4, 1] - t0)) / 1e3,
           'bulk_last_block_start_us': float(np.median(tr[steps, 5, 1] - t0)) / 1e3,
           'busy_us': float(np.median(t1 - t0)) / 1e3, 'gap_to_next_step_us': float(np.median(t0[1:] - t1[:-1])) / 1e3}
    # phase marks of the rows kernel (celllist.cu, cells_phase_mark): the latest warp past each phase
    for name, slot in (('rows_bounds_in_us', 2), ('rows_sweep_done_us', 3), ('rows_reduced_us', 6), ('rows_updated_us', 7)):
        v = tr[steps, slot, 1]
        if (v > 0).any():
            out[name] = float(np.median((v - t0)[v > 0])) / 1e3
    long = period[period > 2 * np.median(period)]
    out['long_periods'] = int(long.size)
    out['long_period_mean_us'] = float(long.mean()) / 1e3 if long.size else 0.0
    print(json.dumps(out))


if __name__ == '__main__':
    main()
