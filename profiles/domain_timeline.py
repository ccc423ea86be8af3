#!/usr/bin/env python
"""In-situ timeline of the decomposed swarm (config C5, weak scaling, one rank per GPU): what every step's kernels did
on rank 0 - bulk update, exchange receive (wait for the neighbours + scatter into the ghosts), rows kernel, exchange
send - from sbc_debug_trace (%globaltimer). Launch with torch.distributed.run, one process per GPU.
usage: python -m torch.distributed.run --nproc-per-node W profiles/domain_timeline.py [--agents N] [--steps K]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--agents', type=int, default=1 << 20, help='agents per GPU')
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--warmup', type=int, default=60)
    args = ap.parse_args()
    import torch, torch.distributed as dist
    from helpers import base_dict, f32
    from sde_burst_coast_b200 import params as PP
    from sde_burst_coast_b200.domain import SlabSwarm, PeerTransport, step_distributed
    world, rank, local = int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    N = args.agents * world
    L = float(np.sqrt(N / 0.01))
    sp, _ = PP.make_sbc_params(base_dict('natPred', N, BC=0, size=L, Npred=0), path=PP.PATH_CELLLIST)
    rng = np.random.default_rng(20261019)
    st = dict(x=f32(L * rng.random(N)), y=f32(L * rng.random(N)), phi=f32(2 * np.pi * rng.random(N) - np.pi), vproj=np.ones(N))
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    slab = SlabSwarm(sp, N, rank, world, seed=20261019, device=local, stream=ts.cuda_stream)
    slab.scatter_initial(st)
    tr = PeerTransport(slab)
    step_distributed(slab, tr, 0, args.warmup)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    slab.swarm.trace(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ts)
    step_distributed(slab, tr, args.warmup, args.steps)
    e1.record(ts)
    torch.cuda.synchronize()
    t = slab.swarm.trace(False, read=True).astype(np.float64)
    steps = np.arange(args.warmup, args.warmup + args.steps) % 1024
    rows, bulk = t[steps, 0], t[steps, 1]
    t0 = np.minimum(rows[:, 0], bulk[:, 0])
    t1 = np.maximum(np.maximum(rows[:, 1], bulk[:, 1]), t[steps, 6, 1])
    period = np.diff(t0)
    med = np.median(period)
    ex = t[:, 2:4]                      # recv, send by exchange count
    recv = ex[:, 0][ex[:, 0, 1] > 0]
    send = ex[:, 1][ex[:, 1, 1] > 0]
    out = {'rank': rank, 'world': world, 'us_per_step_events': 1e3 * e0.elapsed_time(e1) / args.steps,
           'step_period_median_us': med / 1e3, 'long_periods': int((period > 2 * med).sum()),
           'long_period_mean_us': float(period[period > 2 * med].mean()) / 1e3 if (period > 2 * med).any() else 0.0,
           'bulk_us': [float(np.median(bulk[:, 0] - t0)) / 1e3, float(np.median(bulk[:, 1] - t0)) / 1e3],
           'rows_us': [float(np.median(rows[:, 0] - t0)) / 1e3, float(np.median(rows[:, 1] - t0)) / 1e3],
           'recv_duration_us': float(np.median(recv[:, 1] - recv[:, 0])) / 1e3 if len(recv) else None,
           'send_duration_us': float(np.median(send[:, 1] - send[:, 0])) / 1e3 if len(send) else None,
           'edge_rows_us': [float(np.median(t[steps, 6, 0] - t0)) / 1e3, float(np.median(t[steps, 6, 1] - t0)) / 1e3],
           'edge_wait_through_us': [float(np.median(t[steps, 7, 0] - t0)) / 1e3, float(np.median(t[steps, 7, 1] - t0)) / 1e3],
           'busy_us': float(np.median(t1 - t0)) / 1e3, 'gap_to_next_step_us': float(np.median(t0[1:] - t1[:-1])) / 1e3,
           'exchanges_traced': [int(len(recv)), int(len(send))]}
    # where inside a step the exchange kernels sit: match every receive to the step whose rows kernel it precedes
    if len(recv):
        r_start = np.sort(recv[:, 0])
        idx = np.searchsorted(t0, r_start, side='right') - 1
        okm = (idx >= 0) & (idx < len(t0))
        out['recv_start_in_step_us'] = float(np.median(r_start[okm] - t0[idx[okm]])) / 1e3
        r_end = np.sort(recv[:, 1])
        idx = np.searchsorted(t0, r_end, side='right') - 1
        okm = (idx >= 0) & (idx < len(t0))
        out['recv_end_in_step_us'] = float(np.median(r_end[okm] - t0[idx[okm]])) / 1e3
    if len(send):
        out['flags_seen_after_recv_start_us'] = [float(np.median(send[:, 0] - recv[:len(send), 0])) / 1e3, float(np.median(send[:, 1] - recv[:len(send), 0])) / 1e3] if len(send) == len(recv) else None
        s_start = np.sort(send[:, 0])
        idx = np.searchsorted(t0, s_start, side='right') - 1
        okm = (idx >= 0) & (idx < len(t0))
        out['send_start_in_step_us'] = float(np.median(s_start[okm] - t0[idx[okm]])) / 1e3
        s_end = np.sort(send[:, 1])
        idx = np.searchsorted(t0, s_end, side='right') - 1
        okm = (idx >= 0) & (idx < len(t0))
        out['send_end_in_step_us'] = float(np.median(s_end[okm] - t0[idx[okm]])) / 1e3
    if world > 1:
        every = [None] * world
        dist.gather_object(out, every if rank == 0 else None, dst=0)
    else:
        every = [out]
    if rank == 0:
        for o in every:
            print(json.dumps(o))
    slab.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
