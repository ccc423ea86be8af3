#!/usr/bin/env python
"""BASELINE config C5: ONE large periodic swarm (default N = 1,048,576, L = 10240, rho = 0.01) on the
cell-list path, decomposed into x-slabs over the GPUs of one node with NCCL halo / migration exchange.

  python profiles/bench_c5.py [--n N] [--steps K]                                  (1 GPU)
  python -m torch.distributed.run --nproc-per-node W ... profiles/bench_c5.py     (W GPUs)

Prints one JSON line on rank 0: agent-steps/s (strong scaling: total work fixed), pair candidates
examined per second, per-rank owned / ghost counts. Timing: CUDA events on the stream all kernels and
NCCL operations are ordered on, barrier + synchronize on both sides, max over ranks."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--agents', type=int, default=1 << 20, help='total agents (strong scaling) or agents per GPU with --weak')
    ap.add_argument('--weak', action='store_true', help='agents scale with the number of GPUs')
    ap.add_argument('--rho', type=float, default=0.01)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--transport', choices=['peer', 'nccl'], default='peer',
                    help='peer: libsbc kernels store into the neighbours\' inboxes over NVLink; nccl: torch.distributed send/recv')
    args = ap.parse_args()
    import faulthandler
    faulthandler.dump_traceback_later(280, exit=True)
    import torch, torch.distributed as dist
    from helpers import base_dict, f32
    from sde_burst_coast_b200 import params as PP
    from sde_burst_coast_b200.domain import SlabSwarm, DistTransport, PeerTransport, step_distributed
    world, rank, local = int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    N = args.agents * (world if args.weak else 1)
    L = float(np.sqrt(N / args.rho))
    d = base_dict('natPred', N, BC=0, size=L, Npred=0)
    sp, _ = PP.make_sbc_params(d, path=PP.PATH_CELLLIST)
    rng = np.random.default_rng(20261019)
    st = dict(x=f32(L * rng.random(N)), y=f32(L * rng.random(N)), phi=f32(2 * np.pi * rng.random(N) - np.pi), vproj=np.ones(N))
    tstream = torch.cuda.Stream()   # a capturable (non-default) stream shared by libsbc kernels and NCCL
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    slab = SlabSwarm(sp, N, rank, world, seed=20261019, device=local, stream=stream)
    slab.scatter_initial(st)
    tr = PeerTransport(slab) if args.transport == 'peer' else DistTransport(rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_distributed(slab, tr, 0, args.warmup)       # includes the synchronised first burst at s = 1
    barrier()
    s0 = slab.swarm.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    step_distributed(slab, tr, args.warmup, args.steps)
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    s1 = slab.swarm.stats()
    c = slab.counts()
    t = torch.tensor([ms, float(s1['pairs'] - s0['pairs']), float(c['owned']), float(c['ghosts']), float(c['overflow_events'])],
                     dtype=torch.float64, device='cuda')
    allv = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(allv, t)
    else:
        allv = [t]
    if rank == 0:
        ms_max = max(float(v[0]) for v in allv)
        cand = sum(float(v[1]) for v in allv)
        print(json.dumps({
            'metric': 'agent_steps_per_sec', 'workload': 'C5 single swarm, cell list, x-slabs', 'n_agents': N, 'box': L,
            'n_gpus': world, 'transport': args.transport if world > 1 else 'none', 'steps': args.steps, 'scaling': 'weak' if args.weak else 'strong',
            'value': N * args.steps / (ms_max * 1e-3), 'us_per_step': 1e3 * ms_max / args.steps,
            'pair_candidates_per_sec': cand / (ms_max * 1e-3), 'wall_s': wall,
            'owned_per_rank': [int(v[2]) for v in allv], 'ghosts_per_rank': [int(v[3]) for v in allv],
            'overflow_events': int(sum(float(v[4]) for v in allv)),
            'launches_per_step': (sum(s1[k] - s0[k] for k in ('pair_launches', 'update_launches', 'other_launches'))) / args.steps}))
    slab.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
