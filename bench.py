#!/usr/bin/env python
"""bench.py - headline benchmark of the burst-coast hot path on B200 (contract: see the task brief).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
  torchrun --nproc-per-node N bench.py --gpus N ...      (one rank per GPU, NCCL for the barrier)

Workload (BASELINE.json configs[1], SURVEY.md section 8d "C2"): an ensemble of 16384 independent
burst-coast tanks PER GPU (BC=6, L=24.5, selection line 2, N agents each, default 32), metric
neighbour rule. One bench "step" = `--chunk` (default 1000) consecutive time steps of every tank =
one launch of the one-swarm-per-warp kernel. Replicates are independent: ranks share nothing
(weak scaling, no data-path collective); Philox streams are keyed by the global replicate id.

Printed JSON line: `value` = agent-steps/s with the state resident in HBM (CUDA events on the
handle's stream, L2 flushed between timed launches, max over ranks); `e2e` = the same metric through
the C ABI with HOST buffers (sbc_set_state from pinned memory -> sbc_step -> sbc_get_particles),
copies inside the timed region, the ensemble cut into four handles on their own streams / host threads so
that uploads, kernels and downloads of different parts overlap; `roofline` = the ensemble kernel against the FP32 peak; `pair_kernel`
= the dense all-pairs three-zone kernel (north-star kernel) against the FP32 peak;
`cpu_baseline` = the CPU oracle (the reference's own translation units when oracle/_ref is
present) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

SEED = 20261019
_T0 = time.time()


def log(msg):
    if os.environ.get('RANK', '0') == '0':
        print('[bench %.1fs] %s' % (time.time() - _T0, msg), file=sys.stderr, flush=True)
FLOP_PER_PAIR_OPEN, FLOP_PER_PAIR_PERIODIC = 13, 21
# Algorithmic FP32 work of one agent-step of the ensemble kernel:
#   coasting step 16 FLOP, bursting step 58 FLOP, burst start + (N-1) pairs * 13 FLOP;
#   weights = measured branch fractions of the tank (bursting 0.366, burst start 0.0047).
def flop_per_agent_step(n_agents):
    return 0.634 * 16 + 0.366 * 58 + 0.0047 * (60 + 13 * (n_agents - 1))


def fp32_peak_tflops(sm_count, mhz):
    return 2 * 128 * sm_count * mhz * 1e6 / 1e12


class ClockSampler(threading.Thread):
    """nvidia-smi clocks line of the profiling recipe, sampled during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(',')])
        except Exception:
            pass

    def mark(self):
        """Samples taken before this call are outside the timed region."""
        self.first = len(self.rows)

    def stop(self):
        time.sleep(0.25)
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        self.rows = self.rows[max(0, getattr(self, 'first', 0) - 1):]
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            if len(r) >= 9:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith('active'):
                        reasons.add(nm)
        busy = [v for v in sm if v > 0]
        return {'sm_mhz': float(np.median(busy)) if busy else None,
                'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons),
                'samples': len(self.rows)}


def tank_ensemble_state(R, N, sp, seed, first_replicate=0):
    """IC of config C2: uniform in a disc of radius 0.8 L, uniform headings, vproj = 1, every burst
    timer 0 (InitSystem), FP32-representable. Deterministic per global replicate id."""
    from helpers import f32
    out = {k: np.empty((R, N)) for k in ('x', 'y', 'phi', 'vproj')}
    for r in range(R):
        rng = np.random.default_rng([seed, first_replicate + r])
        rad = 0.8 * sp.sizeL * np.sqrt(rng.random(N))
        th = 2 * np.pi * rng.random(N)
        out['x'][r], out['y'][r] = rad * np.cos(th), rad * np.sin(th)
        out['phi'][r] = 2 * np.pi * rng.random(N) - np.pi
        out['vproj'][r] = 1.0
    return {k: f32(v) for k, v in out.items()}


def cpu_sample(sp, N, tanks, steps, kind, threads, seed, s_first=0):
    """Run `tanks` tanks for `steps` steps on the CPU oracle with `threads` host threads (ctypes
    releases the GIL). Returns (agent_steps_per_s, seconds, agent_steps)."""
    from oracle import pyorc
    from concurrent.futures import ThreadPoolExecutor
    worlds = []
    st = tank_ensemble_state(tanks, N, sp, seed)
    for r in range(tanks):
        w = pyorc.World(sp, N, seed=seed, replicate=r, kind=kind)
        w.set_state(**{k: v[r] for k, v in st.items()})
        worlds.append(w)
    t0 = time.perf_counter()
    if threads <= 1:
        for w in worlds:
            w.timed_steps(s_first, steps)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda w: w.timed_steps(s_first, steps), worlds))
    dt = time.perf_counter() - t0
    for w in worlds:
        w.close()
    n = tanks * N * steps
    return n / dt, dt, n


def oracle_kind():
    from oracle import pyorc
    if not os.path.exists(pyorc.PORT_PATH):
        pyorc.build(ref=False)
    return 'reference' if pyorc.have_reference() else 'port'


def run_reference(args, sp, N):
    """--impl reference: the reference's CPU implementation of the path on all host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    kind = oracle_kind()
    cores = len(os.sched_getaffinity(0))
    # calibrate: ~1.5 s of work per bench step on this machine
    rate1, _, _ = cpu_sample(sp, N, max(cores, 1), 200, kind, cores, SEED)
    tanks = max(cores, int(rate1 * 1.5 / (N * args.chunk)) // max(cores, 1) * max(cores, 1))
    for _ in range(args.warmup):
        cpu_sample(sp, N, cores, 100, kind, cores, SEED)
    tot_n, tot_t = 0, 0.0
    for _ in range(args.steps):
        _, dt, n = cpu_sample(sp, N, tanks, args.chunk, kind, cores, SEED)
        tot_n += n
        tot_t += dt
    val = tot_n / tot_t
    sample = '%d tanks x N=%d x %d steps per bench step, %d threads' % (tanks, N, args.chunk, cores)
    print(json.dumps({
        'impl': 'reference', 'metric': 'agent_steps_per_sec', 'value': val, 'unit': 'agent-steps/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * tot_t / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, N, tanks_override=tanks),
        'cpu_baseline': {'value': val, 'unit': 'agent-steps/s', 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': val, 'unit': 'agent-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


def workload_config(args, N, tanks_override=None):
    return {'workload': 'C2 ensemble of independent burst-coast tanks (BC=6, L=24.5, selection line 2, '
                        'metric neighbour rule), one bench step = %d time steps of every tank' % args.chunk,
            'tanks_per_gpu': tanks_override if tanks_override is not None else args.tanks,
            'agents_per_tank': N, 'time_steps_per_bench_step': args.chunk,
            'l2': 'flushed (256 MiB write) between timed launches', 'parallelism': 'replicates sharded, no collective'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--tanks', type=int, default=16384, help='tanks per GPU')
    ap.add_argument('--agents', type=int, default=32, help='agents per tank')
    ap.add_argument('--chunk', type=int, default=1000, help='time steps per bench step')
    ap.add_argument('--pair-n', type=int, default=65536, help='swarm size of the dense pair-kernel measurement')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    args = ap.parse_args()

    from helpers import make_params
    N = args.agents
    sp, _, _ = make_params('burst_coast', N)
    if args.impl == 'reference':
        run_reference(args, sp, N)
        return

    import faulthandler
    faulthandler.dump_traceback_later(240, exit=True)   # a hung rank prints where and dies instead of idling
    local = int(os.environ.get('LOCAL_RANK', '0'))
    sampler = ClockSampler(local)   # nvidia-smi child is started before CUDA / NCCL are initialised
    sampler.start()
    import torch
    import torch.distributed as dist
    from sde_burst_coast_b200.swarm import Swarm, load_library
    from sde_burst_coast_b200.params import PAIR_ALL_ROWS
    load_library()  # fails loudly when libsbc.so is missing: there is no CPU path
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device - the native arm has no CPU fallback')
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    R = args.tanks
    stream = torch.cuda.Stream()
    g = Swarm(sp, N, n_replicates=R, seed=SEED, replicate_offset=rank * R, device=local,
              stream=stream.cuda_stream)
    st = tank_ensemble_state(R, N, sp, SEED, first_replicate=rank * R)
    log('state built')
    g.set_state(**st)
    log('state uploaded')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    s_now = 0

    # ---- device-resident timing --------------------------------------------------------------
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            g.step(s_now, args.chunk)
            s_now += args.chunk
        barrier()
        log('warm-up done')
        stats0 = g.stats()
        sampler.mark()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        for k in range(args.steps):
            flush.zero_()
            ev[k][0].record(stream)
            g.step(s_now, args.chunk)
            ev[k][1].record(stream)
            s_now += args.chunk
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        log('timed region done: %.3f ms' % ms)
        clocks = sampler.stop()
        stats1 = g.stats()
    t = torch.tensor([ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    agent_steps_per_gpu = R * N * args.chunk * args.steps
    value = world * agent_steps_per_gpu / (ms_max * 1e-3)
    launches = sum(stats1[k] - stats0[k] for k in ('pair_launches', 'update_launches', 'block_launches', 'other_launches'))
    pairs = stats1['pairs'] - stats0['pairs']

    # ---- end to end through the C ABI with host buffers ----------------------------------------
    # The tanks are independent, so the public API is driven the way a user with host data would drive it: the
    # ensemble is cut into E2E_PARTS handles, each on its own stream and host thread (ctypes releases the GIL), so
    # that the upload of one part, the kernel of another and the download of a third overlap. Every bench step
    # still uploads every tank's state from pinned host memory and downloads every tank's particle rows.
    # The state every e2e step uploads is the ensemble as the timed region above left it (steady state, timers
    # included), so that both figures describe the same regime of the dynamics.
    snap = g.get_state()
    g.close()
    host = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() for k, v in snap.items()}
    part = torch.empty((R, N, 7), dtype=torch.float64).pin_memory().numpy()
    alive = torch.empty((R, N), dtype=torch.uint8).pin_memory().numpy()
    import ctypes
    dp, u8p = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint8)
    E2E_PARTS = max(1, min(4, R // 1024))
    cuts = [R * k // E2E_PARTS for k in range(E2E_PARTS + 1)]
    parts = []
    for k in range(E2E_PARTS):
        r0, r1 = cuts[k], cuts[k + 1]
        sk = torch.cuda.Stream()
        hk = Swarm(sp, N, n_replicates=r1 - r0, seed=SEED, replicate_offset=rank * R + r0, device=local, stream=sk.cuda_stream)
        parts.append((hk, sk, r0, r1))

    def e2e_run(hk, r0, r1, s_first, n):
        sub = {k: v[r0:r1] for k, v in host.items()}
        for q in range(n):
            hk.set_state(**sub)                      # H2D: the full state in the ABI layout (8 doubles, 2 timers, alive)
            hk.step(s_first + q * args.chunk, args.chunk)
            hk._check(hk.lib.sbc_get_particles(hk.h, part[r0:r1].ctypes.data_as(dp), alive[r0:r1].ctypes.data_as(u8p)))  # D2H

    def e2e_all(s_first, n):
        th = [threading.Thread(target=e2e_run, args=(hk, r0, r1, s_first, n)) for hk, _, r0, r1 in parts]
        for x in th:
            x.start()
        for x in th:
            x.join()

    e2e_all(s_now, 2)
    log('e2e warm-up done')
    barrier()
    t0 = time.perf_counter()
    e2e_all(s_now, args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * agent_steps_per_gpu / float(t.item())
    h2d = sum(v.nbytes for v in host.values())
    d2h = part.nbytes + alive.nbytes
    log('e2e done: %.3f s' % e2e_s)
    for hk, _, _, _ in parts:
        hk.close()

    out = None
    if rank == 0:
        props = torch.cuda.get_device_properties(local)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        mhz = peaks.get('sm_max_mhz', 1965.0)
        peak_fp32 = fp32_peak_tflops(props.multi_processor_count, mhz)
        fl = flop_per_agent_step(N)
        ach = (R * N * args.chunk * args.steps) * fl / (ms * 1e-3) / 1e12   # this rank's kernel
        out = {
            'metric': 'agent_steps_per_sec', 'value': value, 'unit': 'agent-steps/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_max / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': workload_config(args, N),
            'pair_interactions_per_sec': world * pairs / (ms_max * 1e-3),
            'e2e': {'value': e2e_value, 'unit': 'agent-steps/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'note': 'four handles on their own streams and host threads: uploads, kernels and downloads of different '
                            'parts of the ensemble overlap, and no L2 flush separates the launches - which is why this can '
                            'reach the device-resident figure'},
            'gpu_launches': int(launches), 'clocks': clocks,
            'roofline': {'kernel': 'swarm_warp_kernel' if N <= 32 else 'swarm_block_kernel', 'bound': 'fp32',
                         'achieved': ach, 'peak': peak_fp32, 'unit': 'TFLOP/s', 'frac': ach / peak_fp32,
                         # dram__bytes_read+write of this kernel per launch: profiles/r1_ncu_ensemble_kernel.txt
                         # (6.47 MB for 4096 tanks -> scaled to the tanks of this run; writes stay in L2)
                         'traffic': 6.47e6 * R / 4096 * N / 32, 'flop_per_agent_step': fl,
                         'issue_slot_utilisation_ncu': 0.752, 'warp_instructions_per_warp_step_ncu': 214,
                         'peak_source': '2*128 lanes*%d SMs*%.0f MHz (MEASURED_PEAKS.json sm_max_mhz)' % (props.multi_processor_count, mhz),
                         'note': 'state stays in registers for the whole launch (112 B per agent per launch of HBM traffic); the kernel is '
                                 'bound by instruction issue (control, integer, transcendental work), not by FMA throughput'},
        }
        # ---- the north-star kernel: dense all-pairs three-zone pass -------------------------------
        try:
            out['pair_kernel'] = pair_kernel_roofline(args.pair_n, local, peak_fp32)
        except Exception as e:  # keep the headline line even if this leg fails
            out['pair_kernel'] = {'error': str(e)}
        log('pair kernel leg done')
        if world == 1 and not args.no_cpu:
            kind = oracle_kind()
            rate, secs, n = cpu_sample(sp, N, 40, args.chunk * 8, kind, 1, SEED)
            out['cpu_baseline'] = {'value': rate, 'unit': 'agent-steps/s', 'cores': 1, 'kind': kind,
                                   'sample': '40 tanks x N=%d x %d steps, single thread, %.1f s' % (N, args.chunk * 8, secs)}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def pair_kernel_roofline(n, device, peak_fp32):
    """Dense state (every row active, the reference state at s = 1): N(N-1) directed pairs/launch.
    Timed by sbc_bench_pair_pass with CUDA events on the handle's stream (gather + all-pairs kernel +
    split combine; the Hilbert sort of the positions is timed separately as `with_sort`). Cases:
    the probe density of SURVEY section 6 (N=16384 in L=100: ~1960 neighbours inside att_range) in an
    open and a periodic domain, and the sparse density of config C4 (rho = 0.01: ~12 neighbours)."""
    from helpers import make_params, f32
    from sde_burst_coast_b200.swarm import Swarm
    from sde_burst_coast_b200.params import PAIR_ALL_ROWS
    res = {'bound': 'fp32', 'kernel': 'pair_dense_kernel', 'unit': 'TFLOP/s', 'peak': peak_fp32, 'cases': []}
    cases = [('open, dense', n, 'burst_coast', dict(BC=-1), 1.6384, FLOP_PER_PAIR_OPEN),
             ('periodic, dense', n, 'natPred', dict(BC=0, Npred=0), 1.6384, FLOP_PER_PAIR_PERIODIC),
             ('periodic, sparse (C4 density)', n, 'natPred', dict(BC=0, Npred=0), 0.01, FLOP_PER_PAIR_PERIODIC),
             ('open, dense', 16384, 'burst_coast', dict(BC=-1), 1.6384, FLOP_PER_PAIR_OPEN)]
    for name, nn, mode, over, rho, flop in cases:
        L = float(np.sqrt(nn / rho))
        if over.get('BC') == 0:
            over = dict(over, size=L)
        sp, _, _ = make_params(mode, nn, **over)
        rng = np.random.default_rng(SEED)
        st = dict(x=f32(L * rng.random(nn)), y=f32(L * rng.random(nn)), phi=f32(2 * np.pi * rng.random(nn)),
                  vproj=np.ones(nn))
        g = Swarm(sp, nn, device=device)
        g.set_state(**st)
        ms = g.bench_pair_pass(PAIR_ALL_ROWS, iters=10)
        ms_sort = g.bench_pair_pass(PAIR_ALL_ROWS | 2, iters=10)
        stats = g.stats()
        g.close()
        pairs = nn * (nn - 1)
        ach = pairs * flop / (ms * 1e-3) / 1e12
        res['cases'].append({'case': name, 'n_agents': nn, 'box': L, 'pairs_per_launch': pairs, 'ms_per_launch': ms,
                             'ms_per_launch_with_sort': ms_sort, 'pairs_per_sec': pairs / (ms * 1e-3),
                             'flop_per_pair': flop, 'achieved': ach, 'frac': ach / peak_fp32,
                             'frac_at_13_flop': pairs * 13 / (ms * 1e-3) / 1e12 / peak_fp32,
                             'ambiguous_pairs_rechecked_per_launch': stats['ambiguous_pairs'] // 22})
    # headline of this object: the open-domain dense case at the requested N
    res.update({k: res['cases'][0][k] for k in ('n_agents', 'pairs_per_sec', 'achieved', 'frac', 'ms_per_launch')})
    res['note'] = ('every directed pair has its distance evaluated; the zone arithmetic runs only for pairs that '
                   'pass d2 <= att2 (warp-uniform early-out on Hilbert-sorted rows), so sparse states and the '
                   'periodic 21-FLOP convention can read above 1.0; frac_at_13_flop is the conservative figure')
    return res


if __name__ == '__main__':
    main()
