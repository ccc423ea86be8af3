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
`c5_step` (1 GPU) = one large periodic swarm on the cell-list path against the HBM roofline of SURVEY 8(d);
`domain_swarm` (N > 1 GPUs) = that swarm decomposed into x-slabs with the peer-memory halo exchange, 1,048,576 agents
per GPU, plus a small decomposed run whose burst timers are compared bit for bit with the undecomposed swarm;
`cpu_baseline` = the CPU oracle (the reference's own translation units when oracle/_ref is
present) on a bounded sample of the same workload.

Numbers that only a profiler can give (DRAM traffic, issue-slot utilisation, instructions per pair) are NOT literals in
this file: they are read at run time from the ncu summaries committed under profiles/ (named in the output), and
only used when the summary was taken on the same configuration and its kernel duration agrees with this run's.
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
STATE_BYTES_ENSEMBLE = 56   # pv 16 + hv 16 (phi, vproj, cos, sin) + force 8 + timers 8 + (alive 1, rounded) per agent
# Algorithmic FP32 work of one agent-step of the ensemble kernel:
#   coasting step 16 FLOP, bursting step 58 FLOP, burst start + (N-1) pairs * 13 FLOP;
#   weights = measured branch fractions of the tank (bursting 0.366, burst start 0.0047).
def flop_per_agent_step(n_agents):
    return 0.634 * 16 + 0.366 * 58 + 0.0047 * (60 + 13 * (n_agents - 1))


def fp32_peak_tflops(sm_count, mhz):
    return 2 * 128 * sm_count * mhz * 1e6 / 1e12


def read_ncu_summary(name):
    """Parse profiles/<name> (written by profiles/summarize.py report): a `# config: k=v ...` line, then per captured
    launch a `## kernel` header and `metric value unit` rows. Returns (config dict, [ {kernel, metrics...} ]) or None."""
    path = os.path.join(ROOT, 'profiles', name)
    if not os.path.exists(path):
        return None
    cfg, launches, cur = {}, [], None
    for line in open(path):
        line = line.rstrip()
        if line.startswith('# config:'):
            for tok in line[len('# config:'):].split():
                if '=' in tok:
                    k, v = tok.split('=', 1)
                    cfg[k] = v
        elif line.startswith('## '):
            cur = {'kernel': line[3:].strip()}
            launches.append(cur)
        elif cur is not None and line and not line.startswith('#'):
            tok = line.split()
            if len(tok) >= 2:
                try:
                    val = float(tok[1].replace(',', ''))
                except ValueError:
                    continue
                unit = tok[2] if len(tok) > 2 else ''
                scale = {'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 'msecond': 1e-3, 'usecond': 1e-6, 'nsecond': 1e-9, 'second': 1.0,
                         'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}.get(unit, 1.0)
                cur[tok[0]] = val * scale
    return cfg, launches


def recorded_profile(name, want_cfg, kernel_substr, my_seconds_per_launch, tol=0.10):
    """The committed ncu summary `name`, if it belongs to this configuration (every key of want_cfg equal) and its
    kernel duration agrees with this run's own timing within `tol`. Returns a dict for the JSON line (always names
    the file and says why it was or was not used)."""
    got = read_ncu_summary(name)
    out = {'file': 'profiles/' + name}
    if got is None:
        out['used'] = False
        out['why'] = 'no such summary committed'
        return out
    cfg, launches = got
    out['config'] = cfg
    mism = [k for k, v in want_cfg.items() if str(cfg.get(k)) != str(v)]
    rows = [l for l in launches if kernel_substr in l['kernel']]
    if mism or not rows:
        out['used'] = False
        out['why'] = 'summary is for another configuration (%s)' % ', '.join(mism) if mism else 'kernel not in the summary'
        return out
    r = rows[0]
    dur = r.get('gpu__time_duration.sum')
    out['kernel'] = r['kernel'][:80]
    out['ncu_seconds_per_launch'] = dur
    out['bench_seconds_per_launch'] = my_seconds_per_launch
    ok = dur is not None and my_seconds_per_launch > 0 and abs(dur / my_seconds_per_launch - 1.0) <= tol
    out['used'] = bool(ok)
    if not ok:
        out['why'] = 'ncu kernel duration differs from this run by more than %d %%: the summary is stale' % int(tol * 100)
        print('[bench] WARNING: %s does not describe the kernel this run timed (%.4g s vs %.4g s per launch)' %
              (name, dur or float('nan'), my_seconds_per_launch), file=sys.stderr)
        return out
    for k_out, k_in in (('dram_bytes_read', 'dram__bytes_read.sum'), ('dram_bytes_write', 'dram__bytes_write.sum'),
                        ('issue_slot_utilisation', 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
                        ('warp_instructions', 'smsp__inst_executed.sum'),
                        ('threads_per_warp_instruction', 'smsp__thread_inst_executed_per_inst_executed.ratio'),
                        ('fma_pipe_pct', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'),
                        ('registers_per_thread', 'launch__registers_per_thread'),
                        ('waves_per_sm', 'launch__waves_per_multiprocessor')):
        if k_in in r:
            out[k_out] = r[k_in]
    return out


def pin_to_gpu_numa(torch, local):
    """Bind this process (and the threads it starts later) to the CPUs of the NUMA node the GPU hangs on, so that pinned
    buffers are first touched there and the copy threads run there. Returns the node, or None when sysfs does not say."""
    try:
        p = torch.cuda.get_device_properties(local)
        bus = '%04x:%02x:%02x.0' % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bus).read().strip())
        if node < 0:
            return None
        cpus = set()
        for tok in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
            a, _, b = tok.partition('-')
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


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
        'gpu_launches': 0,
        'note': 'the reference\'s own translation units (oracle/_ref) when present, else the port. Bookkeeping of this arm (VERDICT r1 #14): the '
                'decisions come from one Philox block per agent and step computed inside the timed loop (the reference draws from a '
                'serial MT19937), and the Gaussian the reference draws and discards per agent and step (swarmdyn.cpp:188) is left '
                'out - roughly cancelling, a few per cent of the CPU time either way'}))


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
    ap.add_argument('--no-c5', action='store_true', help='skip the single-GPU large-swarm leg (c5_step)')
    ap.add_argument('--no-domain', action='store_true', help='skip the decomposed-swarm leg of a multi-GPU run (domain_swarm)')
    ap.add_argument('--c5-agents', type=int, default=1 << 20, help='agents per GPU of the large-swarm legs')
    args = ap.parse_args()

    from helpers import make_params
    N = args.agents
    sp, _, _ = make_params('burst_coast', N)
    if args.impl == 'reference':
        run_reference(args, sp, N)
        return

    import faulthandler
    faulthandler.dump_traceback_later(420, exit=True)   # a hung rank prints where and dies instead of idling
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
    numa = pin_to_gpu_numa(torch, local)     # before the pinned buffers are allocated (first touch)
    import ctypes
    cores_here = len(os.sched_getaffinity(0))
    ranks_per_node = max(1, min(world, torch.cuda.device_count()))
    E2E_PARTS = max(1, min(4, R // 1024, max(1, cores_here * (2 if numa is not None else 1) // ranks_per_node)))
    cuts = [R * k // E2E_PARTS for k in range(E2E_PARTS + 1)]
    parts = []
    for k in range(E2E_PARTS):
        r0, r1 = cuts[k], cuts[k + 1]
        sk = torch.cuda.Stream()
        hk = Swarm(sp, N, n_replicates=r1 - r0, seed=SEED, replicate_offset=rank * R + r0, device=local, stream=sk.cuda_stream)
        parts.append((hk, sk, r0, r1))

    def e2e_measure(abi, n_steps):
        """Every bench step uploads every tank's state from pinned host memory, steps it, and downloads every tank's
        particle rows; `abi` = 'f64' (the reference's double layout) or 'f32' (the float entry points)."""
        real = np.float64 if abi == 'f64' else np.float32
        rp = ctypes.POINTER(ctypes.c_double if abi == 'f64' else ctypes.c_float)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        host = {}
        for k, v in snap.items():
            a = np.ascontiguousarray(v, dtype=real if v.dtype == np.float64 else v.dtype)
            host[k] = torch.from_numpy(a).pin_memory().numpy()
        part = torch.empty((R, N, 7), dtype=torch.float64 if abi == 'f64' else torch.float32).pin_memory().numpy()
        alive = torch.empty((R, N), dtype=torch.uint8).pin_memory().numpy()

        def run(hk, r0, r1, s_first, n):
            sub = {k: v[r0:r1] for k, v in host.items()}
            getp = hk.lib.sbc_get_particles if abi == 'f64' else hk.lib.sbc_get_particles_f32
            for q in range(n):
                (hk.set_state if abi == 'f64' else hk.set_state_f32)(**sub)      # H2D: 8 reals, 2 timers, alive per agent
                hk.step(s_first + q * args.chunk, args.chunk)
                hk._check(getp(hk.h, part[r0:r1].ctypes.data_as(rp), alive[r0:r1].ctypes.data_as(u8p)))  # D2H

        def run_all(s_first, n):
            th = [threading.Thread(target=run, args=(hk, r0, r1, s_first, n)) for hk, _, r0, r1 in parts]
            for x in th:
                x.start()
            for x in th:
                x.join()

        run_all(s_now, 2)
        barrier()
        t0 = time.perf_counter()
        run_all(s_now, n_steps)
        barrier()
        secs = time.perf_counter() - t0
        tt = torch.tensor([secs], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return (world * R * N * args.chunk * n_steps / float(tt.item()), sum(v.nbytes for v in host.values()),
                part.nbytes + alive.nbytes)

    # The tanks are independent, so the public API is driven the way a user with host data would drive it: the ensemble
    # is cut into E2E_PARTS handles, each on its own stream and host thread (ctypes releases the GIL), so that the upload
    # of one part, the kernel of another and the download of a third overlap. The state every e2e step uploads is the
    # ensemble as the timed region above left it (steady state, timers included). Headline: the float entry points
    # (the device state is FP32; half the bytes on the host link); the reference's double layout is measured beside it.
    e2e_value, h2d, d2h = e2e_measure('f32', args.steps)
    log('e2e (f32 ABI) done')
    e2e_f64_value, h2d64, d2h64 = e2e_measure('f64', max(2, args.steps // 2))
    log('e2e (f64 ABI) done')
    e2e_abi = 'float: sbc_set_state_f32 / sbc_get_particles_f32'
    e2e_note = ('%d handles on their own streams and host threads: uploads, kernels and downloads of different parts of the '
                'ensemble overlap, and no L2 flush separates the launches; host threads and pinned buffers on the NUMA node of '
                'the GPU (%s)' % (E2E_PARTS, 'node %d' % numa if numa is not None else 'not determined'))
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
        kname = 'swarm_warp_kernel' if N <= 32 else 'swarm_block_kernel'
        prof = recorded_profile('r2_ncu_ensemble_kernel.txt', {'tanks': R, 'agents': N, 'chunk': args.chunk}, kname,
                                ms * 1e-3 / args.steps)
        traffic = None
        if prof.get('used') and 'dram_bytes_read' in prof:
            traffic = prof['dram_bytes_read'] + prof.get('dram_bytes_write', 0.0)
            wi = prof.get('warp_instructions')
            if wi:
                prof['warp_instructions_per_warp_step'] = wi / (R * N / 32.0 * args.chunk)
        out = {
            'metric': 'agent_steps_per_sec', 'value': value, 'unit': 'agent-steps/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_max / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': workload_config(args, N),
            'pair_interactions_per_sec': world * pairs / (ms_max * 1e-3),
            'e2e': {'value': e2e_value, 'unit': 'agent-steps/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'abi': e2e_abi, 'parts': E2E_PARTS, 'note': e2e_note,
                    'f64_abi': {'value': e2e_f64_value, 'h2d_bytes_per_step': h2d64, 'd2h_bytes_per_step': d2h64,
                                'abi': 'double (reference layout): sbc_set_state / sbc_get_particles'}},
            'gpu_launches': int(launches), 'clocks': clocks,
            'roofline': {'kernel': kname, 'bound': 'fp32',
                         'achieved': ach, 'peak': peak_fp32, 'unit': 'TFLOP/s', 'frac': ach / peak_fp32,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu summary
                         # named in recorded_profile - null unless that summary is of this configuration and its kernel
                         # duration agrees with this run's
                         'traffic': traffic, 'flop_per_agent_step': fl, 'recorded_profile': prof,
                         'algorithmic_hbm_bytes_per_launch': 2 * STATE_BYTES_ENSEMBLE * R * N,
                         'peak_source': '2*128 lanes*%d SMs*%.0f MHz (MEASURED_PEAKS.json sm_max_mhz)' % (props.multi_processor_count, mhz),
                         'note': 'state stays in registers for the whole launch (HBM is touched at entry and exit only); the kernel is '
                                 'bound by instruction issue (control, integer, transcendental work), not by FMA throughput'},
        }
        # ---- the north-star kernel: dense all-pairs three-zone pass -------------------------------
        try:
            if args.pair_n > 0:
                out['pair_kernel'] = pair_kernel_roofline(args.pair_n, local, peak_fp32, props.multi_processor_count, mhz)
        except Exception as e:  # keep the headline line even if this leg fails
            out['pair_kernel'] = {'error': str(e)}
        log('pair kernel leg done')
        if world == 1 and not args.no_c5:
            try:
                out['c5_step'] = c5_single_gpu(local, peaks.get('hbm_gbs', 6563.2), args.c5_agents)
            except Exception as e:
                out['c5_step'] = {'error': str(e)}
            log('C5 leg done')
        if world == 1 and not args.no_cpu:
            kind = oracle_kind()
            rate, secs, n = cpu_sample(sp, N, 40, args.chunk * 8, kind, 1, SEED)
            out['cpu_baseline'] = {'value': rate, 'unit': 'agent-steps/s', 'cores': 1, 'kind': kind,
                                   'sample': '40 tanks x N=%d x %d steps, single thread, %.1f s' % (N, args.chunk * 8, secs)}
    if world > 1 and not args.no_domain:
        # every rank takes part; rank 0 gets the result
        try:
            dom = domain_swarm_leg(rank, world, local, args.c5_agents)
        except Exception as e:
            dom = {'error': repr(e)} if rank == 0 else None
        if rank == 0:
            out['domain_swarm'] = dom
        log('decomposed-swarm leg done')
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def pair_cases(n):
    """(name, agents, replicates, mode, parameter overrides, density, blob diameter / att_range) of the pair-kernel leg."""
    dense = 1.6384
    return [('open, dense', n, 1, 'burst_coast', dict(BC=-1), dense, None),
            ('periodic, dense', n, 1, 'natPred', dict(BC=0, Npred=0), dense, None),
            ('periodic, sparse (C4 density)', n, 1, 'natPred', dict(BC=0, Npred=0), 0.01, None),
            ('open, dense', 16384, 1, 'burst_coast', dict(BC=-1), dense, None),
            ('open, dense', 4096, 1, 'burst_coast', dict(BC=-1), dense, None),
            ('open, dense', 4096, 16, 'burst_coast', dict(BC=-1), dense, None),
            ('open, dense', 1024, 1, 'burst_coast', dict(BC=-1), dense, None),
            ('open, dense', 1024, 64, 'burst_coast', dict(BC=-1), dense, None),
            ('periodic, dense', 4096, 16, 'natPred', dict(BC=0, Npred=0), dense, None),
            ('open, all near', 4096, 16, 'burst_coast', dict(BC=-1), None, 0.95)]


def pair_case_tag(name, nn, reps, rho, blob):
    return '%s_n%d_r%d' % (name.split(',')[0] + ('_allnear' if blob else ('_sparse' if rho == 0.01 else '')), nn, reps)


def pair_case_state(name, nn, reps, mode, over, rho, blob):
    """Parameters and the seeded state of one pair-kernel case (also used by profiles/pair_case.py for the ncu captures)."""
    from helpers import make_params, f32
    if over.get('BC') == 0:
        over = dict(over, size=float(np.sqrt(nn / rho)))
    sp, _, _ = make_params(mode, nn, **over)
    rng = np.random.default_rng(SEED)
    if blob is not None:     # a disc of diameter blob * att_range: no far pair at all
        L = blob * sp.att_range
        rad, th = 0.5 * L * np.sqrt(rng.random((reps, nn))), 2 * np.pi * rng.random((reps, nn))
        x, y = rad * np.cos(th), rad * np.sin(th)
    else:
        L = float(np.sqrt(nn / rho))
        x, y = L * rng.random((reps, nn)), L * rng.random((reps, nn))
    st = dict(x=f32(x), y=f32(y), phi=f32(2 * np.pi * rng.random((reps, nn))), vproj=np.ones((reps, nn)))
    return sp, st, L, over


def pair_kernel_roofline(n, device, peak_fp32, sm_count, mhz):
    """Dense state (every row active, the reference state at s = 1): N(N-1) directed pairs/launch.
    Timed by sbc_bench_pair_pass with CUDA events on the handle's stream (gather + all-pairs kernel +
    split combine; the Hilbert sort of the positions is timed separately as `with_sort`). Cases:
    the probe density of SURVEY section 6 (N=16384 in L=100: ~1960 neighbours inside att_range) in an
    open and a periodic domain at N = 65536 / 16384 / 4096 / 1024 (the small sizes also as ensembles of
    replicates, so that a launch holds enough pairs to time), the sparse density of config C4 (rho = 0.01:
    ~12 neighbours), and an ALL-NEAR blob (diameter < att_range: every pair executes the full zone arithmetic).
    `frac` is ALWAYS the 13-FLOP figure of SURVEY 8(d) (a periodic pair costs more - `frac_at_21_flop` - but
    a far pair leaves after the distance test, so the larger convention can exceed 1)."""
    from helpers import make_params, f32
    from sde_burst_coast_b200.swarm import Swarm
    from sde_burst_coast_b200.params import PAIR_ALL_ROWS
    res = {'bound': 'fp32', 'kernel': 'pair_dense_kernel', 'unit': 'TFLOP/s', 'peak': peak_fp32, 'cases': []}
    issue_peak = 128.0 * sm_count * mhz * 1e6     # thread-instructions per second the FP32-class pipes can issue
    for name, nn, reps, mode, over, rho, blob in pair_cases(n):
        sp, st, L, over = pair_case_state(name, nn, reps, mode, over, rho, blob)
        g = Swarm(sp, nn, n_replicates=reps, device=device)
        g.set_state(**st)
        s0 = g.stats()
        ms = g.bench_pair_pass(PAIR_ALL_ROWS, iters=10)
        s1 = g.stats()
        ms_sort = g.bench_pair_pass(PAIR_ALL_ROWS | 2, iters=10)
        g.close()
        pairs = reps * nn * (nn - 1)
        pps = pairs / (ms * 1e-3)
        case = {'case': name, 'n_agents': nn, 'replicates': reps, 'box': L, 'pairs_per_launch': pairs, 'ms_per_launch': ms,
                'ms_per_launch_with_sort': ms_sort, 'pairs_per_sec': pps,
                'flop_per_pair': FLOP_PER_PAIR_OPEN, 'achieved': pps * FLOP_PER_PAIR_OPEN / 1e12,
                'frac': pps * FLOP_PER_PAIR_OPEN / 1e12 / peak_fp32,
                'ambiguous_pairs_rechecked_per_launch': (s1['ambiguous_pairs'] - s0['ambiguous_pairs']) // 11}
        if over.get('BC') == 0:
            case['frac_at_21_flop'] = pps * FLOP_PER_PAIR_PERIODIC / 1e12 / peak_fp32
        # issue-slot fraction of SURVEY 8(d): pairs/s * instructions per pair / (128 * SMs * f); the instruction count
        # per pair comes from smsp__thread_inst_executed of the committed ncu capture of this very case
        tag = pair_case_tag(name, nn, reps, rho, blob)
        prof = recorded_profile('r2_ncu_pair_dense_%s.txt' % tag, {'n': nn, 'replicates': reps}, 'pair_dense', ms * 1e-3, tol=0.25)
        if prof.get('used') and prof.get('warp_instructions') and prof.get('threads_per_warp_instruction'):
            ipp = prof['warp_instructions'] * prof['threads_per_warp_instruction'] / pairs
            case['thread_instructions_per_pair'] = ipp
            case['issue_slot_fraction'] = pps * ipp / issue_peak
        if prof.get('used') or 'why' in prof and 'no such' not in prof['why']:
            case['recorded_profile'] = prof
        res['cases'].append(case)
    # headline of this object: the open-domain dense case at the requested N
    res.update({k: res['cases'][0][k] for k in ('n_agents', 'pairs_per_sec', 'achieved', 'frac', 'ms_per_launch')})
    res['note'] = ('every directed pair has its distance evaluated; the zone arithmetic runs only for pairs that '
                   'pass d2 <= att2 (warp-uniform early-out on Hilbert-sorted rows). frac counts 13 FLOP for every '
                   'pair, near or far, periodic or not; the all-near case is the only one where all 13 are executed')
    return res


C5_RHO = 0.01          # BASELINE config C5: N = 1,048,576 in L = 10,240
C5_BYTES_PER_AGENT_STEP = 72   # read + write of the 36-byte persistent state


def c5_state(N, L, seed=SEED):
    from helpers import f32
    rng = np.random.default_rng(seed)
    return dict(x=f32(L * rng.random(N)), y=f32(L * rng.random(N)), phi=f32(2 * np.pi * rng.random(N) - np.pi), vproj=np.ones(N))


def c5_single_gpu(device, hbm_gbs, n_agents, steps=400, warmup=60):
    """Config C5 on ONE GPU: a periodic swarm of n_agents at rho = 0.01 on the cell-list path, `steps` time steps after
    the synchronised first burst, CUDA events on the handle's stream. HBM roofline: 72 B per agent-step."""
    import torch
    from helpers import base_dict
    from sde_burst_coast_b200 import params as PP
    from sde_burst_coast_b200.swarm import Swarm
    N = int(n_agents)
    L = float(np.sqrt(N / C5_RHO))
    sp, _ = PP.make_sbc_params(base_dict('natPred', N, BC=0, size=L, Npred=0), path=PP.PATH_CELLLIST)
    ts = torch.cuda.Stream()
    g = Swarm(sp, N, seed=SEED, device=device, stream=ts.cuda_stream)
    g.set_state(**c5_state(N, L))
    with torch.cuda.stream(ts):
        g.step(0, warmup)
        ts.synchronize()
        s0 = g.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ts)
        g.step(warmup, steps)
        e1.record(ts)
        ts.synchronize()
        ms = e0.elapsed_time(e1)
        s1 = g.stats()
    g.close()
    us = 1e3 * ms / steps
    gbs = C5_BYTES_PER_AGENT_STEP * N / (us * 1e-6) / 1e9
    traffic, prof = None, c5_recorded_profile(N)
    if prof.get('used'):
        traffic = prof['update_bulk_kernel']['dram_bytes_read'] + prof['update_bulk_kernel']['dram_bytes_write']
    return {'workload': 'C5 single periodic swarm, cell list, rho = 0.01', 'n_agents': N, 'box': L, 'steps': steps,
            'us_per_step': us, 'agent_steps_per_sec': N * steps / (ms * 1e-3),
            'pair_candidates_per_sec': (s1['pairs'] - s0['pairs']) / (ms * 1e-3),
            'cell_builds': s1['cell_builds'] - s0['cell_builds'],
            'launches_per_step': sum(s1[k] - s0[k] for k in ('pair_launches', 'update_launches', 'other_launches')) / steps,
            'roofline': {'bound': 'hbm', 'achieved': gbs, 'peak': hbm_gbs, 'unit': 'GB/s', 'frac': gbs / hbm_gbs,
                         'bytes_per_agent_step': C5_BYTES_PER_AGENT_STEP, 'traffic': traffic, 'recorded_profile': prof,
                         'note': 'whole step (pair pass + update + amortised cell build) against the HBM figure of the update alone; '
                                 'traffic = DRAM bytes of ONE launch of the bulk update in the committed ncu capture (cold L2, '
                                 'kernel alone: its 36 MB of stores stay in the L2, hence less than the algorithmic 72 B/agent)'}}


def c5_recorded_profile(n_agents):
    """The committed ncu summaries of the two kernels of a C5 step (profiles/r2_c5_step_breakdown.txt, written by
    profiles/capture_r2.sh c5 + finish_r2.sh), if they are of this swarm size. Inside the graph the two kernels run side by
    side, so this run has no duration of its own to hold against the capture's: the key says so."""
    name = 'r2_c5_step_breakdown.txt'
    out = {'file': 'profiles/' + name, 'used': False}
    try:
        got = read_ncu_summary(name)
        if got is None:
            out['why'] = 'no such summary committed'
            return out
        cfg, launches = got
        out['config'] = cfg
        if str(cfg.get('n_agents')) != str(n_agents):
            out['why'] = 'summary is for another swarm size'
            return out
        for key in ('update_bulk_kernel', 'pair_cells_lanes_kernel'):
            rows = [l for l in launches if key in l['kernel']]
            if not rows:
                out['why'] = 'kernel %s not in the summary' % key
                return out
            r = rows[0]
            out[key] = {'ncu_seconds_per_launch': r.get('gpu__time_duration.sum'),
                        'dram_bytes_read': r.get('dram__bytes_read.sum', 0.0), 'dram_bytes_write': r.get('dram__bytes_write.sum', 0.0),
                        'issue_slot_utilisation': r.get('smsp__issue_active.avg.pct_of_peak_sustained_active'),
                        'registers_per_thread': r.get('launch__registers_per_thread'),
                        'warp_instructions': r.get('smsp__inst_executed.sum')}
        out['used'] = True
        out['check'] = ('none possible from this run: the kernels overlap inside a graph (in-situ timeline: '
                        'profiles/r2_runs/c5_timeline.json)')
    except Exception as e:   # a malformed summary must not cost the bench line
        out['why'] = 'unreadable: %s' % e
    return out


def domain_swarm_leg(rank, world, device, agents_per_gpu, steps=200, warmup=40):
    """The only multi-GPU path that communicates: ONE periodic swarm cut into x-slabs, one rank per GPU, halo /
    migration records stored straight into the neighbours' device memory (CUDA IPC over NVLink, sbc_domain_step).
      (1) parity: a 120,000-agent swarm stepped 130 times decomposed over all ranks vs undecomposed on rank 0 -
          burst timers bit for bit, positions to FP32 summation order (the body of tests/multi_gpu_worker.py) - without
          predator, and with ONE chasing predator (kill rule 3: the closest prey, the sensed-prey count and the centre of
          mass at spawn are all-gathered through the inboxes): the same agents die, the replicated predator is identical;
      (2) weak scaling of config C5: agents_per_gpu agents per GPU (L grows with the GPU count, rho = 0.01), and the
          same with the chasing predator (BASELINE config C5 "plus a variant with one chasing predator")."""
    import torch
    import torch.distributed as dist
    from helpers import base_dict, f32
    from sde_burst_coast_b200 import params as PP
    from sde_burst_coast_b200.domain import SlabSwarm, PeerTransport, step_distributed
    from sde_burst_coast_b200.swarm import Swarm
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    chase = dict(Npred=1, pred_kill=3, pred_move=1, pred_time=0.02, time=50.0, pred_speed0=40.0, kill_rate=400.0, N_confu=500)

    def parity(pred):
        N, L, K = 120000, 3400.0, 130
        sp, npred = PP.make_sbc_params(base_dict('natPred', N, BC=0, size=L, **(chase if pred else dict(Npred=0))), path=PP.PATH_CELLLIST)
        rng = np.random.default_rng(99)
        st = dict(x=f32(L * rng.random(N)), y=f32(L * rng.random(N)), phi=f32(2 * np.pi * rng.random(N) - np.pi), vproj=4.0 * np.ones(N))
        slab = SlabSwarm(sp, N, rank, world, seed=7, device=device, stream=ts.cuda_stream, n_pred=npred)
        slab.scatter_initial(st)
        tr = PeerTransport(slab)
        step_distributed(slab, tr, 0, 40)
        step_distributed(slab, tr, 40, K - 40)
        torch.cuda.synchronize()
        gid, own = slab.owned_state()
        cnt = slab.counts()
        gathered = [None] * world
        dist.gather_object((gid, {k: own[k] for k in ('x', 'y', 'bin_step', 'steps_till_burst')}, cnt, slab.swarm.get_predators()),
                           gathered if rank == 0 else None, dst=0)
        slab.close()
        out = None
        if rank == 0:
            whole = Swarm(sp, N, npred, seed=7, device=device)
            whole.set_state(**st)
            whole.step(0, K)
            ws, wp = whole.get_state(), whole.get_predators()
            whole.close()
            seen = np.zeros(N, dtype=np.int32)
            got = {k: np.zeros(N, dtype=v.dtype) for k, v in gathered[0][1].items()}
            for g_, o_, _c, _p in gathered:
                seen[g_] += 1
                for k, v in o_.items():
                    got[k][g_] = v
            al = ws['alive'] == 1
            dx = np.abs(((got['x'] - ws['x'] + L / 2) % L) - L / 2)[al]
            dy = np.abs(((got['y'] - ws['y'] + L / 2) % L) - L / 2)[al]
            checks = {'every_living_agent_owned_once': bool(np.array_equal(seen == 1, al) and seen.max() <= 1),
                      'steps_till_burst_bit_exact': bool(np.array_equal(got['steps_till_burst'][al], ws['steps_till_burst'][al])),
                      'bin_step_bit_exact': bool(np.array_equal(got['bin_step'][al], ws['bin_step'][al])),
                      'positions_within_1e-3': float(((dx < 1e-3) & (dy < 1e-3)).mean()),
                      'overflow_events': int(sum(c_['overflow_events'] for _, _, c_, _ in gathered))}
            ok = (checks['every_living_agent_owned_once'] and checks['steps_till_burst_bit_exact'] and checks['bin_step_bit_exact'] and
                  checks['positions_within_1e-3'] > 0.995 and checks['overflow_events'] == 0)
            if pred:
                checks['agents_killed'] = int((~al).sum())
                checks['predator_identical_on_all_ranks'] = bool(all(np.array_equal(p_, gathered[0][3]) for _, _, _, p_ in gathered))
                checks['predator_within_2e-3_of_undecomposed'] = bool(np.allclose(gathered[0][3][..., :2], wp[..., :2], atol=2e-3))
                ok = ok and checks['agents_killed'] > 0 and checks['predator_identical_on_all_ranks'] and checks['predator_within_2e-3_of_undecomposed']
            out = ('ok' if ok else 'FAIL', dict(checks, n_agents=N, steps=K, against='undecomposed swarm on rank 0 (same seed, global agent ids)'))
        dist.barrier()
        return out

    def weak(pred):
        N = int(agents_per_gpu) * world
        L = float(np.sqrt(N / C5_RHO))
        sp, npred = PP.make_sbc_params(base_dict('natPred', N, BC=0, size=L, **(chase if pred else dict(Npred=0))), path=PP.PATH_CELLLIST)
        slab = SlabSwarm(sp, N, rank, world, seed=SEED, device=device, stream=ts.cuda_stream, n_pred=npred)
        slab.scatter_initial(c5_state(N, L))
        tr = PeerTransport(slab)
        step_distributed(slab, tr, 0, warmup)        # includes the synchronised first burst at s = 1 (and the predator's spawn at s = 20)
        dist.barrier()
        torch.cuda.synchronize()
        s0 = slab.swarm.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ts)
        step_distributed(slab, tr, warmup, steps)
        e1.record(ts)
        dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        s1 = slab.swarm.stats()
        c = slab.counts()
        t = torch.tensor([ms, float(s1['pairs'] - s0['pairs']), float(c['owned']), float(c['ghosts']), float(c['overflow_events']),
                          float(s1['cell_builds'] - s0['cell_builds'])], dtype=torch.float64, device='cuda')
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        slab.close()
        ms_max = max(float(v[0]) for v in allv)
        return {'n_agents': N, 'box': L, 'steps': steps, 'value': N * steps / (ms_max * 1e-3), 'unit': 'agent-steps/s',
                'us_per_step': 1e3 * ms_max / steps, 'us_per_step_per_rank': [1e3 * float(v[0]) / steps for v in allv],
                'pair_candidates_per_sec': sum(float(v[1]) for v in allv) / (ms_max * 1e-3),
                'owned_per_rank': [int(v[2]) for v in allv], 'ghosts_per_rank': [int(v[3]) for v in allv],
                'overflow_events': int(sum(float(v[4]) for v in allv)), 'full_exchanges': int(allv[0][5]),
                'launches_per_step': sum(s1[k] - s0[k] for k in ('pair_launches', 'update_launches', 'other_launches')) / steps}

    res = {}
    p0, p1 = parity(False), parity(True)
    w0 = weak(False)
    w1 = weak(True)
    torch.cuda.set_stream(torch.cuda.default_stream())
    if rank != 0:
        return None
    res['parity'] = 'ok' if (p0[0] == 'ok' and p1[0] == 'ok') else 'FAIL'
    res['parity_detail'] = p0[1]
    res['parity_detail_with_predator'] = p1[1]
    res.update({'workload': 'C5 single periodic swarm decomposed into x-slabs, peer-memory halo exchange (no NCCL in the step loop)',
                'scaling': 'weak', 'n_gpus': world, 'agents_per_gpu': int(agents_per_gpu)})
    res.update(w0)
    res['with_chasing_predator'] = dict(w1, note='one predator replicated on every rank (natPred, chase, kill rule 3): per step an all-gather of '
                                                 'the ranks\' closest prey and one of (sensed prey, weight sum), 64 bytes per rank each, through the inboxes')
    return res


if __name__ == '__main__':
    main()
