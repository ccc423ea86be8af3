"""Host side of the spatial domain decomposition of ONE large periodic swarm (BASELINE config C5):
slabs along x, one libsbc handle per GPU, per-step halo exchange and agent migration.

The device work (selecting edge agents, building / applying 64-byte records) is in libsbc.so
(csrc/domain.cu, entry points sbc_domain_* of include/sbc.h). This module owns the routing:
which rank is left / right, who gets which buffer, and the transport - `DistTransport` moves the
fixed-size record buffers with torch.distributed point-to-point operations (NCCL over NVLink between
GPUs; the same code runs on CPU tensors with gloo, which is how the routing is tested without GPUs),
`LoopbackTransport` wires several slabs inside one process (single-GPU emulation in the tests).
`PeerTransport` is the fast path between GPUs: it only runs at set-up (the ranks swap the CUDA IPC handles
of their inboxes); afterwards libsbc's own kernels store the records into the neighbours' device memory
over NVLink and no message-library call is made per step (include/sbc.h, sbc_domain_post / _complete).

Nothing here computes the hot path; the reference has no counterpart (it is a single process).
"""
import ctypes

import numpy as np

RECORD_BYTES = 64


def slab_bounds(L, world):
    """[x_lo, x_hi) of every rank; the last slab ends exactly at L."""
    w = L / world
    return [(w * r, L if r == world - 1 else w * (r + 1)) for r in range(world)]


def owner_of(x, L, world):
    """Rank owning position x (array) in a periodic box of length L."""
    r = np.floor(np.asarray(x, dtype=np.float64) / (L / world)).astype(np.int64)
    return np.clip(r, 0, world - 1)


def neighbours(rank, world):
    return (rank - 1) % world, (rank + 1) % world


def check_geometry(L, att_range, world):
    """Slabs must be at least one halo wide (two when both neighbours are the same rank)."""
    w, halo = L / world, att_range * 1.0001
    if world > 1 and w < (2.1 * halo if world == 2 else 1.05 * halo):
        raise ValueError('slab width %.3f is narrower than the halo requirement for att_range %.3f' % (w, att_range))


class DistTransport:
    """Exchange with the two ring neighbours through torch.distributed (nccl for CUDA tensors, gloo for CPU)."""

    def __init__(self, rank, world, group=None):
        import torch.distributed as dist
        self.dist, self.rank, self.world, self.group = dist, rank, world, group

    def exchange(self, send_left, send_right, recv_from_left, recv_from_right):
        """send_left -> left neighbour (arrives there as data from its right), send_right -> right neighbour."""
        d = self.dist
        left, right = neighbours(self.rank, self.world)
        if self.world == 1:
            recv_from_right.copy_(send_left)
            recv_from_left.copy_(send_right)
            return
        if self.world == 2:
            # both neighbours are the same peer: one ordered pair of messages each way. What I send to
            # my left arrives as the peer's "from right" and vice versa; the peer posts the same order.
            ops = [d.P2POp(d.isend, send_left, left, self.group), d.P2POp(d.isend, send_right, right, self.group),
                   d.P2POp(d.irecv, recv_from_right, right, self.group), d.P2POp(d.irecv, recv_from_left, left, self.group)]
        else:
            ops = [d.P2POp(d.isend, send_left, left, self.group), d.P2POp(d.irecv, recv_from_right, right, self.group),
                   d.P2POp(d.isend, send_right, right, self.group), d.P2POp(d.irecv, recv_from_left, left, self.group)]
        for req in d.batch_isend_irecv(ops):
            req.wait()

    def min_int(self, v):
        """Minimum of one integer over the ranks (the rebuild period all slabs can afford)."""
        if self.world == 1:
            return int(v)
        import torch
        dev = 'cuda' if self.dist.get_backend(self.group) == 'nccl' else 'cpu'
        t = torch.tensor([int(v)], dtype=torch.int32, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
        return int(t.item())


class PeerTransport:
    """Set-up of the peer-memory exchange: every rank exports its inbox (CUDA IPC), the handles travel once
    through torch.distributed, each rank maps the inboxes of its two ring neighbours. The per-step traffic
    then never touches this class (sbc_domain_step)."""

    def __init__(self, slab, group=None):
        import torch
        import torch.distributed as dist
        self.dist, self.rank, self.world, self.group = dist, slab.rank, slab.world, group
        lib, h = slab.swarm.lib, slab.swarm.h
        base, nbytes = ctypes.c_void_p(), ctypes.c_size_t()
        handle = (ctypes.c_ubyte * 64)()
        slab.swarm._check(lib.sbc_domain_inbox(h, ctypes.byref(base), ctypes.byref(nbytes), handle))
        if self.world == 1:
            slab.swarm._check(lib.sbc_domain_attach_peers(h, base, base))
            if slab.n_pred > 0:
                slab.swarm._check(lib.sbc_domain_attach_world(h, (ctypes.c_void_p * 1)(base), 1))
            return
        dev = 'cuda' if dist.get_backend(group) == 'nccl' else 'cpu'
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
        every = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(every, mine, group=group)
        left, right = neighbours(self.rank, self.world)
        mapped = {}
        # the halo exchange talks to the two ring neighbours; a predator's all-gathers (closest prey, kill sums, centre
        # of mass) to everybody: with predators every rank maps every inbox
        for peer in (set(range(self.world)) - {self.rank}) if slab.n_pred > 0 else {left, right}:
            raw = (ctypes.c_ubyte * 64)(*every[peer].cpu().tolist())
            p = ctypes.c_void_p()
            slab.swarm._check(lib.sbc_domain_open_peer(h, raw, ctypes.byref(p)))
            mapped[peer] = p
        mapped[self.rank] = base
        slab.swarm._check(lib.sbc_domain_attach_peers(h, mapped[left], mapped[right]))
        if slab.n_pred > 0:
            arr = (ctypes.c_void_p * self.world)(*[mapped[r] for r in range(self.world)])
            slab.swarm._check(lib.sbc_domain_attach_world(h, arr, self.world))
        dist.barrier(group=group)   # every inbox is mapped before anybody pushes

    def min_int(self, v):
        return DistTransport.min_int(self, v)


class LoopbackTransport:
    """All slabs live in this process: collect every rank's send buffers, then deliver."""

    def __init__(self, world):
        self.world = world
        self.pending = {}

    def post(self, rank, send_left, send_right):
        self.pending[rank] = (send_left, send_right)

    def deliver(self, rank, recv_from_left, recv_from_right):
        left, right = neighbours(rank, self.world)
        recv_from_left.copy_(self.pending[left][1])    # left neighbour's "send right"
        recv_from_right.copy_(self.pending[right][0])  # right neighbour's "send left"


class SlabSwarm:
    """One slab of the decomposed swarm on one GPU."""

    def __init__(self, sp, n_global, rank, world, seed=0, device=0, stream=None, own_capacity=None,
                 ghost_capacity=None, max_records=None, n_pred=0):
        import torch
        from .swarm import Swarm
        from . import params as P
        check_geometry(sp.sizeL, sp.att_range, world)
        self.torch = torch
        self.rank, self.world, self.n_global = rank, world, int(n_global)
        mean_own = n_global / world
        halo_frac = min(1.0, sp.att_range * 1.2001 / (sp.sizeL / world))   # band = att_range + skin (<= att_range / 5)
        self.own_cap = int(own_capacity or (mean_own * 1.05 + 4096))   # slack for density fluctuations; every slot costs update and sort time
        self.max_halo = int(max_records or (mean_own * halo_frac * 2 + 1024))      # records per edge band
        self.max_mig = int(max(256, self.max_halo // 8))                            # migrants per edge per step
        self.ghost_cap = 0 if world == 1 else int(ghost_capacity or (2 * self.max_halo + 2 * self.max_mig))
        sp2 = type(sp).from_buffer_copy(sp)
        sp2.path = P.PATH_CELLLIST
        self.sp = sp2
        self.dev = torch.device('cuda', device)
        self.stream = stream   # cudaStream_t (int) the handle's kernels run on; None = the handle's private stream
        self.n_pred = int(n_pred)   # replicated on every rank (sbc_domain_attach_world)
        self.swarm = Swarm(sp2, self.own_cap + self.ghost_cap, self.n_pred, 1, seed=seed, device=device, stream=stream)
        lib, h = self.swarm.lib, self.swarm.h
        self.swarm._check(lib.sbc_domain_config(h, rank, world, self.own_cap, self.max_mig, self.max_halo))
        nbytes = lib.sbc_domain_buffer_bytes(self.max_mig, self.max_halo)
        mk = lambda: torch.zeros(nbytes, dtype=torch.uint8, device=self.dev)
        self.send_l, self.send_r, self.recv_l, self.recv_r = mk(), mk(), mk(), mk()
        self.primed = False   # ghosts of the current positions are in place

    # ---- state in / out (host arrays indexed by GLOBAL agent id) -------------------------------
    def scatter_initial(self, state):
        """Take this rank's agents out of global arrays (x, y, phi, vproj, optional timers...)."""
        N = self.swarm.N
        mine = np.where(owner_of(state['x'], self.sp.sizeL, self.world) == self.rank)[0]
        if mine.size > self.own_cap:
            raise ValueError('own_capacity %d < %d agents of rank %d' % (self.own_cap, mine.size, self.rank))
        local = {}
        for k, v in state.items():
            a = np.zeros(N, dtype=np.asarray(v).dtype)
            a[:mine.size] = np.asarray(v)[mine]
            local[k] = a
        alive = np.zeros(N, dtype=np.uint8)
        alive[:mine.size] = 1
        gid = np.full(N, 0xFFFFFFFF, dtype=np.uint32)
        gid[:mine.size] = mine.astype(np.uint32)
        self.swarm.set_state(alive=alive, **local)
        u32p = ctypes.POINTER(ctypes.c_uint32)
        self.swarm._check(self.swarm.lib.sbc_set_global_ids(self.swarm.h, gid.ctypes.data_as(u32p)))
        self.primed = False

    def local_rebuild_period(self):
        """Refresh exchanges this rank's agents allow between two full exchanges (sbc_domain_rebuild_period)."""
        v = self.swarm.lib.sbc_domain_rebuild_period(self.swarm.h, -1)
        if v < 0:
            self.swarm._check(v)
        return int(v)

    def set_rebuild_period(self, period):
        """The number every rank agreed on (the minimum over the ranks)."""
        v = self.swarm.lib.sbc_domain_rebuild_period(self.swarm.h, int(period))
        if v < 0:
            self.swarm._check(v)
        self.primed = False
        return int(v)

    def owned_state(self):
        """(gids, state dict) of the agents this rank owns now."""
        st = self.swarm.get_state()
        gid = np.zeros(self.swarm.N, dtype=np.uint32)
        u32p = ctypes.POINTER(ctypes.c_uint32)
        self.swarm._check(self.swarm.lib.sbc_get_global_ids(self.swarm.h, gid.ctypes.data_as(u32p)))
        own = np.where(st['alive'][:self.own_cap] == 1)[0]
        return gid[own], {k: v[own] for k, v in st.items()}

    def counts(self):
        a, b, c = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        self.swarm._check(self.swarm.lib.sbc_domain_counts(self.swarm.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return dict(owned=a.value, ghosts=b.value, overflow_events=c.value)

    # ---- the exchange: once before the first step, then after every step -------------------------
    def _p(self, t):
        return ctypes.c_void_p(t.data_ptr())

    def pack(self):
        """Classify the owned agents and fill both send buffers ([header | migrants | halo])."""
        self.swarm._check(self.swarm.lib.sbc_domain_pack(self.swarm.h, self._p(self.send_l), self._p(self.send_r)))

    def unpack(self):
        """Insert the incoming migrants, rebuild the ghosts (incoming halo + my own outgoing migrants)."""
        self.swarm._check(self.swarm.lib.sbc_domain_unpack(self.swarm.h, self._p(self.recv_l), self._p(self.recv_r),
                                                           self._p(self.send_l), self._p(self.send_r)))

    def step_local(self, s):
        self.swarm.step(s, 1)

    def close(self):
        self.swarm.close()


def step_distributed(slab, transport, s_first, n_steps):
    """Advance one rank's slab (world > 1: a DistTransport moves the buffers between ranks). One exchange per
    step: [pack -> exchange -> unpack] once before the first step, then after every step. With a
    PeerTransport the whole loop runs inside libsbc (sbc_domain_step): the records go straight into the
    neighbours' device memory."""
    if isinstance(transport, PeerTransport):
        lib, h = slab.swarm.lib, slab.swarm.h
        if not slab.primed:
            slab.set_rebuild_period(transport.min_int(slab.local_rebuild_period()))
            if slab.world > 1:
                slab.swarm._check(lib.sbc_domain_post(h))
                slab.swarm._check(lib.sbc_domain_complete(h))
            slab.primed = True
        if slab.world > 1:
            slab.swarm._check(lib.sbc_domain_step(h, int(s_first), int(n_steps)))
        else:
            slab.swarm.step(s_first, n_steps)
        return

    # pack / unpack run on the handle's stream, the message library on torch's current stream. When the two are the
    # same stream (the fast set-up: torch.cuda.set_stream(ts); SlabSwarm(..., stream=ts.cuda_stream)) stream order does
    # the job; otherwise the host orders them: the buffers are complete before they are sent and landed before they
    # are unpacked, and the next pack cannot overwrite a buffer still in flight.
    torch = slab.torch
    is_cuda = slab.send_l.is_cuda
    shared = is_cuda and slab.stream is not None and torch.cuda.current_stream(slab.dev).cuda_stream == slab.stream

    def exchange():
        slab.pack()
        if is_cuda and not shared:
            slab.swarm.synchronize()
        transport.exchange(slab.send_l, slab.send_r, slab.recv_l, slab.recv_r)
        if is_cuda and not shared:
            torch.cuda.current_stream(slab.dev).synchronize()
        slab.unpack()
    if not slab.primed:
        slab.set_rebuild_period(transport.min_int(slab.local_rebuild_period()))
        if slab.world > 1:
            exchange()
        slab.primed = True
    for s in range(s_first, s_first + n_steps):
        slab.step_local(s)
        if slab.world > 1:
            exchange()


def attach_loopback_peers(slabs):
    """Slabs of one process on one GPU: the neighbours' inboxes are ordinary device pointers."""
    world = len(slabs)
    bases = []
    for sl in slabs:
        base, nbytes = ctypes.c_void_p(), ctypes.c_size_t()
        sl.swarm._check(sl.swarm.lib.sbc_domain_inbox(sl.swarm.h, ctypes.byref(base), ctypes.byref(nbytes), None))
        bases.append(base)
    for sl in slabs:
        left, right = neighbours(sl.rank, world)
        sl.swarm._check(sl.swarm.lib.sbc_domain_attach_peers(sl.swarm.h, bases[left], bases[right]))
        if sl.n_pred > 0:   # (the slabs then need their own streams: a predator's all-gather waits inside a kernel)
            sl.swarm._check(sl.swarm.lib.sbc_domain_attach_world(sl.swarm.h, (ctypes.c_void_p * world)(*bases), world))


def _step_all(slabs, s):
    """One local step of every slab. With predators a slab's step waits INSIDE its kernels for the other slabs' records
    (and libsbc may synchronise the slab's stream), so the slabs are driven by one host thread each, as one process
    per GPU would."""
    if not any(sl.n_pred > 0 for sl in slabs):
        for sl in slabs:
            sl.step_local(s)
        return
    import threading
    err = []

    def run(sl):
        try:
            sl.step_local(s)
        except Exception as e:   # noqa: BLE001
            err.append(e)
    th = [threading.Thread(target=run, args=(sl,)) for sl in slabs]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if err:
        raise err[0]


def step_loopback(slabs, s_first, n_steps, peer=False):
    """Advance all slabs held in this process in lock step (tests / single-GPU emulation). peer=True uses the
    peer-memory exchange (after attach_loopback_peers): every slab posts, then every slab completes - all on
    one stream, so no wait kernel ever spins."""
    world = len(slabs)
    tr = LoopbackTransport(world)
    if peer:
        def exchange():
            for sl in slabs:
                sl.swarm._check(sl.swarm.lib.sbc_domain_post(sl.swarm.h))
            for sl in slabs:
                sl.swarm._check(sl.swarm.lib.sbc_domain_complete(sl.swarm.h))
        if not all(sl.primed for sl in slabs):
            period = min(sl.local_rebuild_period() for sl in slabs)
            for sl in slabs:
                sl.set_rebuild_period(period)
            if world > 1:
                exchange()
            for sl in slabs:
                sl.primed = True
        for s in range(s_first, s_first + n_steps):
            _step_all(slabs, s)
            if world > 1:
                exchange()
        return

    def exchange():
        for sl in slabs:
            sl.pack()
            tr.post(sl.rank, sl.send_l, sl.send_r)
        slabs[0].torch.cuda.synchronize()
        for sl in slabs:
            tr.deliver(sl.rank, sl.recv_l, sl.recv_r)
            sl.unpack()
    if not all(sl.primed for sl in slabs):
        period = min(sl.local_rebuild_period() for sl in slabs)
        for sl in slabs:
            sl.set_rebuild_period(period)
        if world > 1:
            exchange()
        for sl in slabs:
            sl.primed = True
    for s in range(s_first, s_first + n_steps):
        _step_all(slabs, s)
        if world > 1:
            exchange()


def gather_global(slabs, n_global):
    """Global arrays (by agent id) from slabs held in this process."""
    out = None
    seen = np.zeros(n_global, dtype=np.int32)
    for sl in slabs:
        gid, st = sl.owned_state()
        if out is None:
            out = {k: np.zeros(n_global, dtype=v.dtype) for k, v in st.items()}
        for k, v in st.items():
            out[k][gid] = v
        seen[gid] += 1
    return out, seen
