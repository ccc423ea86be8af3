"""ctypes binding of libsbc.so (include/sbc.h) - the call a Python user of the hot path makes.

`Swarm` mirrors the bracket of the reference's main loop (src/swarmdyn.cpp:64-91): create the
agents (`InitSystem`), give them a state (`ResetSystem`), call `step` where the reference calls
`split_dead` + `Step`, and read particles/predators where it calls `Output`.

The library is CUDA only. Importing this module without the built library, or creating a Swarm
without a GPU, raises - there is no CPU path behind it.
"""
import ctypes
import os

import numpy as np

from .params import SbcParams, PAIR_ACTIVE_ROWS, PAIR_ALL_ROWS  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SBC_LIB_PATH') or os.path.join(_HERE, 'lib', 'libsbc.so')   # (override: A/B runs of two builds)

_dp = ctypes.POINTER(ctypes.c_double)
_fp = ctypes.POINTER(ctypes.c_float)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_u64p = ctypes.POINTER(ctypes.c_uint64)
_vp = ctypes.c_void_p

# every symbol include/sbc.h declares: (name, restype, argtypes)
ABI = [
    ('sbc_abi_version', ctypes.c_int, []),
    ('sbc_create', ctypes.c_int, [ctypes.POINTER(SbcParams), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(_vp)]),
    ('sbc_destroy', ctypes.c_int, [_vp]),
    ('sbc_last_error', ctypes.c_char_p, [_vp]),
    ('sbc_set_stream', ctypes.c_int, [_vp, _vp]),
    ('sbc_synchronize', ctypes.c_int, [_vp]),
    ('sbc_set_state', ctypes.c_int, [_vp] + [_dp] * 8 + [_u32p, _u32p, _u8p]),
    ('sbc_get_state', ctypes.c_int, [_vp] + [_dp] * 8 + [_u32p, _u32p, _u8p]),
    ('sbc_get_particles', ctypes.c_int, [_vp, _dp, _u8p]),
    ('sbc_set_state_f32', ctypes.c_int, [_vp] + [_fp] * 8 + [_u32p, _u32p, _u8p]),
    ('sbc_get_particles_f32', ctypes.c_int, [_vp, _fp, _u8p]),
    ('sbc_get_predators', ctypes.c_int, [_vp, _dp]),
    ('sbc_set_predators', ctypes.c_int, [_vp, _dp, _dp, _dp, ctypes.c_int]),
    ('sbc_get_counts', ctypes.c_int, [_vp, _i32p, _i32p]),
    ('sbc_step', ctypes.c_int, [_vp, ctypes.c_int64, ctypes.c_int64]),
    ('sbc_init_state', ctypes.c_int, [_vp, ctypes.c_int]),
    ('sbc_observables', ctypes.c_int, [_vp, _dp]),
    ('sbc_out_swarm', ctypes.c_int, [_vp, ctypes.c_int, _dp, _dp]),
    ('sbc_out_swarm_fishnet', ctypes.c_int, [_vp, _dp]),
    ('sbc_pair_interactions', ctypes.c_int, [_vp, ctypes.c_int, _i32p, _i32p, _i32p, _dp, _dp, _dp,
                                             _i64p, _i32p, ctypes.c_int64]),
    ('sbc_inject_decisions', ctypes.c_int, [_vp, _dp, _dp, _u32p]),
    ('sbc_philox_decisions', ctypes.c_int, [_vp, ctypes.c_int64, _dp, _dp, _u32p]),
    ('sbc_philox_raw', ctypes.c_int, [_u32p, _u32p, _u32p]),
    ('sbc_debug_flags', ctypes.c_int, [_vp, ctypes.c_int, _u8p]),
    ('sbc_debug_force_multikernel', ctypes.c_int, [_vp, ctypes.c_int]),
    ('sbc_get_stats', ctypes.c_int, [_vp, _u64p]),
    ('sbc_debug_trace', ctypes.c_int, [_vp, ctypes.c_int, _u64p]),
    ('sbc_bench_pair_pass', ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]),
    ('sbc_set_global_ids', ctypes.c_int, [_vp, _u32p]),
    ('sbc_get_global_ids', ctypes.c_int, [_vp, _u32p]),
    ('sbc_domain_config', ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    ('sbc_domain_buffer_bytes', ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    ('sbc_domain_pack', ctypes.c_int, [_vp, _vp, _vp]),
    ('sbc_domain_unpack', ctypes.c_int, [_vp, _vp, _vp, _vp, _vp]),
    ('sbc_domain_counts', ctypes.c_int, [_vp, _i32p, _i32p, _i32p]),
    ('sbc_domain_rebuild_period', ctypes.c_int, [_vp, ctypes.c_int]),
    ('sbc_domain_inbox', ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t), _vp]),
    ('sbc_domain_open_peer', ctypes.c_int, [_vp, _vp, ctypes.POINTER(ctypes.c_void_p)]),
    ('sbc_domain_attach_peers', ctypes.c_int, [_vp, _vp, _vp]),
    ('sbc_domain_attach_world', ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]),
    ('sbc_domain_post', ctypes.c_int, [_vp]),
    ('sbc_domain_complete', ctypes.c_int, [_vp]),
    ('sbc_domain_step', ctypes.c_int, [_vp, ctypes.c_int64, ctypes.c_int64]),
    ('sbc_domain_set_timeout', ctypes.c_int, [_vp, ctypes.c_double]),
]

_lib = None


class SbcError(RuntimeError):
    pass


def load_library(path=None):
    """dlopen libsbc.so and bind every ABI symbol; raises if the library or a symbol is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise SbcError('libsbc.so not built at %s - run `python -c "import __graft_entry__ as g; '
                       'g.build()"` (nvcc, sm_100a); there is no CPU fallback' % path)
    lib = ctypes.CDLL(path)
    for name, res, args in ABI:
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


STATE_FIELDS = ('x', 'y', 'vx', 'vy', 'phi', 'vproj', 'fx', 'fy')


class Swarm:
    """`n_replicates` independent swarms of `n_agents` agents on one GPU."""

    def __init__(self, sp, n_agents, n_pred=0, n_replicates=1, seed=0, replicate_offset=0,
                 device=0, stream=None):
        self.lib = load_library()
        self.sp = sp
        self.N, self.Npred, self.R = int(n_agents), int(n_pred), int(n_replicates)
        h = _vp()
        rc = self.lib.sbc_create(ctypes.byref(sp), self.N, self.Npred, self.R, seed,
                                 replicate_offset, device, ctypes.byref(h))
        if rc != 0:
            raise SbcError('sbc_create failed (%d): %s' % (rc, self.lib.sbc_last_error(None).decode()))
        self.h = h
        if stream is not None:
            self._check(self.lib.sbc_set_stream(self.h, _vp(stream)))

    def _check(self, rc):
        if rc != 0:
            raise SbcError('libsbc error %d: %s' % (rc, self.lib.sbc_last_error(self.h).decode()))

    def close(self):
        if getattr(self, 'h', None):
            self.lib.sbc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state ------------------------------------------------------------------------------
    def _shape(self, a, dtype):
        a = np.ascontiguousarray(a, dtype=dtype).reshape(-1)
        assert a.size == self.R * self.N, 'state arrays are [n_replicates][n_agents]'
        return a

    def set_state(self, **kw):
        args, keep = [], []
        for f in STATE_FIELDS:
            a = kw.get(f)
            if a is not None:
                a = self._shape(a, np.float64)
                keep.append(a)
            args.append(_ptr(a, _dp))
        for f in ('bin_step', 'steps_till_burst'):
            a = kw.get(f)
            if a is not None:
                a = self._shape(a, np.uint32)
                keep.append(a)
            args.append(_ptr(a, _u32p))
        a = kw.get('alive')
        if a is not None:
            a = self._shape(a, np.uint8)
            keep.append(a)
        args.append(_ptr(a, _u8p))
        self._check(self.lib.sbc_set_state(self.h, *args))

    def set_state_f32(self, **kw):
        """The same from float32 arrays (sbc_set_state_f32): no widening on the host, half the bytes on the wire."""
        args, keep = [], []
        for f in STATE_FIELDS:
            a = kw.get(f)
            if a is not None:
                a = self._shape(a, np.float32)
                keep.append(a)
            args.append(_ptr(a, _fp))
        for f in ('bin_step', 'steps_till_burst'):
            a = kw.get(f)
            if a is not None:
                a = self._shape(a, np.uint32)
                keep.append(a)
            args.append(_ptr(a, _u32p))
        a = kw.get('alive')
        if a is not None:
            a = self._shape(a, np.uint8)
            keep.append(a)
        args.append(_ptr(a, _u8p))
        self._check(self.lib.sbc_set_state_f32(self.h, *args))

    def get_particles_f32(self):
        out = np.zeros((self.R, self.N, 7), dtype=np.float32)
        alive = np.zeros((self.R, self.N), dtype=np.uint8)
        self._check(self.lib.sbc_get_particles_f32(self.h, _ptr(out, _fp), _ptr(alive, _u8p)))
        return out, alive

    def get_state(self):
        n = self.R * self.N
        out = {f: np.zeros(n) for f in STATE_FIELDS}
        out['bin_step'] = np.zeros(n, dtype=np.uint32)
        out['steps_till_burst'] = np.zeros(n, dtype=np.uint32)
        out['alive'] = np.zeros(n, dtype=np.uint8)
        self._check(self.lib.sbc_get_state(self.h, *[_ptr(out[f], _dp) for f in STATE_FIELDS],
                                           _ptr(out['bin_step'], _u32p),
                                           _ptr(out['steps_till_burst'], _u32p),
                                           _ptr(out['alive'], _u8p)))
        if self.R > 1:
            out = {k: v.reshape(self.R, self.N) for k, v in out.items()}
        return out

    def get_particles(self):
        out = np.zeros((self.R, self.N, 7))
        alive = np.zeros((self.R, self.N), dtype=np.uint8)
        self._check(self.lib.sbc_get_particles(self.h, _ptr(out, _dp), _ptr(alive, _u8p)))
        return out, alive

    def get_predators(self):
        out = np.zeros((self.R, self.Npred, 6))
        if self.Npred:
            self._check(self.lib.sbc_get_predators(self.h, _ptr(out, _dp)))
        return out

    def set_predators(self, x, y, phi, spawned=1):
        x, y, phi = [np.ascontiguousarray(a, dtype=np.float64).reshape(-1) for a in (x, y, phi)]
        assert x.size == self.R * self.Npred
        self._check(self.lib.sbc_set_predators(self.h, _ptr(x, _dp), _ptr(y, _dp), _ptr(phi, _dp), spawned))

    def counts(self):
        a = np.zeros(self.R, dtype=np.int32)
        d = np.zeros(self.R, dtype=np.int32)
        self._check(self.lib.sbc_get_counts(self.h, _ptr(a, _i32p), _ptr(d, _i32p)))
        return a, d

    # ---- hot path ---------------------------------------------------------------------------
    def step(self, s_first, n_steps=1):
        self._check(self.lib.sbc_step(self.h, int(s_first), int(n_steps)))

    def synchronize(self):
        self._check(self.lib.sbc_synchronize(self.h))

    def observables(self):
        out = np.zeros((self.R, 8))
        self._check(self.lib.sbc_observables(self.h, _ptr(out, _dp)))
        return out

    def init_state(self, ic):
        """Initial condition `ic` of the reference's ResetSystem (settings.cpp:196-387) generated on the device for
        every replicate (Philox protocol; IC 99 is a file and stays on the host)."""
        self._check(self.lib.sbc_init_state(self.h, int(ic)))

    def out_swarm(self, no_pairs=False, want_extra=False):
        """The reference's /swarm row of every replicate (Out_swarm, swarmdyn.cpp:237-344), computed on the device.
        Returns out[R][16], or (out, extra[R][4]) - see include/sbc.h for the columns."""
        out = np.zeros((self.R, 16))
        extra = np.zeros((self.R, 4)) if want_extra else None
        self._check(self.lib.sbc_out_swarm(self.h, 1 if no_pairs else 0, _ptr(out, _dp), _ptr(extra, _dp)))
        return (out, extra) if want_extra else out

    def out_swarm_fishnet(self):
        """The /swarm_fishNet row of every replicate (Out_swarm_fishNet, swarmdyn.cpp:358-410): out[R][4]."""
        out = np.zeros((self.R, 4))
        self._check(self.lib.sbc_out_swarm_fishnet(self.h, _ptr(out, _dp)))
        return out

    # ---- test hooks -------------------------------------------------------------------------
    def pair_interactions(self, flags=PAIR_ACTIVE_ROWS, want_nn=True):
        n = self.R * self.N
        c = [np.zeros(n, dtype=np.int32) for _ in range(3)]
        f = [np.zeros((n, 2)) for _ in range(3)]
        nn_ptr = np.zeros(n + 1, dtype=np.int64)
        cap = n * max(self.N - 1, 1) if want_nn else 0
        nn_idx = np.zeros(max(cap, 1), dtype=np.int32)
        self._check(self.lib.sbc_pair_interactions(
            self.h, flags, *[_ptr(a, _i32p) for a in c], *[_ptr(a, _dp) for a in f],
            _ptr(nn_ptr, _i64p), _ptr(nn_idx, _i32p) if want_nn else None, cap))
        return dict(c_rep=c[0], c_alg=c[1], c_att=c[2], f_rep=f[0], f_alg=f[1], f_att=f[2],
                    nn_ptr=nn_ptr, nn_idx=nn_idx[:nn_ptr[-1]] if want_nn else None)

    def inject(self, u_social=None, u_dir=None, k_next=None):
        us = None if u_social is None else self._shape(u_social, np.float64)
        ud = None if u_dir is None else self._shape(u_dir, np.float64)
        kn = None if k_next is None else self._shape(k_next, np.uint32)
        self._check(self.lib.sbc_inject_decisions(self.h, _ptr(us, _dp), _ptr(ud, _dp), _ptr(kn, _u32p)))

    def philox_decisions(self, step):
        n = self.R * self.N
        us, ud = np.zeros(n), np.zeros(n)
        kn = np.zeros(n, dtype=np.uint32)
        self._check(self.lib.sbc_philox_decisions(self.h, int(step), _ptr(us, _dp), _ptr(ud, _dp),
                                                  _ptr(kn, _u32p)))
        return us, ud, kn

    def debug_flags(self, enable=True, read=False):
        out = np.zeros(self.R * self.N, dtype=np.uint8) if read else None
        self._check(self.lib.sbc_debug_flags(self.h, 1 if enable else 0, _ptr(out, _u8p)))
        return out

    def force_multikernel(self, on=True):
        """True / 1: multi-kernel path; 2: one-block kernel instead of the cluster kernel; False / 0: automatic."""
        self._check(self.lib.sbc_debug_force_multikernel(self.h, int(on)))

    def trace(self, enable=True, read=False):
        """Timeline hook (sbc_debug_trace): returns uint64 [1024][8][2] {start, end} ns by (step mod 1024, kernel) when read."""
        out = np.zeros((1024, 8, 2), dtype=np.uint64) if read else None
        self._check(self.lib.sbc_debug_trace(self.h, 1 if enable else 0, _ptr(out, _u64p)))
        return out

    def stats(self):
        out = np.zeros(8, dtype=np.uint64)
        self._check(self.lib.sbc_get_stats(self.h, _ptr(out, _u64p)))
        keys = ('pair_launches', 'update_launches', 'block_launches', 'pairs', 'ambiguous_pairs',
                'agent_steps', 'cell_builds', 'other_launches')
        return dict(zip(keys, [int(v) for v in out]))

    def bench_pair_pass(self, flags=PAIR_ALL_ROWS, iters=10):
        ms = ctypes.c_float()
        self._check(self.lib.sbc_bench_pair_pass(self.h, flags, iters, ctypes.byref(ms)))
        return ms.value
