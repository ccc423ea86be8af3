"""ctypes binding of the CPU oracles (TEST INFRASTRUCTURE ONLY - see oracle/sbc_oracle.h).

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU legs only; nothing under
sde_burst_coast_b200/ may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

from sde_burst_coast_b200.params import SbcParams

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_PATH = os.path.join(_HERE, 'liborc_port.so')
REF_PATH = os.path.join(_HERE, '_ref', 'liborc_ref.so')

_dp = ctypes.POINTER(ctypes.c_double)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build(ref=True):
    """Compile the oracle libraries (port always; reference build when /root/reference exists)."""
    subprocess.run(['make', '-C', _HERE, 'port'], check=True, stdout=subprocess.DEVNULL)
    if ref and os.path.isdir('/root/reference/src'):
        subprocess.run(['make', '-C', _HERE, '-j4', 'ref'], check=True, stdout=subprocess.DEVNULL)


def _ptr(arr, typ):
    if arr is None:
        return None
    return arr.ctypes.data_as(typ)


class OracleLib:
    def __init__(self, path):
        self.lib = L = ctypes.CDLL(path)
        L.orc_kind.restype = ctypes.c_char_p
        L.orc_create.restype = ctypes.c_void_p
        L.orc_create.argtypes = [ctypes.POINTER(SbcParams), ctypes.c_int, ctypes.c_int,
                                 ctypes.c_uint64, ctypes.c_uint64]
        L.orc_destroy.argtypes = [ctypes.c_void_p]
        L.orc_set_state.argtypes = [ctypes.c_void_p] + [_dp] * 8 + [_u32p, _u32p, _u8p]
        L.orc_get_state.argtypes = [ctypes.c_void_p] + [_dp] * 8 + [_u32p, _u32p, _u8p]
        L.orc_set_predators.argtypes = [ctypes.c_void_p, _dp, _dp, _dp, ctypes.c_int]
        L.orc_get_predators.argtypes = [ctypes.c_void_p, _dp]
        L.orc_get_counts.argtypes = [ctypes.c_void_p, _i32p, _i32p]
        L.orc_pair_interactions.restype = ctypes.c_int64
        L.orc_pair_interactions.argtypes = [ctypes.c_void_p, ctypes.c_int, _i32p, _i32p, _i32p,
                                            _dp, _dp, _dp, _i64p, _i32p, ctypes.c_int64]
        L.orc_inject_decisions.argtypes = [ctypes.c_void_p, _dp, _dp, _u32p]
        L.orc_step.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, _u8p]
        L.orc_philox_decisions.argtypes = [ctypes.c_void_p, ctypes.c_int64, _dp, _dp, _u32p]
        L.orc_philox_raw.argtypes = [_u32p, _u32p, _u32p]
        L.orc_observables.argtypes = [ctypes.c_void_p, _dp]
        L.orc_out_swarm.argtypes = [ctypes.c_void_p, _dp]
        L.orc_init_state.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.orc_out_swarm_fishnet.argtypes = [ctypes.c_void_p, _dp]
        L.orc_ref_voronoi_step.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                           ctypes.c_uint32]
        L.orc_timed_steps.restype = ctypes.c_double
        L.orc_timed_steps.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64]
        self.kind = L.orc_kind().decode()
        if self.kind == 'reference':
            L.orc_ref_out_swarm.argtypes = [ctypes.c_void_p, _dp]
            L.orc_ref_swarmdyn_main.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p)]


_LIBS = {}


def load(kind='port'):
    """kind: 'port' (always available once built) or 'reference' (oracle/_ref)."""
    if kind not in _LIBS:
        path = PORT_PATH if kind == 'port' else REF_PATH
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        _LIBS[kind] = OracleLib(path)
    return _LIBS[kind]


def have_reference():
    return os.path.exists(REF_PATH)


STATE_FIELDS = ('x', 'y', 'vx', 'vy', 'phi', 'vproj', 'fx', 'fy')


class World:
    """One swarm on the CPU oracle."""

    def __init__(self, sp, n_agents, n_pred=0, seed=0, replicate=0, kind='port'):
        self.olib = load(kind)
        self.L = self.olib.lib
        self.sp = sp
        self.N = n_agents
        self.Npred = n_pred
        self.h = self.L.orc_create(ctypes.byref(sp), n_agents, n_pred, seed, replicate)

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, **kw):
        args = []
        keep = []
        for f in STATE_FIELDS:
            a = kw.get(f)
            if a is not None:
                a = np.ascontiguousarray(a, dtype=np.float64)
                assert a.shape == (self.N,)
                keep.append(a)
            args.append(_ptr(a, _dp))
        for f in ('bin_step', 'steps_till_burst'):
            a = kw.get(f)
            if a is not None:
                a = np.ascontiguousarray(a, dtype=np.uint32)
                keep.append(a)
            args.append(_ptr(a, _u32p))
        a = kw.get('alive')
        if a is not None:
            a = np.ascontiguousarray(a, dtype=np.uint8)
            keep.append(a)
        args.append(_ptr(a, _u8p))
        self.L.orc_set_state(self.h, *args)

    def get_state(self):
        out = {f: np.zeros(self.N, dtype=np.float64) for f in STATE_FIELDS}
        out['bin_step'] = np.zeros(self.N, dtype=np.uint32)
        out['steps_till_burst'] = np.zeros(self.N, dtype=np.uint32)
        out['alive'] = np.zeros(self.N, dtype=np.uint8)
        self.L.orc_get_state(self.h, *[_ptr(out[f], _dp) for f in STATE_FIELDS],
                             _ptr(out['bin_step'], _u32p), _ptr(out['steps_till_burst'], _u32p),
                             _ptr(out['alive'], _u8p))
        return out

    def set_predators(self, x, y, phi, spawned=1):
        x, y, phi = [np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, phi)]
        self.L.orc_set_predators(self.h, _ptr(x, _dp), _ptr(y, _dp), _ptr(phi, _dp), spawned)

    def get_predators(self):
        out = np.zeros((self.Npred, 6))
        if self.Npred:
            self.L.orc_get_predators(self.h, _ptr(out, _dp))
        return out

    def counts(self):
        a, d = ctypes.c_int32(), ctypes.c_int32()
        self.L.orc_get_counts(self.h, ctypes.byref(a), ctypes.byref(d))
        return a.value, d.value

    def pair_interactions(self, flags=0, want_nn=True):
        N = self.N
        c = [np.zeros(N, dtype=np.int32) for _ in range(3)]
        f = [np.zeros((N, 2)) for _ in range(3)]
        nn_ptr = np.zeros(N + 1, dtype=np.int64)
        cap = N * max(N - 1, 1) if want_nn else 0
        nn_idx = np.zeros(max(cap, 1), dtype=np.int32)
        total = self.L.orc_pair_interactions(self.h, flags, *[_ptr(a, _i32p) for a in c],
                                             *[_ptr(a, _dp) for a in f], _ptr(nn_ptr, _i64p),
                                             _ptr(nn_idx, _i32p) if want_nn else None, cap)
        return dict(c_rep=c[0], c_alg=c[1], c_att=c[2], f_rep=f[0], f_alg=f[1], f_att=f[2],
                    nn_ptr=nn_ptr, nn_idx=nn_idx[:total] if want_nn else None)

    def inject(self, u_social=None, u_dir=None, k_next=None):
        us = None if u_social is None else np.ascontiguousarray(u_social, dtype=np.float64)
        ud = None if u_dir is None else np.ascontiguousarray(u_dir, dtype=np.float64)
        kn = None if k_next is None else np.ascontiguousarray(k_next, dtype=np.uint32)
        self.L.orc_inject_decisions(self.h, _ptr(us, _dp), _ptr(ud, _dp), _ptr(kn, _u32p))

    def step(self, s_first, n_steps=1, want_flags=False):
        fl = np.zeros(self.N, dtype=np.uint8) if want_flags else None
        self.L.orc_step(self.h, s_first, n_steps, _ptr(fl, _u8p))
        return fl

    def philox_decisions(self, step):
        us, ud = np.zeros(self.N), np.zeros(self.N)
        kn = np.zeros(self.N, dtype=np.uint32)
        self.L.orc_philox_decisions(self.h, step, _ptr(us, _dp), _ptr(ud, _dp), _ptr(kn, _u32p))
        return us, ud, kn

    def observables(self):
        out = np.zeros(8)
        self.L.orc_observables(self.h, _ptr(out, _dp))
        return out

    def init_state(self, ic):
        """ResetSystem (settings.cpp:196-387) with the uniforms of the IC protocol: the port's restatement, or the
        reference's own function fed through the replay queue."""
        rc = self.L.orc_init_state(self.h, int(ic))
        if rc != 0:
            raise RuntimeError('orc_init_state(%d) -> %d' % (ic, rc))

    def out_swarm(self):
        """The 16 columns of Out_swarm (swarmdyn.cpp:237-344): the port's restatement or the reference's own function."""
        out = np.zeros(16)
        self.L.orc_out_swarm(self.h, _ptr(out, _dp))
        return out

    def out_swarm_fishnet(self):
        out = np.zeros(4)
        self.L.orc_out_swarm_fishnet(self.h, _ptr(out, _dp))
        return out

    def ref_out_swarm(self):
        out = np.zeros(16)
        self.L.orc_ref_out_swarm(self.h, _ptr(out, _dp))
        return out

    def voronoi_step(self, s_first, n_steps, mt_seed):
        return self.L.orc_ref_voronoi_step(self.h, s_first, n_steps, mt_seed)

    def timed_steps(self, s_first, n_steps):
        return self.L.orc_timed_steps(self.h, s_first, n_steps)
