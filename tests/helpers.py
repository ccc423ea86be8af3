"""Shared state generators and comparison helpers for the parity tests."""
import numpy as np

from sde_burst_coast_b200 import params as P


def f32(a):
    """Round to FP32-representable doubles: the GPU and the FP64 oracle then see identical states."""
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def base_dict(mode='burst_coast', N=8, SL=2, **over):
    d = P.get_base_params(0, 40, mode=mode)
    P.selection_line_parameter(d, SL)
    d['N'] = N
    if mode == 'burst_coast':
        d['Npred'] = 0
    d.update(over)
    return d


def make_params(mode='burst_coast', N=8, SL=2, **over):
    d = base_dict(mode, N, SL, **over)
    sp, npred = P.make_sbc_params(d)
    return sp, npred, d


def random_state(N, sp, rng, spread=None, mid_burst=True):
    """Random FP32-representable state: positions inside the domain, arbitrary burst phases."""
    L = sp.sizeL
    if sp.BC in (5, 6):
        r = (spread if spread is not None else 0.8 * L) * np.sqrt(rng.random(N))
        th = 2 * np.pi * rng.random(N)
        x, y = r * np.cos(th), r * np.sin(th)
    elif sp.BC == 0:
        x, y = L * rng.random(N), L * rng.random(N)
        if spread is not None:
            x = (L / 2 + spread * (rng.random(N) - 0.5)) % L
            y = (L / 2 + spread * (rng.random(N) - 0.5)) % L
    else:
        s = spread if spread is not None else 40.0
        x, y = s * rng.random(N), s * rng.random(N)
    phi = 2 * np.pi * rng.random(N) - np.pi
    vproj = 0.5 + 20 * rng.random(N)
    st = dict(x=f32(x), y=f32(y), phi=f32(phi), vproj=f32(vproj))
    st['vx'] = f32(st['vproj'] * np.cos(st['phi']))
    st['vy'] = f32(st['vproj'] * np.sin(st['phi']))
    bs = int(sp.burst_steps)
    if mid_burst:
        kind = rng.integers(0, 4, N)  # 0 coast, 1 bursting, 2 burst start, 3 about to re-arm
        bin_step = np.where(kind == 1, rng.integers(1, max(bs, 2), N), 0)
        bin_step = np.where(kind == 2, bs, bin_step)
        stb = rng.integers(1, 400, N)
        stb = np.where(kind == 3, rng.integers(0, 2, N), stb)
        fang = 2 * np.pi * rng.random(N)
        fmag = np.where(bin_step > 0, sp.soc_strength, 0.0)
        st['fx'] = f32(fmag * np.cos(fang))
        st['fy'] = f32(fmag * np.sin(fang))
    else:
        bin_step = np.zeros(N, dtype=np.int64)
        stb = np.zeros(N, dtype=np.int64)
        st['fx'] = np.zeros(N)
        st['fy'] = np.zeros(N)
    st['bin_step'] = bin_step.astype(np.uint32)
    st['steps_till_burst'] = stb.astype(np.uint32)
    return st


def nn_sets(res, N):
    ptr, idx = res['nn_ptr'], res['nn_idx']
    return [tuple(idx[ptr[i]:ptr[i + 1]]) for i in range(N)]


def ang_diff(a, b):
    d = (np.asarray(a) - np.asarray(b) + np.pi) % (2 * np.pi) - np.pi
    return np.abs(d)


def cancellation(pairs):
    """Relative error FP32 may leave in the DIRECTION of a zone sum: c unit vectors summed in FP32 carry an
    absolute error of about c * 2^-23, which matters when they nearly cancel (|sum| << 1). Per row, the worst
    of the zones that have members."""
    cond = np.zeros(len(pairs['c_rep']))
    for c, f in (('c_rep', 'f_rep'), ('c_alg', 'f_alg'), ('c_att', 'f_att')):
        n = pairs[c].astype(np.float64)
        mag = np.hypot(pairs[f][:, 0], pairs[f][:, 1])
        with np.errstate(divide='ignore', invalid='ignore'):
            cond = np.maximum(cond, np.where(n > 0, n * 2.4e-7 / np.maximum(mag, 1e-30), 0.0))
    return np.minimum(cond, 1.0)
