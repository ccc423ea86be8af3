"""Device-side initial conditions (sbc_init_state) against ResetSystem (/root/reference/src/settings.cpp:196-387):
the oracle port per agent, and - when oracle/_ref is present - the reference's own ResetSystem fed with the protocol's
uniforms through the replay queue (ring cases as a sorted set: the reference shuffles with a clock seed).

Bar: the device computes the positions in FP64 in the reference's operation order and rounds the state to FP32, so
x, y, v agree with the rounded oracle values to one FP32 ulp (libm and CUDA sin/cos may differ in the last FP64 bit),
headings to 5e-7 rad; timers, force, alive exact. A replicate's state must not depend on how the ensemble is split."""
import numpy as np
import pytest

from helpers import make_params, ang_diff
from oracle import pyorc

pytestmark = pytest.mark.gpu
KINDS = ['port'] + (['reference'] if pyorc.have_reference() else [])
CASES = [('burst_coast', 8, {}), ('burst_coast', 40, {}), ('natPred', 50, dict(BC=0, size=45.0, Npred=0)),
         ('burst_coast', 30, dict(BC=2, size=20.0)), ('burst_coast', 30, dict(BC=1, size=20.0)),
         ('burst_coast', 30, dict(BC=4, size=20.0)), ('burst_coast', 30, dict(BC=5, size=20.0)), ('burst_coast', 30, dict(BC=-1))]


def _swarm(*a, **k):
    from sde_burst_coast_b200.swarm import Swarm
    return Swarm(*a, **k)


def _ulp32(v):
    return np.spacing(np.abs(np.asarray(v, dtype=np.float32))).astype(np.float64) + 1e-38


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('ic', [0, 1, 2, 3, 4, 5, 7])
@pytest.mark.parametrize('mode,N,over', CASES)
def test_device_initial_conditions_match_oracle(kind, ic, mode, N, over):
    R, off, seed = 3, 5, 424242
    sp, _, _ = make_params(mode, N, **over)
    g = _swarm(sp, N, n_replicates=R, seed=seed, replicate_offset=off)
    g.init_state(ic)
    got = g.get_state()
    g.close()
    for r in range(R):
        w = pyorc.World(sp, N, seed=seed, replicate=off + r, kind=kind)
        w.init_state(ic)
        o = w.get_state()
        kg = ko = np.arange(N)
        if kind == 'reference' and ic in (3, 4, 5):
            kg = np.lexsort((got['phi'][r], got['y'][r], got['x'][r]))   # walls clamp several agents onto one point
            ko = np.lexsort((o['phi'].astype(np.float32), o['y'].astype(np.float32), o['x'].astype(np.float32)))
        for f in ('x', 'y', 'vx', 'vy'):
            assert (np.abs(got[f][r][kg] - o[f][ko]) <= _ulp32(o[f][ko])).all(), (ic, r, f)
        assert (ang_diff(got['phi'][r][kg], o['phi'][ko]) <= 5e-7).all()
        assert np.array_equal(got['vproj'][r], o['vproj']) and not got['fx'][r].any() and not got['fy'][r].any()
        assert np.array_equal(got['bin_step'][r], o['bin_step']) and np.array_equal(got['steps_till_burst'][r], o['steps_till_burst'])
        assert got['alive'][r].all()


def test_initial_conditions_do_not_depend_on_the_split_and_fill_large_swarms():
    N, R = 32, 12
    sp, _, _ = make_params('burst_coast', N)
    whole = _swarm(sp, N, n_replicates=R, seed=9)
    whole.init_state(3)
    a = whole.get_state()
    whole.close()
    for r0, r1 in ((0, 5), (5, 12)):
        part = _swarm(sp, N, n_replicates=r1 - r0, seed=9, replicate_offset=r0)
        part.init_state(3)
        b = part.get_state()
        part.close()
        for f in ('x', 'y', 'vx', 'vy', 'phi'):
            assert np.array_equal(a[f][r0:r1], b[f])
    # the ring shuffle is a bijection for any N (cycle-walked Feistel network): the same places as the port's, each once
    N = 5000
    sp, _, _ = make_params('natPred', N, BC=-1, Npred=0)
    g = _swarm(sp, N, seed=3)
    g.init_state(5)
    s = {k: v.reshape(1, -1) for k, v in g.get_state().items()}
    g.close()
    w = pyorc.World(sp, N, seed=3)
    w.init_state(5)
    o = w.get_state()
    assert (np.abs(s['x'][0] - o['x']) <= _ulp32(o['x'])).all() and (np.abs(s['y'][0] - o['y']) <= _ulp32(o['y'])).all()
    assert len(set(zip(s['x'][0], s['y'][0]))) == N


def test_stepping_from_device_initial_conditions_matches_oracle_timers():
    """ICs on the device, then 300 steps: burst timers stay bit-identical with the oracle started from ITS ICs."""
    N, R = 16, 4
    sp, _, _ = make_params('burst_coast', N)
    g = _swarm(sp, N, n_replicates=R, seed=31)
    g.init_state(0)
    g.step(0, 300)
    got = g.get_state()
    g.close()
    for r in range(R):
        w = pyorc.World(sp, N, seed=31, replicate=r)
        w.init_state(0)
        w.step(0, 300)
        o = w.get_state()
        assert np.array_equal(got['steps_till_burst'][r], o['steps_till_burst'])
