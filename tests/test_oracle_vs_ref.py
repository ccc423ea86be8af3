"""Pin the CPU restatement (oracle/sbc_oracle.cpp, "port") against the reference's own translation
units (oracle/_ref/liborc_ref.so, compiled in place from /root/reference/src): bit-identical
doubles and integers on the same states and the same injected decisions.

Runs wherever oracle/_ref exists (this container builds it; it travels to the GPU box prebuilt).
The golden fixtures in tests/golden/ were produced by the same reference build, so the port stays
pinned even where the reference sources are absent (tests/test_golden.py).
"""
import numpy as np
import pytest

from helpers import make_params, random_state
from oracle import pyorc
from sde_burst_coast_b200.params import PAIR_ACTIVE_ROWS, PAIR_ALL_ROWS

needs_ref = pytest.mark.skipif(not pyorc.have_reference(), reason='oracle/_ref not built')

CASES = [
    ('burst_coast', dict(BC=6), 8), ('burst_coast', dict(BC=6), 30), ('burst_coast', dict(BC=5), 30),
    ('burst_coast', dict(BC=6, alg_range=10.0), 64), ('burst_coast', dict(BC=-1), 40),
    ('natPred', dict(BC=0, size=100.0, Npred=0), 200), ('natPred', dict(BC=0, size=100.0, Npred=0, alg_range=10.0), 200),
    ('burst_coast', dict(BC=1, size=60.0), 50), ('burst_coast', dict(BC=2, size=60.0), 50),
    ('burst_coast', dict(BC=3, size=60.0), 50), ('burst_coast', dict(BC=4, size=60.0), 50),
]


def _same(a, b, what):
    assert np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True), what


@needs_ref
@pytest.mark.parametrize('mode,over,N', CASES)
@pytest.mark.parametrize('flags', [PAIR_ALL_ROWS, PAIR_ACTIVE_ROWS])
def test_pair_pass_bit_identical(mode, over, N, flags):
    sp, _, _ = make_params(mode, N, **over)
    rng = np.random.default_rng(N + 17)
    st = random_state(N, sp, rng, spread=50.0 if sp.BC in (-1, 1, 2, 3, 4) else None)
    res = []
    for kind in ('port', 'reference'):
        w = pyorc.World(sp, N, kind=kind)
        w.set_state(**st)
        res.append(w.pair_interactions(flags))
        w.close()
    for k in res[0]:
        _same(res[0][k], res[1][k], k)


@needs_ref
@pytest.mark.parametrize('mode,over,N', CASES)
def test_steps_bit_identical(mode, over, N):
    """300 steps with the shared Philox decision protocol: every double of the state identical."""
    sp, _, _ = make_params(mode, N, **over)
    rng = np.random.default_rng(N + 29)
    st = random_state(N, sp, rng, spread=50.0 if sp.BC in (-1, 1, 2, 3, 4) else None)
    ws = [pyorc.World(sp, N, seed=99, replicate=3, kind=k) for k in ('port', 'reference')]
    for w in ws:
        w.set_state(**st)
    fl = [np.zeros(N, dtype=np.uint8), np.zeros(N, dtype=np.uint8)]
    for s in range(0, 300, 25):
        for k, w in enumerate(ws):
            fl[k] |= w.step(s, 25, want_flags=True)
        a, b = ws[0].get_state(), ws[1].get_state()
        for key in a:
            _same(a[key], b[key], '%s after step %d' % (key, s + 25))
    # the reference build can only report the branches visible from outside ParticleBurstCoast
    _same(fl[0] & 195, fl[1], 'branch flags (burst start, social, flee, re-arm)')
    assert fl[0].any()
    for w in ws:
        w.close()


@needs_ref
@pytest.mark.parametrize('mode,kill,move,npred', [
    ('natPred', 3, 1, 1), ('natPredNoConfu', 2, 1, 1), ('natPred', 4, 1, 1), ('natPred', 1, 1, 1),
    ('sinFisher', 1, 0, 1), ('mulFisher', 1, 0, 50),
])
def test_predator_phase_bit_identical(mode, kill, move, npred):
    N = 60
    sp, np_, _ = make_params(mode, N, BC=0, size=100.0, Npred=npred, pred_kill=kill, pred_move=move,
                             pred_time=0.02, time=5.0, N_confu=40, kill_rate=20)
    assert np_ == npred
    rng = np.random.default_rng(kill * 10 + move + npred)
    st = random_state(N, sp, rng, spread=40.0)
    ws = [pyorc.World(sp, N, npred, seed=5, replicate=1, kind=k) for k in ('port', 'reference')]
    for w in ws:
        w.set_state(**st)
    for s in range(0, 4000, 200):
        for w in ws:
            w.step(s, 200)
        a, b = ws[0].get_state(), ws[1].get_state()
        alive = a['alive'].astype(bool)
        _same(a['alive'], b['alive'], 'alive after %d' % s)
        for key in a:
            _same(a[key][alive], b[key][alive], '%s after step %d' % (key, s + 100))
        _same(ws[0].get_predators(), ws[1].get_predators(), 'predators')
    assert ws[0].counts()[1] > 0, 'the scenario should kill somebody'
    for w in ws:
        w.close()


@needs_ref
def test_injected_decisions_bit_identical():
    N = 30
    sp, _, _ = make_params('burst_coast', N)
    rng = np.random.default_rng(3)
    st = random_state(N, sp, rng, spread=0.97 * sp.sizeL)
    st['bin_step'][:] = sp.burst_steps  # every row decides
    us, ud = rng.random(N), rng.random(N)
    kn = rng.integers(1, 500, N).astype(np.uint32)
    out = []
    for kind in ('port', 'reference'):
        w = pyorc.World(sp, N, kind=kind)
        w.set_state(**st)
        w.inject(us, ud, kn)
        fl = w.step(0, 1, want_flags=True)
        out.append((w.get_state(), fl))
        w.close()
    for key in out[0][0]:
        _same(out[0][0][key], out[1][0][key], key)
    _same(out[0][1] & 195, out[1][1], 'flags')
    assert (out[0][1] & 4).any(), 'some agent should have been redirected by the wall look-ahead'


@needs_ref
def test_observables_match_reference_out_swarm():
    """orc_observables columns against the reference's own Out_swarm row (swarmdyn.cpp:324-343)."""
    N = 30
    sp, _, _ = make_params('burst_coast', N)
    rng = np.random.default_rng(8)
    st = random_state(N, sp, rng)
    w = pyorc.World(sp, N, kind='reference')
    w.set_state(**st)
    mine = w.observables()
    ref = w.ref_out_swarm()
    assert ref[0] == mine[0] == N
    assert abs(ref[1] - mine[1]) < 1e-12          # pol_order
    assert np.allclose(ref[3:7], mine[2:6], rtol=0, atol=1e-12)  # avg_x, avg_v
    assert abs(ref[7] - mine[6]) < 1e-12          # avg_speed
    p = pyorc.World(sp, N, kind='port')
    p.set_state(**st)
    assert np.array_equal(p.observables(), mine)


@needs_ref
@pytest.mark.parametrize('mode,N,over', [('burst_coast', 30, {}), ('natPred', 200, dict(BC=0, size=100.0, Npred=0))])
def test_port_out_swarm_is_the_reference_row(mode, N, over):
    """orc_out_swarm of the port (all 16 columns, quirks of swarmdyn.cpp:270-296 included) against the reference's own
    Out_swarm on the same state, with and without dead agents: bit-identical."""
    sp, _, _ = make_params(mode, N, **over)
    rng = np.random.default_rng(N)
    st = random_state(N, sp, rng)
    for alive in (None, (rng.random(N) > 0.25).astype(np.uint8)):
        p, w = pyorc.World(sp, N), pyorc.World(sp, N, kind='reference')
        for o in (p, w):
            o.set_state(alive=alive, **st) if alive is not None else o.set_state(**st)
        assert np.array_equal(p.out_swarm(), w.ref_out_swarm(), equal_nan=True)


@needs_ref
@pytest.mark.parametrize('npred', [1, 20])
def test_port_fishnet_row_is_the_reference_row(npred):
    N = 150
    sp, _, _ = make_params('mulFisher' if npred > 1 else 'natPred', N, BC=0, size=100.0, Npred=npred, pred_time=0.0)
    rng = np.random.default_rng(npred)
    st = random_state(N, sp, rng, spread=40.0)
    alive = (rng.random(N) > 0.1).astype(np.uint8)
    px = 40.0 + 0.5 * np.arange(npred, dtype=np.float64)
    p, w = pyorc.World(sp, N, npred), pyorc.World(sp, N, npred, kind='reference')
    for o in (p, w):
        o.set_state(alive=alive, **st)
        o.set_predators(px, np.full(npred, 48.0), np.full(npred, np.pi / 2), 1)
    assert np.array_equal(p.out_swarm_fishnet(), w.out_swarm_fishnet())


@needs_ref
@pytest.mark.parametrize('ic', [0, 1, 2, 3, 4, 5, 7])
@pytest.mark.parametrize('mode,N,over', [('burst_coast', 8, {}), ('burst_coast', 40, {}), ('natPred', 50, dict(BC=0, size=45.0, Npred=0)),
                                         ('burst_coast', 30, dict(BC=2, size=20.0)), ('burst_coast', 30, dict(BC=1, size=20.0)),
                                         ('burst_coast', 30, dict(BC=4, size=20.0)), ('burst_coast', 30, dict(BC=5, size=20.0)),
                                         ('burst_coast', 30, dict(BC=-1))])
def test_port_initial_conditions_are_the_reference_reset_system(ic, mode, N, over):
    """orc_init_state of the port against the reference's own ResetSystem + Boundary (settings.cpp:196-387) fed with
    the same uniforms through the replay queue: every field bit-identical. The ring cases (IC 3-5) shuffle with a clock
    seed in the reference, so WHICH agent holds a ring place differs: rows are compared as a sorted set."""
    sp, _, _ = make_params(mode, N, **over)
    a, b = pyorc.World(sp, N, seed=77, replicate=3), pyorc.World(sp, N, seed=77, replicate=3, kind='reference')
    a.init_state(ic)
    b.init_state(ic)
    sa, sb = a.get_state(), b.get_state()
    ka = kb = np.arange(N)
    if ic in (3, 4, 5):
        ka, kb = (np.lexsort((q['vy'], q['vx'], q['phi'], q['y'], q['x'])) for q in (sa, sb))   # walls clamp several agents onto one point
    for f in ('x', 'y', 'vx', 'vy', 'phi', 'vproj', 'fx', 'fy', 'bin_step', 'steps_till_burst', 'alive'):
        _same(sa[f][ka], sb[f][kb], (ic, f))
    if ic in (3, 4, 5) and over.get('BC') == -1:     # the protocol's shuffle is a bijection: every agent was placed exactly once
        assert len(set(zip(sa['x'], sa['y']))) == N
