"""Edge cases of the hot path on the GPU against the oracle: empty and ragged swarms, a lone agent,
coincident agents, agents on the wall, the path boundaries (N = 32/33, 1024/1025), and the degenerate
arithmetic the reference silently tolerates: zero speed (division by the old vproj),
flee forces that cancel."""
import numpy as np
import pytest

from helpers import make_params, random_state, f32
from oracle import pyorc

pytestmark = pytest.mark.gpu


def _swarm(*a, **k):
    from sde_burst_coast_b200.swarm import Swarm
    return Swarm(*a, **k)


def _run_both(sp, N, st, steps, npred=0, seed=3, alive=None, multikernel=False, preds=None):
    if alive is not None:
        st = dict(st, alive=alive)
    w = pyorc.World(sp, N, npred, seed=seed)
    g = _swarm(sp, N, npred, seed=seed)
    g.force_multikernel(multikernel)
    for obj in (w, g):
        obj.set_state(**st)
        if preds is not None:
            obj.set_predators(*preds, 1)
    w.step(0, steps)
    g.step(0, steps)
    o, gs = w.get_state(), g.get_state()
    cw, cg = w.counts(), g.counts()
    w.close()
    g.close()
    return o, gs, cw, (int(cg[0][0]), int(cg[1][0]))


@pytest.mark.parametrize('multikernel', [False, True])
def test_all_dead_swarm_is_a_no_op(multikernel):
    N = 20
    sp, _, _ = make_params('burst_coast', N)
    st = random_state(N, sp, np.random.default_rng(0))
    o, gs, cw, cg = _run_both(sp, N, st, 50, alive=np.zeros(N, dtype=np.uint8), multikernel=multikernel)
    assert cw == cg == (0, N)
    for k in ('x', 'y', 'phi', 'vproj'):
        assert np.allclose(gs[k], st[k]) and np.allclose(o[k], st[k])   # dead agents keep their last state


@pytest.mark.parametrize('N', [1, 2])
def test_lone_agent_and_coincident_pair(N):
    sp, _, _ = make_params('burst_coast', N)
    st = dict(x=f32(np.full(N, 3.0)), y=f32(np.full(N, -2.0)), phi=f32(np.linspace(0.3, 1.3, N)), vproj=np.ones(N))
    o, gs, _, _ = _run_both(sp, N, st, 400)
    assert np.array_equal(o['steps_till_burst'], gs['steps_till_burst'])
    assert np.all(np.isfinite(gs['x'])) and np.all(np.isfinite(o['x']))
    assert np.all(np.hypot(gs['x'], gs['y']) <= sp.sizeL * (1 + 1e-6))


def test_ragged_ensemble_alive_masks():
    """Replicates with different numbers of alive agents in one handle (ragged input)."""
    N, R = 16, 6
    sp, _, _ = make_params('burst_coast', N)
    rng = np.random.default_rng(5)
    sts = [random_state(N, sp, rng) for _ in range(R)]
    alive = np.ones((R, N), dtype=np.uint8)
    for r in range(R):
        alive[r, rng.choice(N, size=3 * r % N, replace=False)] = 0
    g = _swarm(sp, N, n_replicates=R, seed=8)
    g.set_state(alive=alive, **{k: np.stack([s[k] for s in sts]) for k in sts[0]})
    ref0 = g.pair_interactions(1, want_nn=False)
    g.step(0, 200)
    gs = g.get_state()
    na, nd = g.counts()
    assert np.array_equal(na, alive.sum(axis=1)) and np.array_equal(nd, N - alive.sum(axis=1))
    for r in range(R):
        w = pyorc.World(sp, N, seed=8, replicate=r)
        w.set_state(alive=alive[r], **sts[r])
        pr = w.pair_interactions(1, want_nn=False)
        for k in ('c_rep', 'c_alg', 'c_att'):
            assert np.array_equal(pr[k], ref0[k].reshape(R, N)[r]), (r, k)
        w.step(0, 200)
        o = w.get_state()
        al = alive[r].astype(bool)
        assert np.array_equal(o['steps_till_burst'][al], gs['steps_till_burst'][r][al])
        w.close()
    g.close()


@pytest.mark.parametrize('N', [32, 33, 1024, 1025])
def test_path_boundaries(N):
    """N = 32 | 33: lanes-per-tank kernel vs one block per swarm; 1024 | 1025: block vs multi-kernel."""
    sp, _, _ = make_params('natPred', N, BC=0, size=200.0, Npred=0)
    st = random_state(N, sp, np.random.default_rng(N))
    o, gs, _, _ = _run_both(sp, N, st, 3, seed=11)
    assert np.array_equal(o['steps_till_burst'], gs['steps_till_burst']) and np.array_equal(o['bin_step'], gs['bin_step'])
    ok = (np.abs(o['x'] - gs['x']) < 1e-3) & (np.abs(o['y'] - gs['y']) < 1e-3)
    assert ok.mean() > 0.99


def test_agents_on_and_beyond_the_wall():
    """Positions exactly on the tank wall and outside it (the reference reflects them at the next Boundary call)."""
    N = 8
    sp, _, _ = make_params('burst_coast', N)
    ang = np.linspace(0, 2 * np.pi, N, endpoint=False)
    rad = np.array([24.5, 24.5, 25.0, 26.0, 24.4999, 30.0, 24.5, 10.0])
    st = dict(x=f32(rad * np.cos(ang)), y=f32(rad * np.sin(ang)), phi=f32(ang), vproj=f32(np.full(N, 5.0)))
    o, gs, _, _ = _run_both(sp, N, st, 1)
    assert np.allclose(o['x'], gs['x'], atol=2e-4) and np.allclose(o['y'], gs['y'], atol=2e-4)
    o, gs, _, _ = _run_both(sp, N, st, 300)
    assert np.all(np.hypot(gs['x'], gs['y']) <= 24.5 * (1 + 1e-5))


@pytest.mark.parametrize('multikernel', [False, True])
def test_zero_speed_degenerate_division(multikernel):
    """vproj = 0 makes the reference divide by zero in the heading update (agents_dynamics.cpp:215): the state
    turns non-finite there. Replicated, not fixed: the same agents become non-finite on the GPU."""
    N = 12
    sp, _, _ = make_params('burst_coast', N)
    rng = np.random.default_rng(2)
    st = random_state(N, sp, rng)
    st['vproj'][::3] = 0.0
    st['vx'][::3] = 0.0
    st['vy'][::3] = 0.0
    st['bin_step'][:] = np.where(np.arange(N) % 2 == 0, sp.burst_steps - 5, 0)   # mid burst: a force is acting
    st['fx'], st['fy'] = f32(np.full(N, 100.0)), f32(np.full(N, -60.0))
    o, gs, _, _ = _run_both(sp, N, st, 1, multikernel=multikernel)
    bad_o = ~(np.isfinite(o['x']) & np.isfinite(o['phi']))
    bad_g = ~(np.isfinite(gs['x']) & np.isfinite(gs['phi']))
    assert np.array_equal(bad_o, bad_g), (bad_o, bad_g)
    good = ~bad_o
    assert np.allclose(o['x'][good], gs['x'][good], atol=1e-4)


def test_cancelling_flee_forces_give_nan_like_the_reference():
    """Two predators placed symmetrically around an agent: force_flee sums to ~0 and vec_set_mag divides by
    its length (agents_dynamics.cpp:78-79). Whatever the reference produces (huge or NaN force), the GPU's
    finite/non-finite pattern of the force is the same for agents that took the flee branch."""
    N = 6
    sp, npred, _ = make_params('mulFisher', N, BC=0, size=100.0, Npred=2, pred_kill=0, pred_time=0.0, time=5.0,
                               prob_social=0.0, flee_range=1000.0)
    st = dict(x=f32(np.array([50.0, 20, 30, 70, 80, 10])), y=f32(np.array([50.0, 20, 30, 70, 80, 10])),
              phi=np.zeros(N), vproj=np.ones(N), bin_step=np.full(N, sp.burst_steps, dtype=np.uint32),
              steps_till_burst=np.full(N, 50, dtype=np.uint32))
    preds = (np.array([45.0, 55.0]), np.array([50.0, 50.0]), np.array([0.0, 0.0]))
    o, gs, _, _ = _run_both(sp, N, st, 1, npred=2, preds=preds, multikernel=True)
    fin_o = np.isfinite(o['fx']) & np.isfinite(o['fy'])
    fin_g = np.isfinite(gs['fx']) & np.isfinite(gs['fy'])
    assert np.array_equal(fin_o[1:], fin_g[1:])
    # agent 0 sits exactly between the predators: the sum cancels to 0 (NaN) or to rounding noise (finite)
    assert np.isnan(o['fx'][0]) or abs(np.hypot(o['fx'][0], o['fy'][0]) - sp.soc_strength) < 1e-6


def test_float_abi_equals_double_abi():
    """sbc_set_state_f32 / sbc_get_particles_f32 carry the same numbers as the reference-layout double entry points
    (the state is FP32 on the device either way): identical device state, identical particle rows."""
    from sde_burst_coast_b200.swarm import Swarm
    N, R = 30, 7
    sp, _, _ = make_params('burst_coast', N)
    rng = np.random.default_rng(12)
    st = random_state(N * R, sp, rng)
    state = {k: v.reshape(R, N) for k, v in st.items()}
    state['alive'] = (rng.random((R, N)) > 0.1).astype(np.uint8)
    a = Swarm(sp, N, n_replicates=R, seed=4)
    b = Swarm(sp, N, n_replicates=R, seed=4)
    a.set_state(**state)
    b.set_state_f32(**state)
    sa, sb = a.get_state(), b.get_state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k], equal_nan=True), k
    a.step(0, 300)
    b.step(0, 300)
    pa, aa = a.get_particles()
    pb, ab = b.get_particles_f32()
    assert pb.dtype == np.float32 and np.array_equal(pa.astype(np.float32), pb) and np.array_equal(aa, ab)
    a.close()
    b.close()
