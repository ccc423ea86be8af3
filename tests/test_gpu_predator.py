"""GPU parity of the predator phase (prey<-predator detection, spawn, chase / random walk / fish-net move,
the four kill rules) against the CPU oracle, through the C ABI.

Bars: alive sets and zone/flee counters exact for a given state; predator positions and agent states
within the one-step FP32 tolerance; longer runs are compared statistically (kills vs time)."""
import numpy as np
import pytest

from helpers import make_params, random_state, ang_diff
from oracle import pyorc

pytestmark = pytest.mark.gpu


def _swarm(*a, **k):
    from sde_burst_coast_b200.swarm import Swarm
    return Swarm(*a, **k)


def _setup(mode, N, npred, kill, move, L=100.0, seed=1, **over):
    sp, np_, _ = make_params(mode, N, BC=0, size=L, Npred=npred, pred_kill=kill, pred_move=move,
                             pred_time=0.0, time=50.0, **over)
    assert np_ == npred
    rng = np.random.default_rng(seed)
    st = random_state(N, sp, rng, spread=30.0)
    return sp, st, rng


@pytest.mark.parametrize('kill,move,npred,mode', [(1, 1, 1, 'natPred'), (2, 1, 1, 'natPredNoConfu'), (3, 1, 1, 'natPred'),
                                                   (4, 1, 1, 'natPred'), (1, 0, 1, 'sinFisher'), (1, 0, 50, 'mulFisher')])
@pytest.mark.parametrize('multikernel', [True, False])
def test_one_step_with_predator_present(kill, move, npred, mode, multikernel):
    """Predator(s) placed inside the swarm: one step = detection for burst-starting rows, update,
    predator move, kill. Alive sets exact, everything else within tolerance."""
    N = 200
    sp, st, rng = _setup(mode, N, npred, kill, move, seed=kill * 7 + npred, N_confu=500, kill_rate=4000.0 if kill == 4 else 400.0)
    st['bin_step'] = np.where(rng.random(N) < 0.4, sp.burst_steps, st['bin_step']).astype(np.uint32)
    if npred == 1:
        px, py, pphi = np.array([50.0]), np.array([50.0]), np.array([0.3])
    else:
        px = 40.0 + 0.5 * np.arange(npred, dtype=np.float64)
        py = np.full(npred, 48.0)
        pphi = np.full(npred, np.pi / 2)
    mism = 0
    killed = 0
    for trial in range(3):
        w = pyorc.World(sp, N, npred, seed=100 + trial, replicate=0)
        g = _swarm(sp, N, npred, seed=100 + trial)
        g.force_multikernel(multikernel)
        for obj in (w, g):
            obj.set_state(**st)
            obj.set_predators(px, py, pphi, 1)
        s = 5 + trial   # > pred_time/dt = 0, not the spawn step
        w.step(s, 1)
        g.step(s, 1)
        o, gs = w.get_state(), g.get_state()
        assert np.array_equal(o['alive'], gs['alive']), np.where(o['alive'] != gs['alive'])[0]
        killed += N - int(o['alive'].sum())
        al = o['alive'].astype(bool)
        assert np.array_equal(o['bin_step'][al], gs['bin_step'][al])
        ok = (np.abs(o['x'] - gs['x']) <= 1e-5 * np.maximum(1, np.abs(o['x']))) & \
             (np.abs(o['y'] - gs['y']) <= 1e-5 * np.maximum(1, np.abs(o['y']))) & \
             (np.hypot(o['fx'] - gs['fx'], o['fy'] - gs['fy']) <= 2e-5 * np.maximum(1, np.hypot(o['fx'], o['fy'])))
        mism += int((~ok[al]).sum())
        po, pg = w.get_predators(), g.get_predators()[0]
        assert np.allclose(po, pg, rtol=1e-5, atol=1e-4), (po[:2], pg[:2])
        assert w.counts() == tuple(int(v[0]) for v in g.counts())
        w.close()
        g.close()
    assert killed > 0, 'somebody should have been killed'
    assert mism <= 3, mism   # detection / decision flips at FP32 resolution are rare


@pytest.mark.parametrize('multikernel', [True, False])
@pytest.mark.parametrize('npred', [1, 50])
def test_spawn_step_places_predator_like_oracle(npred, multikernel):
    N = 120
    mode = 'natPred' if npred == 1 else 'mulFisher'
    sp, np_, _ = make_params(mode, N, BC=0, size=100.0, Npred=npred, pred_kill=0, pred_move=1 if npred == 1 else 0,
                             pred_time=0.01, time=50.0)
    rng = np.random.default_rng(3)
    st = random_state(N, sp, rng, spread=30.0)
    w = pyorc.World(sp, N, npred, seed=9, replicate=4)
    g = _swarm(sp, N, npred, seed=9, replicate_offset=4)
    g.force_multikernel(multikernel)
    for obj in (w, g):
        obj.set_state(**st)
    w.step(0, 12)
    g.step(0, 12)   # spawn happens at s = 10
    po, pg = w.get_predators(), g.get_predators()[0]
    assert np.allclose(po[:, :2] % 100.0, pg[:, :2] % 100.0, atol=2e-3), (po[:3], pg[:3])
    assert np.allclose(po[:, 2:4], pg[:, 2:4], atol=2e-3)
    w.close()
    g.close()


def test_survival_curve_matches_oracle_statistically():
    """natPred (chasing predator, confusion kill rule): the number of survivors after 3000 steps over an
    ensemble of swarms agrees with the oracle's ensemble within 4 standard errors."""
    N, R = 30, 48
    sp, _, _ = make_params('natPred', N, BC=0, size=100.0, Npred=1, pred_kill=2, pred_move=1, pred_time=0.05,
                           time=50.0, kill_rate=20.0, pred_speed0=20.0)
    rng = np.random.default_rng(0)
    sts = [random_state(N, sp, rng, spread=25.0, mid_burst=False) for _ in range(R)]
    dead_o, dead_g = [], []
    for r in range(R):
        w = pyorc.World(sp, N, 1, seed=31, replicate=r)
        w.set_state(**sts[r])
        w.step(0, 3000)
        dead_o.append(w.counts()[1])
        w.close()
    g = _swarm(sp, N, 1, n_replicates=R, seed=77)
    g.set_state(**{k: np.stack([s[k] for s in sts]) for k in sts[0]})
    g.step(0, 3000)
    dead_g = g.counts()[1]
    g.close()
    mo, mg = np.mean(dead_o), np.mean(dead_g)
    se = np.sqrt(np.var(dead_o) / R + np.var(dead_g) / R) + 1e-9
    assert mo > 0.5 and mg > 0.5, (mo, mg)
    assert abs(mo - mg) <= 4 * se + 0.5, (mo, mg, se)


def test_block_and_multikernel_paths_agree_over_a_predator_run():
    """Same swarm, same seed, 4000 steps through the predator phase on both schedules: identical burst
    timers and kill counts, predator tracks within tolerance."""
    N = 300
    sp, _, _ = make_params('natPred', N, BC=0, size=100.0, Npred=1, pred_kill=1, pred_move=1, pred_time=0.05,
                           time=50.0, pred_speed0=20.0)
    rng = np.random.default_rng(12)
    st = random_state(N, sp, rng, spread=30.0)
    out = []
    for mk in (False, True):
        g = _swarm(sp, N, 1, seed=5)
        g.force_multikernel(mk)
        g.set_state(**st)
        g.step(0, 4000)
        out.append((g.get_state(), g.get_predators(), g.counts(), g.stats()))
        g.close()
    a, b = out
    assert a[3]['block_launches'] == 1 and b[3]['update_launches'] == 4000
    al = a[0]['alive'].astype(bool) & b[0]['alive'].astype(bool)
    assert np.array_equal(a[0]['steps_till_burst'][al], b[0]['steps_till_burst'][al])
    # trajectories are chaotic over 4000 steps, kill counts agree statistically only
    assert int(a[2][1][0]) > 0 and int(b[2][1][0]) > 0
    assert abs(int(a[2][1][0]) - int(b[2][1][0])) <= 0.5 * max(int(a[2][1][0]), int(b[2][1][0])) + 5


@pytest.mark.parametrize('kill,move,npred,mode', [(1, 1, 1, 'natPred'), (3, 1, 1, 'natPred'), (4, 1, 1, 'natPred'),
                                                   (1, 0, 1, 'sinFisher'), (1, 0, 50, 'mulFisher')])
def test_cluster_kernel_one_step_with_predator(kill, move, npred, mode):
    """Swarms of 257..2048 agents run on a thread-block cluster (swarm_cluster.cu): same checks as above."""
    N = 700
    sp, st, rng = _setup(mode, N, npred, kill, move, seed=kill * 5 + npred, N_confu=2000,
                         kill_rate=4000.0 if kill == 4 else 400.0)
    st['bin_step'] = np.where(rng.random(N) < 0.4, sp.burst_steps, st['bin_step']).astype(np.uint32)
    if npred == 1:
        px, py, pphi = np.array([50.0]), np.array([50.0]), np.array([0.3])
    else:
        px = 40.0 + 0.5 * np.arange(npred, dtype=np.float64)
        py = np.full(npred, 48.0)
        pphi = np.full(npred, np.pi / 2)
    killed = 0
    for trial in range(2):
        w = pyorc.World(sp, N, npred, seed=200 + trial, replicate=0)
        g = _swarm(sp, N, npred, seed=200 + trial)
        for obj in (w, g):
            obj.set_state(**st)
            obj.set_predators(px, py, pphi, 1)
        w.step(7 + trial, 1)
        g.step(7 + trial, 1)
        o, gs = w.get_state(), g.get_state()
        assert g.stats()['block_launches'] == 1
        assert np.array_equal(o['alive'], gs['alive']), np.where(o['alive'] != gs['alive'])[0]
        killed += N - int(o['alive'].sum())
        al = o['alive'].astype(bool)
        assert np.array_equal(o['bin_step'][al], gs['bin_step'][al])
        ok = (np.abs(o['x'] - gs['x']) <= 1e-5 * np.maximum(1, np.abs(o['x']))) & \
             (np.abs(o['y'] - gs['y']) <= 1e-5 * np.maximum(1, np.abs(o['y'])))
        assert (~ok[al]).sum() <= 4
        assert np.allclose(w.get_predators(), g.get_predators()[0], rtol=1e-5, atol=1e-4)
        w.close()
        g.close()
    assert killed > 0


def test_cluster_kernel_spawn_and_long_run():
    N = 1200
    sp, _, _ = make_params('natPred', N, BC=0, size=200.0, Npred=1, pred_kill=1, pred_move=1, pred_time=0.01,
                           time=50.0, pred_speed0=20.0)
    rng = np.random.default_rng(4)
    st = random_state(N, sp, rng, spread=60.0)
    w = pyorc.World(sp, N, 1, seed=9, replicate=0)
    g = _swarm(sp, N, 1, seed=9)
    for obj in (w, g):
        obj.set_state(**st)
    w.step(0, 12)
    g.step(0, 12)
    assert np.allclose(w.get_predators()[:, :2] % 200.0, g.get_predators()[0][:, :2] % 200.0, atol=2e-3)
    g.step(12, 3000)
    gs = g.get_state()
    w.step(12, 3000)
    o = w.get_state()
    both = o['alive'].astype(bool) & gs['alive'].astype(bool)
    assert np.array_equal(o['steps_till_burst'][both], gs['steps_till_burst'][both])
    assert int(g.counts()[1][0]) > 0
    w.close()
    g.close()


@pytest.mark.parametrize('N', [8, 13, 30])
@pytest.mark.parametrize('kill,move,mode', [(1, 1, 'natPred'), (2, 1, 'natPredNoConfu'), (3, 1, 'natPred'),
                                            (4, 1, 'natPred'), (1, 0, 'sinFisher')])
def test_warp_kernel_one_step_with_predator(N, kill, move, mode):
    """Ensembles of small swarms with one predator each run inside warps (swarm_warp_kernel<.., PRED>):
    every replicate against its own oracle world."""
    R = 24
    sp, _, _ = make_params(mode, N, BC=0, size=100.0, Npred=1, pred_kill=kill, pred_move=move, pred_time=0.0,
                           time=50.0, N_confu=50, kill_rate=4000.0 if kill == 4 else 400.0)
    rng = np.random.default_rng(N * 11 + kill)
    sts = []
    for r in range(R):
        st = random_state(N, sp, rng, spread=12.0)
        st['bin_step'] = np.where(rng.random(N) < 0.4, sp.burst_steps, st['bin_step']).astype(np.uint32)
        sts.append(st)
    px = np.full((R, 1), 50.0) + rng.uniform(-3, 3, (R, 1))
    py = np.full((R, 1), 50.0) + rng.uniform(-3, 3, (R, 1))
    pphi = rng.uniform(0, 2 * np.pi, (R, 1))
    g = _swarm(sp, N, 1, n_replicates=R, seed=300)
    g.set_state(**{k: np.stack([s[k] for s in sts]) for k in sts[0]})
    g.set_predators(px, py, pphi, 1)
    g.step(6, 1)
    assert g.stats()['block_launches'] == 1
    gs = g.get_state()
    gp = g.get_predators()
    gc = g.counts()
    killed = mism = 0
    for r in range(R):
        w = pyorc.World(sp, N, 1, seed=300, replicate=r)
        w.set_state(**sts[r])
        w.set_predators(px[r], py[r], pphi[r], 1)
        w.step(6, 1)
        o = w.get_state()
        assert np.array_equal(o['alive'], gs['alive'][r]), (r, np.where(o['alive'] != gs['alive'][r])[0])
        al = o['alive'].astype(bool)
        killed += N - int(al.sum())
        assert np.array_equal(o['bin_step'][al], gs['bin_step'][r][al])
        assert np.array_equal(o['steps_till_burst'][al], gs['steps_till_burst'][r][al])
        ok = (np.abs(o['x'] - gs['x'][r]) <= 1e-5 * np.maximum(1, np.abs(o['x']))) & \
             (np.abs(o['y'] - gs['y'][r]) <= 1e-5 * np.maximum(1, np.abs(o['y']))) & \
             (np.hypot(o['fx'] - gs['fx'][r], o['fy'] - gs['fy'][r]) <= 2e-5 * np.maximum(1, np.hypot(o['fx'], o['fy'])))
        mism += int((~ok[al]).sum())
        po = w.get_predators()
        # position to the one-step tolerance; the chase heading comes from an FP32 difference of positions
        # (|d| ~ 1 at coordinates ~ 50), so speed * (cos, sin) carries ~1e-5 rad * speed
        assert np.allclose(po[:, :2], gp[r][:, :2], rtol=1e-5, atol=1e-4), (r, po, gp[r])
        assert np.allclose(po[:, 2:4], gp[r][:, 2:4], atol=2e-3), (r, po, gp[r])
        assert w.counts() == (int(gc[0][r]), int(gc[1][r]))
        w.close()
    g.close()
    assert killed > 0
    assert mism <= 3, mism


def test_warp_kernel_spawn_like_oracle():
    N, R = 30, 6
    sp, _, _ = make_params('natPred', N, BC=0, size=100.0, Npred=1, pred_kill=0, pred_move=1, pred_time=0.01, time=50.0)
    rng = np.random.default_rng(8)
    sts = [random_state(N, sp, rng, spread=20.0) for _ in range(R)]
    g = _swarm(sp, N, 1, n_replicates=R, seed=9, replicate_offset=2)
    g.set_state(**{k: np.stack([s[k] for s in sts]) for k in sts[0]})
    g.step(0, 12)   # spawn at s = 10
    gp = g.get_predators()
    for r in range(R):
        w = pyorc.World(sp, N, 1, seed=9, replicate=2 + r)
        w.set_state(**sts[r])
        w.step(0, 12)
        po = w.get_predators()
        assert np.allclose(po[:, :2] % 100.0, gp[r][:, :2] % 100.0, atol=2e-3), (r, po, gp[r])
        assert np.allclose(po[:, 2:4], gp[r][:, 2:4], atol=2e-3)
        w.close()
    g.close()


@pytest.mark.parametrize('N,bc', [(30, 0), (16, 0), (30, 5)])
def test_warp_and_block_predator_kernels_agree(N, bc):
    """The in-warp predator path and the one-block-per-swarm kernel (test hook 2) run the same arithmetic:
    same dead sets and timers over a run through spawn, chase and kills."""
    R = 32
    L = 100.0 if bc == 0 else 60.0
    sp, _, _ = make_params('natPred', N, BC=bc, size=L, Npred=1, pred_kill=3, pred_move=1,
                           pred_time=0.0, time=50.0, pred_speed0=20.0, kill_rate=50.0, N_confu=100)
    rng = np.random.default_rng(N + bc)
    sts = [random_state(N, sp, rng, spread=15.0) for _ in range(R)]
    state = {k: np.stack([s[k] for s in sts]) for k in sts[0]}
    c = 50.0 if bc == 0 else 0.0
    px, py = c + rng.uniform(-4, 4, (R, 1)), c + rng.uniform(-4, 4, (R, 1))
    pphi = rng.uniform(0, 2 * np.pi, (R, 1))
    out = []
    for hook in (0, 2):
        g = _swarm(sp, N, 1, n_replicates=R, seed=17)
        g.force_multikernel(hook)
        g.set_state(**state)
        g.set_predators(px, py, pphi, 1)
        g.step(3, 150)
        out.append((g.get_state(), g.get_predators(), g.counts()))
        g.close()
    a, b = out
    assert int(a[2][1].sum()) > 0, 'nobody was killed'
    same = np.array_equal(a[0]['alive'], b[0]['alive'])
    frac = (a[0]['alive'] == b[0]['alive']).mean()
    assert frac >= 0.98, frac
    if same:
        al = a[0]['alive'].astype(bool)
        assert np.array_equal(a[0]['steps_till_burst'][al], b[0]['steps_till_burst'][al])
        assert np.allclose(a[0]['x'][al], b[0]['x'][al], atol=5e-2)


@pytest.mark.parametrize('N,R,hook', [(30, 8, 0), (30, 8, 2), (200, 3, 0), (520, 4, 0), (1100, 2, 0)])
def test_predator_run_independent_of_launch_splits_and_sharding(N, R, hook):
    """natPred through spawn, chase and kills on the warp (N=30), block (hook 2, N=200) and cluster (N>256)
    kernels: (a) one launch == many short launches, bit for bit - the cluster kernel finishes the predator phase
    of a launch's last step before it returns; (b) replicate r of an ensemble == a one-replicate handle
    created with replicate_offset = r."""
    sp, _, _ = make_params('natPred', N, BC=0, size=100.0, Npred=1, pred_kill=3, pred_move=1, pred_time=0.01,
                           time=50.0, pred_speed0=40.0, kill_rate=200.0, N_confu=5000)
    rng = np.random.default_rng(N)
    sts = [random_state(N, sp, rng, spread=30.0) for _ in range(R)]
    state = {k: np.stack([s[k] for s in sts]) for k in sts[0]}

    def run(chunks, r0=0, r1=R):
        g = _swarm(sp, N, 1, n_replicates=r1 - r0, seed=23, replicate_offset=r0)
        g.force_multikernel(hook)
        g.set_state(**{k: v[r0:r1] for k, v in state.items()})
        s = 0
        for c in chunks:
            g.step(s, c)
            s += c
        out = (g.get_state(), g.get_predators(), g.counts())
        g.close()
        return out

    total = 600
    one = run([total])
    many = run([7, 1, 92] + [50] * 10)
    assert int(one[2][1].sum()) > 0, 'nobody was killed'
    for k in one[0]:
        assert np.array_equal(one[0][k], many[0][k], equal_nan=True), k
    assert np.array_equal(one[1], many[1])
    assert np.array_equal(one[2][1], many[2][1])
    r = R - 1
    single = run([total], r, r + 1)
    for k in one[0]:
        assert np.array_equal(np.atleast_2d(one[0][k])[r], np.atleast_2d(single[0][k])[0], equal_nan=True), k
    assert np.array_equal(one[1][r], single[1][0])


@pytest.mark.parametrize('N,npred', [(2600, 50), (4000, 100)])
def test_mid_size_swarm_one_step_with_fish_net(N, npred):
    """Single swarms above 2048 agents with a fish net (multi-kernel path: cell-list pair pass, detection, update
    with the fused net kill, net move) against the oracle."""
    sp, st, rng = _setup('mulFisher', N, npred, 1, 0, L=200.0, seed=N)
    st['bin_step'] = np.where(rng.random(N) < 0.4, sp.burst_steps, st['bin_step']).astype(np.uint32)
    px = 90.0 + 0.4 * np.arange(npred, dtype=np.float64)
    py = np.full(npred, 96.0)
    pphi = np.full(npred, np.pi / 2)
    killed = 0
    for trial in range(2):
        w = pyorc.World(sp, N, npred, seed=400 + trial, replicate=0)
        g = _swarm(sp, N, npred, seed=400 + trial)
        for obj in (w, g):
            obj.set_state(**st)
            obj.set_predators(px, py, pphi, 1)
        w.step(7 + trial, 1)
        g.step(7 + trial, 1)
        o, gs = w.get_state(), g.get_state()
        assert g.stats()['update_launches'] == 1
        assert np.array_equal(o['alive'], gs['alive']), np.where(o['alive'] != gs['alive'])[0]
        killed += N - int(o['alive'].sum())
        al = o['alive'].astype(bool)
        assert np.array_equal(o['bin_step'][al], gs['bin_step'][al])
        assert np.array_equal(o['steps_till_burst'][al], gs['steps_till_burst'][al])
        ok = (np.abs(o['x'] - gs['x']) <= 1e-5 * np.maximum(1, np.abs(o['x']))) & \
             (np.abs(o['y'] - gs['y']) <= 1e-5 * np.maximum(1, np.abs(o['y'])))
        assert (~ok[al]).sum() <= 4 + N // 1000
        assert np.allclose(w.get_predators(), g.get_predators()[0], rtol=1e-5, atol=1e-4)
        assert w.counts() == tuple(int(v[0]) for v in g.counts())
        w.close()
        g.close()
    assert killed > 0


def test_mid_size_net_run_launch_splits_and_graphs():
    """mulFisher run through net spawn, re-spawn and kills: one sbc_step call == many short ones (bit for bit);
    graph replay (non-default stream) and plain launches give the same result."""
    N, npred = 4000, 100
    sp, _, _ = make_params('mulFisher', N, BC=0, size=200.0, Npred=npred, pred_kill=1, pred_move=0, pred_time=0.01,
                           time=50.0, pred_speed0=60.0)
    rng = np.random.default_rng(6)
    st = random_state(N, sp, rng, spread=80.0)

    def run(chunks, stream):
        g = _swarm(sp, N, npred, seed=31, stream=stream)
        g.set_state(**st)
        s = 0
        for c in chunks:
            g.step(s, c)
            s += c
        g.synchronize()
        out = (g.get_state(), g.get_predators(), g.counts(), g.stats())
        g.close()
        return out

    import torch
    ts = torch.cuda.Stream()
    one = run([500], ts.cuda_stream)          # CUDA-graph replay of the step
    many = run([3, 9, 88] + [100] * 4, ts.cuda_stream)
    plain = run([500], None)                  # handle-owned stream... also graph-capable; legacy stream is not
    assert one[3]['update_launches'] == 500
    assert int(one[2][1][0]) > 0, 'the net should have caught somebody'
    for other in (many, plain):
        for k in one[0]:
            assert np.array_equal(one[0][k], other[0][k], equal_nan=True), k
        assert np.array_equal(one[1], other[1])
        assert np.array_equal(one[2][1], other[2][1])
