"""GPU parity of the whole step (interaction + burst decision + burst-coast update + boundary) against
the CPU oracle, through the C ABI.

Stated tolerances (FP32 device vs FP64 oracle on FP32-representable states):
  one step with identical decisions : |dx| <= 1e-5 max(1,|x|), |dvproj| <= 1e-5 max(1,vproj),
                                      |dphi| <= 2e-5 rad (mod 2 pi), force within 1e-5 relative;
  burst timers (bin_step, steps_till_burst) : bit-exact, for any number of steps;
  branch decisions : identical except inside a documented guard band; mismatches are counted and
                     bounded, never hidden.
"""
import numpy as np
import pytest

from helpers import make_params, random_state, ang_diff, cancellation as _cancellation
from oracle import pyorc

pytestmark = pytest.mark.gpu

TOL_X, TOL_V, TOL_PHI, TOL_F = 1e-5, 1e-5, 2e-5, 2e-5


def _swarm(*a, **k):
    from sde_burst_coast_b200.swarm import Swarm
    return Swarm(*a, **k)


def _close_state(o, g, sp, what='', cond=None):
    """o: oracle state, g: GPU state (same shapes). Returns the mask of agents that agree. `cond` (from
    _cancellation) widens the bars of rows whose social force is an ill-conditioned sum."""
    fmag = np.maximum(1.0, np.hypot(o['fx'], o['fy']))
    extra = np.zeros_like(fmag) if cond is None else cond
    kick = extra * fmag * sp.dt     # what the force error does to speed / heading / position within one step
    okx = (np.abs(o['x'] - g['x']) <= TOL_X * np.maximum(1, np.abs(o['x'])) + kick) & \
          (np.abs(o['y'] - g['y']) <= TOL_X * np.maximum(1, np.abs(o['y'])) + kick)
    okv = np.abs(o['vproj'] - g['vproj']) <= TOL_V * np.maximum(1, np.abs(o['vproj'])) + kick
    okp = ang_diff(o['phi'], g['phi']) <= TOL_PHI + kick
    okf = np.hypot(o['fx'] - g['fx'], o['fy'] - g['fy']) <= (TOL_F + extra) * fmag
    # the lab-frame velocity (what Boundary<> edits at a wall and particle::out() reports): heading error times speed
    vmag = np.maximum(1.0, np.hypot(o['vx'], o['vy']))
    okw = np.hypot(o['vx'] - g['vx'], o['vy'] - g['vy']) <= (TOL_PHI + TOL_V) * vmag + kick * (1 + vmag)
    return okx & okv & okp & okf & okw


def _send_through_walls(st, sp, rng):
    """Box boundaries (Boundary<> cases 1-4, /root/reference/src/agents_dynamics.cpp:300-430): put a third of the agents
    a fraction of one step's travel inside a wall (or a corner), heading outwards, so that this step crosses it."""
    from helpers import f32
    N, L = st['x'].size, sp.sizeL
    pick = np.where(rng.random(N) < 0.34)[0]
    for i in pick:
        walls = rng.integers(0, 4, size=1 + (rng.random() < 0.25))   # sometimes two walls: a corner
        vp = 8.0 + 10 * rng.random()
        reach = vp * sp.dt
        ang = None
        for w_ in walls:
            if w_ == 0: st['x'][i] = L - 0.3 * reach; ang = 0.0
            if w_ == 1: st['x'][i] = 0.3 * reach; ang = np.pi
            if w_ == 2: st['y'][i] = L - 0.3 * reach; ang = 0.5 * np.pi
            if w_ == 3: st['y'][i] = 0.3 * reach; ang = -0.5 * np.pi
        if len(walls) == 2 and walls[0] // 2 != walls[1] // 2:
            ang = np.arctan2(1.0 if walls[1] == 2 else -1.0, 1.0 if walls[0] == 0 else -1.0)
        st['phi'][i] = ang + 0.4 * (rng.random() - 0.5)
        st['vproj'][i] = vp
        st['bin_step'][i] = 0         # coasting: the heading does not turn during this step
        st['fx'][i] = st['fy'][i] = 0.0
        st['steps_till_burst'][i] = max(int(st['steps_till_burst'][i]), 5)
    for k in ('x', 'y', 'phi', 'vproj'):
        st[k] = f32(st[k])
    st['vx'] = f32(st['vproj'] * np.cos(st['phi']))
    st['vy'] = f32(st['vproj'] * np.sin(st['phi']))


@pytest.mark.parametrize('mode,over,N', [
    ('burst_coast', dict(BC=6), 8), ('burst_coast', dict(BC=6), 30), ('burst_coast', dict(BC=5), 64),
    ('burst_coast', dict(BC=6, alg_range=10.0), 100), ('burst_coast', dict(BC=-1), 40),
    ('natPred', dict(BC=0, size=100.0, Npred=0), 300),
    ('natPred', dict(BC=0, size=100.0, Npred=0, alg_range=10.0), 1024),
    ('natPred', dict(BC=0, size=200.0, Npred=0), 2000),
    ('natPred', dict(BC=0, size=300.0, Npred=0), 3000),            # cell-list pair pass from N > 2048
    ('burst_coast', dict(BC=5, size=120.0), 2500),
    ('burst_coast', dict(BC=2, size=60.0), 50), ('burst_coast', dict(BC=3, size=60.0), 50),
    ('burst_coast', dict(BC=1, size=60.0), 50), ('burst_coast', dict(BC=4, size=60.0), 50),   # inelastic box / x periodic + y inelastic
    ('burst_coast', dict(BC=1, size=60.0), 300), ('burst_coast', dict(BC=4, size=60.0), 300),
])
@pytest.mark.parametrize('multikernel', [False, True, 2])   # automatic (warp / block / cluster kernel), multi-kernel, block kernel
@pytest.mark.parametrize('kind', ['port', 'reference'])
def test_one_step_injected(mode, over, N, multikernel, kind):
    """kind: the CPU side is the restatement (`port`) or the reference's own translation units compiled in place
    (`reference`, oracle/_ref - present wherever build() ran with /root/reference mounted; the library travels)."""
    if kind == 'reference' and not pyorc.have_reference():
        pytest.skip('oracle/_ref/liborc_ref.so not built (needs /root/reference at build time)')
    sp, _, _ = make_params(mode, N, **over)
    rng = np.random.default_rng(N * 3 + sp.BC + 100)
    bad_total = 0
    flag_mismatch = 0
    hits_total = 0
    trials = 6
    for t in range(trials):
        wide = sp.BC in (5, 6) and t % 2 == 1
        st = random_state(N, sp, rng, spread=(0.97 * sp.sizeL if wide else (50.0 if sp.BC in (-1, 1, 2, 3, 4) else None)))
        if sp.BC in (1, 2, 3, 4):
            _send_through_walls(st, sp, rng)
        us, ud = rng.random(N), rng.random(N)
        ud = np.asarray(ud, dtype=np.float32).astype(np.float64)
        kn = rng.integers(1, 600, N).astype(np.uint32)
        w = pyorc.World(sp, N)
        w.set_state(**st)
        cond = _cancellation(w.pair_interactions(0, want_nn=False))   # burst-starting rows of this state
        w.inject(us, ud, kn)
        fo = w.step(0, 1, want_flags=True)
        o = w.get_state()
        if kind == 'reference':
            # the reference's ParticleBurstCoast does not tell which internal branch it took (wall redirect, clamp,
            # boundary hit): the branch flags come from the restatement, the STATE the GPU is held to comes from the
            # reference's own translation units - and the two must be the same doubles
            wr = pyorc.World(sp, N, kind='reference')
            wr.set_state(**st)
            wr.inject(us, ud, kn)
            fr = wr.step(0, 1, want_flags=True)
            o_ref = wr.get_state()
            wr.close()
            assert np.array_equal(fr, fo & 195)   # burst start, social cue, flee cue, re-arm: the observable flags
            for k in o:
                assert np.array_equal(o[k], o_ref[k], equal_nan=True), 'port and reference differ in ' + k
            o = o_ref
        g = _swarm(sp, N)
        if multikernel:
            g.force_multikernel(multikernel)
        g.debug_flags(True)
        g.set_state(**st)
        g.inject(us, ud, kn)
        g.step(0, 1)
        gs = g.get_state()
        fg = g.debug_flags(True, read=True)
        assert np.array_equal(o['bin_step'], gs['bin_step'])
        assert np.array_equal(o['steps_till_burst'], gs['steps_till_burst'])
        ok = _close_state(o, gs, sp, cond=cond)
        same_flags = (fo == fg)
        hits_total += int(((fo & 32) != 0).sum())
        # an agent may only disagree if one of its branch decisions flipped (guard band)
        assert np.all(ok | ~same_flags), 'state differs although every branch agreed: %s' % np.where(~ok & same_flags)[0]
        bad_total += int((~ok).sum())
        flag_mismatch += int((~same_flags).sum())
        g.close()
        w.close()
    # branch flips are rare events (ties within FP32 rounding of a threshold)
    assert flag_mismatch <= max(1, trials * N // 200), (flag_mismatch, bad_total)
    if sp.BC in (1, 2, 3, 4):
        assert hits_total >= trials * N // 10, 'the wall cases of Boundary<> were not exercised'


@pytest.mark.parametrize('N,R', [(8, 64), (16, 32), (30, 16), (64, 8), (200, 3)])
def test_timers_exact_over_many_steps(N, R):
    """bin_step / steps_till_burst depend only on the Philox protocol: exact after 3000 steps."""
    sp, _, _ = make_params('burst_coast', N)
    rng = np.random.default_rng(N)
    sts = [random_state(N, sp, rng, mid_burst=False) for _ in range(R)]
    g = _swarm(sp, N, n_replicates=R, seed=42, replicate_offset=5)
    g.set_state(**{k: np.stack([s[k] for s in sts]) for k in sts[0]})
    g.step(0, 3000)
    gs = g.get_state()
    for r in range(min(R, 4)):
        w = pyorc.World(sp, N, seed=42, replicate=5 + r)
        w.set_state(**sts[r])
        w.step(0, 3000)
        o = w.get_state()
        assert np.array_equal(o['bin_step'], np.atleast_2d(gs['bin_step'])[r])
        assert np.array_equal(o['steps_till_burst'], np.atleast_2d(gs['steps_till_burst'])[r])
    assert np.all(np.isfinite(gs['x'])) and np.all(np.hypot(gs['x'], gs['y']) <= sp.sizeL * (1 + 1e-6))
    g.close()


def test_replicates_independent_of_sharding():
    """Replicate r of an ensemble handle == a single-replicate handle created with
    replicate_offset = r (bit-exact): results do not depend on how replicates are split over GPUs."""
    N, R = 16, 12
    sp, _, _ = make_params('burst_coast', N)
    rng = np.random.default_rng(77)
    sts = [random_state(N, sp, rng) for _ in range(R)]
    g = _swarm(sp, N, n_replicates=R, seed=9)
    g.set_state(**{k: np.stack([s[k] for s in sts]) for k in sts[0]})
    g.step(0, 500)
    full = g.get_state()
    g.close()
    for r0, r1 in ((0, 5), (5, 12)):
        h = _swarm(sp, N, n_replicates=r1 - r0, seed=9, replicate_offset=r0)
        h.set_state(**{k: np.stack([s[k] for s in sts[r0:r1]]) for k in sts[0]})
        h.step(0, 500)
        part = h.get_state()
        for k in full:
            assert np.array_equal(full[k][r0:r1], part[k]), k
        h.close()


def test_split_launches_equal_one_launch():
    """sbc_step(s, a) ; sbc_step(s+a, b) == sbc_step(s, a+b) bit-exactly (state lives in FP32 SoA)."""
    N, R = 30, 4
    sp, _, _ = make_params('burst_coast', N)
    rng = np.random.default_rng(1)
    sts = [random_state(N, sp, rng) for _ in range(R)]
    state = {k: np.stack([s[k] for s in sts]) for k in sts[0]}
    a = _swarm(sp, N, n_replicates=R, seed=3)
    b = _swarm(sp, N, n_replicates=R, seed=3)
    a.set_state(**state)
    b.set_state(**state)
    a.step(0, 700)
    for s in range(0, 700, 70):
        b.step(s, 70)
    sa, sb = a.get_state(), b.get_state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    a.close()
    b.close()


@pytest.mark.parametrize('N', [30, 64, 300, 1500])
def test_block_path_matches_multikernel_path(N):
    """The one-block-per-swarm kernel and the pair-pass + update kernels are two schedules of the same
    arithmetic: identical timers, states within the one-step tolerance for a short run."""
    over = dict(BC=0, size=100.0, Npred=0) if N > 64 else dict(BC=6)
    sp, _, _ = make_params('natPred' if N > 64 else 'burst_coast', N, **over)
    rng = np.random.default_rng(N)
    st = random_state(N, sp, rng)
    a = _swarm(sp, N, seed=11)
    b = _swarm(sp, N, seed=11)
    b.force_multikernel(True)
    a.set_state(**st)
    b.set_state(**st)
    a.step(0, 3)
    b.step(0, 3)
    sa, sb = a.get_state(), b.get_state()
    assert np.array_equal(sa['bin_step'], sb['bin_step'])
    assert np.array_equal(sa['steps_till_burst'], sb['steps_till_burst'])
    ok = _close_state(sa, sb, sp)
    assert ok.mean() >= 0.97, ok.mean()
    assert a.stats()['block_launches'] == 1 and b.stats()['update_launches'] == 3
    if N > 256:   # the automatic path is the cluster kernel there: check the one-block kernel as well
        c = _swarm(sp, N, seed=11)
        c.force_multikernel(2)
        c.set_state(**st)
        c.step(0, 3)
        sc = c.get_state()
        assert np.array_equal(sa['steps_till_burst'], sc['steps_till_burst'])
        assert _close_state(sa, sc, sp).mean() >= 0.97
        c.close()
    a.close()
    b.close()


def test_trajectory_divergence_reported():
    """Many steps with the shared Philox decisions: report steps-to-divergence (chaotic once a
    branch flips); the first 100 steps of a tank must still agree to 1e-3."""
    N = 8
    sp, _, _ = make_params('burst_coast', N)
    rng = np.random.default_rng(2)
    st = random_state(N, sp, rng, mid_burst=False)
    w = pyorc.World(sp, N, seed=21, replicate=0)
    w.set_state(**st)
    g = _swarm(sp, N, seed=21)
    g.set_state(**st)
    first_div = None
    for s in range(0, 2000, 50):
        w.step(s, 50)
        g.step(s, 50)
        o, gs = w.get_state(), g.get_state()
        err = max(np.abs(o['x'] - gs['x']).max(), np.abs(o['y'] - gs['y']).max())
        if first_div is None and err > 1e-3:
            first_div = s + 50
        assert np.array_equal(o['steps_till_burst'], gs['steps_till_burst'])
    print('steps to 1e-3 divergence:', first_div)
    assert first_div is None or first_div > 100
    g.close()
    w.close()


def test_queue_mode_equals_one_unit_per_warp():
    """Large ensembles are stepped in (tanks, time slice) units handed out through a queue (swarm_block.cu, EnsQueue): the
    result is bit-identical to the classic launch, in which every warp takes its tanks through all the steps."""
    import os
    N, R, K = 32, 6000, 640      # 6000 warps > the 4736 resident ones: the queue mode engages
    sp, _, _ = make_params('burst_coast', N)
    rng = np.random.default_rng(8)
    st = random_state(N * R, sp, rng)
    state = {k: v.reshape(R, N) for k, v in st.items()}
    out = []
    for env in ('1', '0'):
        os.environ['SBC_ENS_QUEUE'] = env
        g = _swarm(sp, N, n_replicates=R, seed=3)
        g.set_state(**state)
        g.step(0, K)
        g.step(K, 100)           # a second launch carries on from the first
        out.append((g.get_state(), g.stats()))
        g.close()
    os.environ.pop('SBC_ENS_QUEUE')
    for k in out[0][0]:
        assert np.array_equal(out[0][0][k], out[1][0][k]), k
    assert out[0][1]['pairs'] == out[1][1]['pairs'] and out[0][1]['agent_steps'] == (K + 100) * N * R
