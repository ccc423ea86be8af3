"""Same-state parity of the device observables with the reference's Out_swarm / Out_swarm_fishNet
(/root/reference/src/swarmdyn.cpp:237-344, :358-410): the GPU (sbc_out_swarm, sbc_out_swarm_fishnet) against the
oracle - the port's restatement, and the reference's own functions when oracle/_ref is present - on identical states.

Bars: N, Ndead, Nfront, NfrontClose exact; selections (NND, aspect ratio, elongation, hull area) 1e-12 relative - the
device makes them in FP64 in the reference's operation order; means 1e-12 (summation order only); IID 1e-6 relative
(distances summed from FP32 square roots); a_com_maxIID 1e-6 absolute (acos near +-1 amplifies summation-order noise);
ND is NaN on both sides whenever a neighbour list is empty (every state here)."""
import numpy as np
import pytest

from helpers import make_params, random_state
from oracle import pyorc

pytestmark = pytest.mark.gpu
KINDS = ['port'] + (['reference'] if pyorc.have_reference() else [])
COLS = ['N', 'pol', 'L_norm', 'x0', 'x1', 'v0', 'v1', 'speed', 'v2', 'elong', 'aspect', 'a_com', 'hull', 'NND', 'IID', 'ND']
RTOL = np.array([0, 1e-12, 1e-10, 1e-12, 1e-12, 1e-12, 1e-12, 1e-12, 1e-12, 1e-12, 1e-12, 0, 1e-12, 1e-12, 1e-6, 0])
ATOL = np.array([0, 1e-14, 1e-12, 1e-12, 1e-12, 1e-12, 1e-12, 0, 0, 0, 0, 1e-6, 1e-9, 0, 0, 0])


def _swarm(*a, **k):
    from sde_burst_coast_b200.swarm import Swarm
    return Swarm(*a, **k)


def _check_row(got, ref, what):
    for k in range(16):
        if np.isnan(ref[k]):
            assert np.isnan(got[k]), (what, COLS[k], got[k])
            continue
        if got[k] == ref[k]:        # also +-inf (a zero-width swarm has an infinite elongation on both sides)
            continue
        assert abs(got[k] - ref[k]) <= ATOL[k] + RTOL[k] * abs(ref[k]), (what, COLS[k], got[k], ref[k])


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('mode,N,R,over', [('burst_coast', 8, 5, {}), ('burst_coast', 32, 7, {}), ('natPred', 30, 4, dict(Npred=0)),
                                           ('natPred', 300, 3, dict(BC=0, size=100.0, Npred=0)),
                                           ('burst_coast', 700, 2, dict(BC=-1)),
                                           ('natPred', 3000, 1, dict(BC=0, size=400.0, Npred=0))])
def test_out_swarm_matches_oracle_on_the_same_state(kind, mode, N, R, over):
    sp, _, _ = make_params(mode, N, **over)
    rng = np.random.default_rng(N + R)
    sts = [random_state(N, sp, rng) for _ in range(R)]
    g = _swarm(sp, N, n_replicates=R)
    g.set_state(**{k: np.stack([s[k] for s in sts]) for k in sts[0]})
    got, extra = g.out_swarm(want_extra=True)
    sub = g.observables()
    for r in range(R):
        w = pyorc.World(sp, N, kind=kind)
        w.set_state(**sts[r])
        _check_row(got[r], w.out_swarm(), (mode, N, r))
        o8 = w.observables()
        assert np.allclose(sub[r], o8, rtol=1e-12, atol=1e-12), (sub[r], o8)
        assert abs(extra[r, 0] - o8[7]) <= 1e-12 * o8[7]            # the true nearest-neighbour mean
        assert abs(extra[r, 1] - 2 * got[r, 14]) <= 1e-12 * extra[r, 1]
    nop = g.out_swarm(no_pairs=True)
    for k in (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 12):
        assert np.array_equal(nop[:, k], got[:, k]), COLS[k]
    assert np.isnan(nop[:, [10, 11, 13, 14, 15]]).all()
    g.close()


@pytest.mark.parametrize('kind', KINDS)
def test_out_swarm_with_dead_agents_and_degenerate_sets(kind):
    """Dead agents drop out and the `j < i-1` quirk follows the compacted order; a lattice has many exactly equal
    distances (first-in-order selection of the most distant pair); collinear agents have a zero-area hull."""
    N = 64
    sp, _, _ = make_params('natPred', N, BC=0, size=60.0, Npred=0)
    rng = np.random.default_rng(11)
    st = random_state(N, sp, rng)
    alive = (rng.random(N) > 0.3).astype(np.uint8)
    alive[:3] = [1, 0, 1]
    lattice = random_state(N, sp, rng)
    lattice['x'], lattice['y'] = 10.0 + 5.0 * (np.arange(N) % 8), 10.0 + 5.0 * (np.arange(N) // 8)
    line = random_state(N, sp, rng)
    line['x'], line['y'] = 2.0 + 0.75 * np.arange(N), np.full(N, 7.0)
    for name, s, al in (('dead', st, alive), ('lattice', lattice, None), ('line', line, None)):
        g = _swarm(sp, N)
        w = pyorc.World(sp, N, kind=kind)
        for o in (g, w):
            o.set_state(alive=al, **s) if al is not None else o.set_state(**s)
        got = g.out_swarm()[0]
        ref = w.out_swarm()
        if name == 'line':     # elongation across a line is x/0: both sides give inf
            assert got[12] == 0.0 and ref[12] == 0.0
        _check_row(got, ref, name)
        g.close()


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('npred', [1, 50])
def test_fishnet_row_matches_oracle(kind, npred):
    N = 400
    sp, np_, _ = make_params('mulFisher' if npred > 1 else 'natPred', N, BC=0, size=100.0, Npred=npred, pred_time=0.0)
    rng = np.random.default_rng(5 + npred)
    st = random_state(N, sp, rng, spread=40.0)
    alive = (rng.random(N) > 0.1).astype(np.uint8)
    px = 40.0 + 0.5 * np.arange(npred, dtype=np.float64)
    py, pphi = np.full(npred, 48.0), np.full(npred, np.pi / 2)
    g = _swarm(sp, N, npred)
    w = pyorc.World(sp, N, npred, kind=kind)
    for o in (g, w):
        o.set_state(alive=alive, **st)
        o.set_predators(px, py, pphi, 1)
    got, ref = g.out_swarm_fishnet()[0], w.out_swarm_fishnet()
    assert got[0] == ref[0] and got[2] == ref[2] and got[3] == ref[3], (got, ref)
    assert abs(got[1] - ref[1]) <= 1e-12 * ref[1]
    g.close()


def test_out_swarm_large_swarm_runs_and_agrees_with_properties():
    """Config-C5-like density at a size the oracle cannot do in seconds: the O(N^2) pass against what the cell-list
    pair pass knows (every agent's nearest neighbour at rho = 0.01 lies well inside att_range), N and means exact."""
    N = 65536
    L = float(np.sqrt(N / 0.01))
    sp, _, _ = make_params('natPred', N, BC=0, size=L, Npred=0)
    rng = np.random.default_rng(9)
    st = random_state(N, sp, rng)
    g = _swarm(sp, N)
    g.set_state(**st)
    out, extra = g.out_swarm(want_extra=True)
    g.close()
    assert out[0, 0] == N
    assert abs(out[0, 3] - st['x'].mean()) < 1e-9 * L and abs(out[0, 7] - np.hypot(st['vx'], st['vy']).mean()) < 1e-9
    # a sample of agents: brute-force nearest neighbour in double
    idx = rng.choice(N, 200, replace=False)
    dx = st['x'][None, :] - st['x'][idx, None]
    dy = st['y'][None, :] - st['y'][idx, None]
    dx -= L * np.round(dx / L); dy -= L * np.round(dy / L)
    d = np.hypot(dx, dy)
    d[np.arange(200), idx] = np.inf
    assert 0.4 / np.sqrt(0.01) < extra[0, 0] < 0.6 / np.sqrt(0.01)    # 0.5 / sqrt(rho) for a Poisson field
    assert abs(d.min(axis=1).mean() - extra[0, 0]) < 0.5               # the sample mean is near the population mean
    assert extra[0, 2] <= L / np.sqrt(2) * (1 + 1e-12)
