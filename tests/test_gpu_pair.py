"""GPU parity of the three-zone pair pass against the CPU oracle, through the C ABI.

Bar: zone counters and neighbour sets bit-exact; accumulated forces within
||d|| <= 4 * 2^-24 * sqrt(n) * max(1, sum|terms|) for n contributing pairs.
"""
import numpy as np
import pytest

from helpers import make_params, random_state, nn_sets, f32
from oracle import pyorc
from sde_burst_coast_b200.params import PAIR_ACTIVE_ROWS, PAIR_ALL_ROWS

pytestmark = pytest.mark.gpu


KINDS = ['port', 'reference']


def _need(kind):
    if kind == 'reference' and not pyorc.have_reference():
        pytest.skip('oracle/_ref/liborc_ref.so not built (needs /root/reference at build time)')


def _compare(sp, N, st, flags, alive=None, kind='port'):
    """CUDA path against the CPU side of `kind`: the restatement, or the reference's own IntCalcPrey / SFM_4Zone
    driven by InteractionGlobal (oracle/_ref)."""
    from sde_burst_coast_b200.swarm import Swarm
    _need(kind)
    w = pyorc.World(sp, N, 0, kind=kind)
    if alive is not None:
        st = dict(st, alive=alive)
    w.set_state(**st)
    ref = w.pair_interactions(flags)
    g = Swarm(sp, N)
    g.set_state(**st)
    got = g.pair_interactions(flags)
    for k in ('c_rep', 'c_alg', 'c_att'):
        assert np.array_equal(ref[k], got[k]), k
    assert np.array_equal(ref['nn_ptr'], got['nn_ptr'])
    assert nn_sets(ref, N) == nn_sets(got, N)
    eps = 2.0 ** -24
    for k, c, scale in (('f_rep', 'c_rep', 1.0), ('f_att', 'c_att', 1.0), ('f_alg', 'c_alg', 25.0)):
        n = ref[c].astype(np.float64)
        tol = 4 * eps * np.sqrt(np.maximum(n, 1)) * np.maximum(1.0, n * scale)
        err = np.linalg.norm(ref[k] - got[k], axis=1)
        assert np.all(err <= tol), (k, float(err.max()), float(tol[np.argmax(err)]))
    g.close()
    w.close()
    return ref


@pytest.mark.parametrize('mode,BC,L,N', [
    ('burst_coast', 6, 24.5, 8), ('burst_coast', 6, 24.5, 30), ('burst_coast', 6, 24.5, 64),
    ('burst_coast', 5, 24.5, 33), ('burst_coast', -1, 0, 257),
    ('natPred', 0, 100.0, 300), ('natPred', 0, 584.0, 1024), ('natPred', 0, 100.0, 1500),
])
@pytest.mark.parametrize('alg', [None, 10.0])
@pytest.mark.parametrize('flags', [PAIR_ALL_ROWS, PAIR_ACTIVE_ROWS])
@pytest.mark.parametrize('kind', KINDS)
def test_pair_parity(mode, BC, L, N, alg, flags, kind):
    over = dict(BC=BC, Npred=0)
    if L:
        over['size'] = L
    if alg is not None:
        over['alg_range'] = alg
    sp, _, _ = make_params(mode, N, **over)
    rng = np.random.default_rng(1000 + N + BC)
    st = random_state(N, sp, rng, spread=60.0 if BC == -1 else None)
    ref = _compare(sp, N, st, flags, kind=kind)
    if flags == PAIR_ALL_ROWS:
        assert ref['c_rep'].sum() + ref['c_att'].sum() > 0


@pytest.mark.parametrize('kind', KINDS)
def test_pair_threshold_ties_and_coincident(kind):
    """Distances exactly on the closed upper bounds, coincident agents, a lone agent."""
    sp, _, _ = make_params('burst_coast', 8, BC=-1, rep_range=4.0, alg_range=8.0, att_range=16.0)
    x = np.array([0.0, 4.0, 8.0, 16.0, 0.0, -4.0000005, 8.000001, 1000.0])
    y = np.zeros(8)
    st = dict(x=f32(x), y=f32(y), phi=np.zeros(8), vproj=np.ones(8))
    ref = _compare(sp, 8, st, PAIR_ALL_ROWS, kind=kind)
    assert ref['c_rep'][0] == 2 and ref['c_alg'][0] == 2 and ref['c_att'][0] == 2  # ties are inside
    assert ref['c_rep'][7] + ref['c_alg'][7] + ref['c_att'][7] == 0


def test_pair_dead_agents_and_single():
    sp, _, _ = make_params('natPred', 40, BC=0, size=100.0, Npred=0)
    rng = np.random.default_rng(7)
    st = random_state(40, sp, rng, spread=30.0)
    alive = (rng.random(40) > 0.3).astype(np.uint8)
    _compare(sp, 40, st, PAIR_ALL_ROWS, alive=alive)
    sp1, _, _ = make_params('burst_coast', 1)
    _compare(sp1, 1, random_state(1, sp1, rng), PAIR_ALL_ROWS)


def test_pair_band_recheck_exercised():
    """Pairs placed within a few FP32 ulps of every threshold must come out exactly as in FP64."""
    sp, _, _ = make_params('natPred', 8, BC=0, size=100.0, Npred=0, alg_range=10.0)
    rng = np.random.default_rng(11)
    N = 512
    thr = np.array([sp.rep_range, sp.alg_range, sp.att_range])
    x = np.zeros(N)
    y = np.zeros(N)
    x[0], y[0] = 50.0, 50.0
    ang = 2 * np.pi * rng.random(N)
    r = thr[rng.integers(0, 3, N)] * (1 + (rng.integers(-6, 7, N)) * 2.0 ** -24)
    x[1:], y[1:] = 50.0 + (r * np.cos(ang))[1:], 50.0 + (r * np.sin(ang))[1:]
    st = dict(x=f32(x % 100.0), y=f32(y % 100.0), phi=np.zeros(N), vproj=np.ones(N))
    from sde_burst_coast_b200.swarm import Swarm
    sp2, _, _ = make_params('natPred', N, BC=0, size=100.0, Npred=0, alg_range=10.0)
    _compare(sp2, N, st, PAIR_ALL_ROWS)
    g = Swarm(sp2, N)
    g.set_state(**st)
    g.pair_interactions(PAIR_ALL_ROWS, want_nn=False)
    assert g.stats()['ambiguous_pairs'] > 0
    g.close()


@pytest.mark.parametrize('BC,L,N,spread', [(0, 300.0, 3000, None), (0, 100.0, 2500, None), (0, 61.0, 500, None),
                                          (-1, 0, 2000, 400.0), (6, 150.0, 2000, None)])
@pytest.mark.parametrize('alg', [None, 10.0])
def test_cell_list_path_matches_oracle(BC, L, N, spread, alg):
    """Burst-gated rows through the cell list (celllist.cu): same integers as the all-pairs oracle."""
    from sde_burst_coast_b200 import params as PP
    from sde_burst_coast_b200.swarm import Swarm
    from helpers import base_dict
    over = dict(BC=BC, Npred=0)
    if L:
        over['size'] = L
    if alg is not None:
        over['alg_range'] = alg
    d = base_dict('natPred' if BC == 0 else 'burst_coast', N, **over)
    sp, _ = PP.make_sbc_params(d, path=PP.PATH_CELLLIST)
    rng = np.random.default_rng(N + BC)
    st = random_state(N, sp, rng, spread=spread)
    st['bin_step'] = np.where(rng.random(N) < 0.3, sp.burst_steps, st['bin_step']).astype(np.uint32)
    w = pyorc.World(sp, N, 0, kind='port')
    w.set_state(**st)
    ref = w.pair_interactions(PAIR_ACTIVE_ROWS)
    g = Swarm(sp, N)
    g.set_state(**st)
    got = g.pair_interactions(PAIR_ACTIVE_ROWS)
    assert g.stats()['cell_builds'] == 1
    for k in ('c_rep', 'c_alg', 'c_att', 'nn_ptr', 'nn_idx'):
        assert np.array_equal(ref[k], got[k]), k
    assert ref['c_att'].sum() > 0
    for k in ('f_rep', 'f_alg', 'f_att'):
        assert np.allclose(ref[k], got[k], rtol=2e-5, atol=2e-5), k
    # candidates examined are far fewer than all pairs once the grid has more than 3 x 3 cells
    if L >= 300.0 or BC != 0:
        assert g.stats()['pairs'] < 0.5 * (st['bin_step'] == sp.burst_steps).sum() * (N - 1)
    g.close()
    w.close()


def test_cell_list_steps_match_all_pairs_path():
    from sde_burst_coast_b200 import params as PP
    from sde_burst_coast_b200.swarm import Swarm
    from helpers import base_dict
    N = 9000
    d = base_dict('natPred', N, BC=0, size=300.0, Npred=0)
    rng = np.random.default_rng(4)
    out = []
    for path in (PP.PATH_CELLLIST, PP.PATH_ALLPAIRS):
        sp, _ = PP.make_sbc_params(d, path=path)
        if not out:
            st = random_state(N, sp, rng)
        g = Swarm(sp, N, seed=8)
        g.set_state(**st)
        g.step(0, 80)
        out.append((g.get_state(), g.stats()))
        g.close()
    a, b = out[0][0], out[1][0]
    # one cell build serves several steps (Verlet skin); the all-pairs path never builds
    assert 2 <= out[0][1]['cell_builds'] < 40 and out[1][1]['cell_builds'] == 0
    assert np.array_equal(a['bin_step'], b['bin_step']) and np.array_equal(a['steps_till_burst'], b['steps_till_burst'])
    close = (np.abs(a['x'] - b['x']) < 1e-3) & (np.abs(a['y'] - b['y']) < 1e-3)
    assert close.mean() > 0.99


@pytest.mark.parametrize('mode,over,N', [('natPred', dict(BC=0, size=100.0, Npred=0, alg_range=10.0), 300),
                                        ('burst_coast', dict(BC=6, size=60.0), 200), ('burst_coast', dict(BC=-1), 150)])
def test_dense_pass_on_an_ensemble(mode, over, N):
    """All rows of several swarms in one launch (every swarm put in Hilbert order by one 64-bit-key sort): integers exact
    per swarm against the oracle."""
    from sde_burst_coast_b200.swarm import Swarm
    R = 5
    sp, _, _ = make_params(mode, N, **over)
    rng = np.random.default_rng(31)
    sts = [random_state(N, sp, rng, spread=60.0 if sp.BC == -1 else None) for _ in range(R)]
    alive = (rng.random((R, N)) > 0.15).astype(np.uint8)
    g = Swarm(sp, N, n_replicates=R)
    g.set_state(alive=alive, **{k: np.stack([s[k] for s in sts]) for k in sts[0]})
    got = g.pair_interactions(PAIR_ALL_ROWS)
    g.close()
    for r in range(R):
        w = pyorc.World(sp, N)
        w.set_state(alive=alive[r], **sts[r])
        ref = w.pair_interactions(PAIR_ALL_ROWS)
        w.close()
        for k in ('c_rep', 'c_alg', 'c_att'):
            assert np.array_equal(ref[k], got[k][r * N:(r + 1) * N]), (r, k)
        lo, hi = got['nn_ptr'][r * N], got['nn_ptr'][(r + 1) * N]
        assert np.array_equal(ref['nn_idx'], got['nn_idx'][lo:hi]), r
        for k in ('f_rep', 'f_att'):
            assert np.allclose(ref[k], got[k][r * N:(r + 1) * N], rtol=2e-5, atol=2e-5), (r, k)
