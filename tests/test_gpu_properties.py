"""Size-independent properties of the three-zone pass at BASELINE's FULL sizes, where the CPU oracle is
too slow to be the checker (N = 65,536 and N = 1,048,576):
  * two independent CUDA paths - tiled all-pairs (pair_dense.cu) and cell list (celllist.cu) - give
    identical zone counters for every agent;
  * the metric neighbour relation is symmetric: every zone holds an even number of directed pairs, and
    the repulsion / attraction sums cancel over the swarm (u_ij = -u_ji);
  * counters are invariant under a relabelling of the agents and, in a periodic box, under an exactly
    representable translation.
Plus the ensemble at full size: 16,384 tanks x 32 agents stepped on the GPU, sample tanks replayed by the
oracle (burst timers exact)."""
import numpy as np
import pytest

from helpers import base_dict, f32, make_params
from oracle import pyorc
from sde_burst_coast_b200 import params as PP

pytestmark = pytest.mark.gpu


def _counts(sp, N, st, path_all_rows):
    from sde_burst_coast_b200.swarm import Swarm
    g = Swarm(sp, N)
    g.set_state(**st)
    r = g.pair_interactions(1 if path_all_rows else 0, want_nn=False)
    stats = g.stats()
    g.close()
    return r, stats


@pytest.mark.parametrize('N,rho', [(65536, 1.6384), (1 << 20, 0.01)])
def test_dense_and_cell_paths_agree_at_full_size(N, rho):
    L = float(np.sqrt(N / rho))
    rng = np.random.default_rng(N)
    st = dict(x=f32(L * rng.random(N)), y=f32(L * rng.random(N)), phi=f32(2 * np.pi * rng.random(N) - np.pi),
              vproj=np.ones(N))
    d = base_dict('natPred', N, BC=0, size=L, Npred=0, alg_range=10.0)
    sp_dense, _ = PP.make_sbc_params(d, path=PP.PATH_ALLPAIRS)
    sp_cells, _ = PP.make_sbc_params(d, path=PP.PATH_CELLLIST)
    dense, s_dense = _counts(sp_dense, N, st, True)                                  # every row, all N(N-1) pairs
    st_all = dict(st, bin_step=np.full(N, sp_cells.burst_steps, dtype=np.uint32))    # every row starts a burst
    cells, s_cells = _counts(sp_cells, N, st_all, False)                             # rows through the cell list
    assert s_cells['cell_builds'] == 1 and s_cells['pairs'] < 0.2 * N * (N - 1)
    for k in ('c_rep', 'c_alg', 'c_att'):
        assert np.array_equal(dense[k], cells[k]), k
        assert int(dense[k].sum()) % 2 == 0, 'a symmetric relation holds an even number of directed pairs'
    assert dense['c_att'].sum() > N          # the state is not trivial
    for k, c in (('f_rep', 'c_rep'), ('f_att', 'c_att')):
        for r in (dense, cells):
            tot = np.abs(r[k].sum(axis=0)).max()
            assert tot <= 1e-6 * max(1, r[c].sum()), (k, tot)                        # unit vectors cancel pairwise
        assert np.allclose(dense[k], cells[k], rtol=1e-4, atol=5e-3)   # ~2000 unit vectors summed in two different orders


def test_relabelling_and_translation_invariance():
    N, L = 20000, 300.0
    rng = np.random.default_rng(3)
    st = dict(x=f32(L * rng.random(N)), y=f32(L * rng.random(N)), phi=np.zeros(N), vproj=np.ones(N))
    d = base_dict('natPred', N, BC=0, size=L, Npred=0, alg_range=10.0)
    sp, _ = PP.make_sbc_params(d, path=PP.PATH_ALLPAIRS)
    base, _ = _counts(sp, N, st, True)
    perm = rng.permutation(N)
    st_p = {k: v[perm] for k, v in st.items()}
    rp, _ = _counts(sp, N, st_p, True)
    for k in ('c_rep', 'c_alg', 'c_att'):
        assert np.array_equal(rp[k], base[k][perm]), k
    # positions are multiples of 2^-15 below 512: adding 64 and folding by 300 (= 75 * 4) stays exact
    q = 2.0 ** -15
    st_q = dict(st, x=np.floor(st['x'] / q) * q, y=np.floor(st['y'] / q) * q)
    b2, _ = _counts(sp, N, st_q, True)
    st_t = dict(st_q, x=(st_q['x'] + 64.0) % L, y=(st_q['y'] + 128.0) % L)
    rt, _ = _counts(sp, N, st_t, True)
    for k in ('c_rep', 'c_alg', 'c_att'):
        assert np.array_equal(rt[k], b2[k]), k


def test_full_size_ensemble_sample_tanks_replayed_by_the_oracle():
    """BASELINE config C2 at full size: 16,384 tanks x 32 agents, 2,000 steps on the GPU; a sample of tanks is
    replayed by the oracle: burst timers exact, everybody inside the tank, finite."""
    from sde_burst_coast_b200.swarm import Swarm
    N, R, steps = 32, 16384, 2000
    sp, _, _ = make_params('burst_coast', N)
    rng = np.random.default_rng(16384)
    rad = 0.8 * sp.sizeL * np.sqrt(rng.random((R, N)))
    th = 2 * np.pi * rng.random((R, N))
    st = dict(x=f32(rad * np.cos(th)), y=f32(rad * np.sin(th)), phi=f32(2 * np.pi * rng.random((R, N)) - np.pi),
              vproj=np.ones((R, N)))
    g = Swarm(sp, N, n_replicates=R, seed=2026)
    g.set_state(**st)
    g.step(0, steps)
    gs = g.get_state()
    assert g.stats()['agent_steps'] == N * R * steps
    g.close()
    assert np.all(np.isfinite(gs['x'])) and np.all(np.hypot(gs['x'], gs['y']) <= sp.sizeL * (1 + 1e-6))
    for r in (0, 1, 4095, 8191, 16383):
        w = pyorc.World(sp, N, seed=2026, replicate=r)
        w.set_state(**{k: v[r] for k, v in st.items()})
        w.step(0, steps)
        o = w.get_state()
        assert np.array_equal(o['steps_till_burst'], gs['steps_till_burst'][r])
        assert np.array_equal(o['bin_step'], gs['bin_step'][r])
        w.close()
