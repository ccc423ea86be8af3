"""Voronoi neighbour mode (the shipped default of the reference: InteractionVoronoiF2F,
agents_interact.cpp:26-77): candidates are the Delaunay neighbours, then the three-zone filter.

CPU: the port against the reference's own driver running on the brute-force Delaunay stand-in
(bit-identical). GPU: the CUDA kernels against the port - neighbour lists and zone counters exact,
one step within the FP32 tolerance, burst timers exact over long runs."""
import numpy as np
import pytest

from helpers import base_dict, random_state, nn_sets, ang_diff
from oracle import pyorc
from sde_burst_coast_b200 import params as PP

needs_ref = pytest.mark.skipif(not pyorc.have_reference(), reason='oracle/_ref not built')


def _sp(N, **over):
    d = base_dict('burst_coast', N, **over)
    return PP.make_sbc_params(d, neighbour_mode=PP.NEIGHBOUR_VORONOI)[0]


@needs_ref
@pytest.mark.parametrize('N,over', [(3, {}), (8, {}), (30, dict(alg_range=10.0)), (64, dict(BC=5)), (50, dict(BC=-1))])
def test_port_matches_reference_voronoi_driver(N, over):
    sp = _sp(N, **over)
    rng = np.random.default_rng(N)
    st = random_state(N, sp, rng, spread=60.0 if sp.BC == -1 else None)
    res = []
    for kind in ('port', 'reference'):
        w = pyorc.World(sp, N, seed=3, kind=kind)
        w.set_state(**st)
        r = [w.pair_interactions(1), w.pair_interactions(0)]
        w.step(0, 400)
        res.append((r, w.get_state()))
        w.close()
    for a, b in zip(res[0][0], res[1][0]):
        for k in a:
            assert np.array_equal(a[k], b[k]), k
    for k in res[0][1]:
        assert np.array_equal(res[0][1][k], res[1][1][k], equal_nan=True), k


def test_voronoi_has_fewer_neighbours_than_metric():
    N = 30
    rng = np.random.default_rng(1)
    spv, spm = _sp(N), PP.make_sbc_params(base_dict('burst_coast', N))[0]
    st = random_state(N, spv, rng)
    out = []
    for sp in (spv, spm):
        w = pyorc.World(sp, N)
        w.set_state(**st)
        r = w.pair_interactions(1)
        out.append((r['c_rep'] + r['c_alg'] + r['c_att']).mean())
        sets = nn_sets(r, N)
        out.append(sets)
        w.close()
    assert out[0] < out[2]
    assert all(set(v) <= set(m) for v, m in zip(out[1], out[3]))   # Delaunay-filtered subset of the metric set


@pytest.mark.gpu
@pytest.mark.parametrize('N,over', [(8, {}), (16, {}), (30, dict(alg_range=10.0)), (64, dict(BC=5)), (200, dict(BC=-1))])
def test_gpu_voronoi_neighbours_exact(N, over):
    from sde_burst_coast_b200.swarm import Swarm
    sp = _sp(N, **over)
    rng = np.random.default_rng(N + 5)
    st = random_state(N, sp, rng, spread=80.0 if sp.BC == -1 else None)
    w = pyorc.World(sp, N)
    w.set_state(**st)
    g = Swarm(sp, N)
    g.set_state(**st)
    for flags in (1, 0):
        ref, got = w.pair_interactions(flags), g.pair_interactions(flags)
        for k in ('c_rep', 'c_alg', 'c_att', 'nn_ptr', 'nn_idx'):
            assert np.array_equal(ref[k], got[k]), (flags, k)
        for k in ('f_rep', 'f_alg', 'f_att'):
            assert np.allclose(ref[k], got[k], rtol=1e-5, atol=2e-5), (flags, k)
    g.close()
    w.close()


@pytest.mark.gpu
@pytest.mark.parametrize('N,R', [(8, 16), (13, 8), (30, 8), (64, 4)])
def test_gpu_voronoi_step_parity(N, R):
    """One step with injected decisions in every tank, then a long run with the Philox protocol."""
    from sde_burst_coast_b200.swarm import Swarm
    sp = _sp(N)
    rng = np.random.default_rng(N * 3)
    sts = [random_state(N, sp, rng) for _ in range(R)]
    for s in sts:
        s['bin_step'] = np.where(rng.random(N) < 0.5, sp.burst_steps, s['bin_step']).astype(np.uint32)
    us, ud = rng.random((R, N)), np.asarray(rng.random((R, N)), dtype=np.float32).astype(np.float64)
    kn = rng.integers(1, 600, (R, N)).astype(np.uint32)
    g = Swarm(sp, N, n_replicates=R, seed=4)
    g.set_state(**{k: np.stack([s[k] for s in sts]) for k in sts[0]})
    g.inject(us, ud, kn)
    g.step(0, 1)
    gs = g.get_state()
    bad = 0
    for r in range(R):
        w = pyorc.World(sp, N, seed=4, replicate=r)
        w.set_state(**sts[r])
        w.inject(us[r], ud[r], kn[r])
        w.step(0, 1)
        o = w.get_state()
        assert np.array_equal(o['bin_step'], gs['bin_step'][r]) and np.array_equal(o['steps_till_burst'], gs['steps_till_burst'][r])
        ok = (np.abs(o['x'] - gs['x'][r]) <= 1e-5 * np.maximum(1, np.abs(o['x']))) & \
             (np.abs(o['y'] - gs['y'][r]) <= 1e-5 * np.maximum(1, np.abs(o['y']))) & \
             (np.abs(o['vproj'] - gs['vproj'][r]) <= 1e-5 * np.maximum(1, o['vproj'])) & (ang_diff(o['phi'], gs['phi'][r]) <= 2e-5) & \
             (np.hypot(o['fx'] - gs['fx'][r], o['fy'] - gs['fy'][r]) <= 2e-5 * np.maximum(1, np.hypot(o['fx'], o['fy'])))
        bad += int((~ok).sum())
        w.step(1, 999)
        w.close()
    assert bad <= max(1, R * N // 100), bad
    g.step(1, 999)
    gs = g.get_state()
    w = pyorc.World(sp, N, seed=4, replicate=R - 1)
    w.set_state(**sts[R - 1])
    w.inject(us[R - 1], ud[R - 1], kn[R - 1])
    w.step(0, 1000)
    o = w.get_state()
    assert np.array_equal(o['steps_till_burst'], gs['steps_till_burst'][R - 1])
    assert np.all(np.hypot(gs['x'], gs['y']) <= sp.sizeL * (1 + 1e-6))
    g.close()


@pytest.mark.gpu
def test_voronoi_mode_rejected_where_not_implemented():
    from sde_burst_coast_b200.swarm import Swarm, SbcError
    d = base_dict('natPred', 2000, BC=0, size=400.0, Npred=0)
    sp, _ = PP.make_sbc_params(d, neighbour_mode=PP.NEIGHBOUR_VORONOI)
    with pytest.raises(SbcError, match='Voronoi'):
        Swarm(sp, 2000)


def _vsp(mode, N, **over):
    d = base_dict(mode, N, **over)
    return PP.make_sbc_params(d, neighbour_mode=PP.NEIGHBOUR_VORONOI)


@pytest.mark.gpu
@pytest.mark.parametrize('N,L', [(30, 60.0), (200, 100.0), (700, 150.0)])
def test_gpu_periodic_voronoi_neighbours_exact(N, L):
    """Periodic box (InteractionVoronoiF2FP's tiling, agents_operation.cpp:240-293, as the well-defined nearest-image
    triangulation): neighbour lists and zone counters of the CUDA kernels identical to the port's, and the edge set is
    the Delaunay triangulation of the 3 x 3 tiled point set (scipy) wherever that one is unambiguous."""
    from sde_burst_coast_b200.swarm import Swarm
    sp, _ = _vsp('natPred', N, BC=0, size=L, Npred=0, alg_range=10.0)
    rng = np.random.default_rng(N)
    st = random_state(N, sp, rng)
    w = pyorc.World(sp, N)
    w.set_state(**st)
    g = Swarm(sp, N)
    g.set_state(**st)
    ref, got = w.pair_interactions(1), g.pair_interactions(1)
    for k in ('c_rep', 'c_alg', 'c_att', 'nn_ptr', 'nn_idx'):
        assert np.array_equal(ref[k], got[k]), k
    g.close()
    w.close()
    # independent check of the criterion: Qhull on the tiled set; an agent's metric neighbours that are Delaunay
    # neighbours of one of its images must be exactly the port's list
    from scipy.spatial import Delaunay
    pts = np.stack([st['x'], st['y']], axis=1)
    tiles = np.concatenate([pts + np.array([a * L, b * L]) for a in (-1, 0, 1) for b in (-1, 0, 1)])
    tri = Delaunay(tiles)
    indptr, indices = tri.vertex_neighbor_vertices
    centre = 4 * N    # the (0, 0) tile
    sets = nn_sets(ref, N)
    for i in range(N):
        nb = set(int(v) % N for v in indices[indptr[centre + i]:indptr[centre + i + 1]]) - {i}
        dx = (pts[:, 0] - pts[i, 0] + L / 2) % L - L / 2
        dy = (pts[:, 1] - pts[i, 1] + L / 2) % L - L / 2
        near = set(np.where(np.hypot(dx, dy) <= sp.att_range)[0]) - {i}
        assert set(sets[i]) == (nb & near), i


@pytest.mark.gpu
@pytest.mark.parametrize('mode,N,npred,kill,move,L', [('natPred', 30, 1, 3, 1, 100.0), ('natPred', 200, 1, 1, 0, 100.0),
                                                      ('mulFisher', 300, 20, 1, 0, 100.0)])
def test_gpu_voronoi_predator_phase_matches_port(mode, N, npred, kill, move, L):
    """Voronoi candidates in the predator phase (InteractionVoronoiF2FP, agents_interact.cpp:79-141): the predators are
    vertices of the triangulation, both directions of an edge interact. CUDA kernels against the port on
    the same states: the same agents die, timers bit-identical, one step within the FP32 tolerance."""
    from sde_burst_coast_b200.swarm import Swarm
    sp, np_ = _vsp(mode, N, BC=0, size=L, Npred=npred, pred_kill=kill, pred_move=move, pred_time=0.0, time=50.0,
                   kill_rate=400.0, N_confu=500)
    assert np_ == npred
    rng = np.random.default_rng(N + kill)
    st = random_state(N, sp, rng, spread=40.0)
    st['bin_step'] = np.where(rng.random(N) < 0.4, sp.burst_steps, st['bin_step']).astype(np.uint32)
    px = 45.0 + 0.5 * np.arange(npred, dtype=np.float64)
    py, pphi = np.full(npred, 48.0), np.full(npred, np.pi / 2)
    w = pyorc.World(sp, N, npred, seed=9)
    g = Swarm(sp, N, npred, seed=9)
    for o in (w, g):
        o.set_state(**st)
        o.set_predators(px, py, pphi, 1)
    w.step(5, 1)
    g.step(5, 1)
    o, gs = w.get_state(), g.get_state()
    assert np.array_equal(o['alive'], gs['alive'])
    al = o['alive'] == 1
    assert np.array_equal(o['bin_step'][al], gs['bin_step'][al]) and np.array_equal(o['steps_till_burst'][al], gs['steps_till_burst'][al])
    ok = (np.abs(o['x'] - gs['x']) <= 1e-4) & (np.abs(o['y'] - gs['y']) <= 1e-4) & (ang_diff(o['phi'], gs['phi']) <= 1e-4)
    assert ok[al].mean() > 0.97, ok[al].mean()
    w.step(6, 300)
    g.step(6, 300)
    o, gs = w.get_state(), g.get_state()
    al = (o['alive'] == 1) & (gs['alive'] == 1)
    assert np.array_equal(o['steps_till_burst'][al], gs['steps_till_burst'][al])
    assert abs(int(o['alive'].sum()) - int(gs['alive'].sum())) <= max(2, N // 20)   # trajectories diverge in FP32: kills agree in number
    g.close()
    w.close()
