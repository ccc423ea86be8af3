"""Statistical parity of ensemble observables (polarisation, nearest-neighbour distance, mean speed,
distance from the tank centre) between the CUDA ensemble path and the CPU oracle in the same
(metric) neighbour mode: independent tanks, different random streams, means must agree within
4 standard errors. This is the third leg of the north-star correctness check."""
import numpy as np
import pytest

from helpers import make_params, f32
from oracle import pyorc

pytestmark = pytest.mark.gpu


def _ic(R, N, sp, seed):
    rng = np.random.default_rng(seed)
    rad = 0.8 * sp.sizeL * np.sqrt(rng.random((R, N)))
    th = 2 * np.pi * rng.random((R, N))
    return dict(x=f32(rad * np.cos(th)), y=f32(rad * np.sin(th)), phi=f32(2 * np.pi * rng.random((R, N)) - np.pi),
                vproj=np.ones((R, N)))


@pytest.mark.parametrize('N,steps', [(8, 6000), (30, 3000)])
def test_tank_observables_match_oracle(N, steps):
    from sde_burst_coast_b200.swarm import Swarm
    sp, _, _ = make_params('burst_coast', N)
    Rg, Ro = 2048, 96
    ic = _ic(Rg, N, sp, 1)
    g = Swarm(sp, N, n_replicates=Rg, seed=123)
    g.set_state(**ic)
    samples_g = []
    for k in range(4):                      # four snapshots, 500 steps apart, after the transient
        g.step(0 if k == 0 else steps + 500 * (k - 1), steps if k == 0 else 500)
        obs = g.observables()
        st = g.get_state()
        r = np.hypot(st['x'], st['y']).mean(axis=1)
        samples_g.append(np.column_stack([obs[:, 1], obs[:, 6], obs[:, 7], r]))
    g.close()
    G = np.mean(samples_g, axis=0)          # per-tank time averages [Rg, 4]
    ico = _ic(Ro, N, sp, 2)
    O = []
    for r in range(Ro):
        w = pyorc.World(sp, N, seed=999, replicate=r)
        w.set_state(**{k: v[r] for k, v in ico.items()})
        snaps = []
        for k in range(4):
            w.step(0 if k == 0 else steps + 500 * (k - 1), steps if k == 0 else 500)
            o = w.observables()
            s = w.get_state()
            snaps.append([o[1], o[6], o[7], np.hypot(s['x'], s['y']).mean()])
        O.append(np.mean(snaps, axis=0))
        w.close()
    O = np.array(O)
    names = ['polarisation', 'mean speed', 'nearest-neighbour distance', 'distance from centre']
    for k, name in enumerate(names):
        mg, mo = G[:, k].mean(), O[:, k].mean()
        se = np.sqrt(G[:, k].var() / Rg + O[:, k].var() / Ro)
        assert abs(mg - mo) <= 4 * se + 1e-3 * abs(mo), (name, mg, mo, se)
    assert 0.0 < G[:, 0].mean() < 1.0 and G[:, 1].mean() > 1.0   # a sane ensemble: polarisation in (0,1), fish swim
