"""The cell list of the large-swarm path is rebuilt when an agent may have moved skin/2 since the build; the host
bounds the per-step displacement by max(initial speed, max force / beta) * dt (x3 where walls reflect), see
set_cells_max_age in csrc/sbc_api.cu. This runs the CPU oracle and checks that no agent ever moves further in
one step than that bound."""
import numpy as np
import pytest

from helpers import make_params, random_state
from oracle import pyorc


@pytest.mark.parametrize('mode,over,factor', [
    ('natPred', dict(BC=0, size=100.0, Npred=0), 1.0),
    ('burst_coast', dict(BC=6), 3.0),
    ('burst_coast', dict(BC=5, size=40.0), 3.0),
    ('burst_coast', dict(BC=2, size=60.0), 3.0),
])
def test_per_step_displacement_stays_below_the_host_bound(mode, over, factor):
    N, steps = 48, 2500
    sp, _, _ = make_params(mode, N, **over)
    rng = np.random.default_rng(5)
    st = random_state(N, sp, rng, mid_burst=True)
    v0 = 6.0                                   # faster than anything the forces sustain: the bound must cover it
    st['vproj'] = np.full(N, v0)
    if 'vx' in st:
        sc = v0 / np.maximum(np.hypot(st['vx'], st['vy']), 1e-12)
        st['vx'], st['vy'] = st['vx'] * sc, st['vy'] * sc
    w = pyorc.World(sp, N, seed=3)
    w.set_state(**st)
    L = sp.sizeL
    vmax = max(v0, max(sp.soc_strength, sp.env_strength) / sp.beta)
    bound = factor * vmax * sp.dt * 1.01
    prev = w.get_state()
    worst = 0.0
    for s in range(steps):
        w.step(s, 1)
        cur = w.get_state()
        dx, dy = cur['x'] - prev['x'], cur['y'] - prev['y']
        if sp.BC == 0:
            dx, dy = dx - L * np.rint(dx / L), dy - L * np.rint(dy / L)
        worst = max(worst, float(np.hypot(dx, dy).max()))
        prev = cur
    w.close()
    assert worst <= bound, (worst, bound)
    assert worst > 0.2 * bound / factor         # the bound is not vacuous
