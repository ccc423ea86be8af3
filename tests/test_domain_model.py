"""Geometry of the decomposed swarm with refresh exchanges, checked on a numpy model that
is independent of the CUDA code: slabs along x in a periodic box; at a FULL exchange agents are re-assigned to
the slab they are in and every rank mirrors the neighbours' agents within att_range + skin of its edges; for the
next K = floor(skin / 2 / max_step) steps nobody changes rank and the ghost SETS stay fixed (only positions are
refreshed). Claim: at every step every owned agent - even one that has drifted out of its slab - finds all its
true neighbours (distance <= att_range) among the owned agents and ghosts of its rank."""
import numpy as np
import pytest

from sde_burst_coast_b200 import domain


def _min_image(d, L):
    return d - L * np.rint(d / L)


@pytest.mark.parametrize('world', [2, 3, 5])
@pytest.mark.parametrize('skin_frac', [0.2, 0.5])
def test_fixed_ghost_sets_cover_every_neighbour_between_full_exchanges(world, skin_frac):
    rng = np.random.default_rng(world * 10 + int(skin_frac * 10))
    L, att, N = 400.0, 19.5, 1200
    skin = skin_frac * att
    max_step = 0.31                                   # largest displacement per step (any direction)
    K = int(np.floor(0.5 * skin / max_step))          # refresh exchanges between two full ones
    assert K >= 1
    domain.check_geometry(L, att + skin, world)
    pos = rng.random((N, 2)) * L
    bounds = domain.slab_bounds(L, world)
    owner = ghosts = None
    worst_outside = 0.0
    for step in range(3 * (K + 1)):
        if step % (K + 1) == 0:                       # full exchange: migration + new ghost sets
            owner = domain.owner_of(pos[:, 0], L, world)
            ghosts = []
            for r in range(world):
                lo, hi = bounds[r]
                x = pos[:, 0]
                to_lo = np.abs(_min_image(x - lo, L))   # distance to the slab's edges along the ring
                to_hi = np.abs(_min_image(x - hi, L))
                ghosts.append(np.where((owner != r) & (np.minimum(to_lo, to_hi) <= att + skin))[0])
        # every rank sweeps its owned rows against owned + ghosts (current positions)
        for r in range(world):
            own = np.where(owner == r)[0]
            visible = np.zeros(N, dtype=bool)
            visible[own] = True
            visible[ghosts[r]] = True
            d = np.hypot(_min_image(pos[own, None, 0] - pos[None, :, 0], L), _min_image(pos[own, None, 1] - pos[None, :, 1], L))
            needed = d <= att
            missing = needed & ~visible[None, :]
            assert not missing.any(), (step, r, np.argwhere(missing)[:3])
            lo, hi = bounds[r]
            inside = (pos[own, 0] >= lo) & (pos[own, 0] < hi)
            out = np.where(inside, 0.0, np.minimum(np.abs(_min_image(pos[own, 0] - lo, L)), np.abs(_min_image(pos[own, 0] - hi, L))))
            worst_outside = max(worst_outside, float(out.max()))
        # move: worst case for the argument - everybody at full speed in a persistent random direction
        if step % (K + 1) == 0:
            ang = rng.random(N) * 2 * np.pi
        pos = (pos + max_step * np.stack([np.cos(ang), np.sin(ang)], axis=1)) % L
    assert 0 < worst_outside <= 0.5 * skin + 1e-9     # agents do drift out of their slab, by at most skin/2
