"""Kill predicates on the device (sbc_kill_classify, csrc/sbc_device.cuh) are decided on the FP32 squared distance
and re-decided in FP64 only when d2 lies within 8e-6 (relative) of a threshold. This CPU check emulates the FP32
arithmetic with numpy (no FMA: at most an ulp away from the device) on pairs concentrated around the thresholds
and demands that OUTSIDE a four times narrower band the FP32 decision already equals the reference's FP64 one
(agents_dynamics.cpp:845-877: d <= kill_range, d < kill_range, d <= 4 kill_range with mathtools.h:54-83 distances)."""
import numpy as np
import pytest

f32 = np.float32


def _pairs(L, thr, periodic, n, rng):
    px = (rng.random(n) * L).astype(f32)
    py = (rng.random(n) * L).astype(f32)
    ang = rng.random(n) * 2 * np.pi
    d = thr * (1.0 + rng.uniform(-3e-5, 3e-5, n))          # distances hugging the threshold
    ax = px.astype(np.float64) + d * np.cos(ang)
    ay = py.astype(np.float64) + d * np.sin(ang)
    if periodic:
        ax, ay = np.mod(ax, L), np.mod(ay, L)
    ax, ay = ax.astype(f32), ay.astype(f32)
    if periodic:   # rounding may land on L itself
        ax, ay = np.where(ax >= f32(L), f32(0), ax), np.where(ay >= f32(L), f32(0), ay)
    return ax, ay, px, py


def _exact_distance(ax, ay, px, py, L, periodic):
    r0 = ax.astype(np.float64) - px.astype(np.float64)
    r1 = ay.astype(np.float64) - py.astype(np.float64)
    if periodic:
        for r in (r0, r1):
            h = np.sign(r) * L / 2.0
            r[...] = np.fmod(r + h, L) - h
    return np.sqrt(r0 * r0 + r1 * r1)


def _delta_pbc32(xi, xj, L):
    """sbc_delta_pbc: fold count from the rounded difference, then an exact-by-Sterbenz difference."""
    Lf, inv = f32(L), f32(1.0 / L)
    n = np.rint((xj - xi) * inv).astype(f32)
    a = xj - np.maximum(n, f32(0)) * Lf
    b = xi + np.minimum(n, f32(0)) * Lf
    L_lo = f32(L - float(Lf))
    return (a - b) - n * L_lo


@pytest.mark.parametrize('periodic,L', [(True, 100.0), (True, 584.0), (True, 10240.0), (False, 120.0)])
@pytest.mark.parametrize('kill_range', [5.0, 0.7, 19.3])
def test_fp32_decisions_agree_outside_a_quarter_of_the_band(periodic, L, kill_range):
    if periodic and 4 * kill_range * 1.1 >= L / 2:
        pytest.skip('sensing range reaches the half box')
    rng = np.random.default_rng(int(L + 1000 * kill_range))
    kr2, rs2 = f32(kill_range) * f32(kill_range), f32(16) * (f32(kill_range) * f32(kill_range))
    checked = 0
    for thr, thr2 in ((kill_range, kr2), (4 * kill_range, rs2)):
        ax, ay, px, py = _pairs(L, thr, periodic, 300000, rng)
        d = _exact_distance(ax, ay, px, py, L, periodic)
        if periodic:
            dx, dy = _delta_pbc32(px, ax, L), _delta_pbc32(py, ay, L)
        else:
            dx, dy = ax - px, ay - py
        d2 = dy * dy + dx * dx                                   # float32 throughout
        clear = np.abs(d2 - thr2) > f32(2e-6) * thr2               # device band is 8e-6
        assert clear.sum() > 1000
        le32, lt32 = d2 <= thr2, d2 < thr2
        le64, lt64 = d <= thr, d < thr
        assert np.array_equal(le32[clear], le64[clear]), int((le32 != le64)[clear].sum())
        assert np.array_equal(lt32[clear], lt64[clear]), int((lt32 != lt64)[clear].sum())
        checked += int(clear.sum())
    assert checked > 10000


@pytest.mark.parametrize('L', [100.0, 584.0, 10240.0, 14481.546878700494])
def test_periodic_difference_is_accurate_to_one_rounding(L):
    """sbc_delta_pbc (emulated above) against the exact minimum image of the same FP32 coordinates: the fold
    happens before the subtraction, so the result carries one rounding of the difference itself, not an ulp of L."""
    rng = np.random.default_rng(int(L))
    xi = (rng.random(500000) * L).astype(f32)
    xj = (rng.random(500000) * L).astype(f32)
    xi, xj = np.where(xi >= f32(L), f32(0), xi), np.where(xj >= f32(L), f32(0), xj)
    got = _delta_pbc32(xi, xj, L).astype(np.float64)
    r = xj.astype(np.float64) - xi.astype(np.float64)
    exact = r - L * np.rint(r / L)
    away_from_tie = np.abs(np.abs(r) - L / 2) > 1e-3 * L          # at |r| = L/2 either image is a minimum image
    err = np.abs(got - exact)[away_from_tie]
    bound = 2.0 ** -23 * np.abs(exact[away_from_tie]) + 2.0 ** -24 * np.abs(L - float(f32(L))) + 1e-30
    assert np.all(err <= bound), float((err / bound).max())
    # near neighbours (the pairs the zone tests care about) come out EXACT: both coordinates within a factor two
    close = away_from_tie & (np.abs(exact) < 0.25 * np.minimum(xi, xj).astype(np.float64)) & (np.rint(r / L) == 0)
    assert close.sum() > 1000 and np.all(got[close] == exact[close])
