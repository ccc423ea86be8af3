"""The FP64 minimum image of the reference, fmod(r + sgn(r) L/2, L) - sgn(r) L/2 (mathtools.h:64-71), is evaluated
on the device without fmod (sbc_fmod_near in csrc/sbc_device.cuh): for |t| < 2L, fmod(t, L) is t or the exact
difference t -/+ L (Sterbenz). This checks the claimed bit-identity on the CPU in IEEE double arithmetic."""
import numpy as np
import pytest


def _reference(r, L):
    s = np.sign(r)
    h = s * L / 2.0
    return np.fmod(r + h, L) - h


def _device_form(r, L):
    s = np.sign(r)
    h = s * L / 2.0
    t = r + h
    at = np.abs(t)
    near = np.where(at < L, t, np.where(t > 0, t - L, t + L))
    out = np.where(at < 2 * L, near, np.fmod(t, L))
    return out - h


@pytest.mark.parametrize('L', [100.0, 24.5, 10240.0, 14481.546878700494, 1e-3, 3.0000000000000004])
def test_conditional_subtraction_equals_fmod_bitwise(L):
    rng = np.random.default_rng(int(L * 1000) % 2**32)
    r = np.concatenate([
        rng.uniform(-1.5 * L, 1.5 * L, 400000),                      # every difference of two coordinates in [0, L)
        rng.uniform(-1.5 * L, 1.5 * L, 100000).astype(np.float32).astype(np.float64),   # FP32-representable inputs
        np.array([0.0, -0.0, L / 2, -L / 2, L, -L, np.nextafter(L / 2, 0), np.nextafter(L / 2, L), np.nextafter(L, 0),
                  -np.nextafter(L, 0), 1.5 * L, -1.5 * L, np.nextafter(1.5 * L, 0), 5e-324, -5e-324]),
        rng.uniform(-40 * L, 40 * L, 20000),                          # beyond 2L: the library path is kept
    ])
    a, b = _reference(r, L), _device_form(r, L)
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), np.where(a.view(np.uint64) != b.view(np.uint64))[0][:5]
