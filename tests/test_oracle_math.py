"""The oracle's fixed-sequence transcendentals (oracle/oracle_math.h) against glibc/numpy."""
import numpy as np

from oracle_lib import oracle_math


def _ulps(a, ref):
    a = a.astype(np.float64)
    ref64 = ref.astype(np.float64)
    spacing = np.spacing(np.abs(ref.astype(np.float32))).astype(np.float64)
    return np.abs(a - ref64) / spacing


def test_expf():
    x = np.concatenate([np.linspace(-87, 0, 400001), -np.logspace(-8, 1.9, 20000)]).astype(np.float32)
    got = oracle_math(0, x)
    ref = np.exp(x.astype(np.float64))
    assert _ulps(got, ref.astype(np.float32)).max() <= 4
    assert oracle_math(0, np.array([-100.0, 0.0], np.float32)).tolist() == [0.0, 1.0]


def test_atan2f():
    rng = np.random.default_rng(1)
    y = rng.standard_normal(500000).astype(np.float32)
    x = rng.standard_normal(500000).astype(np.float32)
    y[:1000] = 0
    x[1000:2000] = 0
    got = oracle_math(1, y, x)
    ref = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    assert np.abs(got - ref).max() <= 4 * np.spacing(np.float32(np.pi))
    assert oracle_math(1, np.zeros(1, np.float32), np.zeros(1, np.float32))[0] == 0.0
    assert np.all(got <= np.float32(np.pi)) and np.all(got >= -np.float32(np.pi))


def test_sincosf():
    x = np.linspace(-7, 7, 700001).astype(np.float32)
    s, c = oracle_math(2, x), oracle_math(3, x)
    assert np.abs(s - np.sin(x.astype(np.float64))).max() <= 2e-7
    assert np.abs(c - np.cos(x.astype(np.float64))).max() <= 2e-7


def test_exp2f():
    x = np.linspace(-2, 3, 200001).astype(np.float32)
    got = oracle_math(4, x)
    ref = np.exp2(x.astype(np.float64))
    assert (np.abs(got - ref) / ref).max() <= 5e-7
