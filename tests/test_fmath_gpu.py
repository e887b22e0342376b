"""csrc/fmath.cuh on the DEVICE (real MUFU seeds) against libm, through the C ABI (include/acs.h: acs_debug_fmath).

The host test (tests/test_fmath_host.py) pins the algebra of the sequences with emulated seeds; this one pins the device
build: the guard-free division / reciprocal / square root must be correctly rounded (bit-equal to IEEE) on operands of
the magnitudes the simulator has, the elementary functions within a few ulp of libm.  Tolerances in ulp of the result."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
N = 1 << 20


def _ulps(got, want):
    want = np.asarray(want, dtype=np.float64)
    ulp = np.spacing(np.abs(want))
    return np.abs(got - want) / np.where(ulp > 0, ulp, 1.0)


def _probe(op, a, b=None):
    from aircombat_selfplay_b200.capi import fmath_probe
    ta = torch.tensor(a, device="cuda")
    tb = None if b is None else torch.tensor(b, device="cuda")
    o, o2 = fmath_probe(op, ta, tb)
    torch.cuda.synchronize()
    return o.cpu().numpy(), o2.cpu().numpy()


def _wide(rng, n, lo=-80, hi=80, signed=True):
    x = np.ldexp(rng.uniform(0.5, 1.0, n), rng.integers(lo, hi, n))
    return x * rng.choice([-1.0, 1.0], n) if signed else x


def test_division_reciprocal_and_roots_are_correctly_rounded_on_the_device():
    rng = np.random.default_rng(0)
    a, b = _wide(rng, N), _wide(rng, N)
    q, _ = _probe("div", a, b)
    assert np.array_equal(q, a / b), float(_ulps(q, a / b).max())
    r, _ = _probe("rcp", b)
    assert _ulps(r, 1.0 / b).max() <= 1.0
    x = _wide(rng, N, signed=False)
    s, s0 = _probe("sqrt", x)
    assert np.array_equal(s, np.sqrt(x)) and np.array_equal(s0, np.sqrt(x))
    rs, _ = _probe("rsqrt", x)
    want = (1.0 / np.sqrt(x.astype(np.longdouble)))          # 1 / sqrt in double rounds twice: compare with extended precision
    assert float((np.abs(rs.astype(np.longdouble) - want) / np.spacing(np.abs(want.astype(np.float64)))).max()) <= 1.01
    # operands the *0 variant exists for: zero, subnormal (flushed seed), and a NaN that must stay one
    _, s0 = _probe("sqrt", np.array([0.0, 1e-310, 4.0, np.nan]))
    assert s0[0] == 0.0 and s0[1] == 0.0 and s0[2] == 2.0 and np.isnan(s0[3])
    q, _ = _probe("div", np.array([0.0, -0.0, 6.0]), np.array([3.0, 3.0, 3.0]))
    assert q[0] == 0.0 and q[2] == 2.0


def test_small_integer_quotients_match_ieee_bit_for_bit():
    # the action decode and the observation normalisation divide small integers / physical values by literals
    k = np.arange(0, 41, dtype=np.float64)
    for d in (20.0, 40.0, 58.0, 340.0, 1000.0, 5000.0):
        q, _ = _probe("div", k, np.full_like(k, d))
        assert np.array_equal(q, k / d)


def test_trigonometric_functions_on_the_device():
    rng = np.random.default_rng(1)
    x = rng.uniform(-1.0, 1.0, N) * np.where(rng.random(N) < 0.5, 7.0, 2.0e4)
    s, c = _probe("sincos", x)
    assert np.abs(s - np.sin(x)).max() <= 3e-16 and np.abs(c - np.cos(x)).max() <= 3e-16
    s1, c1 = _probe("sin", x)
    assert np.abs(s1 - np.sin(x)).max() <= 4e-16 and np.abs(c1 - c).max() <= 2.3e-16
    xs = rng.uniform(-0.78, 0.78, N)
    s, c = _probe("sincos_small", xs)
    assert _ulps(s, np.sin(xs)).max() <= 1.5 and _ulps(c, np.cos(xs)).max() <= 2.0
    yy, xx = _wide(rng, N, -30, 30), _wide(rng, N, -30, 30)
    yy[:1000] = 0.0; xx[1000:2000] = 0.0
    at, _ = _probe("atan2", yy, xx)
    want = np.arctan2(yy, xx)
    nz = want != 0
    assert _ulps(at[nz], want[nz]).max() <= 5.0 and np.abs(at - want).max() <= 7e-16
    # angle of attack / sideslip form: across = R sin, along = R cos
    ang = rng.uniform(-1.0, 1.0, N) * np.where(rng.random(N) < 0.9, 0.78, 3.1)
    R = rng.uniform(100.0, 1200.0, N)
    got, _ = _probe("angle_sc", R * np.sin(ang), R * np.cos(ang))
    assert _ulps(got, np.arctan2(R * np.sin(ang), R * np.cos(ang))).max() <= 6.0
    xa = np.concatenate([rng.uniform(-1.0, 1.0, N // 2), 1.0 - np.ldexp(rng.uniform(0.5, 1.0, N // 2), rng.integers(-50, 0, N // 2)),
                         [1.0, -1.0, 0.0]])
    ac, _ = _probe("acos", xa)
    want = np.arccos(xa)
    nz = want > 0
    assert _ulps(ac[nz], want[nz]).max() <= 8.0 and ac[-3] == 0.0 and abs(ac[-2] - np.pi) <= 5e-16


def test_exponential_family_on_the_device():
    rng = np.random.default_rng(2)
    z = rng.uniform(-1.0, 1.0, N) * np.where(rng.random(N) < 0.5, 3.0, 600.0)
    e, _ = _probe("exp", z)
    assert _ulps(e, np.exp(z)).max() <= 3.0
    lx = _wide(rng, N, -100, 100, signed=False)
    lg, _ = _probe("log", lx)
    ok = lx != 1.0
    assert _ulps(lg[ok], np.log(lx[ok])).max() <= 5.0
    t = rng.uniform(-12.0, 12.0, N)
    th, _ = _probe("tanh", t)
    assert np.abs(th - np.tanh(t)).max() <= 5e-16
    y = rng.uniform(-0.9999, 0.9999, N)
    _, ath = _probe("tanh", y)
    assert np.abs(ath - np.arctanh(y)).max() <= 4e-15
    # ISA pressure law: temperature ratio inside one layer, exponents of the troposphere (-5.2559) and of the +1 K/km layer (34.163)
    den = rng.uniform(180.0, 300.0, N)
    num = den * rng.uniform(0.75, 1.33, N)
    p1, p2 = _probe("pow_ratio", num, den)
    assert np.abs(p1 / np.power(num / den, -5.255876113278518) - 1.0).max() <= 2e-15
    assert np.abs(p2 / np.power(num / den, 34.16319474407325) - 1.0).max() <= 1e-14
