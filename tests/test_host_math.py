"""Accuracy of the fast-mode elementary functions (csrc/tfg_math.cuh), compiled for the HOST with g++ and compared
with 80-bit long-double libm.  The device build evaluates the same C++ (MUFU seeds are emulated by their documented
precision: the reciprocal / rsqrt of the high word only), so these bounds carry over to the kernel up to FMA
contraction.  No GPU needed.
"""

import ctypes as C
import subprocess
from fractions import Fraction
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
L = np.longdouble


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = tmp_path_factory.mktemp("hostmath") / "mathcheck.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so),
                    str(ROOT / "tests" / "host_math" / "mathcheck.cpp")], check=True)
    return C.CDLL(str(so))


def call(lib, name, x, y=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty_like(x)
    f = getattr(lib, name)
    f.argtypes = [C.c_void_p] * 3 + [C.c_long]
    f.restype = None
    f(x.ctypes.data, None if y is None else y.ctypes.data, out.ctypes.data, x.size)
    return out


def ulps(got, want):
    u = np.spacing(np.abs(want.astype(np.float64))).astype(L)
    return float(np.max(np.abs((got.astype(L) - want) / u)))


def test_exp(lib):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-700, 700, 200_000), rng.uniform(-30, 30, 200_000), rng.uniform(-1, 1, 200_000),
                        rng.uniform(-1e-3, 1e-3, 1000), [0.0]])
    want = np.exp(x.astype(L))
    assert ulps(call(lib, "mc_exp_tab", x), want) < 1.1
    assert ulps(call(lib, "mc_exp_core", x), want) < 1.2


def test_n_at_once_variants_equal_scalar_routines(lib):
    rng = np.random.default_rng(7)
    x = rng.uniform(-50, 50, 6000)
    assert np.array_equal(call(lib, "mc_exp_tab", x), call(lib, "mc_exp_tab_2", x))
    x = np.exp(x)
    assert np.array_equal(call(lib, "mc_log_tab", x), call(lib, "mc_log_tab_2", x))
    assert np.array_equal(call(lib, "mc_rcp3", x), call(lib, "mc_rcp3_3", x))


def test_log(lib):
    rng = np.random.default_rng(1)
    x = np.concatenate([np.exp(rng.uniform(-700, 700, 200_000)), rng.uniform(0.5, 2, 400_000),
                        1 + rng.uniform(-1e-2, 1e-2, 200_000), 1 + rng.uniform(-1e-6, 1e-6, 2000),
                        rng.uniform(1e-6, 100, 200_000), [1.0, 0.6875, 1.375, 2.0, 0.5]])
    want = np.log(x.astype(L))
    got = call(lib, "mc_log_tab", x)
    assert got[x == 1.0][0] == 0.0
    assert ulps(got, want) < 1.9          # relative accuracy holds through x ~ 1 (the two bins next to 1 use c = 1)
    assert ulps(call(lib, "mc_log_core", x), want) < 2.1


def test_reciprocal_division_sqrt(lib):
    rng = np.random.default_rng(2)
    b = np.concatenate([rng.uniform(-1e3, 1e3, 300_000), np.exp(rng.uniform(-600, 600, 100_000)),
                        1 + rng.uniform(0, 1, 300_000)])
    a = rng.uniform(-1e3, 1e3, b.size)
    assert ulps(call(lib, "mc_rcp3", b), 1 / b.astype(L)) < 0.52
    assert ulps(call(lib, "mc_rcp", b), 1 / b.astype(L)) < 0.51
    assert ulps(call(lib, "mc_div_fast", a, b), a.astype(L) / b.astype(L)) < 1.6
    assert ulps(call(lib, "mc_div", a, b), a.astype(L) / b.astype(L)) < 0.51
    w = np.concatenate([rng.uniform(0, 4, 300_000), np.exp(rng.uniform(-300, 300, 100_000))])
    assert ulps(call(lib, "mc_sqrt_pos", w), np.sqrt(w.astype(L))) < 0.51


def test_arctangents_and_wet_bulb(lib):
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-100, 100, 200_000), rng.uniform(-1, 1, 200_000)])
    assert ulps(call(lib, "mc_atan_core", x), np.arctan(x.astype(L))) < 3.0
    s = rng.uniform(0, 1, 300_000)
    assert ulps(call(lib, "mc_asin01", s), np.arcsin(s.astype(L))) < 2.1
    T, RH = rng.uniform(-90, 90, 500_000), rng.uniform(0, 2, 500_000)
    a, b = T + RH, RH - 1.676331
    want = np.arctan(a.astype(L)) - np.arctan(b.astype(L))
    assert np.max(np.abs(call(lib, "mc_atan_diff", a, b).astype(L) - want)) < 1e-15
    b2 = rng.uniform(-1.68, 0.33, 200_000)   # a*b ~ -1: the difference passes through +-pi/2
    a2 = -1 / b2 * (1 + rng.uniform(-1e-9, 1e-9, b2.size))
    want = np.arctan(a2.astype(L)) - np.arctan(b2.astype(L))
    assert np.max(np.abs(call(lib, "mc_atan_diff", a2, b2).astype(L) - want)) < 1e-15
    # Stull (2011) wet bulb as the reference evaluates it (bmi_topoflow_glacier.py:1514-1520), RH a fraction
    Tl, Rl = T.astype(L), RH.astype(L)
    want = (Tl * np.arctan(L(0.151977) * np.sqrt(Rl + L(8.313659))) + np.arctan(Tl + Rl) - np.arctan(Rl - L(1.676331))
            + L(0.00391838) * Rl ** L(1.5) * np.arctan(L(0.023101) * Rl) - L(4.86035))
    assert np.max(np.abs(call(lib, "mc_stull", T, RH).astype(L) - want)) < 5e-14   # |T| * 1 ulp of the arctangent


def test_table_driven_wet_bulb_and_air_mass(lib):
    """Round 2: piecewise-polynomial tables (scripts/gen_math_coeffs.py) for the Stull wet bulb and the Kasten-Young
    air mass against the closed forms in 80-bit arithmetic."""
    rng = np.random.default_rng(7)
    T = np.concatenate([rng.uniform(-95, 95, 400_000), rng.uniform(-3, 3, 200_000)])
    RH = np.concatenate([rng.uniform(3 / 64, 2, 500_000), rng.uniform(0.2, 1.1, 100_000)])
    RH[:66] = np.concatenate([np.arange(1, 33) / 16.0, (np.arange(1, 33) + 0.5) / 16.0, [3 / 64, 2.0]])  # bin centres / edges
    Tl, Rl = T.astype(L), RH.astype(L)
    want = (Tl * np.arctan(L(0.151977) * np.sqrt(Rl + L(8.313659))) + np.arctan(Tl + Rl) - np.arctan(Rl - L(1.676331))
            + L(0.00391838) * Rl ** L(1.5) * np.arctan(L(0.023101) * Rl) - L(4.86035))
    got = call(lib, "mc_stull_tab", T, RH)
    assert np.max(np.abs(got.astype(L) - want)) < 3e-14
    # ... and it agrees with the closed-form fast routine it replaces to the same level
    assert np.max(np.abs(got - call(lib, "mc_stull", T, RH))) < 6e-14
    # Kasten & Young (1989): 1/M = sin(gamma) + 0.50572 (gamma + 6.07995)^-1.6364, gamma = asin(c) in degrees
    c = np.concatenate([rng.uniform(0, 1, 600_000), rng.uniform(0, 0.05, 100_000), rng.uniform(0.93, 1.0, 100_000),
                        np.arange(0, 129) / 128.0, (np.arange(0, 128) + 0.5) / 128.0, [0.9375, 0.93750001, 1.0, 0.0]])
    cl = c.astype(L)
    gamma = np.arcsin(cl) * (L(180) / L(np.pi)) if False else np.arcsin(cl) * L(180) / np.arccos(L(-1))
    want = cl + L(0.50572) * (gamma + L(6.07995)) ** L(-1.6364)
    got = call(lib, "mc_inv_air_mass", c).astype(L)
    rel = np.abs(got - want) / want
    assert rel.max() < 6e-14, rel.max()
    assert rel[c > 0.05].max() < 4e-15, rel[c > 0.05].max()


def test_fused_exp_pair_and_seventh_root_equal_separate_routines(lib):
    rng = np.random.default_rng(11)
    n = 30000
    x = rng.uniform(-8, 3, n)
    y = 10.0 ** rng.uniform(-4, 0, n)          # argument of the 7th root: (e_air / 10) / T_K ~ 1e-3
    got = call(lib, "mc_exp2_root7", x, y)
    e = call(lib, "mc_exp_tab", x)
    r = call(lib, "mc_root7", y, np.ones(n))
    idx = np.arange(n - n % 3).reshape(-1, 3)
    assert np.array_equal(got[idx[:, 0]], e[idx[:, 0]]) and np.array_equal(got[idx[:, 1]], e[idx[:, 1]])
    assert np.array_equal(got[idx[:, 2]], r[idx[:, 2]])


def test_seventh_root(lib):
    """x**(1/7) by one Householder step from a float32 seed: <= ~8 ulp with the seed's relative error up to 3e-6
    (MUFU lg2/ex2 deliver ~4e-7)."""
    rng = np.random.default_rng(5)
    x = np.exp(rng.uniform(np.log(1e-12), np.log(10.0), 300_000))
    want = x.astype(L) ** (L(1) / L(7))
    for scale, bound in ((1.0, 8.0), (1 + 1e-6, 9.0), (1 - 1e-6, 9.0), (1 + 3e-6, 32.0)):
        got = call(lib, "mc_root7", x, np.full(x.size, scale))
        assert float(np.max(np.abs((got.astype(L) - want) / want))) < bound * 2.3e-16, scale


def test_markstein_division_by_3600_is_correctly_rounded():
    """q = a*y, r = fma(-3600, q, a), q' = fma(r, y, q), y = RN(1/3600)  ==  RN(a/3600)  (tfg_num.cuh div3600)."""
    rng = np.random.default_rng(4)
    y = 1.0 / 3600.0

    def fma(a, b, c):  # exact: float(Fraction) rounds to nearest-even
        return float(Fraction(a) * Fraction(b) + Fraction(c))

    vals = np.concatenate([rng.uniform(0, 1, 20_000), np.exp(rng.uniform(-600, 20, 20_000)),
                           3600.0 * rng.uniform(0, 1e-3, 10_000), [3600.0 * k * 2.0 ** -60 for k in range(1, 500)], [0.0]])
    for a in map(float, vals):
        q = a * y
        assert fma(fma(-3600.0, q, a), y, q) == a / 3600.0
