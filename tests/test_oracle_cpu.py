"""CPU tests of the oracle itself (the reference ships no golden vectors — PARITY UNPINNED;
these are the independent anchors of SURVEY.md Appendix C and analytic properties)."""
import numpy as np
from conftest import random_spec


def test_gauss_weights(oracle):
    wt = oracle.table("wt", oracle.iy)
    assert abs(wt.sum() - 1.0) < 2e-15                      # legendre.f90:162
    assert abs(wt[0] - 0.0031533460523054) < 1e-15
    x, w = np.polynomial.legendre.leggauss(oracle.il)
    assert np.abs(wt - w[::-1][: oracle.iy]).max() < 1e-14    # true Gauss weights


def test_latitudes_are_approximate(oracle):
    """F9: geometry.f90:68 uses the first-guess latitudes, not the Gauss nodes."""
    sh = oracle.table("sia_half", oracle.iy)
    x, _ = np.polynomial.legendre.leggauss(oracle.il)
    err = np.abs(sh - x[::-1][: oracle.iy]).max()
    assert 4e-5 < err < 6e-5


def test_nsh2_and_fft_factors(oracle):
    nsh2 = oracle.itable("nsh2", oracle.nx)
    assert list(nsh2[:4]) == [62, 62, 60, 58] and nsh2[-1] == 2 and nsh2.sum() == 1054
    ifac = oracle.itable("ifac", 15)
    assert list(ifac[:6]) == [96, 4, 2, 4, 4, 3]                # fftpack.f90:1-36


def test_fftpack_is_perturbed_dft(oracle):
    """F13: single-precision 2*pi / sqrt(3) / sqrt(2) => ~1e-7 from the exact DFT."""
    rng = np.random.default_rng(0)
    v = rng.standard_normal(96)
    y = v.copy()
    oracle.L.orc_rfftf(oracle.p(y))
    ref = np.fft.rfft(v)
    hc = np.zeros(96)
    hc[0] = ref[0].real
    hc[1:95:2] = ref[1:48].real
    hc[2:95:2] = ref[1:48].imag
    hc[95] = ref[48].real
    err = np.abs(y - hc).max() / np.abs(hc).max()
    assert 1e-9 < err < 1e-6
    z = y.copy()
    oracle.L.orc_rfftb(oracle.p(z))
    rt = np.abs(z / 96 - v).max()
    assert 1e-10 < rt < 1e-6       # an exact round trip would mean the constants were "fixed"


def test_y00_and_roundtrip(oracle):
    spec = np.zeros((oracle.nx, oracle.mx), dtype=complex)
    spec[0, 0] = 1.0
    g = oracle.spec_to_grid(spec)
    assert np.allclose(g, np.float32(np.sqrt(np.float32(0.5))), rtol=0, atol=1e-15)   # legendre.f90:212
    rng = np.random.default_rng(1)
    s = random_spec(rng, (), oracle.nx, oracle.mx, oracle.trunc, full_triangle=False)
    s[:, 0] = s[:, 0].real
    s2 = oracle.grid_to_spec(oracle.spec_to_grid(s))
    err = np.abs(s2 - s).max()
    assert 1e-5 < err < 5e-2     # F9: not an identity


def test_linearity(oracle):
    rng = np.random.default_rng(2)
    a = random_spec(rng, (), oracle.nx, oracle.mx, oracle.trunc)
    b = random_spec(rng, (), oracle.nx, oracle.mx, oracle.trunc)
    ga, gb, gab = oracle.spec_to_grid(a), oracle.spec_to_grid(b), oracle.spec_to_grid(2 * a - 3 * b)
    assert np.abs(gab - (2 * ga - 3 * gb)).max() < 1e-12


def test_uvspec_vdspec_consistency(oracle):
    """(vor,div) -> uvspec -> grid (u,v) -> vdspec(kcos=2) returns (vor,div) up to the
    quadrature defect of the approximate latitudes (F9): a structure/sign check."""
    rng = np.random.default_rng(3)
    n = np.arange(oracle.nx)[:, None]
    m = np.arange(oracle.mx)[None, :]
    low = (m + n) <= 10
    vor = random_spec(rng, (), oracle.nx, oracle.mx, oracle.trunc) * low
    div = random_spec(rng, (), oracle.nx, oracle.mx, oracle.trunc) * low
    vor[:, 0] = vor[:, 0].real
    div[:, 0] = div[:, 0].real
    vor[0, 0] = div[0, 0] = 0
    uc, vc = oracle.op2("uvspec", vor, div)
    ug, vg = oracle.spec_to_grid(uc, 2), oracle.spec_to_grid(vc, 2)
    vo2 = np.empty_like(vor)
    di2 = np.empty_like(div)
    oracle.L.orc_vdspec(oracle.p(ug), oracle.p(vg), oracle.p(vo2), oracle.p(di2), 2)
    assert np.abs((vo2 - vor) * low).max() < 2e-2 * np.abs(vor).max()
    assert np.abs((di2 - div) * low).max() < 2e-2 * np.abs(div).max()


def test_legendre_table_matches_scipy(oracle):
    """legendre.f90:194-237 against an independent implementation: the table is the orthonormal associated
    Legendre function sqrt((2l+1)/2 (l-m)!/(l+m)!) P_l^m(x) without the Condon-Shortley phase, l = m + n,
    evaluated at the model's (approximate) latitudes; the recurrence is seeded in real32 -> 1e-6 agreement"""
    from math import factorial
    from scipy.special import lpmv
    o = oracle
    cpol = o.table("cpol", 2 * o.mx * o.nx * o.iy).reshape(o.iy, o.nx, 2 * o.mx)
    x = o.table("sia_half", o.iy)
    worst = 0.0
    for m in range(0, o.mx, 3):
        for n in range(0, o.nx):
            l = m + n
            if l > o.trunc + 1:
                continue
            norm = np.sqrt((2 * l + 1) / 2.0 * factorial(l - m) / factorial(l + m))
            ref = (-1) ** m * norm * lpmv(m, l, x)
            got = cpol[:, n, 2 * m]
            assert np.array_equal(got, cpol[:, n, 2 * m + 1])      # re and im slots hold the same value (legendre.f90:50-56)
            worst = max(worst, np.abs(got - ref).max())
    assert worst < 2e-6, worst
