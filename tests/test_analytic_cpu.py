"""Oracle-independent analytic checks of the start-up tables — applied to BOTH the oracle and the
product's host table builders, so that a shared misreading of the reference's formulas has to survive a closed form too:
  * implicit.f90:144-158   xj(:,:,l) is the inverse of xf(:,:,l) = I + xi^2 l(l+1)/a^2 (R tref dhs^T - xd xc) for every l;
  * geopotential.f90:33-57 the hydrostatic integration reproduces the closed form for an isothermal column and a linear-in-log-sigma profile;
  * legendre.f90:158-191   get_weights = the Gauss-Legendre weights (numpy leggauss), hemisphere weights summing to 1;
  * spectral.f90:45-60     el2 = l(l+1)/a^2 and the triangular filter."""
import ctypes
import numpy as np
import pytest
from conftest import load_pkg

REARTH, RGAS_CP = 6.371e6, None


def _prod(name, trunc=30):
    return load_pkg().host_table(trunc, name)


def _consts():
    akap = float(np.float32(2.0) / np.float32(7.0))          # physical_constants.f90:23 (real32 quotient)
    cp = 1004.0
    return {"akap": akap, "cp": cp, "rgas": akap * cp, "grav": float(np.float32(9.81)), "gamma": 6.0}


@pytest.mark.parametrize("trunc", [30, 47])
def test_implicit_xj_inverts_xf(trunc):
    """product tables of initialize_implicit as the host builder leaves them"""
    c = _consts()
    kx = 8
    mx, nx = trunc + 1, trunc + 2
    nl = mx + nx + 1
    xj = _prod("xj", trunc).reshape(nl, kx, kx).transpose(0, 2, 1)       # Fortran (k,k1,l) -> [l][k][k1]
    xc = _prod("xc", trunc).reshape(kx, kx).T                             # already multiplied by xi (implicit.f90:160-161)
    xd = _prod("xd", trunc).reshape(kx, kx).T
    tref, dhs, dhsx = _prod("tref", trunc), _prod("dhs", trunc), _prod("dhsx", trunc)
    xi = dhsx[0] / dhs[0]                                                  # dhsx = xi * dhs (implicit.f90:74)
    assert xi in (600.0, 1200.0, 2400.0)                                   # xi = dt * alph, alph = 0.5, dt in {delt/2, delt, 2 delt} (time_stepping.f90:15-23)
    xe = xd @ (xc / xi)
    worst = 0.0
    for l in range(1, nl + 1):
        xxx = l * (l + 1) / REARTH ** 2
        xf = np.eye(kx) + xi * xi * xxx * (c["rgas"] * np.outer(tref, dhs) - xe)
        worst = max(worst, np.abs(xj[l - 1] @ xf - np.eye(kx)).max())
    assert worst < 1e-12, worst


def test_reference_temperature_profile():
    """implicit.f90:62-67: tref = 288 max(0.2, sigma)^(R gamma / (1000 g)), real32 0.2"""
    c = _consts()
    fsg, tref = _prod("fsg"), _prod("tref")
    rgam = c["rgas"] * c["gamma"] / (1000.0 * c["grav"])
    ref = 288.0 * np.maximum(float(np.float32(0.2)), fsg) ** rgam
    assert np.abs(tref - ref).max() < 1e-12
    assert np.all(np.diff(tref) >= 0) and 180 < tref[0] < 230 and 280 < tref[-1] < 288


def test_hydrostatic_geopotential_closed_forms(oracle):
    """geopotential.f90:33-57 on horizontally uniform columns (only the m = n = 0 coefficient): an isothermal atmosphere gives
    phi(k) = phis + R T ln(1 / sigma_k); the xgeop tables telescope to that exactly"""
    c = _consts()
    o = oracle
    o.model_init()                                   # initialize_geopotential (initialization.f90:53)
    fsg = o.table("fsg", o.kx)
    hsg = o.table("hsg", o.kx + 1)
    T0, phis0 = 250.0, 1234.5
    t = np.zeros((o.kx, o.nx, o.mx), complex)
    t[:, 0, 0] = T0
    phis = np.zeros((o.nx, o.mx), complex)
    phis[0, 0] = phis0
    phi = np.zeros_like(t)
    o.L.orc_get_geopotential(o.p(t), o.p(phis), o.p(phi))
    ref = phis0 + c["rgas"] * T0 * np.log(1.0 / fsg)
    assert np.abs(phi[:, 0, 0].real - ref).max() < 1e-9 * ref.max()
    assert np.abs(phi[:, 1:, :]).max() == 0 and np.abs(phi[:, 0, 1:]).max() == 0
    # the product's tables are the oracle's: xgeop1(k) = R ln(hsg(k+1)/fsg(k)), xgeop2(k) = R ln(fsg(k)/hsg(k)) (geopotential.f90:22-27)
    x1, x2 = _prod("xgeop1"), _prod("xgeop2")
    assert np.abs(x1 - c["rgas"] * np.log(hsg[1:] / fsg)).max() < 1e-12
    assert np.abs(x2[1:] - c["rgas"] * np.log(fsg[1:] / hsg[1:-1])).max() < 1e-12


@pytest.mark.parametrize("trunc", [30, 47])
def test_product_gauss_weights_and_latitudes(trunc):
    il = 48 if trunc == 30 else 72
    wt, sh = _prod("wt", trunc), _prod("sia_half", trunc)
    x, w = np.polynomial.legendre.leggauss(il)
    assert np.abs(wt - w[::-1][: il // 2]).max() < 1e-14           # legendre.f90:158-191 = true Gauss weights
    assert abs(wt.sum() - 1.0) < 2e-15
    assert 1e-5 < np.abs(sh - x[::-1][: il // 2]).max() < 1e-4      # geometry.f90:68: first-guess latitudes, NOT the nodes (F9)


def test_spectral_tables_closed_forms(oracle):
    o = oracle
    n = np.arange(o.nx)[:, None]; m = np.arange(o.mx)[None, :]
    l = m + n
    el2 = o.table("el2", o.mx * o.nx).reshape(o.nx, o.mx)
    assert np.abs(el2 - l * (l + 1) / REARTH ** 2).max() < 1e-24    # spectral.f90:45-46
    tf = o.table("trfilt", o.mx * o.nx).reshape(o.nx, o.mx)
    assert np.array_equal(tf, (l <= o.trunc).astype(float))          # spectral.f90:47-51
    assert np.array_equal(_prod("el2"), el2.ravel()) and np.array_equal(_prod("trfilt"), tf.ravel())
