"""Parity of the CUDA transforms / spectral operators (through the C ABI) against the oracle.
Tolerance: fp64, 1e-12 relative RMS per call (north_star: 1e-10 after 48 h of steps)."""
import numpy as np
import pytest
from conftest import random_spec, rel_rms

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _pair(request, res):
    return (request.getfixturevalue("ctx"), request.getfixturevalue("oracle")) if res == 30 else \
        (request.getfixturevalue("ctx47"), request.getfixturevalue("oracle47"))


@pytest.mark.parametrize("res", [30, 47])
@pytest.mark.parametrize("nb", [1, 8, 91])
def test_spec_to_grid(request, res, nb):
    c, o = _pair(request, res)
    rng = np.random.default_rng(1234)
    s = random_spec(rng, (nb,), o.nx, o.mx, o.trunc)
    # garbage outside the active triangle must be ignored (legendre.f90:38 nsh2)
    n = np.arange(o.nx)[:, None]; m = np.arange(o.mx)[None, :]
    s_dirty = s + 1e6 * ((m + n) > o.trunc + 1)
    kcos = np.where(np.arange(nb) % 2 == 0, 1, 2).astype(np.int32)
    ref = o.spec_to_grid(s, kcos)
    got = c.spec_to_grid(s_dirty, kcos)
    assert rel_rms(got, ref) < TOL
    assert np.abs(got - ref).max() < 1e-11 * np.abs(ref).max()


@pytest.mark.parametrize("res", [30, 47])
@pytest.mark.parametrize("nb", [1, 8, 73])
def test_grid_to_spec(request, res, nb):
    c, o = _pair(request, res)
    rng = np.random.default_rng(4321)
    g = rng.uniform(-1, 1, size=(nb, o.il, o.ix))
    ref = o.grid_to_spec(g)
    got = c.grid_to_spec(g)
    assert rel_rms(got, ref) < TOL
    # structural zeros: row nx, l > trunc+1, Im(m=0)  (legendre.f90:142-154, fourier.f90:76)
    n = np.arange(o.nx)[:, None]; m = np.arange(o.mx)[None, :]
    dead = ((m + n) > o.trunc + 1) | (n > o.trunc)
    assert np.all(got[:, dead] == 0)
    assert np.all(got[:, :, 0].imag == 0)


def test_legendre_and_fourier_stages(ctx, oracle):
    rng = np.random.default_rng(7)
    o = oracle
    x = random_spec(rng, (3,), o.nx, o.mx, o.trunc).view(np.float64).reshape(3, o.nx, 2 * o.mx)
    assert rel_rms(ctx.legendre_inv(x), o.legendre_inv(x)) < TOL
    f = rng.uniform(-1, 1, size=(3, o.il, 2 * o.mx))
    assert rel_rms(ctx.fourier_inv(f, 1), o.fourier_inv(f, 1)) < TOL
    assert rel_rms(ctx.fourier_inv(f, 2), o.fourier_inv(f, 2)) < TOL
    assert rel_rms(ctx.legendre_dir(f), o.legendre_dir(f)) < TOL
    g = rng.uniform(-1, 1, size=(3, o.il, o.ix))
    assert rel_rms(ctx.fourier_dir(g), o.fourier_dir(g)) < TOL


def test_edge_cases(ctx, oracle):
    o = oracle
    z = np.zeros((1, o.nx, o.mx), dtype=complex)
    assert np.all(ctx.spec_to_grid(z) == 0)
    assert np.all(ctx.grid_to_spec(np.zeros((1, o.il, o.ix))) == 0)
    assert ctx.spec_to_grid(np.zeros((0, o.nx, o.mx), dtype=complex)).shape == (0, o.il, o.ix)   # empty batch
    y00 = z.copy(); y00[0, 0, 0] = 1
    assert np.allclose(ctx.spec_to_grid(y00), np.float32(np.sqrt(np.float32(0.5))), rtol=0, atol=1e-15)
    big = 1e150 * random_spec(np.random.default_rng(0), (1,), o.nx, o.mx, o.trunc)
    assert rel_rms(ctx.spec_to_grid(big), o.spec_to_grid(big)) < TOL


def test_linearity_large_batch(ctx, oracle):
    """size-independent property at the micro-benchmark's full batch (B = 5824)"""
    o = oracle
    rng = np.random.default_rng(11)
    nb = 5824
    a = random_spec(rng, (nb,), o.nx, o.mx, o.trunc)
    perm = rng.permutation(nb)
    ga = ctx.spec_to_grid(a)
    gp = ctx.spec_to_grid(a[perm])
    assert np.array_equal(gp, ga[perm])            # batch entries independent and deterministic
    gs = ctx.spec_to_grid(a[:64].sum(axis=0, keepdims=True))
    assert rel_rms(gs[0], ga[:64].sum(axis=0)) < 1e-13
    idx = rng.choice(nb, 16, replace=False)
    assert rel_rms(ga[idx], o.spec_to_grid(a[idx])) < TOL


def test_spectral_operators(ctx, oracle):
    o = oracle
    rng = np.random.default_rng(5)
    a = random_spec(rng, (4,), o.nx, o.mx, o.trunc)
    b = random_spec(rng, (4,), o.nx, o.mx, o.trunc)
    for name, two_in in (("uvspec", True), ("vds", True), ("grad", False)):
        ref = o.op2(name, a, b if two_in else None)
        got = getattr(ctx, name)(a, b) if two_in else getattr(ctx, name)(a)
        for r, g in zip(ref, got):
            assert rel_rms(g, r) < 1e-14, name
    assert rel_rms(ctx.laplacian(a), o.op2("laplacian", a, nout=1)) < 1e-15
    assert rel_rms(ctx.inverse_laplacian(a), o.op2("inverse_laplacian", a, nout=1)) < 1e-15
    t = a.copy()
    for i in range(4):
        o.L.orc_trunct(o.p(t[i]))
    assert np.array_equal(ctx.trunct(a), t)


def test_vdspec(ctx, oracle):
    o = oracle
    rng = np.random.default_rng(6)
    ug = rng.uniform(-1, 1, size=(2, o.il, o.ix)); vg = rng.uniform(-1, 1, size=(2, o.il, o.ix))
    for kcos in (2, 1):
        vo, di = ctx.vdspec(ug, vg, kcos)
        for i in range(2):
            rv = np.empty((o.nx, o.mx), complex); rd = np.empty_like(rv)
            o.L.orc_vdspec(o.p(np.ascontiguousarray(ug[i])), o.p(np.ascontiguousarray(vg[i])), o.p(rv), o.p(rd), kcos)
            assert rel_rms(vo[i], rv) < TOL and rel_rms(di[i], rd) < TOL


def test_implicit_terms_and_horizontal_diffusion(ctx, oracle):
    """the operator-level drop-ins of implicit.f90:168-217 and horizontal_diffusion.f90:86-105 (the main loop runs both fused inside the
    spectral step): random tendencies on the triangle, matrices of initialize_implicit(dt) for the three time steps first_step uses"""
    import ctypes
    o = oracle
    rng = np.random.default_rng(21)
    kx = 8
    for dt in (0.5 * 2400.0, 2400.0, 2 * 2400.0):
        o.L.orc_initialize_implicit(ctypes.c_double(dt))
        ctx.initialize_implicit(dt)
        divdt = 1e-9 * random_spec(rng, (kx,), o.nx, o.mx, o.trunc)
        tdt = 1e-4 * random_spec(rng, (kx,), o.nx, o.mx, o.trunc)
        psdt = 1e-8 * random_spec(rng, (), o.nx, o.mx, o.trunc)
        a, b, c = ctx.implicit_terms(divdt, tdt, psdt)
        ra, rb, rc = divdt.copy(), tdt.copy(), psdt.copy()
        assert o.L.orc_implicit_terms(o.p(ra), o.p(rb), o.p(rc)) == 0
        assert rel_rms(a, ra) < 1e-14 and rel_rms(b, rb) < 1e-14 and rel_rms(c, rc) < 1e-14
        assert not np.array_equal(a, divdt)
        assert np.all(a[:, 0, 0] == 0)                          # l = 0: divdt(1,1,:) stays zero (implicit.f90:199)
    for name, name1, nlev in (("dmp", "dmp1", kx), ("dmpd", "dmp1d", kx), ("dmps", "dmp1s", 1)):
        d = ctx.table(name, o.nx * o.mx).reshape(o.nx, o.mx)
        d1 = ctx.table(name1, o.nx * o.mx).reshape(o.nx, o.mx)
        lead = (nlev,) if nlev > 1 else ()
        field = random_spec(rng, lead, o.nx, o.mx, o.trunc)
        fdt = 1e-5 * random_spec(rng, lead, o.nx, o.mx, o.trunc)
        got = ctx.do_horizontal_diffusion(field, fdt, d, d1)
        ref = fdt.copy()
        assert o.L.orc_do_horizontal_diffusion(o.p(np.ascontiguousarray(field)), o.p(ref), o.p(np.ascontiguousarray(d)), o.p(np.ascontiguousarray(d1)), nlev) == 0
        assert np.array_equal(got, ref), name                   # one multiply-subtract-multiply per coefficient: bit for bit
