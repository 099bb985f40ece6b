"""Parity of every transform kernel a context can be switched to (speedy_set_option): grid->spec of
ensemble batches through the quad kernel (default: FFT + DMMA Legendre, four fields at a time), the
whole-field FFT kernel ("k2_field") and the streaming kernel with the dense operator ("k2_quad" = 0);
spec->grid with the dense-operator Fourier stage ("dense_inverse").  Same oracle, same tolerances as the default kernels:
1e-12 relative RMS per transform call, 1e-10 on the prognostic coefficients after 48 h."""
import os
import numpy as np
import pytest
from conftest import ROOT, random_spec, rel_rms

pytestmark = pytest.mark.gpu
BC = os.path.join(ROOT, "data", "bc_t30.bin")
PROG = ("vor", "div", "t", "tr", "ps")


VARIANTS = {"quad": {"k2_quad": 1, "k1_quad": 1}, "field": {"k2_quad": 0, "k2_field": 1}, "stream": {"k2_quad": 0, "k2_field": 0, "k1_quad": 0}}


def _select(c, variant):
    for k, v in VARIANTS[variant].items():
        c.set_option(k, v)


@pytest.mark.parametrize("variant", ["quad", "field"])
@pytest.mark.parametrize("nb", [584, 1201, 1202, 1203])
def test_grid_to_spec_whole_field(pkg, oracle, nb, variant):
    o = oracle
    c = pkg.Speedy(trunc=30)
    _select(c, variant)
    rng = np.random.default_rng(99)
    g = rng.uniform(-1, 1, size=(nb, o.il, o.ix))
    got = c.grid_to_spec(g)
    _select(c, "stream")
    base = c.grid_to_spec(g)
    idx = np.r_[0:8, nb // 2:nb // 2 + 8, nb - 8:nb]
    ref = o.grid_to_spec(g[idx])
    assert rel_rms(got[idx], ref) < 1e-12
    assert rel_rms(got, base) < 1e-12
    n = np.arange(o.nx)[:, None]; m = np.arange(o.mx)[None, :]
    dead = ((m + n) > o.trunc + 1) | (n > o.trunc)
    assert np.all(got[:, dead] == 0)
    assert np.all(got[:, :, 0].imag == 0)
    c.close()


@pytest.mark.parametrize("nb", [584, 1201, 1202, 1203])
def test_spec_to_grid_quad(pkg, oracle, nb):
    """spec->grid of a large batch through the quad kernel (FFT + DMMA Legendre, four fields at a time) vs the oracle and the streaming kernel"""
    o = oracle
    c = pkg.Speedy(trunc=30)
    rng = np.random.default_rng(97)
    s = random_spec(rng, (nb,), o.nx, o.mx, o.trunc)
    n = np.arange(o.nx)[:, None]; m = np.arange(o.mx)[None, :]
    s_dirty = s + 1e6 * ((m + n) > o.trunc + 1)                      # garbage outside the triangle must be ignored (legendre.f90:38)
    kcos = np.where(np.arange(nb) % 3 == 0, 1, 2).astype(np.int32)
    _select(c, "quad")
    got = c.spec_to_grid(s_dirty, kcos)
    _select(c, "stream")
    base = c.spec_to_grid(s_dirty, kcos)
    idx = np.r_[0:8, nb // 2:nb // 2 + 8, nb - 8:nb]
    assert rel_rms(got[idx], o.spec_to_grid(s[idx], kcos[idx])) < 1e-12
    assert rel_rms(got, base) < 1e-12
    c.close()


@pytest.mark.parametrize("nb", [8, 91])
def test_spec_to_grid_dense_inverse(pkg, oracle, nb):
    o = oracle
    c = pkg.Speedy(trunc=30)
    c.set_option("dense_inverse", 1)
    rng = np.random.default_rng(98)
    s = random_spec(rng, (nb,), o.nx, o.mx, o.trunc)
    kcos = np.where(np.arange(nb) % 2 == 0, 1, 2).astype(np.int32)
    assert rel_rms(c.spec_to_grid(s, kcos), o.spec_to_grid(s, kcos)) < 1e-12
    c.close()


@pytest.mark.parametrize("variant", ["quad", "field", "stream", "dense_inverse"])
def test_48h_run_variant(pkg, oracle, variant):
    """eight identical members (the batch variants of the kernels) for 48 h against the oracle"""
    oracle.model_init(BC)
    assert oracle.run(72) == 0
    c = pkg.Speedy(trunc=30, nmembers=8)
    if variant == "dense_inverse":
        c.set_option("dense_inverse", 1)
    else:
        _select(c, variant)
    c.model_init(BC)
    assert c.run_steps(72) == 0
    ref = oracle.state()
    for n in PROG:
        f = c.get_field(n, all_members=True)
        assert np.array_equal(f[0], f[7]), n
        e = rel_rms(f[0], ref[n])
        assert e < 1e-10, (n, e)
    for n in ("iptop", "icnv", "icltop"):
        assert np.array_equal(c.get_field(n), oracle.ifield(n)), n
    c.close()


@pytest.mark.parametrize("sppt", [0, 1])
def test_step_plumbing_is_bitwise_neutral(pkg, sppt):
    """what only changes WHEN or WHERE the ensemble step moves its data must not change a bit: the per-member hand-off between the quad
    spec->grid kernel and the column kernel (member_ready.cuh), the L2 discards of the transient grid fields, the shared transient buffer
    (grid tendencies over the staged grid fields, coefficients over their own grid rows), the SPPT pattern drawn in the spectral step's prologue
    instead of a kernel of its own, CUDA graphs vs plain launches.
    72 steps of 8 members (SPPT members differ from each other)."""
    out = {}
    cases = {"plain": dict(member_ready=0, l2_discard=0, transient_alias=0, sppt_fold=0, graphs=1),
             "all": dict(member_ready=1, l2_discard=1, transient_alias=1, sppt_fold=1, graphs=1),
             "all, no graphs": dict(member_ready=1, l2_discard=1, transient_alias=1, sppt_fold=1, graphs=0),
             "no alias": dict(member_ready=1, l2_discard=1, transient_alias=0, sppt_fold=1, graphs=1),
             "no hand-off, no fold": dict(member_ready=0, l2_discard=1, transient_alias=1, sppt_fold=0, graphs=1)}
    for name, opts in cases.items():
        c = pkg.Speedy(trunc=30, nmembers=8, sppt_on=sppt, seed=11)
        for k, v in opts.items():
            c.set_option(k, v)
        c.model_init(BC)
        assert c.run_steps(72) == 0
        out[name] = {n: c.get_field(n, all_members=True) for n in PROG + ("sst_om", "tau2", "hfluxn", "precnv")}
        c.close()
    base = out["plain"]
    if sppt:
        assert not np.array_equal(base["t"][0], base["t"][7])
    for name in cases:
        for n, v in base.items():
            assert np.array_equal(out[name][n], v), (name, n)


def test_step_plumbing_is_bitwise_neutral_over_a_month_boundary(pkg):
    """the same over 35 days (1260 steps, into February: daily forcing and its gated humidity-correction transform 35 times, a new
    climatology month), 8 SPPT members, everything on against everything off"""
    out = {}
    for name, v in (("plain", 0), ("all", 1)):
        c = pkg.Speedy(trunc=30, nmembers=8, sppt_on=1, seed=5)
        for k in ("member_ready", "l2_discard", "transient_alias", "sppt_fold"):
            c.set_option(k, v)
        c.model_init(BC)
        assert c.run_steps(35 * 36) == 0
        out[name] = {n: c.get_field(n, all_members=True) for n in PROG + ("sst_om", "stl_am", "tau2", "qcorh")}
        c.close()
    for n, v in out["plain"].items():
        assert np.isfinite(v).all(), n
        assert np.array_equal(out["all"][n], v), n
