"""Product host tables (speedy.f90_b200/csrc/host/tables.cpp) against the oracle's
independently written start-up code; no GPU needed."""
import numpy as np
import pytest

NAMES = ["wt", "sia", "coa", "cosgr", "cosgr2", "coriol", "radang", "hsg", "fsg", "dhs", "cpol", "epsi",
         "fft_work", "el2", "elm2", "trfilt", "gradx", "gradym", "gradyp", "uvdx", "uvdym", "uvdyp", "vddym", "vddyp"]


@pytest.mark.parametrize("res", [30, 47])
def test_tables_bit_identical(pkg, oracle, oracle47, res):
    orc = oracle if res == 30 else oracle47
    for name in NAMES:
        a = pkg.host_table(res, name)
        b = orc.table(name, a.size)
        assert np.array_equal(a, b), name


@pytest.mark.parametrize("res", [30, 47])
def test_dense_fourier_operators_match_fftpack(pkg, oracle, oracle47, res):
    """F13: the dense operators applied on the GPU must be FFTPACK's operator."""
    orc = oracle if res == 30 else oracle47
    kp = (2 * orc.mx + 3) // 4 * 4
    finv = pkg.host_table(res, "finv").reshape(orc.ix, kp)
    ffwd = pkg.host_table(res, "ffwd").reshape(kp, orc.ix)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((orc.il, 2 * orc.mx))
    ref = orc.fourier_inv(x)
    mine = x @ finv[:, : 2 * orc.mx].T
    assert np.abs(mine - ref).max() / np.abs(ref).max() < 5e-15
    g = rng.standard_normal((orc.il, orc.ix))
    ref = orc.fourier_dir(g)
    mine = g @ ffwd[: 2 * orc.mx].T
    assert np.abs(mine - ref).max() / np.abs(ref).max() < 5e-15
    assert np.all(finv[:, 1] == 0) and np.all(ffwd[1] == 0)      # Im(m=0): fourier.f90:33,76
    exact = np.fft.irfft(np.eye(orc.ix // 2 + 1)[2], n=orc.ix) * orc.ix
    assert 1e-10 < np.abs(finv[:, 4] - exact).max() < 1e-5        # perturbed, not the exact DFT
