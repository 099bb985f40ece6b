"""csrc/fft96.cuh, csrc/fft144.cuh — FFTPACK's backward passes regrouped into two register-resident stages (K1's Fourier stage).
The header is built for the host (device qualifiers compiled away) and must reproduce the oracle's pass-by-pass rfftb1
(fftpack.f90:69-134) bit for bit: same butterflies, same twiddle table, only the order of independent work differs."""
import ctypes
import os
import subprocess

import numpy as np

from conftest import ROOT, load_pkg


def test_regrouped_backward_fft_is_fftpack_bit_for_bit(oracle, tmp_path):
    so = tmp_path / "fft96_host.so"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so),
                           os.path.join(ROOT, "tests", "helpers", "fft96_host.cpp")])
    L = ctypes.CDLL(str(so))
    wa = np.ascontiguousarray(load_pkg().host_table(30, "fft_work"))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rng = np.random.default_rng(7)
    for trial in range(40):
        c = rng.standard_normal(96) * 10.0 ** rng.integers(-3, 4)
        if trial % 2:
            c[61:] = 0.0                      # what fourier_inv feeds it: wavenumbers above the truncation are zero
        ref = c.copy()
        oracle.L.orc_rfftb(P(ref))
        out = np.zeros(96)
        L.fft96_backward(P(c), P(wa), P(out))
        assert np.array_equal(out, ref)


def test_regrouped_backward_fft_t47(oracle47, tmp_path):
    so = tmp_path / "fft144_host.so"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so),
                           os.path.join(ROOT, "tests", "helpers", "fft144_host.cpp")])
    L = ctypes.CDLL(str(so))
    wa = np.ascontiguousarray(load_pkg().host_table(47, "fft_work"))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rng = np.random.default_rng(8)
    for trial in range(40):
        c = rng.standard_normal(144) * 10.0 ** rng.integers(-3, 4)
        if trial % 2:
            c[95:] = 0.0
        ref = c.copy()
        oracle47.L.orc_rfftb(P(ref))
        out = np.zeros(144)
        L.fft144_backward(P(c), P(wa), P(out))
        assert np.array_equal(out, ref)


def test_regrouped_forward_fft_is_fftpack_bit_for_bit(oracle, tmp_path):
    """csrc/fft96f.cuh (building block of a whole-field grid->spec kernel, not yet used on the device): the forward passes
    radf3/radf4/radf4/radf2 regrouped into two stages == the oracle's pass-by-pass rfftf1 (fftpack.f90:136-202)"""
    so = tmp_path / "fft96f_host.so"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so),
                           os.path.join(ROOT, "tests", "helpers", "fft96f_host.cpp")])
    L = ctypes.CDLL(str(so))
    wa = np.ascontiguousarray(load_pkg().host_table(30, "fft_work"))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rng = np.random.default_rng(9)
    for trial in range(40):
        g = rng.standard_normal(96) * 10.0 ** rng.integers(-3, 4)
        ref = g.copy()
        oracle.L.orc_rfftf(P(ref))
        out = np.full(96, np.nan)
        L.fft96_forward(P(g), P(wa), P(out))
        assert np.array_equal(out, ref)


def test_whole_field_grid_to_spec_dataflow(oracle, tmp_path):
    """The data flow of the experimental whole-field grid->spec kernel (k_g2s_field, transforms.cu), emulated on the host with
    the same pieces: regrouped forward FFT per row, fourier_dir's real32 1/ix, half-complex position <-> coefficient row,
    Gaussian-weighted even/odd folds, direct Legendre sums over the triangle-packed P table with two n of equal parity per
    task.  Must equal the oracle's grid_to_spec (spectral.f90:112-122) exactly."""
    so = tmp_path / "fft96f_host.so"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so),
                           os.path.join(ROOT, "tests", "helpers", "fft96f_host.cpp")])
    L = ctypes.CDLL(str(so))
    S = load_pkg()
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    IX, IL, IY, MX, NX, TRUNC = 96, 48, 24, 31, 32, 30
    wa = np.ascontiguousarray(S.host_table(30, "fft_work"))
    poly = S.host_table(30, "poly").reshape(IY, NX, MX)
    wt = S.host_table(30, "wt")
    tri_cnt = lambda n: min(MX, MX - n + 1)
    tri_off = np.concatenate([[0], np.cumsum([tri_cnt(n) for n in range(NX)])])
    polyt = np.zeros((IY, (tri_off[NX] + 1) // 2 * 2))
    for n in range(NX):
        polyt[:, tri_off[n]:tri_off[n] + tri_cnt(n)] = poly[:, n, :tri_cnt(n)]
    g = np.random.default_rng(5).standard_normal((IL, IX))
    Y = np.zeros((IL, IX))
    for r in range(IL):
        row, out = np.ascontiguousarray(g[r]), np.zeros(IX)
        L.fft96_forward(P(row), P(wa), P(out))
        Y[r] = out
    scale = float(np.float32(1.0) / np.float32(IX))
    EO = np.zeros((2, IY, 64))
    for c in range(2 * MX):
        if c == 1:
            continue                              # Im(m = 0) = 0 (fourier.f90:76)
        pos = 0 if c == 0 else c - 1
        south, north = Y[:IY, pos] * scale, Y[IL - 1 - np.arange(IY), pos] * scale
        EO[0, :, c] = (north + south) * wt
        EO[1, :, c] = (north - south) * wt
    spec = np.zeros((NX, MX), complex)
    for n in range(NX):
        for m in range(MX):
            if n <= TRUNC and m + n <= MX:
                re = im = 0.0
                for jh in range(IY):              # ascending latitude, as legendre.f90:144,152
                    p = polyt[jh, tri_off[n] + m]
                    re += p * EO[n & 1, jh, 2 * m]
                    im += p * EO[n & 1, jh, 2 * m + 1]
                spec[n, m] = complex(re, im)
    assert np.array_equal(spec, oracle.grid_to_spec(g[None])[0])
