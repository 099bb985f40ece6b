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
