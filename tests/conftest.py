import ctypes
import importlib.util
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_pkg():
    name = "speedy_f90_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(ROOT, "speedy.f90_b200", "__init__.py"),
        submodule_search_locations=[os.path.join(ROOT, "speedy.f90_b200")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


class Oracle:
    """ctypes view of oracle/liboracle_<res>.so — the CPU checker (test infrastructure)."""

    def __init__(self, res="t30"):
        path = os.path.join(ROOT, "oracle", f"liboracle_{res}.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), f"liboracle_{res}.so"])
        self.L = ctypes.CDLL(path)
        self.L.orc_init_transforms()
        d = (ctypes.c_int * 8)()
        self.L.orc_dims(d)
        self.trunc, self.ix, self.iy, self.il, self.kx, self.nx, self.mx, self.ntr = list(d)

    @staticmethod
    def p(a):
        return a.ctypes.data_as(ctypes.c_void_p)

    def table(self, name, n):
        a = np.zeros(n)
        rc = self.L.orc_get_table(name.encode(), self.p(a), n)
        assert rc == 0, (name, rc)
        return a

    def itable(self, name, n):
        a = np.zeros(n, dtype=np.int32)
        assert self.L.orc_get_itable(name.encode(), self.p(a), n) == 0
        return a

    def spec_to_grid(self, spec, kcos=1):
        spec = np.ascontiguousarray(spec, dtype=np.complex128)
        lead = spec.shape[:-2]
        nb = int(np.prod(lead)) if lead else 1
        k = np.ascontiguousarray(np.broadcast_to(np.asarray(kcos, np.int32), lead).ravel() if lead else np.array([kcos]), dtype=np.int32)
        out = np.empty(lead + (self.il, self.ix))
        self.L.orc_spec_to_grid(self.p(spec), nb, self.p(k), self.p(out))
        return out

    def grid_to_spec(self, grid):
        grid = np.ascontiguousarray(grid, dtype=np.float64)
        lead = grid.shape[:-2]
        nb = int(np.prod(lead)) if lead else 1
        out = np.empty(lead + (self.nx, self.mx), dtype=np.complex128)
        self.L.orc_grid_to_spec(self.p(grid), nb, self.p(out))
        return out

    def _each(self, fn, x, out_tail, *extra):
        x = np.ascontiguousarray(x, dtype=np.float64)
        lead = x.shape[:-2]
        xs = x.reshape((-1,) + x.shape[-2:])
        out = np.empty((xs.shape[0],) + out_tail)
        for b in range(xs.shape[0]):
            fn(self.p(xs[b]), *extra, self.p(out[b]))
        return out.reshape(lead + out_tail)

    def legendre_inv(self, x):
        return self._each(self.L.orc_legendre_inv, x, (self.il, 2 * self.mx))

    def legendre_dir(self, x):
        return self._each(self.L.orc_legendre_dir, x, (self.nx, 2 * self.mx))

    def fourier_inv(self, x, kcos=1):
        return self._each(self.L.orc_fourier_inv, x, (self.il, self.ix), kcos)

    def fourier_dir(self, x):
        return self._each(self.L.orc_fourier_dir, x, (self.il, 2 * self.mx))

    def op2(self, name, a, b=None, nout=2):
        fn = getattr(self.L, "orc_" + name)
        a = np.ascontiguousarray(a, dtype=np.complex128)
        o1 = np.empty_like(a)
        o2 = np.empty_like(a)
        a2 = a.reshape((-1,) + a.shape[-2:])
        b2 = None if b is None else np.ascontiguousarray(b, dtype=np.complex128).reshape(a2.shape)
        v1 = o1.reshape(a2.shape)
        v2 = o2.reshape(a2.shape)
        for i in range(a2.shape[0]):
            args = [self.p(a2[i])]
            if b2 is not None:
                args.append(self.p(b2[i]))
            args.append(self.p(v1[i]))
            if nout == 2:
                args.append(self.p(v2[i]))
            fn(*args)
        return (o1, o2) if nout == 2 else o1

    # ---- model-level surface (o_api.cpp) ------------------------------------------------
    def model_init(self, bc=None, y=1982, m=1, d=1, h=0, mi=0):
        bc = bc or os.path.join(ROOT, "data", "bc_t30.bin")
        rc = self.L.orc_model_init(bc.encode(), y, m, d, h, mi)
        assert rc == 0, rc

    def run(self, nsteps):
        return self.L.orc_model_run(nsteps)

    def field(self, name, shape=None, dtype=np.float64):
        self.L.orc_field_len.restype = ctypes.c_longlong
        n = self.L.orc_field_len(name.encode())
        assert n > 0, name
        a = np.zeros(n)
        assert self.L.orc_get_field(name.encode(), self.p(a), ctypes.c_longlong(n)) == 0
        if dtype == np.complex128:
            a = a.view(np.complex128)
        return a.reshape(shape) if shape is not None else a

    def set_field(self, name, arr):
        a = np.ascontiguousarray(arr)
        a = a.view(np.float64).ravel() if a.dtype == np.complex128 else np.ascontiguousarray(a, dtype=np.float64).ravel()
        rc = self.L.orc_set_field(name.encode(), self.p(a), ctypes.c_longlong(a.size))
        assert rc == 0, (name, rc)

    def ifield(self, name):
        a = np.zeros(self.ix * self.il, dtype=np.int32)
        assert self.L.orc_get_ifield(name.encode(), self.p(a), ctypes.c_longlong(a.size)) == 0
        return a.reshape(self.il, self.ix)

    def vec(self, name, n):
        a = np.zeros(n)
        assert self.L.orc_get_vec(name.encode(), self.p(a), n) == 0, name
        return a

    def state(self):
        k, nx, mx = self.kx, self.nx, self.mx
        c = np.complex128
        return {"vor": self.field("vor", (2, k, nx, mx), c), "div": self.field("div", (2, k, nx, mx), c),
                "t": self.field("t", (2, k, nx, mx), c), "tr": self.field("tr", (2, k, nx, mx), c),
                "ps": self.field("ps", (2, nx, mx), c), "phi": self.field("phi", (k, nx, mx), c)}

    def step(self, j1, j2, dt, csw=True):
        self.L.orc_step(j1, j2, ctypes.c_double(dt), int(csw))

    def initialize_implicit(self, dt):
        self.L.orc_initialize_implicit(ctypes.c_double(dt))

    def get_tendencies(self, j2, csw=True):
        k, nx, mx = self.kx, self.nx, self.mx
        o = [np.zeros((k, nx, mx), np.complex128), np.zeros((k, nx, mx), np.complex128), np.zeros((k, nx, mx), np.complex128),
             np.zeros((nx, mx), np.complex128), np.zeros((k, nx, mx), np.complex128)]
        self.L.orc_get_tendencies(j2, int(csw), *[self.p(x) for x in o])
        return tuple(o)

    def physics(self, vor, div, t, q, phi, psl, ut, vt, tt, qt, csw=True):
        a = [np.ascontiguousarray(x, dtype=np.complex128) for x in (vor, div, t, q, phi, psl)]
        g = [np.ascontiguousarray(x, dtype=np.float64).copy() for x in (ut, vt, tt, qt)]
        self.L.orc_get_physical_tendencies(*[self.p(x) for x in a], *[self.p(x) for x in g], int(csw))
        return tuple(g)

    def check_diagnostics(self, level=2):
        d = np.zeros(24)
        rc = self.L.orc_check_diagnostics(level, self.p(d))
        return rc, d.reshape(3, 8)

    def output_fields(self):
        k, il, ix = self.kx, self.il, self.ix
        o = [np.empty((k, il, ix), np.float32) for _ in range(5)] + [np.empty((il, ix), np.float32)]
        self.L.orc_output_fields(*[self.p(x) for x in o])
        return dict(zip(("u", "v", "t", "q", "phi", "ps"), o))

    def date(self):
        d = (ctypes.c_int * 5)()
        s = ctypes.c_longlong()
        self.L.orc_model_date(d, ctypes.byref(s))
        return tuple(d), s.value


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def oracle():
    return Oracle("t30")


@pytest.fixture(scope="session")
def oracle47():
    return Oracle("t47")


@pytest.fixture(scope="session")
def oracle47_n72():
    return Oracle("t47_n72")


@pytest.fixture(scope="session")
def ctx(pkg):
    c = pkg.Speedy(trunc=30)
    yield c
    c.close()


@pytest.fixture(scope="session")
def ctx47(pkg):
    c = pkg.Speedy(trunc=47)
    yield c
    c.close()


def bc_t47():
    """synthetic T47 boundary pack (nearest-neighbour resampling of the T30 reference files), built on demand"""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_t47_boundary
    return make_t47_boundary.ensure()


def random_spec(rng, lead, nx, mx, trunc, full_triangle=True):
    """uniform(-1,1) on the nsh2-active triangle (m+n <= trunc+1), zeros elsewhere"""
    s = rng.uniform(-1, 1, size=lead + (nx, mx)) + 1j * rng.uniform(-1, 1, size=lead + (nx, mx))
    n = np.arange(nx)[:, None]
    m = np.arange(mx)[None, :]
    s = s * ((m + n) <= (trunc + 1 if full_triangle else trunc))
    return s


def rel_rms(a, b):
    return float(np.sqrt(np.mean(np.abs(a - b) ** 2)) / max(np.sqrt(np.mean(np.abs(b) ** 2)), 1e-300))
