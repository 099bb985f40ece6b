"""speedy.f90_b200 — host-side Python mirror of the reference's module interfaces
(`spectral`, `legendre`, `fourier`, `tendencies`, `time_stepping`, `physics`) on top of the
C ABI in include/speedy_b200.h.  The compute path is libspeedy_b200.so (hand-written CUDA
for sm_100a); there is no CPU fallback: if the library or a GPU is missing, calls raise.

Arrays follow the reference's Fortran layout.  In numpy (C order) that means spectral
fields are complex128 arrays of shape (..., nx, mx) and grid fields float64 (..., il, ix).
"""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.environ.get("SPEEDY_B200_LIB") or os.path.join(_HERE, "libspeedy_b200.so")   # override: A/B builds of the same library
_lib = None
BC_T30 = os.path.join(os.path.dirname(_HERE), "data", "bc_t30.bin")   # packed reference boundary files (tools/pack_boundary.py)

c_double_p = ctypes.POINTER(ctypes.c_double)


class SpeedyError(RuntimeError):
    pass


class Cfg(ctypes.Structure):
    _fields_ = [("trunc", ctypes.c_int), ("kx", ctypes.c_int), ("ntr", ctypes.c_int),
                ("nmembers", ctypes.c_int), ("device", ctypes.c_int), ("sppt_on", ctypes.c_int),
                ("seed", ctypes.c_ulonglong), ("member_offset", ctypes.c_int), ("nsteps", ctypes.c_int), ("precision", ctypes.c_int)]


class Namelist(ctypes.Structure):
    """speedy_namelist: the &params / &date groups of namelist.nml (params.f90:46-70, date.f90:54-71)"""
    _fields_ = [("nsteps_out", ctypes.c_int), ("nstdia", ctypes.c_int),
                ("start_datetime", ctypes.c_int * 5), ("end_datetime", ctypes.c_int * 5)]

    def as_dict(self):
        return {"nsteps_out": self.nsteps_out, "nstdia": self.nstdia,
                "start_datetime": tuple(self.start_datetime), "end_datetime": tuple(self.end_datetime)}


def lib():
    """Load libspeedy_b200.so (built in-tree by `make -C speedy.f90_b200`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIBPATH):
            raise SpeedyError(f"{_LIBPATH} not built: run __graft_entry__.build() / make -C speedy.f90_b200")
        _lib = ctypes.CDLL(_LIBPATH)
        for name, rt in (("speedy_last_error", ctypes.c_char_p), ("speedy_launch_count", ctypes.c_longlong),
                         ("speedy_stream", ctypes.c_void_p), ("speedy_host_table_len", ctypes.c_longlong),
                         ("speedy_output_len", ctypes.c_size_t), ("speedy_state_len", ctypes.c_size_t),
                         ("speedy_field_names", ctypes.c_char_p), ("speedy_steps_between", ctypes.c_longlong),
                         ("speedy_host_boundary", ctypes.c_longlong)):
            getattr(_lib, name).restype = rt
        _lib.speedy_set_field.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t]
        _lib.speedy_get_field.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t]
        _lib.speedy_get_ifield.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t]
        _lib.speedy_get_physical_tendencies.argtypes = [ctypes.c_void_p] * 11 + [ctypes.c_int]
        _lib.speedy_output_fields.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 6
        _lib.speedy_step_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int]
        _lib.speedy_run_steps_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
        _lib.speedy_model_init.argtypes = [ctypes.c_void_p, ctypes.c_char_p] + [ctypes.c_int] * 5
        _lib.speedy_write_output.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
        _lib.speedy_write_output_file.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 6
        _lib.speedy_save_restart.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        _lib.speedy_load_restart.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        _lib.speedy_set_option.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
        _lib.speedy_write_output_async.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_longlong]
        _lib.speedy_host_boundary.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t]
        _lib.speedy_read_namelist.argtypes = [ctypes.c_char_p, ctypes.POINTER(Namelist)]
        _lib.speedy_namelist_defaults.argtypes = [ctypes.POINTER(Namelist)]
        _lib.speedy_steps_between.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _lib.speedy_main_loop.argtypes = [ctypes.c_void_p, ctypes.POINTER(Namelist), ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _chk(rc):
    if rc < 0:
        raise SpeedyError(lib().speedy_last_error().decode())
    return rc


def host_table(trunc, name):
    """Start-up table `name` for truncation `trunc`, built on the host (no GPU needed)."""
    L = lib()
    n = L.speedy_host_table_len(trunc, name.encode())
    if n < 0:
        raise SpeedyError(f"unknown table {name}")
    out = np.zeros(n)
    _chk(L.speedy_host_table(trunc, name.encode(), _p(out), ctypes.c_size_t(n)))
    return out


def host_boundary(bc_path, name, trunc=30, n=None):
    """Start-up boundary field `name` (host-only) from a packed .bin or a directory of the reference's NetCDF-4 files; `n`: leading values only."""
    L = lib()
    total = L.speedy_host_boundary(str(bc_path).encode(), int(trunc), name.encode(), None, 0)
    if total < 0:
        raise SpeedyError(L.speedy_last_error().decode())
    out = np.zeros(total if n is None else min(int(n), total))
    if L.speedy_host_boundary(str(bc_path).encode(), int(trunc), name.encode(), _p(out), ctypes.c_size_t(out.size)) < 0:
        raise SpeedyError(L.speedy_last_error().decode())
    return out


def read_namelist(path=None):
    """initialize_params + the namelist part of initialize_date (host-only): `path` None or missing -> the reference's defaults."""
    nml = Namelist()
    _chk(lib().speedy_read_namelist(None if path is None else str(path).encode(), ctypes.byref(nml)))
    return nml


def steps_between(start, end, nsteps=36):
    """trips of the reference's main loop from `start` to `end` ((y, m, d, h, mi)); raises if the end date is never met."""
    a = np.ascontiguousarray(start, dtype=np.int32)
    b = np.ascontiguousarray(end, dtype=np.int32)
    n = lib().speedy_steps_between(_p(a), _p(b), int(nsteps))
    if n < 0:
        raise SpeedyError(lib().speedy_last_error().decode())
    return n


def write_output_file(path, u, v, t, q, phi, ps, trunc=30, nsteps=36, start=(1982, 1, 1, 0, 0), timestep=0):
    """Host-only writer of one output() file (input_output.f90:95-217) from float32 fields in C order (kx, il, ix) / (il, ix)."""
    st = np.ascontiguousarray(start, dtype=np.int32)
    f = [np.ascontiguousarray(a, dtype=np.float32) for a in (u, v, t, q, phi, ps)]
    _chk(lib().speedy_write_output_file(str(path).encode(), int(trunc), int(nsteps), _p(st), int(timestep), *[_p(a) for a in f]))


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Speedy:
    """One context = one GPU + one batch of ensemble members (speedy_ctx)."""

    def __init__(self, trunc=30, nmembers=1, device=0, sppt_on=0, seed=0, member_offset=0, nsteps=0, precision=0):
        L = lib()
        cfg = Cfg(trunc, 8, 1, nmembers, device, sppt_on, seed, member_offset, nsteps, precision)
        self.nsteps = nsteps or 36
        h = ctypes.c_void_p()
        _chk(L.speedy_create(ctypes.byref(cfg), ctypes.byref(h)))
        self.h = h
        self.L = L
        d = (ctypes.c_int * 8)()
        _chk(L.speedy_dims(h, d))
        self.trunc, self.ix, self.iy, self.il, self.kx, self.nx, self.mx, self.ntr = list(d)
        self.nmembers = nmembers

    def close(self):
        if self.h:
            self.L.speedy_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- tables -------------------------------------------------------------------
    def table(self, name, n):
        out = np.zeros(n)
        _chk(self.L.speedy_get_table(self.h, name.encode(), _p(out), ctypes.c_size_t(n)))
        return out

    def set_table(self, name, arr):
        a = _c(arr, np.float64).ravel()
        _chk(self.L.speedy_set_table(self.h, name.encode(), _p(a), ctypes.c_size_t(a.size)))

    # ---- module spectral / legendre / fourier ----------------------------------------
    def _batch(self, a, tail):
        a = np.asarray(a)
        if a.shape[-len(tail):] != tuple(tail):
            raise ValueError(f"expected trailing shape {tail}, got {a.shape}")
        lead = a.shape[:-len(tail)]
        return lead, int(np.prod(lead)) if lead else 1

    def spec_to_grid(self, vorm, kcos=1):
        """spectral.f90:98 — vorm complex (..., nx, mx) -> grid (..., il, ix)."""
        lead, nb = self._batch(vorm, (self.nx, self.mx))
        a = _c(vorm, np.complex128)
        k = np.broadcast_to(np.asarray(kcos, dtype=np.int32), lead).astype(np.int32).ravel() if lead else np.array([kcos], np.int32)
        k = _c(k, np.int32)
        out = np.empty(lead + (self.il, self.ix))
        _chk(self.L.speedy_spec_to_grid(self.h, _p(a), nb, _p(k), _p(out)))
        return out

    def grid_to_spec(self, vorg):
        """spectral.f90:112 — grid (..., il, ix) -> complex (..., nx, mx)."""
        lead, nb = self._batch(vorg, (self.il, self.ix))
        a = _c(vorg, np.float64)
        out = np.empty(lead + (self.nx, self.mx), dtype=np.complex128)
        _chk(self.L.speedy_grid_to_spec(self.h, _p(a), nb, _p(out)))
        return out

    def legendre_inv(self, x):
        """legendre.f90:74 — real (..., nx, 2mx) -> (..., il, 2mx)."""
        lead, nb = self._batch(x, (self.nx, 2 * self.mx))
        a = _c(x, np.float64)
        out = np.empty(lead + (self.il, 2 * self.mx))
        _chk(self.L.speedy_legendre_inv(self.h, _p(a), nb, _p(out)))
        return out

    def legendre_dir(self, x):
        """legendre.f90:114 — real (..., il, 2mx) -> (..., nx, 2mx)."""
        lead, nb = self._batch(x, (self.il, 2 * self.mx))
        a = _c(x, np.float64)
        out = np.empty(lead + (self.nx, 2 * self.mx))
        _chk(self.L.speedy_legendre_dir(self.h, _p(a), nb, _p(out)))
        return out

    def fourier_inv(self, x, kcos=1):
        """fourier.f90:23 — real (..., il, 2mx) -> (..., il, ix)."""
        lead, nb = self._batch(x, (self.il, 2 * self.mx))
        a = _c(x, np.float64)
        k = np.full(nb, kcos, dtype=np.int32)
        out = np.empty(lead + (self.il, self.ix))
        _chk(self.L.speedy_fourier_inv(self.h, _p(a), nb, _p(k), _p(out)))
        return out

    def fourier_dir(self, x):
        """fourier.f90:56 — real (..., il, ix) -> (..., il, 2mx)."""
        lead, nb = self._batch(x, (self.il, self.ix))
        a = _c(x, np.float64)
        out = np.empty(lead + (self.il, 2 * self.mx))
        _chk(self.L.speedy_fourier_dir(self.h, _p(a), nb, _p(out)))
        return out

    def _op2(self, fn, a, b, two_out=True):
        lead, nb = self._batch(a, (self.nx, self.mx))
        a = _c(a, np.complex128)
        o1 = np.empty_like(a)
        o2 = np.empty_like(a)
        if b is None:
            _chk(fn(self.h, _p(a), nb, _p(o1), _p(o2)) if two_out else fn(self.h, _p(a), nb, _p(o1)))
        else:
            b = _c(b, np.complex128)
            _chk(fn(self.h, _p(a), _p(b), nb, _p(o1), _p(o2)))
        return (o1, o2) if two_out else o1

    def laplacian(self, x):
        return self._op2(self.L.speedy_laplacian, x, None, two_out=False)

    def inverse_laplacian(self, x):
        return self._op2(self.L.speedy_inverse_laplacian, x, None, two_out=False)

    def grad(self, psi):
        return self._op2(self.L.speedy_grad, psi, None)

    def vds(self, ucosm, vcosm):
        return self._op2(self.L.speedy_vds, ucosm, vcosm)

    def uvspec(self, vorm, divm):
        return self._op2(self.L.speedy_uvspec, vorm, divm)

    def trunct(self, x):
        lead, nb = self._batch(x, (self.nx, self.mx))
        a = _c(x, np.complex128).copy()
        _chk(self.L.speedy_trunct(self.h, _p(a), nb))
        return a

    def vdspec(self, ug, vg, kcos=2):
        lead, nb = self._batch(ug, (self.il, self.ix))
        a = _c(ug, np.float64)
        b = _c(vg, np.float64)
        o1 = np.empty(lead + (self.nx, self.mx), dtype=np.complex128)
        o2 = np.empty_like(o1)
        _chk(self.L.speedy_vdspec(self.h, _p(a), _p(b), nb, kcos, _p(o1), _p(o2)))
        return o1, o2


    # ---- model state (prognostics.f90 / physics / slab module state) -------------------
    def field_len(self, name):
        shapes = self.field_shapes()
        return int(np.prod(shapes[name][0]))

    def field_shapes(self):
        """name -> (numpy shape, dtype) per member, C order (= reversed Fortran shape)."""
        nx, mx, kx, il, ix = self.nx, self.mx, self.kx, self.il, self.ix
        c, f = np.complex128, np.float64
        sh = {"vor": ((2, kx, nx, mx), c), "div": ((2, kx, nx, mx), c), "t": ((2, kx, nx, mx), c), "tr": ((2, kx, nx, mx), c),
              "ps": ((2, nx, mx), c), "phi": ((kx, nx, mx), c), "phis": ((nx, mx), c), "tcorh": ((nx, mx), c), "qcorh": ((nx, mx), c),
              "vordt": ((kx, nx, mx), c), "divdt": ((kx, nx, mx), c), "tdt": ((kx, nx, mx), c), "trdt": ((kx, nx, mx), c), "psdt": ((nx, mx), c),
              "sppt_spec": ((kx, nx, mx), c), "sppt_eta": ((kx, nx, mx), c), "phi_next": ((kx, nx, mx), c), "sout": ((74, nx, mx), c),
              "gin": ((99, il, ix), f), "gout": ((74, il, ix), f), "sstan3": ((3, il, ix), f), "tau2": ((4, kx, il, ix), f),
              "stratc": ((2, il, ix), f), "tt_rsw": ((kx, il, ix), f)}
        for n in ("slru", "ustr", "vstr", "shf", "evap", "hfluxn"):
            sh[n] = ((3, il, ix), f)
        for n in ("phis0 fmask_l fmask_s forog alb0 fsol ozone ozupp zenit stratz alb_l alb_s albsfc snowc stl_am stl_lm snowd_am "
                  "soilw_am sst_am sice_am tice_am ssti_om sst_om tice_om sice_om sstcl_ob sicecl_ob ticecl_ob stlcl_ob ssrd ssr tsr "
                  "precnv precls cbmf slrd slr olr ts tskin u0 v0 t0 qcloud cloudc clstr qcorh_g").split():
            sh[n] = ((il, ix), f)
        for n in ("iptop", "icltop", "icnv"):
            sh[n] = ((il, ix), np.int32)
        return sh

    def get_field(self, name, all_members=False):
        shape, dt = self.field_shapes()[name]
        lead = (self.nmembers,) if all_members else ()
        out = np.empty(lead + shape, dtype=dt)
        if dt == np.int32:
            _chk(self.L.speedy_get_ifield(self.h, name.encode(), _p(out), ctypes.c_size_t(out.size)))
        else:
            _chk(self.L.speedy_get_field(self.h, name.encode(), _p(out), ctypes.c_size_t(out.size * (2 if dt == np.complex128 else 1))))
        return out

    def set_field(self, name, arr):
        """arr has the per-member shape (broadcast to all members) or a leading nmembers axis."""
        shape, dt = self.field_shapes()[name]
        a = _c(arr, dt)
        if a.shape != shape and a.shape != (self.nmembers,) + shape:
            raise ValueError(f"{name}: expected shape {shape}, got {a.shape}")
        _chk(self.L.speedy_set_field(self.h, name.encode(), _p(a), ctypes.c_size_t(a.size * (2 if dt == np.complex128 else 1))))

    # ---- module time_stepping / tendencies / physics / coupler -------------------------------
    def model_init(self, bc_path, year=1982, month=1, day=1, hour=0, minute=0):
        """initialization.f90:12 — boundary data, rest state, coupler, forcing, first_step."""
        _chk(self.L.speedy_model_init(self.h, str(bc_path).encode(), year, month, day, hour, minute))

    def implicit_terms(self, divdt, tdt, psdt):
        """implicit.f90:168-217 — returns the corrected (divdt, tdt, psdt); complex (kx, nx, mx) x 2 and (nx, mx)."""
        a, b, c = (np.array(x, dtype=np.complex128, order="C") for x in (divdt, tdt, psdt))
        assert a.shape == (self.kx, self.nx, self.mx) and b.shape == a.shape and c.shape == (self.nx, self.mx)
        _chk(self.L.speedy_implicit_terms(self.h, _p(a), _p(b), _p(c)))
        return a, b, c

    def do_horizontal_diffusion(self, field, fdt, dmp, dmp1):
        """horizontal_diffusion.f90:86-105 — (fdt - dmp*field)*dmp1 for a complex (nx, mx) or (nlev, nx, mx) field."""
        f = _c(field, np.complex128)
        t = np.array(fdt, dtype=np.complex128, order="C")
        d, d1 = _c(dmp, np.float64), _c(dmp1, np.float64)
        assert f.shape == t.shape and d.shape == (self.nx, self.mx) and d1.shape == d.shape
        nlev = 1 if f.ndim == 2 else f.shape[0]
        _chk(self.L.speedy_do_horizontal_diffusion(self.h, _p(f), _p(t), _p(d), _p(d1), int(nlev)))
        return t

    def initialize_implicit(self, dt):
        _chk(self.L.speedy_initialize_implicit(self.h, ctypes.c_double(dt)))

    def step(self, j1, j2, dt, compute_shortwave=True):
        """time_stepping.f90:35 on the resident state."""
        _chk(self.L.speedy_step(self.h, j1, j2, ctypes.c_double(dt), int(compute_shortwave)))

    def first_step(self):
        _chk(self.L.speedy_first_step(self.h))

    def get_tendencies(self, j2, compute_shortwave=True):
        """tendencies.f90:11 — returns (vordt, divdt, tdt, psdt, trdt) of member 0."""
        _chk(self.L.speedy_get_tendencies(self.h, j2, int(compute_shortwave)))
        return tuple(self.get_field(n) for n in ("vordt", "divdt", "tdt", "psdt", "trdt"))

    def get_physical_tendencies(self, vor, div, t, q, phi, psl, utend, vtend, ttend, qtend, compute_shortwave=True):
        """physics.f90:43 — spectral inputs (kx,nx,mx) [psl (nx,mx)], grid tendencies (kx,il,ix) returned updated."""
        a = [_c(x, np.complex128) for x in (vor, div, t, q, phi, psl)]
        g = [_c(x, np.float64).copy() for x in (utend, vtend, ttend, qtend)]
        _chk(self.L.speedy_get_physical_tendencies(self.h, *[_p(x) for x in a], *[_p(x) for x in g], int(compute_shortwave)))
        return tuple(g)

    def run_steps(self, nsteps):
        """speedy.f90:27-54 main-loop body, nsteps times; returns 1 if check_diagnostics tripped."""
        return _chk(self.L.speedy_run_steps(self.h, int(nsteps)))

    def enqueue_steps(self, nsteps):
        """main-loop body nsteps times, enqueue only (no host synchronisation); pair with finish()"""
        _chk(self.L.speedy_enqueue_steps(self.h, int(nsteps)))

    def finish(self):
        """drain the stream; returns 1 if check_diagnostics tripped since the last finish / run_steps"""
        return _chk(self.L.speedy_finish(self.h))

    def couple_sea_land(self, day):
        _chk(self.L.speedy_couple_sea_land(self.h, int(day)))

    def set_forcing(self, imode=1):
        _chk(self.L.speedy_set_forcing(self.h, int(imode)))

    def check_diagnostics(self, time_level=2):
        d = np.zeros(3 * self.kx)
        rc = _chk(self.L.speedy_check_diagnostics(self.h, time_level, _p(d)))
        return rc, d.reshape(3, self.kx)

    def model_date(self):
        d = (ctypes.c_int * 5)()
        s = ctypes.c_longlong()
        _chk(self.L.speedy_model_date(self.h, d, ctypes.byref(s)))
        return tuple(d), s.value

    def output_fields(self, member=0):
        """input_output.f90:184-206 — float32 u,v,t,q,phi (kx,il,ix) and ps (il,ix)."""
        k, il, ix = self.kx, self.il, self.ix
        o = [np.empty((k, il, ix), np.float32) for _ in range(5)] + [np.empty((il, ix), np.float32)]
        _chk(self.L.speedy_output_fields(self.h, member, *[_p(x) for x in o]))
        return dict(zip(("u", "v", "t", "q", "phi", "ps"), o))

    def write_output(self, directory=".", member=0):
        """output() (input_output.f90:95-217): writes `yyyymmddhhmm.nc` (NetCDF classic) for the resident state; returns the path."""
        buf = ctypes.create_string_buffer(4096)
        _chk(self.L.speedy_write_output(self.h, int(member), str(directory).encode(), buf, ctypes.c_size_t(len(buf))))
        return buf.value.decode()

    def write_output_async(self, directory, ymdhm, timestep, member=0):
        """output() without waiting: conversions enqueued on the stream, the file written by the library's host threads.  `ymdhm`,
        `timestep`: date and step counter - 1 of the enqueued state (the caller knows the calendar).  output_drain() waits for the files."""
        d = (ctypes.c_int * 5)(*[int(x) for x in ymdhm])
        _chk(self.L.speedy_write_output_async(self.h, int(member), str(directory).encode(), d, ctypes.c_longlong(int(timestep))))

    def output_drain(self):
        _chk(self.L.speedy_output_drain(self.h))

    def main_loop(self, nml, out_dir=None, member=0, verbose=False):
        """program speedy (speedy.f90:24-54) after model_init(start date of `nml`): output every nsteps_out steps into `out_dir`
        (None: no files), diagnostics every nstdia steps to stdout when verbose.  Returns (rc, main-loop steps done); rc 1 = range error."""
        done = ctypes.c_longlong()
        rc = _chk(self.L.speedy_main_loop(self.h, ctypes.byref(nml), None if out_dir is None else str(out_dir).encode(), int(member), int(bool(verbose)), ctypes.byref(done)))
        return rc, done.value

    def run_info(self):
        d = (ctypes.c_int * 4)()
        _chk(self.L.speedy_run_info(self.h, d))
        return {"nmembers": d[0], "nsteps": d[1], "sppt_on": d[2], "precision": d[3]}

    def save_restart(self, path):
        _chk(self.L.speedy_save_restart(self.h, str(path).encode()))

    def load_restart(self, path):
        _chk(self.L.speedy_load_restart(self.h, str(path).encode()))

    def step_host(self, state, j1, j2, dt, compute_shortwave=True):
        a = _c(state, np.float64).copy()
        _chk(self.L.speedy_step_host(self.h, _p(a), ctypes.c_size_t(a.size), j1, j2, ctypes.c_double(dt), int(compute_shortwave)))
        return a

    def time_kernels(self, nsteps=36, flush_l2=False):
        """mean CUDA-event ms per kernel of the main-loop body -> dict name -> ms"""
        self.L.speedy_kernel_names.restype = ctypes.c_char_p
        names = self.L.speedy_kernel_names().decode().split()
        ms = np.zeros(len(names))
        _chk(self.L.speedy_time_kernels(self.h, int(nsteps), int(flush_l2), _p(ms)))
        return dict(zip(names, ms.tolist()))

    def trace(self, on=True):
        """in-graph timeline of the four main-loop kernels (GPU global timer); read with trace_read()"""
        _chk(self.L.speedy_trace(self.h, int(bool(on))))

    def trace_read(self):
        o = np.zeros(9)
        _chk(self.L.speedy_trace_read(self.h, _p(o)))
        self.L.speedy_kernel_names.restype = ctypes.c_char_p
        names = self.L.speedy_kernel_names().decode().split()
        return {"us": dict(zip(names, o[:4].tolist())), "gap_before_us": dict(zip(names, o[4:8].tolist())), "steps": int(o[8])}

    def state_len(self):
        return int(self.L.speedy_state_len(self.h))

    def run_steps_host(self, state, nsteps, out=None):
        """Main loop with the prognostic state resident in HOST memory: `state` (float64, nmembers*state_len,
        [vor,div,t,tr,ps] per member) is uploaded, advanced nsteps steps and downloaded IN PLACE;
        `out` (float32, output_len) receives member 0's output() fields.  Pinned arrays copy asynchronously."""
        assert state.dtype == np.float64 and state.flags.c_contiguous
        po = None
        if out is not None:
            assert out.dtype == np.float32 and out.flags.c_contiguous
            po = _p(out)
        return _chk(self.L.speedy_run_steps_host(self.h, _p(state), ctypes.c_size_t(state.size), int(nsteps), po))

    def set_sppt_draw(self, on):
        """sppt.f90:45-99 — draw eta on the device (default) or read it from the `sppt_eta` field"""
        _chk(self.L.speedy_set_sppt_draw(self.h, int(bool(on))))

    def set_option(self, name, value):
        """kernel-selection switch of the context (speedy_set_option): "k2_field", "dense_inverse", "graphs"."""
        _chk(self.L.speedy_set_option(self.h, name.encode(), int(value)))

    def set_graphs(self, on):
        _chk(self.L.speedy_set_graphs(self.h, int(bool(on))))

    # ---- misc -----------------------------------------------------------------------
    def synchronize(self):
        _chk(self.L.speedy_synchronize(self.h))

    @property
    def launch_count(self):
        return self.L.speedy_launch_count(self.h)

    @property
    def stream(self):
        return self.L.speedy_stream(self.h)
