"""The caller of the path — `program speedy` (speedy.f90:1-54): namelist groups (params.f90:46-70, date.f90:54-71) and the trip count of
the main loop, host-only parts (no GPU)."""
import ctypes
import os
import subprocess

import pytest
from conftest import ROOT, Oracle, load_pkg

# the two groups as the reference's namelist.nml spells them (component-wise datetimes, comments between the groups)
SHIPPED_STYLE = """! nsteps_out = model variables are output every nsteps_out timesteps
&params
nsteps_out = 36
nstdia     = 90
/
! the integration period
&date
start_datetime%year   = 1982
start_datetime%month  = 1
start_datetime%day    = 1
start_datetime%hour   = 0
start_datetime%minute = 0
end_datetime%year     = 1982
end_datetime%month    = 1
end_datetime%day      = 3
end_datetime%hour     = 0
end_datetime%minute   = 0
/
"""


def test_defaults_and_missing_file(tmp_path):
    """no namelist.nml: nsteps_out = 1, nstdia = 36*5 (params.f90:58-59), 1982-01-01 .. 1982-02-01 (date.f90:62-63)"""
    pkg = load_pkg()
    want = {"nsteps_out": 1, "nstdia": 180, "start_datetime": (1982, 1, 1, 0, 0), "end_datetime": (1982, 2, 1, 0, 0)}
    assert pkg.read_namelist(None).as_dict() == want
    assert pkg.read_namelist(tmp_path / "absent.nml").as_dict() == want


def test_namelist_forms(tmp_path):
    pkg = load_pkg()
    p = tmp_path / "namelist.nml"
    p.write_text(SHIPPED_STYLE)
    want = {"nsteps_out": 36, "nstdia": 90, "start_datetime": (1982, 1, 1, 0, 0), "end_datetime": (1982, 1, 3, 0, 0)}
    assert pkg.read_namelist(p).as_dict() == want
    # Fortran namelist input is case-insensitive, groups come in any order, a derived type may be assigned as a whole, a variable
    # that is not named keeps its default, blanks or commas separate values
    p.write_text("&DATE\n START_DATETIME = 1982, 1, 1, 0, 0   ! whole derived type\n End_Datetime%Day = 3, end_datetime%month=1 end_datetime%year = 1982 /\n"
                 "&Params nsteps_out=36, NSTDIA = 90 /\n")
    assert pkg.read_namelist(p).as_dict() == want
    p.write_text("&params\n/\n&date\nend_datetime%year = 1983\n/\n")
    got = pkg.read_namelist(p).as_dict()
    assert got["nsteps_out"] == 1 and got["nstdia"] == 180 and got["end_datetime"] == (1983, 2, 1, 0, 0)


@pytest.mark.parametrize("text", ["&params\nnsteps_out = 1\n/\n",                       # group &date missing: the reference's read hits end of file
                                  "&params\nnsteps = 3\n/\n&date\n/\n",                   # not a member of the group
                                  "&params\n/\n&date\nstart_datetime%second = 1\n/\n",
                                  "&params\nnsteps_out 4\n/\n&date\n/\n"])
def test_namelist_errors(tmp_path, text):
    pkg = load_pkg()
    p = tmp_path / "namelist.nml"
    p.write_text(text)
    with pytest.raises(pkg.SpeedyError):
        pkg.read_namelist(p)


def test_trip_count_follows_newdate():
    """the loop ends when the date EQUALS the end date (speedy.f90:27): count newdate calls with the oracle's calendar"""
    pkg = load_pkg()
    o = Oracle("t30")
    for start, end in (((1982, 1, 1, 0, 0), (1982, 1, 3, 0, 0)), ((1982, 1, 1, 0, 0), (1982, 1, 10, 0, 0)), ((1983, 12, 30, 12, 0), (1984, 3, 1, 0, 40)),
                       ((1982, 2, 27, 0, 0), (1982, 3, 2, 8, 0))):
        o.L.orc_calendar_init(*start)
        d = (ctypes.c_int * 5)()
        tm, ty, im = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
        n = 0
        while True:
            o.L.orc_get_date(d, ctypes.byref(tm), ctypes.byref(ty), ctypes.byref(im))
            if tuple(d) == end:
                break
            o.L.orc_newdate()
            n += 1
            assert n < 10 ** 5
        assert pkg.steps_between(start, end) == n
    assert pkg.steps_between((1982, 1, 1, 0, 0), (1982, 1, 2, 0, 0), nsteps=72) == 72
    with pytest.raises(pkg.SpeedyError):                      # 00:10 is never a model time at 40-minute steps: the reference would loop forever
        pkg.steps_between((1982, 1, 1, 0, 0), (1982, 1, 3, 0, 10))
    with pytest.raises(pkg.SpeedyError):
        pkg.steps_between((1982, 1, 3, 0, 0), (1982, 1, 1, 0, 0))


def test_namelist_struct_layout_matches_the_bind_c_type():
    """`type, bind(C) :: speedy_namelist` = 2 + 5 + 5 c_int in this order (the Fortran source declares exactly that)"""
    pkg = load_pkg()
    assert ctypes.sizeof(pkg.Namelist) == 12 * 4
    src = open(os.path.join(ROOT, "fortran", "speedy_b200_c.f90")).read()
    blk = src[src.index("type, bind(C) :: speedy_namelist"):]
    blk = blk[:blk.index("end type")]
    assert "integer(c_int) :: nsteps_out, nstdia" in blk and "integer(c_int) :: start_datetime(5), end_datetime(5)" in blk
    hdr = open(os.path.join(ROOT, "include", "speedy_b200.h")).read()
    blk = hdr[hdr.index("typedef struct speedy_namelist"):hdr.index("} speedy_namelist;")]
    order = [blk.index(n) for n in ("nsteps_out;", "nstdia;", "start_datetime[5]", "end_datetime[5]")]
    assert order == sorted(order)


def test_executable_fails_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    exe = os.path.join(ROOT, "speedy.f90_b200", "bin", "speedy_b200")
    assert os.path.exists(exe), "built by __graft_entry__.build() / make -C speedy.f90_b200"
    (tmp_path / "namelist.nml").write_text(SHIPPED_STYLE)
    r = subprocess.run([exe, "--bc", os.path.join(ROOT, "data", "bc_t30.bin")], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 2 and "no CUDA device" in r.stderr
    assert "nsteps_out (frequency of output)  =    36" in r.stdout and "  End date: 1982/01/03 00:00" in r.stdout      # params.f90:69, date.f90:79-81
    assert not list(tmp_path.glob("*.nc"))


REF_BC = "/root/reference/data/bc/t30"


@pytest.mark.skipif(not os.path.isdir(REF_BC), reason="the reference's boundary files are not on this machine (GPU box)")
def test_reference_netcdf4_tree_read_directly(tmp_path):
    """speedy_model_init's other boundary source: a directory of the reference's own NetCDF-4 files, read without an HDF5 library.  Every
    start-up field equals the one from the packed file bit for bit; the anomaly record is complete (420 months, the pack carries 72); the
    flat layout of a run directory (run.sh links the files side by side) works like the clim/ + anom/ tree"""
    import numpy as np
    pkg = load_pkg()
    pack = os.path.join(ROOT, "data", "bc_t30.bin")
    for name in ("phi0", "fmask", "alb0", "stl12", "snowd12", "soilw12", "sst12", "sice12", "fmask_l", "bmask_l", "fmask_s", "rhcapl", "cdland", "cdsea", "cdice", "solar"):
        a, b = pkg.host_boundary(REF_BC, name), pkg.host_boundary(pack, name)
        assert a.size == b.size and np.array_equal(a, b), name
    a, b = pkg.host_boundary(REF_BC, "ssta"), pkg.host_boundary(pack, "ssta")
    assert a.size == 420 * 96 * 48 and b.size == 72 * 96 * 48 and np.array_equal(a[:b.size], b)
    flat = tmp_path / "rundir"
    flat.mkdir()
    for sub in ("clim", "anom"):
        for f in os.listdir(os.path.join(REF_BC, sub)):
            os.symlink(os.path.join(REF_BC, sub, f), flat / f)
    assert np.array_equal(pkg.host_boundary(flat, "sst12"), pkg.host_boundary(pack, "sst12"))
    # errors are loud: a file missing from the directory, a file that is not HDF5
    os.remove(flat / "soil.nc")
    with pytest.raises(pkg.SpeedyError, match="soil.nc not found"):
        pkg.host_boundary(flat, "phi0")
    (flat / "soil.nc").write_bytes(b"CDF\x01" + b"\0" * 4096)
    with pytest.raises(pkg.SpeedyError, match="not a NetCDF-4 / HDF5 file"):
        pkg.host_boundary(flat, "phi0")
