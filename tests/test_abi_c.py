"""The drop-in boundary from compiled C, as the Fortran bind(C) shim (fortran/speedy_b200_c.f90) would use it.  No Fortran
compiler exists in this image; what can be pinned is (i) the layout of the one derived type that crosses the boundary and
(ii) the call sequence by reference from a compiled host language."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
from conftest import ROOT, load_pkg


def _build(tmp_path):
    exe = tmp_path / "abi_drive"
    libdir = os.path.join(ROOT, "speedy.f90_b200")
    load_pkg().lib()                                         # the library must exist (built by __graft_entry__.build())
    subprocess.check_call(["gcc", "-O1", "-Wall", "-o", str(exe), os.path.join(ROOT, "tests", "helpers", "abi_drive.c"),
                           "-L" + libdir, "-lspeedy_b200", "-Wl,-rpath," + libdir])
    return str(exe)


def test_speedy_cfg_layout_matches_the_bind_c_type(tmp_path):
    """`type, bind(C) :: speedy_cfg` = its components in sequence with C's natural alignment: 6 x c_int, c_long_long, 3 x c_int"""
    out = subprocess.check_output([_build(tmp_path), "layout"], text=True)
    got = dict(l.split() for l in out.strip().splitlines())

    class F(ctypes.Structure):                               # the Fortran declaration, component by component
        _fields_ = [(n, ctypes.c_int) for n in ("trunc", "kx", "ntr", "nmembers", "device", "sppt_on")] + [("seed", ctypes.c_longlong)] + \
                   [(n, ctypes.c_int) for n in ("member_offset", "nsteps", "precision")]
    assert int(got["sizeof"]) == ctypes.sizeof(F) == ctypes.sizeof(load_pkg().Cfg)
    for name, _ in F._fields_:
        assert int(got[name]) == getattr(F, name).offset == getattr(load_pkg().Cfg, name).offset, name
    # and the Fortran source declares exactly these components in this order
    src = open(os.path.join(ROOT, "fortran", "speedy_b200_c.f90")).read()
    blk = src[src.index("type, bind(C) :: speedy_cfg"):src.index("end type")]
    decl = [w.strip() for line in blk.splitlines()[1:] if "::" in line for w in line.split("::")[1].split(",")]
    assert decl == [n for n, _ in F._fields_]


def test_every_entry_point_has_a_fortran_interface():
    hdr = open(os.path.join(ROOT, "include", "speedy_b200.h")).read()
    import re
    names = sorted(set(re.findall(r"\b(speedy_[a-z0-9_]+)\s*\(", hdr)) - {"speedy_b200"})
    src = open(os.path.join(ROOT, "fortran", "speedy_b200_c.f90")).read()
    missing = [n for n in names if f'name="{n}"' not in src]
    assert not missing, missing


@pytest.mark.gpu
def test_main_loop_driven_from_compiled_c(tmp_path, pkg):
    """create -> model_init -> run_steps_host -> destroy from C == the same calls through the Python layer"""
    bc = os.path.join(ROOT, "data", "bc_t30.bin")
    out = subprocess.check_output([_build(tmp_path), "run", bc, "40"], text=True)
    lines = dict((l.split()[0], l.split()[1:]) for l in out.strip().splitlines())
    c = pkg.Speedy(trunc=30)
    c.model_init(bc)
    st = np.concatenate([c.get_field(n).view(np.float64).ravel() for n in ("vor", "div", "t", "tr", "ps")])
    o = np.empty((5 * c.kx + 1) * c.il * c.ix, np.float32)
    assert c.run_steps_host(st, 40, o) == 0
    (y, m, d, h, mi), step = c.model_date()
    assert lines["date"] == [str(v) for v in (y, m, d, h, mi)] + ["step", str(step)]
    # the C driver sums in index order: the same order here
    s1 = s2 = 0.0
    s1 = float(np.add.reduce(st))      # pairwise in numpy: compare to rounding, the element values are bit-identical runs
    s2 = float(np.add.reduce(st * st))
    assert abs(float(lines["checksum"][0]) - s1) <= 1e-9 * max(1.0, abs(s1))
    assert abs(float(lines["checksum"][1]) - s2) <= 1e-9 * s2
    c.close()
