"""The caller of the path on the GPU: `program speedy` (speedy.f90:1-54) as the speedy_b200 executable and as speedy_main_loop —
BASELINE configs[0]: the reference's namelist, a 2-day run, NetCDF output compared field by field."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
from scipy.io import netcdf_file

from conftest import ROOT
from test_program_cpu import SHIPPED_STYLE

pytestmark = pytest.mark.gpu
BC = os.path.join(ROOT, "data", "bc_t30.bin")
EXE = os.path.join(ROOT, "speedy.f90_b200", "bin", "speedy_b200")
FIELDS = ("u", "v", "t", "q", "phi", "ps")
PROG = ("vor", "div", "t", "tr", "ps")


@pytest.fixture(scope="module", autouse=True)
def _executable():
    """the executable is built in-tree by __graft_entry__.build(); a checkout without built artefacts builds it here (g++, C ABI only)"""
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "speedy.f90_b200"), "bin/speedy_b200"])


def _read(path):
    nc = netcdf_file(str(path), "r", mmap=False)
    out = {n: np.array(nc.variables[n][0]) for n in FIELDS}
    hours = float(nc.variables["time"][0])
    nc.close()
    return out, hours


def _close(a, b, tol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.sqrt(np.mean((a - b) ** 2)) <= tol * max(np.sqrt(np.mean(b ** 2)), 1e-30)


def test_executable_runs_the_namelist(pkg, oracle, tmp_path):
    """namelist: 2 days, output every 36 steps, diagnostics every 90: three files named after the model date, the reference's print-out,
    every file equal to the library's own output() at that step and to the oracle's within float32 rounding"""
    (tmp_path / "namelist.nml").write_text(SHIPPED_STYLE)
    r = subprocess.run([EXE, "--bc", BC], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert sorted(p.name for p in tmp_path.glob("*.nc")) == ["198201010000.nc", "198201020000.nc", "198201030000.nc"]
    # diagnostics.f90:71-73: ' step =',i6,' reke =',(10f8.2) / 13x,' deke =' / 13x,' temp =' at steps 0 (prognostics.f90:120) and 90 > 72: only 0
    steps = [int(m.group(1)) for m in re.finditer(r"^ step =\s*(\d+) reke =((?:\s*-?\d+\.\d\d){8})$", r.stdout, re.M)]
    assert steps == [0]
    assert len(re.findall(r"^ {13} deke =(?:\s*-?\d+\.\d\d){8}$", r.stdout, re.M)) == 1
    temp0 = [float(x) for x in re.search(r"^ {13} temp =(.*)$", r.stdout, re.M).group(1).split()]
    o = oracle
    o.model_init(BC)
    c = pkg.Speedy(trunc=30)
    c.model_init(BC)
    # step 0: the file holds the INITIAL state (prognostics.f90:123-126): first_step (eps = 0) leaves time level 1 and its geopotential as they were
    rc, d0 = c.check_diagnostics(1)
    assert rc == 0 and np.allclose(temp0, d0[2], atol=0.006)
    f0, h0 = _read(tmp_path / "198201010000.nc")
    assert h0 == 0.0
    assert not f0["u"].any() and not f0["v"].any()             # the reference starts from rest (prognostics.f90:34-50)
    want0, ref0 = c.output_fields(), o.output_fields()
    for n in FIELDS:
        assert np.array_equal(f0[n], want0[n]), n
        assert _close(f0[n], ref0[n], 2e-7), n
    # days 1 and 2
    for k, name in ((36, "198201020000.nc"), (72, "198201030000.nc")):
        assert c.run_steps(36) == 0 and o.run(36) == 0
        f, hours = _read(tmp_path / name)
        assert hours == 24.0 * k / 36
        mine, ref = c.output_fields(), o.output_fields()
        for n in FIELDS:
            assert np.array_equal(f[n], mine[n]), (name, n)  # the events of the loop (files, prints) do not touch the trajectory
            assert _close(f[n], ref[n], 2e-6), (name, n)
    c.close()


def test_main_loop_events_do_not_touch_the_trajectory(pkg, tmp_path, capfd):
    """speedy_main_loop with a file after EVERY step (the shipped nsteps_out = 1) and a print every 4 steps == plain run_steps, bit for bit;
    two members, files of member 1"""
    nml = pkg.read_namelist(None)
    nml.nsteps_out, nml.nstdia = 1, 4
    nml.end_datetime[:] = (1982, 1, 1, 8, 0)                   # 12 steps
    a = pkg.Speedy(trunc=30, nmembers=2, sppt_on=1, seed=3)
    a.model_init(BC)
    rc, done = a.main_loop(nml, tmp_path, member=1, verbose=True)
    assert (rc, done) == (0, 12) and a.model_date() == ((1982, 1, 1, 8, 0), 13)
    out = capfd.readouterr().out
    assert [int(x) for x in re.findall(r"^ step =\s*(\d+) reke", out, re.M)] == [0, 4, 8, 12]
    files = sorted(p.name for p in tmp_path.glob("*.nc"))
    assert len(files) == 13 and files[0] == "198201010000.nc" and files[1] == "198201010040.nc" and files[-1] == "198201010800.nc"
    b = pkg.Speedy(trunc=30, nmembers=2, sppt_on=1, seed=3)
    b.model_init(BC)
    assert b.run_steps(12) == 0
    for n in PROG + ("phi", "tau2", "stl_am", "ts"):
        assert np.array_equal(a.get_field(n, all_members=True), b.get_field(n, all_members=True)), n
    last, hours = _read(tmp_path / files[-1])
    assert hours == 8.0
    want = b.output_fields(member=1)
    for n in FIELDS:
        assert np.array_equal(last[n], want[n]), n
    rcd, d = b.check_diagnostics(2)
    temp = [float(x) for x in re.findall(r"^ {13} temp =(.*)$", out, re.M)[-1].split()]
    assert np.allclose(temp, d[2], atol=0.006)
    # a context that is not at the namelist's start date is refused
    with pytest.raises(pkg.SpeedyError):
        a.main_loop(nml, None)
    a.close(); b.close()


def test_executable_reports_a_range_failure(tmp_path):
    """T47 at the reference's 36 steps per day leaves the accepted range on day 33: exit code 1, the reference's message, the failing step's lines"""
    from conftest import bc_t47
    (tmp_path / "namelist.nml").write_text("&params\nnsteps_out = 360\nnstdia = 100000\n/\n&date\nend_datetime%month = 3\n/\n")
    r = subprocess.run([EXE, "--bc", bc_t47(), "--trunc", "47"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1, (r.stdout, r.stderr)
    assert "Model variables out of accepted range" in r.stderr
    steps = [int(x) for x in re.findall(r"^ step =\s*(\d+) reke", r.stdout, re.M)]
    assert len(steps) == 2 and steps[0] == 0 and 36 * 25 < steps[1] < 36 * 45
    # the reference stops inside check_diagnostics of the failing step: no file from that step on (the loop enqueues ahead of the guard)
    hours = sorted(_read(p)[1] for p in tmp_path.glob("*.nc"))
    assert hours[0] == 0.0 and len(hours) >= 3 and all(h * 36 / 24 < steps[1] for h in hours)
    assert hours == [240.0 * k for k in range(len(hours))]


def test_async_output_equals_the_synchronous_writer(pkg, tmp_path):
    """speedy_write_output_async (conversions enqueued, pinned ring, host writer threads) while the device runs on: 40 files, more than the
    ring holds, byte-identical to speedy_write_output of the same steps"""
    a = pkg.Speedy(trunc=30, nmembers=2)
    a.model_init(BC)
    b = pkg.Speedy(trunc=30, nmembers=2)
    b.model_init(BC)
    da, db = tmp_path / "async", tmp_path / "sync"
    da.mkdir(); db.mkdir()
    ymdhm = [1982, 1, 1, 0, 0]
    for k in range(1, 41):
        a.enqueue_steps(1)
        d = (ctypes.c_int * 5)(*ymdhm)
        assert pkg.lib().speedy_host_calendar(d, 1, None, None, None) == 0      # newdate on the host: no device round trip
        ymdhm = list(d)
        a.write_output_async(da, ymdhm, k, member=k % 2)
        assert b.run_steps(1) == 0
        b.write_output(db, member=k % 2)
    assert a.finish() == 0
    a.output_drain()
    names = sorted(p.name for p in db.glob("*.nc"))
    assert len(names) == 40 and names == sorted(p.name for p in da.glob("*.nc"))
    for n in names:
        assert (da / n).read_bytes() == (db / n).read_bytes(), n
    for n in PROG:
        assert np.array_equal(a.get_field(n, all_members=True), b.get_field(n, all_members=True)), n
    with pytest.raises(pkg.SpeedyError):                       # a write that fails is reported by the drain
        a.write_output_async(tmp_path / "no" / "such" / "dir", ymdhm, 41)
        a.output_drain()
    a.close(); b.close()


def test_executable_ensemble_members_into_their_own_directories(pkg, tmp_path):
    """--members 3 --sppt --member -1: every member's files under member<e>/, equal to speedy_write_output of the same members"""
    (tmp_path / "namelist.nml").write_text("&params\nnsteps_out = 18\nnstdia = 36\n/\n&date\nend_datetime%month = 1\nend_datetime%day = 2\n/\n")
    r = subprocess.run([EXE, "--bc", BC, "--members", "3", "--sppt", "--seed", "11", "--member", "-1"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    c = pkg.Speedy(trunc=30, nmembers=3, sppt_on=1, seed=11)
    c.model_init(BC)
    assert c.run_steps(36) == 0
    last = []
    for e in range(3):
        names = sorted(p.name for p in (tmp_path / f"member{e}").glob("*.nc"))
        assert names == ["198201010000.nc", "198201011200.nc", "198201020000.nc"]
        f, hours = _read(tmp_path / f"member{e}" / names[-1])
        want = c.output_fields(member=e)
        for n in FIELDS:
            assert np.array_equal(f[n], want[n]), (e, n)
        last.append(f["t"])
    assert not np.array_equal(last[0], last[1])                 # SPPT: the members differ
    c.close()
