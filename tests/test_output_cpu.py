"""Output-file writer (reference: input_output.f90:95-217) — host-only part, no GPU.

The reference writes one NetCDF classic file per output time through the NetCDF library; the library writes the same
format itself.  The file is read back with an independent reader (scipy.io.netcdf_file) and its structure compared
with what the reference defines: dimension and variable names, definition order, attributes, coordinate values."""
import numpy as np
import pytest
from scipy.io import netcdf_file

from conftest import load_pkg


def fields(kx, il, ix, seed=0):
    rng = np.random.default_rng(seed)
    f3 = [rng.standard_normal((kx, il, ix)).astype(np.float32) for _ in range(5)]
    return f3, rng.standard_normal((il, ix)).astype(np.float32)


@pytest.mark.parametrize("trunc,ix,il", [(30, 96, 48), (47, 144, 72)])
def test_output_file_structure(tmp_path, trunc, ix, il):
    S = load_pkg()
    f3, ps = fields(8, il, ix)
    path = tmp_path / "198201030000.nc"
    S.write_output_file(path, *f3, ps, trunc=trunc, nsteps=36, start=(1982, 1, 1, 0, 0), timestep=72)
    raw = path.read_bytes()
    assert raw[:4] == b"CDF\x01"                     # classic format, as nf90_create(nf90_clobber) makes
    nc = netcdf_file(str(path), "r", mmap=False)
    # input_output.f90:136-170: definition order and shapes
    assert list(nc.dimensions.items()) == [("time", None), ("lon", ix), ("lat", il), ("lev", 8)]
    assert list(nc.variables) == ["time", "lon", "lat", "lev", "u", "v", "t", "q", "phi", "ps"]
    for name, f in zip(("u", "v", "t", "q", "phi"), f3):
        v = nc.variables[name]
        assert v.dimensions == ("time", "lev", "lat", "lon") and v.data.dtype == np.dtype(">f4")
        assert np.array_equal(v[0], f)               # values pass through bit for bit
    assert nc.variables["ps"].dimensions == ("time", "lat", "lon")
    assert np.array_equal(nc.variables["ps"][0], ps)
    att = {n: (nc.variables[n].long_name.decode(), getattr(nc.variables[n], "units", b"").decode()) for n in nc.variables if n != "time"}
    assert att == {"lon": ("longitude", ""), "lat": ("latitude", ""), "lev": ("atmosphere_sigma_coordinate", ""),
                   "u": ("eastward_wind", "m/s"), "v": ("northward_wind", "m/s"), "t": ("air_temperature", "K"),
                   "q": ("specific_humidity", "1"), "phi": ("geopotential_height", "m"), "ps": ("surface_air_pressure", "Pa")}
    assert nc.variables["time"].units == b"hours since 1982-01-01 00:00:0.0"
    # input_output.f90:178-181
    assert nc.variables["time"][:].tolist() == [48.0]
    assert np.array_equal(nc.variables["lon"][:], (np.float32(360.0 / ix) * np.arange(ix, dtype=np.float32)))
    radang = S.host_table(trunc, "radang")
    lat = (radang * 90.0 / float(np.arcsin(np.float32(1.0)))).astype(np.float32)
    assert np.array_equal(nc.variables["lat"][:], lat)
    assert lat[0] < 0 and np.all(np.diff(lat) > 0)   # south to north (geometry.f90:66)
    assert np.array_equal(nc.variables["lev"][:], S.host_table(trunc, "fsg").astype(np.float32))
    nc.close()


def test_output_file_time_axis(tmp_path):
    # timestep*24.0/real(nsteps): real32 arithmetic, any step / steps-per-day / start date
    S = load_pkg()
    f3, ps = fields(8, 48, 96, 1)
    for nsteps, step in ((36, 1), (36, 37), (72, 5)):
        p = tmp_path / f"x{nsteps}_{step}.nc"
        S.write_output_file(p, *f3, ps, nsteps=nsteps, start=(1999, 12, 31, 6, 30), timestep=step)
        nc = netcdf_file(str(p), "r", mmap=False)
        assert nc.variables["time"][0] == np.float32(step) * np.float32(24.0) / np.float32(nsteps)
        assert nc.variables["time"].units == b"hours since 1999-12-31 06:30:0.0"
        nc.close()


def test_output_file_errors(tmp_path):
    S = load_pkg()
    f3, ps = fields(8, 48, 96, 2)
    with pytest.raises(S.SpeedyError):
        S.write_output_file(tmp_path / "no_such_dir" / "a.nc", *f3, ps)


def test_reference_run_comparison_harness(tmp_path):
    """tools/compare_reference_run.py end to end on files in the reference's output format (written by the library's host-side writer
    from the oracle's fields): the harness that pins the oracle to a gfortran run once one exists must read the files, step the oracle
    to each file's time and report a zero difference here"""
    import subprocess, sys, json, os
    from conftest import ROOT
    out = tmp_path / "cmp.json"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "compare_reference_run.py"), "--self-test", str(tmp_path / "run"), "--json", str(out)])
    res = json.load(open(out))
    assert len(res["files"]) == 4 and [r["step"] for r in res["files"]] == [0, 1, 2, 3]
    assert res["worst"]["oracle"] == 0.0
