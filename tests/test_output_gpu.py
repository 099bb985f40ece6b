"""SURVEY.md §8(f) N2 — the step after the path: output files (input_output.f90:95-217) and restart files."""
import os

import numpy as np
import pytest
from scipy.io import netcdf_file

from conftest import ROOT

pytestmark = pytest.mark.gpu
BC = os.path.join(ROOT, "data", "bc_t30.bin")
PROG = ("vor", "div", "t", "tr", "ps")


def test_output_file_of_a_run(pkg, oracle, tmp_path):
    """48 h from rest, then output(): file named after the model date, time axis from the step counter, fields equal to
    speedy_output_fields bit for bit and to the oracle's output() within float32 rounding of a 1e-10 state difference"""
    c = pkg.Speedy(trunc=30)
    c.model_init(BC)
    assert c.run_steps(72) == 0
    path = c.write_output(tmp_path)
    assert os.path.basename(path) == "198201030000.nc"          # README.md:12 end date of the shipped run
    nc = netcdf_file(path, "r", mmap=False)
    assert nc.variables["time"][:].tolist() == [48.0]
    assert nc.variables["time"].units == b"hours since 1982-01-01 00:00:0.0"
    out = c.output_fields()
    for name in ("u", "v", "t", "q", "phi", "ps"):
        assert np.array_equal(nc.variables[name][0], out[name]), name
    o = oracle
    o.model_init(BC)
    assert o.run(72) == 0
    ref = o.output_fields()
    for name in ("u", "v", "t", "q", "phi", "ps"):
        a, b = np.asarray(nc.variables[name][0], np.float64), np.asarray(ref[name], np.float64)
        assert np.sqrt(np.mean((a - b) ** 2)) <= 2e-7 * np.sqrt(np.mean(b ** 2)), name
    nc.close()
    c.close()


@pytest.mark.parametrize("sppt", [0, 1])
def test_restart_continues_bit_for_bit(pkg, tmp_path, sppt):
    """day 1 + restart file + day 2 in a fresh context == two uninterrupted days, for every state field and the calendar
    (with SPPT: the AR(1) pattern and the draw counter travel with the file)"""
    kw = dict(trunc=30, nmembers=2, sppt_on=sppt, seed=5)
    a = pkg.Speedy(**kw)
    a.model_init(BC)
    assert a.run_steps(40) == 0                                  # not a multiple of 3: the short-wave cadence carries over
    rst = tmp_path / "day1.rst"
    a.save_restart(rst)
    assert a.run_steps(32) == 0
    b = pkg.Speedy(**kw)
    b.model_init(BC)
    b.load_restart(rst)
    assert b.model_date() == ((1982, 1, 2, 2, 40), 41)
    assert b.run_steps(32) == 0
    assert a.model_date() == b.model_date()
    for n in PROG + ("phi", "stl_am", "sst_om", "tt_rsw", "tau2", "ssrd", "ts", "qcloud"):
        assert np.array_equal(a.get_field(n, all_members=True), b.get_field(n, all_members=True)), n
    if sppt:
        x = a.get_field("vor", all_members=True)
        assert not np.array_equal(x[0], x[1])                    # the members did diverge
    # a file from another configuration is refused
    c = pkg.Speedy(trunc=30, nmembers=1)
    c.model_init(BC)
    with pytest.raises(pkg.SpeedyError):
        c.load_restart(rst)
    others = [c]
    if sppt:
        # ... and so is another SPPT stream (seed / member_offset) or noise source: the continuation would silently differ
        for kw2, draw in ((dict(kw, seed=6), True), (dict(kw, member_offset=2), True), (kw, False)):
            d = pkg.Speedy(**kw2)
            d.model_init(BC)
            d.set_sppt_draw(draw)
            with pytest.raises(pkg.SpeedyError):
                d.load_restart(rst)
            others.append(d)
    for s in [a, b] + others:
        s.close()


def test_real32_transform_mode_initialises_at_t47(pkg):
    """precision = 1 at T47: the first transform of model_init is the real32 grid->spec kernel, whose shared-memory opt-in
    (50.7 KB > the 48 KB default) must be in place whichever launcher runs first"""
    from conftest import bc_t47
    c = pkg.Speedy(trunc=47, precision=1)
    c.model_init(bc_t47())
    assert c.run_steps(4) == 0
    assert np.isfinite(c.get_field("vor")).all()
    c.close()
