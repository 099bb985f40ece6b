"""The CUDA path against the committed known-answer vectors (tests/golden/, generated from the CPU oracle by make_golden.py): no oracle
library involved at test time.  Tolerances as everywhere: transforms 1e-12, prognostic spectral coefficients after 48 h 1e-10 (north_star)."""
import os

import numpy as np
import pytest
from conftest import ROOT, rel_rms

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(ROOT, "tests", "golden", "t30_golden.npz"))
BC = os.path.join(ROOT, "data", "bc_t30.bin")


def test_transforms_match_golden(ctx):
    assert rel_rms(ctx.spec_to_grid(G["spec"], G["kcos"]), G["grid"]) < 1e-12
    assert rel_rms(ctx.grid_to_spec(G["gin"]), G["back"]) < 1e-12


@pytest.mark.parametrize("members", [1, 8])
def test_48h_run_matches_golden(pkg, members):
    """BASELINE configs[0] from rest; with 8 members (no SPPT: identical members) through the ensemble-step kernels"""
    c = pkg.Speedy(trunc=30, nmembers=members)
    c.model_init(BC)
    assert c.run_steps(72) == 0
    for n in ("vor", "div", "t", "tr", "ps"):
        a = c.get_field(n, all_members=True)
        for e in (0, members - 1):
            assert rel_rms(a[e][0], G[n]) < 1e-10, (n, e)
    rc, diag = c.check_diagnostics(2)
    assert rc == 0 and np.allclose(diag, G["diag"], rtol=1e-8)
    date, step = c.model_date()
    assert tuple(G["date"]) == date + (step,)
    out = c.output_fields()
    stats = np.array([[out[n].astype(np.float64).mean(), out[n].astype(np.float64).std(), out[n].min(), out[n].max()] for n in ("u", "v", "t", "q", "phi", "ps")])
    assert np.allclose(stats, G["out_stats"], rtol=1e-5, atol=1e-6)
    c.close()
