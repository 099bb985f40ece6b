"""The kernels a step runs depend on the member count (single-member latency variants, batch variants of the streaming transforms,
the quad transforms with the per-member hand-off, the L2 discards and the shared transient buffer from three fields per SM on): every
regime against the same oracle trajectory, identical members being identical bit for bit."""
import os
import numpy as np
import pytest
from conftest import ROOT, rel_rms, bc_t47

pytestmark = pytest.mark.gpu
BC = os.path.join(ROOT, "data", "bc_t30.bin")
PROG = ("vor", "div", "t", "tr", "ps")


@pytest.fixture(scope="module")
def ref24(oracle):
    oracle.model_init(BC)
    assert oracle.run(24) == 0
    return {**oracle.state(), "iptop": oracle.ifield("iptop"), "icltop": oracle.ifield("icltop")}


@pytest.mark.parametrize("members", [2, 3, 5, 6, 7, 12, 20])
def test_24_steps_at_every_member_count(pkg, ref24, members):
    c = pkg.Speedy(trunc=30, nmembers=members)
    c.model_init(BC)
    assert c.run_steps(24) == 0
    for n in PROG:
        f = c.get_field(n, all_members=True)
        assert np.array_equal(f[0], f[members - 1]), n
        e = rel_rms(f[0], ref24[n])
        assert e < 1e-10, (members, n, e)
    for n in ("iptop", "icltop"):
        assert np.array_equal(c.get_field(n), ref24[n]), n
    c.close()


def test_t47_two_members_follow_the_single_member_run(pkg):
    """T47 has no quad kernels: two members take the batch variants of the streaming transforms and of the column kernel (with the L2
    discards of the grid fields), one member the latency variants"""
    out = []
    t47_bc = bc_t47()
    for members in (1, 2):
        c = pkg.Speedy(trunc=47, nmembers=members, nsteps=72)
        c.model_init(t47_bc)
        assert c.run_steps(18) == 0
        out.append({n: c.get_field(n, all_members=True) for n in PROG})
        c.close()
    for n in PROG:
        assert np.array_equal(out[1][n][0], out[1][n][1]), n
        e = rel_rms(out[1][n][0], out[0][n][0])
        assert e < 1e-11, (n, e)
