"""Throughput of the secondary BASELINE configs on one GPU (graph-replayed main loop, wall clock):
configs[2] 8 SPPT members per GPU at T30, configs[3] T47 single member (synthetic boundaries, 48 h window
repeated from the rest state: at the reference's time step T47 leaves the diagnostics bounds on day 33)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from __graft_entry__ import _load_pkg
import make_t47_boundary
pkg = _load_pkg()
res = {}
def timed(c, days):
    c.run_steps(36 * 2); c.synchronize()
    t0 = time.perf_counter()
    for _ in range(days):
        assert c.run_steps(36) == 0
    c.synchronize()
    return time.perf_counter() - t0
c = pkg.Speedy(trunc=30, nmembers=8, sppt_on=1, seed=1); c.model_init(pkg.BC_T30)
dt = timed(c, 20); res["t30_8members_sppt"] = {"member_days_per_s": 8 * 20 / dt, "us_per_step": 1e6 * dt / (20 * 36), "members": 8}; c.close()
c = pkg.Speedy(trunc=30, nmembers=8); c.model_init(pkg.BC_T30)
dt = timed(c, 20); res["t30_8members"] = {"member_days_per_s": 8 * 20 / dt, "us_per_step": 1e6 * dt / (20 * 36), "members": 8}; c.close()
c = pkg.Speedy(trunc=47); c.model_init(make_t47_boundary.ensure())
dt = timed(c, 20); res["t47_single"] = {"sim_days_per_s": 20 / dt, "us_per_step": 1e6 * dt / (20 * 36)}; c.close()
print(json.dumps(res))
