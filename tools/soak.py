"""long run of an SPPT ensemble through the ensemble step (per-member hand-off, L2 discards, shared transient buffer, folded SPPT update):
python tools/soak.py [members] [days]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import _load_pkg
pkg = _load_pkg()
m = int(sys.argv[1]) if len(sys.argv) > 1 else 8
days = int(sys.argv[2]) if len(sys.argv) > 2 else 400
c = pkg.Speedy(trunc=30, nmembers=m, sppt_on=1, seed=3)
c.model_init(pkg.BC_T30)
t0 = time.perf_counter()
for d0 in range(0, days, 50):
    n = min(50, days - d0)
    c.enqueue_steps(36 * n)
    assert c.finish() == 0, "range guard tripped"
dt = time.perf_counter() - t0
t = c.get_field("t", all_members=True)
ps = c.get_field("ps", all_members=True)
assert np.isfinite(t).all() and np.isfinite(ps).all()
spread = float(np.std(t[:, 0, -1, 0, 0].real))
print("soak ok: %d members x %d days in %.2f s (%.1f member-days/s), date %s, spread of the lowest-level mean T coefficient %.3e" % (
    m, days, dt, m * days / dt, c.date() if hasattr(c, "date") else "?", spread))
