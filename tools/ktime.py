"""Quick per-kernel timing of the main-loop body (CUDA events, warm L2, plain launches) and the
graph-replayed step time.  usage: python tools/ktime.py [members] [trunc] [sppt]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import _load_pkg
pkg = _load_pkg()
members = int(sys.argv[1]) if len(sys.argv) > 1 else 1
trunc = int(sys.argv[2]) if len(sys.argv) > 2 else 30
bc = pkg.BC_T30
if trunc == 47:
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_t47_boundary
    bc = make_t47_boundary.ensure()
sppt = 1 if len(sys.argv) > 3 and sys.argv[3] == "sppt" else 0
c = pkg.Speedy(trunc=trunc, nmembers=members, sppt_on=sppt, seed=7)
c.model_init(bc)
for _ in range(3):
    c.run_steps(36)
c.synchronize()
t0 = time.perf_counter()
days = 20
for _ in range(days):
    c.run_steps(36)
c.synchronize()
dt = time.perf_counter() - t0
c.trace(True)
c.run_steps(36 * 5)
tr = c.trace_read()
c.trace(False)
print("in-graph timeline (us/step):", {k: round(v, 2) for k, v in tr["us"].items()}, "gaps:", {k: round(v, 2) for k, v in tr["gap_before_us"].items()}, "steps", tr["steps"])
kt = c.time_kernels(36, False) if not sppt else {}
print("members %d T%d: %.2f us/step (graph, wall)  %.1f member-days/s | kernels warm (us): %s | sum %.1f" % (
    members, trunc, 1e6 * dt / (days * 36), days * members / dt, {k: round(1e3 * v, 2) for k, v in kt.items()}, 1e3 * sum(kt.values())))
