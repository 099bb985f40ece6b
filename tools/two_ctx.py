"""aggregate throughput of several contexts of M members each, advanced concurrently on their own streams (each context replays its
own CUDA graph): python tools/two_ctx.py [members per context] [contexts] [sppt]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import _load_pkg
pkg = _load_pkg()
m = int(sys.argv[1]) if len(sys.argv) > 1 else 4
nctx = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sppt = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cs = [pkg.Speedy(trunc=30, nmembers=m, sppt_on=sppt, seed=1, member_offset=i * m) for i in range(nctx)]
for c in cs:
    c.model_init(pkg.BC_T30)
for _ in range(3):
    for c in cs:
        c.enqueue_steps(36)
for c in cs:
    assert c.finish() == 0
days = 20
t0 = time.perf_counter()
for _ in range(days):
    for c in cs:
        c.enqueue_steps(36)
for c in cs:
    assert c.finish() == 0
dt = time.perf_counter() - t0
print("%d contexts x %d members%s: %.2f us per step of all, %.1f member-days/s" % (nctx, m, " sppt" if sppt else "", 1e6 * dt / (days * 36), days * m * nctx / dt))
