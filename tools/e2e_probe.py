"""where the host-resident day goes: python tools/e2e_probe.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import _load_pkg
import torch
pkg = _load_pkg()
c = pkg.Speedy(trunc=30)
c.model_init(pkg.BC_T30)
n = c.state_len() if callable(c.state_len) else c.state_len
st = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
st[:] = np.concatenate([c.get_field(k).view(np.float64).ravel() for k in ("vor", "div", "t", "tr", "ps")])
out = torch.empty((5 * c.kx + 1) * c.il * c.ix, dtype=torch.float32).pin_memory().numpy()
def timeit(f, reps=30):
    for _ in range(5): f()
    c.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    c.synchronize()
    return 1e6 * (time.perf_counter() - t0) / reps
print("run_steps(36) + sync          %.1f us" % timeit(lambda: c.run_steps(36)))
print("enqueue_steps(36) back to back %.1f us" % timeit(lambda: c.enqueue_steps(36)))
print("run_steps_host no out         %.1f us" % timeit(lambda: c.run_steps_host(st, 36, None)))
print("run_steps_host with out       %.1f us" % timeit(lambda: c.run_steps_host(st, 36, out)))
print("run_steps_host 0 steps, out   %.1f us" % timeit(lambda: c.run_steps_host(st, 0, out)))
print("run_steps_host 0 steps no out %.1f us" % timeit(lambda: c.run_steps_host(st, 0, None)))
print("output_fields                 %.1f us" % timeit(lambda: c.output_fields()))
