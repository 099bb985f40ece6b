"""phase stamps of the quad grid->spec kernel on a large batch (debug aid): python tools/qstamp.py [nb]"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import _load_pkg
pkg = _load_pkg()
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4672
c = pkg.Speedy(trunc=30)
c.set_option("k2_quad", 1)
grid = torch.rand((nb, c.il, c.ix), dtype=torch.float64, device="cuda")
spec = torch.empty((nb, c.nx, c.mx, 2), dtype=torch.float64, device="cuda")
for _ in range(3):
    c.L.speedy_grid_to_spec_dev(c.h, ctypes.c_void_p(grid.data_ptr()), nb, ctypes.c_void_p(spec.data_ptr()))
c.synchronize()
c.trace(True)
c.L.speedy_grid_to_spec_dev(c.h, ctypes.c_void_p(grid.data_ptr()), nb, ctypes.c_void_p(spec.data_ptr()))
c.synchronize()
buf = (ctypes.c_ulonglong * 64)()
import ctypes as C
# raw read of the trace buffer through trace_read's stderr dump
os.environ["SPEEDY_TRACE_STAMPS"] = "1"
print(c.trace_read())
