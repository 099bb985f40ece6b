"""phase stamps of the quad spec->grid kernel on a large batch (debug aid): python tools/qstamp1.py [nb]"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import _load_pkg
pkg = _load_pkg()
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 5824
c = pkg.Speedy(trunc=30)
grid = torch.empty((nb, c.il, c.ix), dtype=torch.float64, device="cuda")
spec = torch.rand((nb, c.nx, c.mx, 2), dtype=torch.float64, device="cuda")
for _ in range(3):
    c.L.speedy_spec_to_grid_dev(c.h, ctypes.c_void_p(spec.data_ptr()), nb, None, ctypes.c_void_p(grid.data_ptr()))
c.synchronize()
c.trace(True)
c.L.speedy_spec_to_grid_dev(c.h, ctypes.c_void_p(spec.data_ptr()), nb, None, ctypes.c_void_p(grid.data_ptr()))
c.synchronize()
os.environ["SPEEDY_TRACE_STAMPS"] = "1"
c.trace_read()
