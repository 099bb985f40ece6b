"""a few steps with the real32 transforms (profiling target): python tools/run_f32.py [members] [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import _load_pkg
pkg = _load_pkg()
m = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
c = pkg.Speedy(trunc=30, nmembers=m, precision=1)
c.set_graphs(False)
c.model_init(pkg.BC_T30)
assert c.run_steps(n) == 0
print("ok", c.launch_count)
