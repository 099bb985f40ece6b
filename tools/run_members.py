"""run a few simulated days with M members (profiling target): python tools/run_members.py [members] [days] [sppt]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import _load_pkg
pkg = _load_pkg()
m = int(sys.argv[1]) if len(sys.argv) > 1 else 8
days = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sppt = int(sys.argv[3]) if len(sys.argv) > 3 else 0
c = pkg.Speedy(trunc=30, nmembers=m, sppt_on=sppt, seed=1)
c.set_graphs(False)
c.model_init(pkg.BC_T30)
assert c.run_steps(36 * days) == 0
print("ok", c.launch_count)
