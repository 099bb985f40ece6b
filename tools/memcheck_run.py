"""Short run of each configuration for compute-sanitizer (memcheck / racecheck): plain launches, 4 steps.
usage: compute-sanitizer --tool memcheck python tools/memcheck_run.py [t30|t30x4|t30x8|t47]   (t30x8: the quad transform kernels)"""
import sys, os
sys.path.insert(0, os.getcwd())
from __graft_entry__ import _load_pkg
pkg = _load_pkg()
sel = sys.argv[1] if len(sys.argv) > 1 else "all"
for name, trunc, members in (("t30", 30, 1), ("t30x4", 30, 4), ("t30x8", 30, 8), ("t47", 47, 1)):
    if sel not in ("all", name):
        continue
    bc = pkg.BC_T30
    if trunc == 47:
        sys.path.insert(0, "tools"); import make_t47_boundary; bc = make_t47_boundary.ensure()
    c = pkg.Speedy(trunc=trunc, nmembers=members, sppt_on=1 if members > 1 else 0)
    c.model_init(bc)
    c.set_graphs(False)
    assert c.run_steps(4) == 0
    c.output_fields()
    c.close()
    print("ok", trunc, members, flush=True)
