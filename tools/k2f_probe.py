"""probe: grid->spec through the alternative batch kernel on random fields vs the default kernel (debug aid)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import _load_pkg
pkg = _load_pkg()
opt = sys.argv[1] if len(sys.argv) > 1 else "k2_field"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 584
c = pkg.Speedy(trunc=30)
rng = np.random.default_rng(99)
g = rng.uniform(-1, 1, size=(nb, c.il, c.ix))
c.set_option("k2_quad", 0)
base = c.grid_to_spec(g)
c.set_option(opt, 1)
got = c.grid_to_spec(g)
err = np.sqrt(np.mean(np.abs(got - base) ** 2) / np.mean(np.abs(base) ** 2))
print(opt, "nb", nb, "rel rms vs default kernel:", err)
