"""probe: spec->grid through the quad kernel on random fields vs the streaming kernel (debug aid)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import _load_pkg
pkg = _load_pkg()
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 584
c = pkg.Speedy(trunc=30)
rng = np.random.default_rng(99)
s = rng.uniform(-1, 1, size=(nb, c.nx, c.mx)) + 1j * rng.uniform(-1, 1, size=(nb, c.nx, c.mx))
n = np.arange(c.nx)[:, None]; m = np.arange(c.mx)[None, :]
s = s * ((m + n) <= c.trunc + 1)
kcos = np.where(np.arange(nb) % 3 == 0, 1, 2).astype(np.int32)
c.set_option("k1_quad", 0)
base = c.spec_to_grid(s, kcos)
c.spec_to_grid(np.zeros_like(s), kcos)          # the library's scratch buffer now holds zeros: `got` must be written afresh
c.set_option("k1_quad", 1)
got = c.spec_to_grid(s, kcos)
err = np.sqrt(np.mean(np.abs(got - base) ** 2) / np.mean(np.abs(base) ** 2))
print("k1_quad nb", nb, "rel rms vs streaming kernel:", err)
if err > 1e-12:
    d = np.abs(got - base).reshape(nb, -1).max(axis=1)
    print("worst fields:", np.argsort(d)[-8:], d[np.argsort(d)[-8:]])
    f = int(np.argmax(d)); e = np.abs(got[f] - base[f])
    print("field", f, "rows with error:", np.nonzero(e.max(axis=1) > 1e-9)[0][:48], "cols:", np.nonzero(e.max(axis=0) > 1e-9)[0][:24])
