"""SPPT pattern with and without the fold into the spectral step: the field each step CONSUMES must be the same"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import _load_pkg
pkg = _load_pkg()
m = int(sys.argv[1]) if len(sys.argv) > 1 else 1
res = {}
for fold in (0, 1):
    c = pkg.Speedy(trunc=30, nmembers=m, sppt_on=1, seed=11)
    c.set_option("sppt_fold", fold)
    c.set_option("graphs", 0)
    c.model_init(pkg.BC_T30)
    seq = [c.get_field("sppt_spec", all_members=True).copy()]
    for s in range(4):
        assert c.run_steps(1) == 0
        seq.append(c.get_field("sppt_spec", all_members=True).copy())
    res[fold] = (seq, c.get_field("t", all_members=True).copy())
    c.close()
a, b = res[0][0], res[1][0]
for i in range(len(a)):
    print("after init + %d steps: unfolded[i] == folded[i]: %s; unfolded[i+1] == folded[i]: %s" % (
        i, np.array_equal(a[i], b[i]), np.array_equal(a[i + 1], b[i]) if i + 1 < len(a) else "-"))
for i in range(len(a)):
    print(i, [bool(np.array_equal(a[j], b[i])) for j in range(len(a))], [float("%.3e" % np.abs(a[j] - b[i]).max()) for j in range(len(a))])
print(a[1][0, 0, 1, :3], b[0][0, 0, 1, :3], b[1][0, 0, 1, :3])
print("t equal:", np.array_equal(res[0][1], res[1][1]), np.abs(res[0][1] - res[1][1]).max())
