"""BASELINE configs[4] tolerance study: the same T30/L8 run with fp64 transforms (precision = 0) and with real32
spherical-harmonic transforms (precision = 1; grid-point columns, semi-implicit solve and time stepping stay fp64),
relative RMS difference of the prognostic spectral fields every 6 h over 48 h (rest start, 1982-01-01), for a single
member and for the mean over an 8-member SPPT ensemble (same noise in both precisions).

  python tools/precision_study.py > profiles/<round>_precision_study.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import _load_pkg  # noqa: E402

PROG = ("vor", "div", "t", "tr", "ps")


def rel_rms(a, b):
    return float(np.sqrt(np.mean(np.abs(a - b) ** 2)) / max(np.sqrt(np.mean(np.abs(b) ** 2)), 1e-300))


def curve(pkg, nmembers, sppt):
    runs = [pkg.Speedy(trunc=30, nmembers=nmembers, sppt_on=sppt, seed=5, precision=p) for p in (0, 1)]
    for r in runs:
        r.model_init(pkg.BC_T30)
    out = []
    for h in range(6, 49, 6):
        for r in runs:
            assert r.run_steps(9) == 0           # 9 steps of 40 min = 6 h
        row = {"hours": h}
        for n in PROG:
            a = runs[1].get_field(n, all_members=True)[:, 0]      # time level 1
            b = runs[0].get_field(n, all_members=True)[:, 0]
            row[n] = rel_rms(a, b)
            if nmembers > 1:
                row[n + "_ensemble_mean"] = rel_rms(a.mean(0), b.mean(0))
        out.append(row)
    for r in runs:
        r.close()
    return out


def main():
    pkg = _load_pkg()
    res = {"what": "rel. RMS of prognostic spectral coefficients (time level 1), real32-transform run vs fp64 run, T30/L8 from the rest state",
           "single_member": curve(pkg, 1, 0), "sppt_8_members": curve(pkg, 8, 1)}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
