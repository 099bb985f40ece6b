"""Global-mean climate of a one-year T30 run (area weights = the Gaussian weights): a physical plausibility check of
the whole path that does not depend on the oracle.  usage: python tools/climate_check.py [days]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import _load_pkg

pkg = _load_pkg()
days = int(sys.argv[1]) if len(sys.argv) > 1 else 365
c = pkg.Speedy(trunc=30)
c.model_init(pkg.BC_T30)
wt = pkg.host_table(30, "wt")                      # iy half-table weights, hemispheres sum to 1
w = np.concatenate([wt, wt[::-1]]) / 2.0
gm = lambda f: float((f.mean(axis=-1) * w).sum())  # area-weighted global mean of an (il, ix) field
res = {}
acc = {k: 0.0 for k in ("olr", "tsr", "ssr", "slr", "precip_mm_day", "t_low", "ps_hpa")}
nacc = 0
out0 = c.output_fields()
res["ps0_hpa"] = gm(out0["ps"].astype(np.float64)) / 100.0
for d in range(days):
    assert c.run_steps(36) == 0
    if d >= days - 360 and d % 5 == 0:             # instantaneous samples every 5 days over the last 360 days
        o = c.output_fields()
        acc["olr"] += gm(c.get_field("olr")); acc["tsr"] += gm(c.get_field("tsr")); acc["ssr"] += gm(c.get_field("ssr")); acc["slr"] += gm(c.get_field("slr"))
        acc["precip_mm_day"] += gm(c.get_field("precnv") + c.get_field("precls")) * 86.4   # g/(m^2 s) -> mm/day
        acc["t_low"] += gm(o["t"][-1].astype(np.float64)); acc["ps_hpa"] += gm(o["ps"].astype(np.float64)) / 100.0
        nacc += 1
res.update({k: v / nacc for k, v in acc.items()})
res["toa_net"] = res["tsr"] - res["olr"]
res["samples"] = nacc
res["date"] = c.model_date()[0]
print(json.dumps(res))
