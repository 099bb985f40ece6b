"""Micro-benchmark for BASELINE metric (2): Legendre/transform GB/s vs the HBM roofline.
Times K1 (spec_to_grid) and K2 (grid_to_spec) with device-resident inputs at the batch
sizes of SURVEY.md §8(d), CUDA events on the library's stream, L2 flushed between reps."""
import json
import os
import sys
import ctypes

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import _load_pkg  # noqa: E402


def main():
    pkg = _load_pkg()
    trunc = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    c = pkg.Speedy(trunc=trunc)
    L = c.L
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    stream = torch.cuda.ExternalStream(c.stream)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    nact = sum(2 * min(c.mx, c.trunc + 2 - n) for n in range(c.nx))          # active reals (nsh2 sum)
    bytes_inv = 8 * nact + 8 * c.ix * c.il
    bytes_dir = 8 * c.ix * c.il + 8 * (nact - 2)
    res = []
    for inverse, batches in ((True, [1, 8, 91, 728, 5824]), (False, [1, 8, 73, 584, 4672])):
        for nb in batches:
            g = torch.Generator(device="cuda").manual_seed(1234)
            spec = torch.rand((nb, c.nx, c.mx, 2), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
            grid = torch.rand((nb, c.il, c.ix), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
            torch.cuda.synchronize()
            times = []
            for rep in range(8):
                flush.zero_()
                torch.cuda.synchronize()
                with torch.cuda.stream(stream):
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    if inverse:
                        rc = L.speedy_spec_to_grid_dev(c.h, ctypes.c_void_p(spec.data_ptr()), nb, None, ctypes.c_void_p(grid.data_ptr()))
                    else:
                        rc = L.speedy_grid_to_spec_dev(c.h, ctypes.c_void_p(grid.data_ptr()), nb, ctypes.c_void_p(spec.data_ptr()))
                    assert rc == 0, L.speedy_last_error()
                    e1.record(stream)
                e1.synchronize()
                if rep >= 3:
                    times.append(e0.elapsed_time(e1) * 1e-3)
            t = float(np.median(times))
            by = (bytes_inv if inverse else bytes_dir) * nb
            res.append({"kernel": "spec_to_grid" if inverse else "grid_to_spec", "batch": nb, "us": t * 1e6,
                        "GBps": by / t / 1e9, "frac_hbm": by / t / 1e9 / hbm, "transforms_per_s": nb / t})
            print(json.dumps(res[-1]))
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    tag = sys.argv[2] if len(sys.argv) > 2 else ""
    json.dump({"trunc": trunc, "hbm_gbs_peak": hbm, "env": {k: v for k, v in os.environ.items() if k.startswith("SPEEDY_")}, "results": res},
              open(os.path.join(out, f"bench_transforms_t{trunc}{tag}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
