"""Generates tests/golden/t30_golden.npz: known-answer vectors for the T30 hot path.

The reference ships no golden outputs and cannot be built here, so these vectors come from the repo's own CPU
oracle (oracle/, the line-by-line restatement of speedy.f90) — they pin the ORACLE against drifting between rounds and give the GPU
tests a fixed target that does not need the oracle library at all; they do NOT pin anything to a gfortran build (DESIGN.md §5,
tools/compare_reference_run.py is the harness for that).  Regenerate only on purpose:

    python tests/golden/make_golden.py        # CPU only, ~10 s

Content: (a) spec_to_grid / grid_to_spec of seeded fields (seed 1234, uniform(-1,1) on the triangle);
(b) the model after 48 h from rest on the reference's T30 boundary files (BASELINE configs[0]): prognostic spectral fields at time
level 1, check_diagnostics, the date, and the float32 output() fields' global statistics."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import Oracle, random_spec  # noqa: E402


def main():
    o = Oracle("t30")
    rng = np.random.default_rng(1234)
    spec = random_spec(rng, (4,), o.nx, o.mx, o.trunc)
    kcos = np.array([1, 2, 1, 2], np.int32)
    grid = o.spec_to_grid(spec, kcos)
    gin = rng.uniform(-1, 1, size=(3, o.il, o.ix))
    back = o.grid_to_spec(gin)
    o.model_init(os.path.join(ROOT, "data", "bc_t30.bin"))
    assert o.run(72) == 0
    st = o.state()
    rc, diag = o.check_diagnostics(2)
    date, step = o.date()
    out = o.output_fields()
    stats = np.array([[out[n].astype(np.float64).mean(), out[n].astype(np.float64).std(), out[n].min(), out[n].max()] for n in ("u", "v", "t", "q", "phi", "ps")])
    np.savez_compressed(os.path.join(HERE, "t30_golden.npz"), spec=spec, kcos=kcos, grid=grid, gin=gin, back=back,
                        vor=st["vor"][0], div=st["div"][0], t=st["t"][0], tr=st["tr"][0], ps=st["ps"][0],
                        diag=diag, date=np.array(date + (step,)), out_stats=stats)
    print("wrote t30_golden.npz", os.path.getsize(os.path.join(HERE, "t30_golden.npz")), "bytes; date", date, "step", step)


if __name__ == "__main__":
    main()
