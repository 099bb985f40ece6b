#!/usr/bin/env python
"""Pin the oracle (and the GPU path) to a REAL run of the reference — the harness that closes "parity unpinned" the moment someone
has a Fortran toolchain.

The reference writes one NetCDF classic file `yyyymmddhhmm.nc` per output step (input_output.f90:95-217: float32 u, v, t, q, phi
(lon, lat, lev, time) and ps (lon, lat, time)).  Given a directory of such files from a `gfortran -Ofast` build of
samhatfield/speedy.f90, this tool re-runs the same dates with the CPU oracle (oracle/, the C++ restatement every GPU test is checked
against) and, with --gpu, with the B200 library, and prints per file and field the relative RMS and maximum difference.

Recipe for the reference side (any Linux box with gfortran + NetCDF-Fortran; nothing here can build it: no Fortran compiler in the image):
    cd speedy.f90 && NETCDF=/usr bash build.sh          # source/gfortran.makefile: -Ofast -fconvert=swap
    # namelist.nml (the shipped file already has nsteps_out = 1): start 1982-01-01 00:00, end 1982-01-03 00:00 = BASELINE configs[0]
    sed -i 's/end_datetime%day *= *10/end_datetime%day    = 3/' namelist.nml
    bash run.sh && ls rundir/*.nc                       # 198201010000.nc ... 198201030000.nc (73 files)
    python tools/compare_reference_run.py rundir --gpu  # on a B200 box; without --gpu the oracle alone is compared (CPU)

Expected, if the restatement is faithful: float32 fields agree to a few ulp early in the run (the reference is built -Ofast:
reassociation / FMA contraction make it non-reproducible across compilers) and drift apart at the rate two
reference builds with different compilers would; an O(1e-3) relative difference at step 1 means a misread formula.

    --self-test DIR   writes files for the first steps with the library's own host-side writer from the ORACLE's fields into DIR and
                      compares them: exercises this harness end to end without a reference build (pytest runs it on the CPU)."""
import argparse
import ctypes
import glob
import os
import re
import sys

import numpy as np
from scipy.io import netcdf_file

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
FIELDS = ("u", "v", "t", "q", "phi", "ps")


def read_reference_file(path):
    """-> dict of float32 arrays in C order (lev, lat, lon) / (lat, lon), and the file's time axis value (hours since start)"""
    nc = netcdf_file(path, "r", mmap=False)
    out = {n: np.array(nc.variables[n][0], dtype=np.float32) for n in FIELDS}
    hours = float(nc.variables["time"][0])
    nc.close()
    return out, hours


def stats(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    d = a - b
    ref = max(float(np.sqrt(np.mean(b * b))), 1e-300)
    return float(np.sqrt(np.mean(d * d))) / ref, float(np.abs(d).max())


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("directory", nargs="?")
    ap.add_argument("--gpu", action="store_true", help="also run the B200 library and compare it")
    ap.add_argument("--steps-per-day", type=int, default=36)
    ap.add_argument("--max-files", type=int, default=0)
    ap.add_argument("--self-test", metavar="DIR")
    ap.add_argument("--json", metavar="FILE")
    args = ap.parse_args()
    from conftest import Oracle, load_pkg
    pkg = load_pkg()
    o = Oracle("t30")
    bc = os.path.join(ROOT, "data", "bc_t30.bin")

    if args.self_test:
        os.makedirs(args.self_test, exist_ok=True)
        o.model_init(bc)
        for step in range(0, 4):
            if step:
                assert o.run(1) == 0
            f = o.output_fields()
            (y, m, d, h, mi), _ = o.date()
            pkg.write_output_file(os.path.join(args.self_test, f"{y:04d}{m:02d}{d:02d}{h:02d}{mi:02d}.nc"), f["u"], f["v"], f["t"], f["q"], f["phi"], f["ps"],
                                  trunc=30, nsteps=args.steps_per_day, start=(1982, 1, 1, 0, 0), timestep=step)
        args.directory = args.self_test

    files = sorted(p for p in glob.glob(os.path.join(args.directory, "*.nc")) if re.fullmatch(r"\d{12}\.nc", os.path.basename(p)))
    if not files:
        raise SystemExit(f"no yyyymmddhhmm.nc files in {args.directory}")
    if args.max_files:
        files = files[:args.max_files]
    name0 = os.path.basename(files[0])
    start = tuple(int(name0[a:b]) for a, b in ((0, 4), (4, 6), (6, 8), (8, 10), (10, 12)))
    o.model_init(bc, *start)
    g = None
    if args.gpu:
        g = pkg.Speedy(trunc=30)
        g.model_init(bc, *start)
    hours_per_step = 24.0 / args.steps_per_day
    done, rows, worst = 0, [], {"oracle": 0.0, "gpu": 0.0}
    print(f"{'file':>16} {'step':>5}  " + "  ".join(f"{n:>22}" for n in FIELDS) + ("   [rel-RMS oracle vs reference" + (" | GPU vs reference]" if g else "]")))
    for path in files:
        ref, hours = read_reference_file(path)
        step = int(round(hours / hours_per_step))
        if step < done:
            raise SystemExit(f"{path}: time axis goes backwards")
        if step > done:
            assert o.run(step - done) == 0, "oracle: model variables out of accepted range"
            if g:
                assert g.run_steps(step - done) == 0
            done = step
        fo = o.output_fields()
        fg = g.output_fields() if g else None
        cells = []
        row = {"file": os.path.basename(path), "step": step}
        for n in FIELDS:
            r_o, m_o = stats(fo[n], ref[n])
            worst["oracle"] = max(worst["oracle"], r_o)
            row[n] = {"oracle_rel_rms": r_o, "oracle_max_abs": m_o}
            cell = f"{r_o:9.2e}"
            if fg is not None:
                r_g, m_g = stats(fg[n], ref[n])
                worst["gpu"] = max(worst["gpu"], r_g)
                row[n].update({"gpu_rel_rms": r_g, "gpu_max_abs": m_g})
                cell += f" | {r_g:9.2e}"
            cells.append(f"{cell:>22}")
        rows.append(row)
        print(f"{os.path.basename(path):>16} {step:5d}  " + "  ".join(cells))
    print(f"worst relative RMS over {len(files)} files: oracle {worst['oracle']:.3e}" + (f", GPU {worst['gpu']:.3e}" if g else ""))
    if args.json:
        import json
        json.dump({"files": rows, "worst": worst}, open(args.json, "w"), indent=1)
    return worst


if __name__ == "__main__":
    main()
