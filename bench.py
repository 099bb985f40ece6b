#!/usr/bin/env python
"""bench.py — simulated-days/sec of the speedy.f90 hot path at T30/L8 (BASELINE.json metric).

A bench "step" is ONE SIMULATED DAY = 36 model time steps (params.f90:30) of the main-loop body
(speedy.f90:27-54: daily forcing, step(2,2,2*delt) with dynamics + physics, diagnostics, calendar,
land/sea slabs) for every ensemble member resident on this rank's GPU, replayed from a CUDA graph.

  python bench.py --gpus N --steps K --warmup W [--members M] [--impl reference]

N > 1 (torchrun, one rank per GPU): members shard across ranks with no data-path collective
 -> weak scaling; the only collective is the timing max/barrier.
`--impl reference` times the reference's CPU algorithm (the oracle restatement built -Ofast:
no Fortran compiler exists in this image) on the host, one core (the
reference is serial), rank 0 only.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
BC = os.path.join(ROOT, "data", "bc_t30.bin")
NSTEPS_PER_DAY = 36
METRIC = "simulated-days/sec at T30/L8"
UNIT = "sim-days/s"


def _args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60, help="timed simulated days")
    ap.add_argument("--warmup", type=int, default=5, help="untimed simulated days")
    ap.add_argument("--members", type=int, default=1, help="ensemble members per GPU (1 = BASELINE configs[1])")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def _config(args):
    return {"workload": "T30/L8 (96x48x8, 36 steps/day) single-member integration from the rest state, fp64 "
                        "(BASELINE configs[1]); one bench step = 1 simulated day",
            "members_per_gpu": args.members, "total_members": args.members * args.gpus,
            "model_steps_per_bench_step": NSTEPS_PER_DAY, "start_date": "1982-01-01",
            "parallelism": f"ensemble members sharded {args.members}/GPU x {args.gpus} GPU(s), no data-path collective",
            "l2": "256 MiB buffer overwritten between timed steps (L2 flush); within a simulated day the 15 MB state is L2-resident by construction",
            "output_cadence": "none inside the timed region (nsteps_out > run length)"}


# ------------------------------------------------------------------------------------------
# CPU side: the oracle restatement, fast build (test infrastructure used as the reported baseline)
# ------------------------------------------------------------------------------------------
def _oracle_fast():
    path = os.path.join(ROOT, "oracle", "liboracle_t30_fast.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle_t30_fast.so"])
    L = ctypes.CDLL(path)
    rc = L.orc_model_init(BC.encode(), 1982, 1, 1, 0, 0)
    if rc:
        raise RuntimeError(f"oracle init failed rc={rc}")
    return L


def _cpu_days_per_sec(L, days):
    t0 = time.perf_counter()
    rc = L.orc_model_run(days * NSTEPS_PER_DAY)
    dt = time.perf_counter() - t0
    if rc:
        raise RuntimeError("oracle diagnostics out of range")
    return days / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    L = _oracle_fast()
    _cpu_days_per_sec(L, max(args.warmup, 1))
    v, dt = _cpu_days_per_sec(L, args.steps)
    sample = f"{args.steps} simulated days ({args.steps * NSTEPS_PER_DAY} time steps) after {max(args.warmup, 1)} warm-up days, single member"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "reference T30 boundary files (packed), rest-state initial condition",
            "config": _config(args),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                             "note": "C++ restatement of speedy.f90 (g++ -Ofast -march=native), not a gfortran build: no Fortran compiler in this image; the reference is serial"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def _ncu_traffic(kernel, members):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the newest committed
    `ncu --set full` summary under profiles/ (captured at 1 member; None for other batch sizes)"""
    import glob
    if members != 1:
        return None, "no ncu capture at this member count"
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full_summary.json")), reverse=True):
        try:
            k = json.load(open(path))["kernels"][kernel]
            return k["traffic_bytes"], os.path.relpath(path, ROOT)
        except Exception:
            continue
    return None, "no ncu summary committed"


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from __graft_entry__ import _load_pkg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = _load_pkg()
    c = pkg.Speedy(trunc=30, nmembers=args.members, device=local)
    c.model_init(BC)
    stream = torch.cuda.ExternalStream(c.stream, device=torch.device("cuda", local))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`) ----------------
    for _ in range(max(args.warmup, 3)):
        assert c.run_steps(NSTEPS_PER_DAY) == 0
    c.synchronize()
    launches0 = c.launch_count
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)               # L2 flush between timed steps (untimed)
        torch.cuda.synchronize()
        ev[k][0].record(stream)
        rc = c.run_steps(NSTEPS_PER_DAY)    # one simulated day; returns after the stream drained (diagnostics guard read back)
        ev[k][1].record(stream)
        assert rc == 0, "check_diagnostics: model variables out of accepted range"
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    dev_s = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
    launches = c.launch_count - launches0
    tt = torch.tensor([dev_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_s = float(tt.item())
    total_days = args.steps * args.members * world
    value = total_days / dev_s

    # ---------------- end-to-end through the C ABI with host buffers (`e2e`) ----------------
    # every step: prognostic state host(pinned) -> device, one simulated day, state + the 41 output levels device -> host,
    # all inside ONE reference-facing call: speedy_run_steps_host (the main loop with the module arrays left on the host)
    names = ("vor", "div", "t", "tr", "ps")
    st = np.concatenate([np.concatenate([c.get_field(n, all_members=True)[e].view(np.float64).ravel() for n in names]) for e in range(args.members)])
    state = torch.from_numpy(st).pin_memory()
    outb = torch.empty((5 * c.kx + 1) * c.il * c.ix, dtype=torch.float32).pin_memory()
    assert state.numel() == c.state_len() * args.members
    h2d = state.numel() * 8
    d2h = state.numel() * 8 + outb.numel() * 4
    e2e_days = max(3, min(args.steps, 30))
    s_np, o_np = state.numpy(), outb.numpy()
    assert c.run_steps_host(s_np, NSTEPS_PER_DAY, o_np) == 0      # warm-up of the host path
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_days):
        assert c.run_steps_host(s_np, NSTEPS_PER_DAY, o_np) == 0
    barrier()
    e2e_s = time.perf_counter() - t0
    t2 = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = e2e_days * args.members * world / float(t2.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (rank 0, N=1 semantics) ----------------
    peaks, peak_src = {}, "fallback (B200_PROFILING.md)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    kt_cold = c.time_kernels(NSTEPS_PER_DAY, flush_l2=True)
    kt_warm = c.time_kernels(NSTEPS_PER_DAY, flush_l2=False)
    # the same kernels on the GPU's own nanosecond timer inside the replayed graph (first CTA start -> last CTA end,
    # programmatic dependent launch off so that the kernels do not overlap): no event / launch overhead in these
    c.trace(True)
    c.run_steps(3 * NSTEPS_PER_DAY)
    timeline = c.trace_read()
    c.trace(False)
    M = args.members
    nact = sum(2 * min(c.mx, c.trunc + 2 - n) for n in range(c.nx))      # 1054 active reals (legendre.f90:33-41)
    NG = c.ix * c.il
    alg = {  # algorithmic bytes per launch (SURVEY.md §8d per-unit figures x units per launch)
        "spec_to_grid": M * 77 * (8 * nact + 8 * NG),                   # 91 of the reference minus the 14 level-1 wind fields nobody reads
        "grid_to_spec": M * 73 * (8 * NG + 8 * (nact - 2)),
        "grid_columns": M * NG * 8 * (77 + 73 + 45 + 8 + 32 + 2),       # 77 fields in, 73 out, ~45 2-D surface/slab state, tt_rsw, tau2, stratc
        "spec_step": M * c.mx * c.nx * 16 * 165,
    }
    # BASELINE metric (2), "Legendre GB/s vs roofline": the two transform kernels alone on a large device-resident batch
    # (5824 = 91 fields x 8 levels x 8 members), CUDA events on the library stream, L2 flushed between reps
    def _transform_batch(inverse, nb):
        spec = torch.rand((nb, c.nx, c.mx, 2), dtype=torch.float64, device="cuda") * 2 - 1
        grid = torch.rand((nb, c.il, c.ix), dtype=torch.float64, device="cuda") * 2 - 1
        times = []
        for rep in range(6):
            flush.fill_(rep)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            if inverse:
                rc = c.L.speedy_spec_to_grid_dev(c.h, ctypes.c_void_p(spec.data_ptr()), nb, None, ctypes.c_void_p(grid.data_ptr()))
            else:
                rc = c.L.speedy_grid_to_spec_dev(c.h, ctypes.c_void_p(grid.data_ptr()), nb, ctypes.c_void_p(spec.data_ptr()))
            assert rc == 0
            e1.record(stream)
            e1.synchronize()
            if rep >= 2:
                times.append(e0.elapsed_time(e1) * 1e-3)
        t = sorted(times)[len(times) // 2]
        by = nb * ((8 * nact + 8 * NG) if inverse else (8 * NG + 8 * (nact - 2)))
        return {"batch": nb, "us": 1e6 * t, "GBps": by / t / 1e9, "frac_hbm": by / t / 1e9 / hbm, "transforms_per_s": nb / t}
    try:
        legendre_batched = {"spec_to_grid": _transform_batch(True, 5824), "grid_to_spec": _transform_batch(False, 4672),
                            "note": "algorithmic bytes per transform over the measured HBM copy bandwidth; at this batch the transforms are bound by shared-memory wavefronts of the Legendre stages and, in grid->spec, by the dense DFT on the FP64 tensor pipe (profiles/README.md, r1n), not by HBM"}
    except Exception as ex:
        legendre_batched = {"error": str(ex)}
    tot = sum(kt_warm.values())
    dom = max(alg, key=lambda k: timeline["us"][k])     # the longest kernel of the step on the GPU's own timer
    ach = alg[dom] / (kt_cold[dom] * 1e-3) / 1e9
    traffic, traffic_src = _ncu_traffic(dom, M)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": traffic,
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": alg[dom],
                "peak_source": peak_src, "timing": "CUDA events on the library stream around each launch, L2 flushed before each launch, mean of 36 steps",
                "achieved_warm_l2": alg[dom] / (kt_warm[dom] * 1e-3) / 1e9,
                "share_of_step": kt_warm[dom] / tot,
                "kernel_ms_cold": kt_cold, "kernel_ms_warm": kt_warm,
                "kernel_us_gpu_timer": timeline["us"], "gap_before_us_gpu_timer": timeline["gap_before_us"],
                "gpu_timer_note": "in-graph durations on %globaltimer with programmatic dependent launch disabled; the CUDA-event figures above carry ~5 us of event+launch overhead per kernel",
                "legendre": {k: {"GBps_cold": alg[k] / (kt_cold[k] * 1e-3) / 1e9, "frac_hbm_cold": alg[k] / (kt_cold[k] * 1e-3) / 1e9 / hbm,
                                 "GBps_warm": alg[k] / (kt_warm[k] * 1e-3) / 1e9} for k in ("spec_to_grid", "grid_to_spec")},
                "legendre_batched": legendre_batched,
                "note": "single-member T30 is latency-bound (SURVEY.md F12): one step moves ~19 MB (3 us of HBM time); each kernel is a chain of dependent FP64 operations (DFMA 8.7 cycles, exp 160 cycles dependent issue on B200, tools/dmma_probe.cu)"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "reference T30 boundary files (packed), rest-state initial condition",
            "config": _config(args), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "speedy_run_steps_host(ctx, state, n, 36, out): pinned host state -> device, 36 steps, state + output() fields -> host, per simulated day; wall clock",
                    "days": e2e_days},
            "gpu_launches": int(launches), "us_per_model_step": 1e6 * dev_s / (args.steps * NSTEPS_PER_DAY),
            "wall_s_timed_region": t_wall, "roofline": roofline}
    if world == 1 and args.members == 1:
        # BASELINE configs[2] in passing (not the headline): 8 SPPT members resident on this GPU, 5 simulated days
        try:
            c8 = pkg.Speedy(trunc=30, nmembers=8, device=local, sppt_on=1, seed=1)
            c8.model_init(BC)
            for _ in range(2):
                assert c8.run_steps(NSTEPS_PER_DAY) == 0
            c8.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                assert c8.run_steps(NSTEPS_PER_DAY) == 0
            c8.synchronize()
            dt8 = time.perf_counter() - t0
            line["ensemble_8_members_per_gpu"] = {"member_days_per_s": 8 * 5 / dt8, "us_per_step": 1e6 * dt8 / (5 * NSTEPS_PER_DAY),
                                                  "note": "SPPT on, wall clock, device-resident; per-GPU figure of the 64-member / 8-GPU config"}
            c8.close()
        except Exception as ex:
            line["ensemble_8_members_per_gpu"] = {"error": str(ex)}
    if not args.no_cpu_baseline and world == 1:
        try:
            L = _oracle_fast()
            _cpu_days_per_sec(L, 2)
            v, dt = _cpu_days_per_sec(L, 25)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"25 simulated days (900 time steps, {dt:.1f} s) after 2 warm-up days, single member",
                                    "note": "C++ restatement of speedy.f90 built -Ofast -march=native (no Fortran compiler in this image); the reference is serial"}
        except Exception as ex:  # the baseline leg must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"failed: {ex}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = _args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
