#!/usr/bin/env python
"""bench.py — simulated-days/sec of the speedy.f90 hot path at T30/L8 (BASELINE.json metric).

A bench "step" is ONE SIMULATED DAY = 36 model time steps (params.f90:30) of the main-loop body
(speedy.f90:27-54: daily forcing, step(2,2,2*delt) with dynamics + physics, diagnostics, calendar,
land/sea slabs) for every ensemble member resident on this rank's GPU, replayed from a CUDA graph.

  python bench.py --gpus N --steps K --warmup W [--members M] [--impl reference]

N > 1 (torchrun, one rank per GPU): members shard across ranks with no data-path collective
 -> weak scaling; the only collective is the timing max/barrier.
For every N the same run also measures BASELINE configs[2] at that N (8 SPPT members per GPU, the
ensemble-moment all-reduce over NCCL once per simulated day): key `ensemble_sppt_8_per_gpu`.
`--impl reference` times the reference's CPU algorithm — the C++ PORT in oracle/ built -Ofast (no
Fortran compiler exists in this image; cpu_baseline.kind = "port") — on the host,
one core (the reference is serial), rank 0 only.
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
    ap.add_argument("--no-ensemble", action="store_true", help="skip the BASELINE configs[2] leg (8 SPPT members per GPU + daily moment all-reduce)")
    return ap.parse_args()


def _config(args):
    return {"workload": "T30/L8 (96x48x8, 36 steps/day) single-member integration from the rest state, fp64 "
                        "(BASELINE configs[1]); one bench step = 1 simulated day",
            "members_per_gpu": args.members, "total_members": args.members * args.gpus,
            "model_steps_per_bench_step": NSTEPS_PER_DAY, "start_date": "1982-01-01",
            "parallelism": f"ensemble members sharded {args.members}/GPU x {args.gpus} GPU(s), no data-path collective",
            "l2": "256 MiB buffer overwritten between timed steps (L2 flush, enqueued on the same stream, untimed); within a simulated day the 15 MB state is L2-resident by construction",
            "host_sync": "none inside the timed region: all timed days are enqueued (CUDA graph replays), the range guard is polled once after the last day",
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
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25", "-i", str(index)],
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


def _pin_rank_to_cores(local, nlocal):
    """one disjoint block of host cores per rank: the per-day launch work of N processes must not share cores"""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(nlocal, 1))
        mine = cores[local * per:(local + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        return len(mine)
    except Exception:
        return None


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
    ncores = _pin_rank_to_cores(local, int(os.environ.get("LOCAL_WORLD_SIZE", str(world))))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = _load_pkg()
    c = pkg.Speedy(trunc=30, nmembers=args.members, device=local)
    c.model_init(BC)
    stream = torch.cuda.ExternalStream(c.stream, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def rank_max(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident throughput (`value`) ----------------
    # every timed day is ENQUEUED (graph replay) behind an untimed L2 flush on the same stream, bracketed by CUDA events on that
    # stream; the host synchronises once, after the last day (the range guard is polled there): no host round trip per day
    for _ in range(max(args.warmup, 3)):
        assert c.run_steps(NSTEPS_PER_DAY) == 0
    c.synchronize()
    launches0 = c.launch_count
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for k in range(args.steps):
            flush.fill_(k & 0xFF)               # L2 flush between timed steps (untimed, same stream)
            ev[k][0].record(stream)
            c.enqueue_steps(NSTEPS_PER_DAY)     # one simulated day
            ev[k][1].record(stream)
    rc = c.finish()
    assert rc == 0, "check_diagnostics: model variables out of accepted range"
    barrier()
    t_wall = time.perf_counter() - t_wall0
    dev_s = rank_max(sum(a.elapsed_time(b) for a, b in ev) * 1e-3)
    launches = c.launch_count - launches0
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
    e2e_value = e2e_days * args.members * world / rank_max(time.perf_counter() - t0)

    # ---------------- BASELINE configs[2] at this N: 8 SPPT members per GPU, ensemble-moment all-reduce once per simulated day ----------------
    ens_line = None
    if args.members == 1 and not args.no_ensemble:
        try:
            ens_line = _ensemble_leg(pkg, torch, dist, world, rank, local, dev, barrier, rank_max)
        except Exception as ex:
            ens_line = {"error": repr(ex)}
    # clocks / throttle reasons sampled through ALL the timed regions (device-resident days, host-buffer days, the ensemble leg): the
    # first alone lasts ~20 ms at the driver's --steps 20
    clocks = sampler.stop() if sampler else None

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
    # programmatic dependent launch off so that the kernels do not overlap): no event / launch overhead in these.
    # ONE timer for the roofline: the kernel ranking and achieved / frac both come from these durations.
    c.trace(True)
    c.run_steps(3 * NSTEPS_PER_DAY)
    timeline = c.trace_read()
    c.trace(False)
    M = args.members
    alg = _algorithmic_bytes(c, M)
    try:
        legendre_batched = {"spec_to_grid": _transform_batch(c, torch, stream, flush, True, 5824, hbm),
                            "grid_to_spec": _transform_batch(c, torch, stream, flush, False, 4672, hbm),
                            "note": "BASELINE metric (2): algorithmic bytes per transform over the measured HBM copy bandwidth, device-resident random fields, L2 flushed "
                                    "between repetitions; both directions run the quad kernels (four fields at a time: FFTPACK FFT, Legendre sums as FP64 tensor-core "
                                    "tiles with the P fragments in tensor memory, tensor-map loads and stores); both are bound by shared-memory wavefronts and the "
                                    "FP64 pipe before HBM (profiles/README.md r2)"}
    except Exception as ex:
        legendre_batched = {"error": str(ex)}
    dom = max(alg, key=lambda k: timeline["us"][k])     # the longest kernel of the step on the GPU's own timer
    dom_us = timeline["us"][dom]
    ach = alg[dom] / (dom_us * 1e-6) / 1e9
    traffic, traffic_src = _ncu_traffic(dom, M)
    step_us = 1e6 * dev_s / (args.steps * NSTEPS_PER_DAY)
    whole_step = {f"members_{M}": {"algorithmic_bytes_per_step": sum(alg.values()), "us_per_step": step_us,
                                   "GBps": sum(alg.values()) / (step_us * 1e-6) / 1e9, "frac_hbm": sum(alg.values()) / (step_us * 1e-6) / 1e9 / hbm}}
    if ens_line and "us_per_step" in ens_line:
        a8 = sum(_algorithmic_bytes(c, ens_line["members_per_gpu"], sppt=True).values())
        whole_step[f"members_{ens_line['members_per_gpu']}_sppt"] = {"algorithmic_bytes_per_step": a8, "us_per_step": ens_line["us_per_step"],
                                                                      "GBps": a8 / (ens_line["us_per_step"] * 1e-6) / 1e9,
                                                                      "frac_hbm": a8 / (ens_line["us_per_step"] * 1e-6) / 1e9 / hbm}
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": traffic,
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": alg[dom],
                "peak_source": peak_src,
                "timing": "in-graph duration of the kernel on %globaltimer (first CTA start -> last CTA end), mean over 108 replayed steps, programmatic dependent launch off; "
                          "the kernel ranking uses the same timer",
                "kernel_us_gpu_timer": timeline["us"], "gap_before_us_gpu_timer": timeline["gap_before_us"],
                "share_of_step": dom_us / sum(timeline["us"].values()),
                "frac_by_kernel": {k: alg[k] / (timeline["us"][k] * 1e-6) / 1e9 / hbm for k in alg},
                "cuda_event_timing": {"note": "secondary: CUDA events around each plain launch carry ~5 us of event + launch overhead per kernel",
                                      "kernel_ms_cold_l2": kt_cold, "kernel_ms_warm_l2": kt_warm,
                                      "achieved_cold_l2": alg[dom] / (kt_cold[dom] * 1e-3) / 1e9},
                "whole_step": whole_step,
                "legendre_batched": legendre_batched,
                "note": "single-member T30 is latency-bound (SURVEY.md F12): one step moves ~19 MB (3 us of HBM time); each kernel is a chain of dependent FP64 operations (DFMA 8.7 cycles, exp 160 cycles dependent issue on B200, tools/dmma_probe.cu)"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "reference T30 boundary files (packed), rest-state initial condition",
            "config": _config(args), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "speedy_run_steps_host(ctx, state, n, 36, out): pinned host state -> device, 36 steps, state + output() fields -> host, per simulated day; wall clock",
                    "days": e2e_days},
            "gpu_launches": int(launches), "us_per_model_step": step_us,
            "wall_s_timed_region": t_wall, "host_cores_per_rank": ncores, "roofline": roofline}
    if ens_line is not None:
        line["ensemble_sppt_8_per_gpu"] = ens_line
    if not args.no_cpu_baseline and world == 1:
        try:
            L = _oracle_fast()
            _cpu_days_per_sec(L, 2)
            v, dt = _cpu_days_per_sec(L, 25)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"25 simulated days (900 time steps, {dt:.1f} s) after 2 warm-up days, single member",
                                    "note": "C++ PORT of speedy.f90 (oracle/, g++ -Ofast -march=native), not a gfortran build of the reference: no Fortran compiler in this image; the reference is serial"}
        except Exception as ex:  # the baseline leg must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"failed: {ex}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _algorithmic_bytes(c, M, sppt=False):
    """algorithmic bytes per launch (SURVEY.md §8d per-unit figures x units per launch)"""
    nact = sum(2 * min(c.mx, c.trunc + 2 - n) for n in range(c.nx))      # 1054 active reals (legendre.f90:33-41)
    NG = c.ix * c.il
    ninv = 77 + (8 if sppt else 0)                                        # 91 of the reference minus the 14 level-1 wind fields nobody reads (+8 SPPT)
    return {"spec_to_grid": M * ninv * (8 * nact + 8 * NG),
            "grid_to_spec": M * 73 * (8 * NG + 8 * (nact - 2)),
            "grid_columns": M * NG * 8 * (ninv + 73 + 45 + 8 + 32 + 2),   # fields in, 73 out, ~45 2-D surface/slab state, tt_rsw, tau2, stratc
            "spec_step": M * c.mx * c.nx * 16 * 165}


def _transform_batch(c, torch, stream, flush, inverse, nb, hbm):
    """BASELINE metric (2), "Legendre GB/s vs roofline": a transform kernel alone on a large device-resident batch
    (5824 = 91 fields x 8 levels x 8 members), CUDA events on the library stream, L2 flushed between reps"""
    nact = sum(2 * min(c.mx, c.trunc + 2 - n) for n in range(c.nx))
    NG = c.ix * c.il
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


def _ensemble_leg(pkg, torch, dist, world, rank, local, dev, barrier, rank_max, members=8, days=8):
    """BASELINE configs[2] at N GPUs: `members` SPPT members per GPU (global member index = rank * members + e keys the noise),
    the whole ensemble advanced `days` simulated days; once per simulated day the sum / sum of squares of the 41 output levels over
    the resident members (speedy_ensemble_sums_dev) is all-reduced over the ranks (NCCL, one flat fp64 buffer: the ensemble
    mean / spread diagnostic).  Everything is enqueued; one host synchronisation at the end."""
    ce = pkg.Speedy(trunc=30, nmembers=members, device=local, sppt_on=1, seed=1, member_offset=rank * members)
    ce.model_init(BC)
    es = torch.cuda.ExternalStream(ce.stream, device=dev)
    n41 = 41 * ce.il * ce.ix
    flat = torch.zeros(2 * n41 + 1, dtype=torch.float64, device=dev)
    comm_seen = dist.get_world_size() if world > 1 else 1

    def one_day():
        ce.enqueue_steps(NSTEPS_PER_DAY)
        rc = ce.L.speedy_ensemble_sums_dev(ce.h, ctypes.c_void_p(flat.data_ptr()), ctypes.c_void_p(flat.data_ptr() + 8 * n41))
        assert rc == 0
        flat[-1] = float(members)
        if world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    with torch.cuda.stream(es):
        for _ in range(2):
            one_day()
    assert ce.finish() == 0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(es):
        e0.record(es)
        for _ in range(days):
            one_day()
        e1.record(es)
    assert ce.finish() == 0
    barrier()
    dt = rank_max(e0.elapsed_time(e1) * 1e-3)
    total = int(round(float(flat[-1].item())))
    # steady-state cost of the collective alone
    ar_us = None
    if world > 1:
        with torch.cuda.stream(es):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(es)
            for _ in range(20):
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            a1.record(es)
        torch.cuda.synchronize()
        ar_us = rank_max(a0.elapsed_time(a1) * 1e3 / 20)
    ce.close()
    return {"config": "BASELINE configs[2]: T30/L8 SPPT ensemble, 8 members per GPU", "members_per_gpu": members, "total_members": members * world,
            "members_counted_by_allreduce": total, "comm_nranks_seen": comm_seen, "days": days,
            "member_days_per_s": members * world * days / dt, "us_per_step": 1e6 * dt / (days * NSTEPS_PER_DAY),
            "allreduce_us_steady_state": ar_us, "allreduce_bytes": int(flat.numel() * 8),
            "timing": "CUDA events on the library stream around all days (incl. the daily moment reduction and all-reduce), max over ranks; weak scaling"}


def main():
    args = _args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
