"""BASELINE configs[4] as it is named: an 8-member SPPT ensemble on the GPUs of one node (one member per GPU on 8 GPUs), run twice —
fp64 everywhere, and with the real32 spherical-harmonic transforms (precision = 1: Legendre + Fourier + uvspec / grad in real32; grid-point
columns, semi-implicit solve and time stepping fp64) — with the SAME noise (the SPPT stream is keyed by the global member index).
Every 6 h to 48 h: relative RMS of each member's prognostic spectral fields between the two runs (max and mean over the members) and of the
ensemble-mean output fields (the NCCL moment all-reduce of both ensembles); then the throughput of both modes over the same days.

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/configs4_ensemble.py [members] > out.json
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import _load_pkg  # noqa: E402

PROG = ("vor", "div", "t", "tr", "ps")


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = _load_pkg()
    from speedy_f90_b200 import ensemble
    runs = [ensemble.Ensemble(pkg, total, device=local, sppt_on=1, seed=5, rank=rank, world=world, precision=p) for p in (0, 1)]
    for r in runs:
        r.model_init(pkg.BC_T30)
    dev = torch.device("cuda", local)
    rows = []
    for h in range(6, 49, 6):
        for r in runs:
            assert r.run_steps(9) == 0
        row = {"hours": h}
        for n in PROG:
            a = runs[1].ctx.get_field(n, all_members=True)[:, 0]
            b = runs[0].ctx.get_field(n, all_members=True)[:, 0]
            num = np.sum(np.abs(a - b) ** 2, axis=tuple(range(1, a.ndim)))
            den = np.sum(np.abs(b) ** 2, axis=tuple(range(1, a.ndim)))
            rel = torch.tensor(np.sqrt(num / np.maximum(den, 1e-300)), dtype=torch.float64, device=dev)
            if world > 1:
                parts = [torch.empty_like(rel) for _ in range(world)]
                dist.all_gather(parts, rel)
                rel = torch.cat(parts)
            row[n] = {"max_over_members": float(rel.max()), "mean_over_members": float(rel.mean()), "members": int(rel.numel())}
        m0, s0 = runs[0].mean_spread()
        m1, s1 = runs[1].mean_spread()
        names = ("u", "v", "t", "q", "phi")
        em = {}
        for i, n in enumerate(names):
            d, b = m1[8 * i:8 * i + 8] - m0[8 * i:8 * i + 8], m0[8 * i:8 * i + 8]
            em[n] = float(torch.sqrt((d * d).mean() / (b * b).mean().clamp_min(1e-300)))
        d, b = m1[40] - m0[40], m0[40]
        em["ps"] = float(torch.sqrt((d * d).mean() / (b * b).mean()))
        row["ensemble_mean_output_fields"] = em
        row["spread_t_lowest_level_K"] = {"fp64": float(s0[23].mean()), "real32_transforms": float(s1[23].mean())}
        rows.append(row)
    # throughput of both modes: 10 days each, all enqueued, one synchronisation; max over ranks
    perf = {}
    for name, r in (("fp64", runs[0]), ("real32_transforms", runs[1])):
        c = r.ctx
        c.enqueue_steps(36); c.finish()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        days = 10
        for _ in range(days):
            c.enqueue_steps(36)
        assert c.finish() == 0
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        perf[name] = {"member_days_per_s": total * days / float(dt), "us_per_step": 1e6 * float(dt) / (days * 36), "members_per_gpu": r.hi - r.lo}
    if rank == 0:
        print(json.dumps({"config": "BASELINE configs[4]: T30/L8, %d-member SPPT ensemble on %d GPU(s), real32 transforms + fp64 columns / implicit solve vs fp64" % (total, world),
                          "n_gpus": world, "members": total, "what": "rel. RMS real32-transform run vs fp64 run, same SPPT noise; per-member prognostic spectral coefficients (time level 1) and ensemble-mean output fields (NCCL moment all-reduce)",
                          "curves": rows, "throughput": perf,
                          "note": "precision = 1 is a tolerance-study mode: its transforms are plain FFMA kernels (transforms_f32.cu), slower than the fp64 production kernels"}, indent=1))
    for r in runs:
        r.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
