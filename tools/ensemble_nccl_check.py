"""Multi-GPU check of the ensemble path: under torchrun (one rank per GPU, NCCL) every rank integrates its
block of an 8-member SPPT ensemble for one day and the ensemble mean / spread is formed with ONE all-reduce of the on-device
moments; rank 0 then repeats the whole ensemble in a single context and compares (members must be bit-identical, moments equal
to summation-order rounding).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/ensemble_nccl_check.py
"""
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import _load_pkg  # noqa: E402

TOTAL, STEPS = 8, 36


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = _load_pkg()
    ens = importlib.import_module("speedy_f90_b200.ensemble")
    E = ens.Ensemble(pkg, TOTAL, device=local, sppt_on=1, seed=11, rank=rank, world=world)
    E.model_init(pkg.BC_T30)
    assert E.run_steps(STEPS) == 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    mean, spread = E.mean_spread()                    # NCCL all-reduce inside
    ev1.record(); torch.cuda.synchronize()
    mine = E.ctx.get_field("t", all_members=True)
    res = {"world": world, "rank_blocks": None, "allreduce_ms": ev0.elapsed_time(ev1)}
    if rank == 0:
        W = ens.Ensemble(pkg, TOTAL, device=local, sppt_on=1, seed=11, rank=0, world=1)
        W.model_init(pkg.BC_T30)
        assert W.run_steps(STEPS) == 0
        sw, sw2 = W.local_moments()                    # all 8 members are resident here: no collective
        m1, s1 = ens.moments_to_mean_spread(sw, sw2, TOTAL)
        allm = W.ctx.get_field("t", all_members=True)
        res["members_bit_identical"] = bool(np.array_equal(mine, allm[E.lo:E.hi]))
        scale = m1.abs().amax(dim=(1, 2), keepdim=True).clamp_min(1e-30)
        res["mean_max_rel_diff"] = float(((mean - m1).abs() / scale).max())
        res["spread_max_rel_diff"] = float(((spread - s1).abs() / scale).max())
        res["rank_blocks"] = [list(ens.block_partition(TOTAL, world, r)) for r in range(world)]
        res["t_spread_max"] = float(spread[16:24].max())
        W.close()
        print(json.dumps(res), flush=True)
        assert res["members_bit_identical"] and res["mean_max_rel_diff"] < 1e-12 and res["spread_max_rel_diff"] < 1e-6
    E.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
