"""Host-side ensemble logic (speedy.f90_b200/ensemble.py) on CPU: the member partition and the
ensemble-mean/spread reduction over a world_size-2 gloo process group (members
shard across ranks, the only collective is the moment all-reduce on output steps)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
from conftest import ROOT, load_pkg


def _ens():
    load_pkg()
    import importlib
    return importlib.import_module("speedy_f90_b200.ensemble")


def test_block_partition_covers_every_member_once():
    ens = _ens()
    for total in (1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = ens.block_partition(total, world, r)
                assert 0 <= lo <= hi <= total
                seen += list(range(lo, hi))
            assert seen == list(range(total))                       # contiguous, ordered, disjoint, complete
            sizes = [np.subtract(*ens.block_partition(total, world, r)[::-1]) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert ens.block_partition(64, 8, 3) == (24, 32)                # BASELINE configs[2]: 8 members per GPU
    assert ens.owner_of(27, 64, 8) == (3, 3)
    with pytest.raises(ValueError):
        ens.block_partition(4, 2, 2)


def test_moments_to_mean_spread_numpy():
    ens = _ens()
    rng = np.random.default_rng(0)
    x = rng.standard_normal((6, 41, 4, 5)) * 3 + 2
    m, sd = ens.moments_to_mean_spread(x.sum(0), (x * x).sum(0), 6)
    assert np.allclose(m, x.mean(0)) and np.allclose(sd, x.std(0))


_WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(ROOT_PLACEHOLDER, "tests"))
from conftest import load_pkg
load_pkg()
import importlib
ens = importlib.import_module("speedy_f90_b200.ensemble")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
TOTAL, SHAPE = 5, (41, 6, 12)
def member(e):          # the "output fields" of global member e (synthetic, seeded by the GLOBAL index)
    return np.random.default_rng(1000 + e).standard_normal(SHAPE) * (1 + e) + e
lo, hi = ens.block_partition(TOTAL, world, rank)
mine = np.stack([member(e) for e in range(lo, hi)])
s, s2 = torch.from_numpy(mine.sum(0)), torch.from_numpy((mine * mine).sum(0))
n = ens.allreduce_moments(s, s2, hi - lo)
mean, spread = ens.moments_to_mean_spread(s, s2, n)
allm = np.stack([member(e) for e in range(TOTAL)])
assert n == TOTAL, n
assert np.allclose(mean.numpy(), allm.mean(0), rtol=1e-12, atol=1e-12)
assert np.allclose(spread.numpy(), allm.std(0), rtol=1e-9, atol=1e-9)
# every rank holds the same reduced moments (bitwise: one all-reduce result)
chk = torch.tensor([float(mean.sum()), float(spread.sum())], dtype=torch.float64)
both = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(both, chk)
assert all(torch.equal(both[0], b) for b in both)
dist.destroy_process_group()
print(f"rank{rank}ok\n", end="", flush=True)
"""


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_allreduce_moments_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.replace("ROOT_PLACEHOLDER", repr(ROOT)))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rank0ok" in r.stdout and "rank1ok" in r.stdout


def test_bench_reference_arm_nonzero_rank_is_silent():
    """under torchrun (N>1) rank 0 alone runs the CPU reference arm; other ranks exit 0 without work"""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
