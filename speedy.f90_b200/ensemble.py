"""Ensemble driver: members of an SPPT / initial-condition ensemble sharded over the GPUs of
one node, one process per GPU.

The reference has no ensemble driver (it is one serial trajectory, speedy.f90:21-54); an
ensemble is N independent copies of that loop.  Members never exchange data inside the time
loop, so the only collective is the ensemble-mean / spread diagnostic on output steps: each
rank reduces its resident members on the device (speedy_ensemble_sums_dev: sum and sum of
squares of the 41 output levels, fp64, in output() units, input_output.f90:201-206) and the
partial moments are all-reduced (NCCL over NVLink on GPUs; any torch.distributed backend
works, which is how the CPU tests drive this logic with gloo).

Partition: contiguous blocks — rank r owns global members [lo, hi) and holds them as one
batched context with cfg.member_offset = lo.  The SPPT noise of a member is keyed by its
GLOBAL index (seed, step, member_offset + local, coefficient), so a member's trajectory does
not depend on how many GPUs the ensemble is spread over (tested on the B200 box).
"""
import numpy as np

N_OUT_LEVELS = 41   # u, v, t, q, phi on kx = 8 levels + ps


def block_partition(total, world, rank):
    """Contiguous block partition: rank r owns members [lo, hi).  Blocks differ by at most one
    member; with total % world == 0 every rank holds total // world members (weak scaling)."""
    if total < 0 or world < 1 or not (0 <= rank < world):
        raise ValueError("bad partition arguments")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owner_of(member, total, world):
    """rank that owns global member `member` under block_partition, and its local index"""
    for r in range(world):
        lo, hi = block_partition(total, world, r)
        if lo <= member < hi:
            return r, member - lo
    raise ValueError("member out of range")


def moments_to_mean_spread(s, s2, n):
    """ensemble mean and (population) standard deviation from sum, sum of squares, count"""
    mean = s / n
    var = s2 / n - mean * mean
    return mean, (var.clamp_min(0.0) if hasattr(var, "clamp_min") else np.maximum(var, 0.0)) ** 0.5


def allreduce_moments(s, s2, count, group=None):
    """all-reduce (sum) the partial moments over the ranks; tensors are reduced in place.
    Returns the global member count.  Without an initialised process group this is a no-op."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        # one flat buffer -> one collective (latency-bound: 2 x 1.5 MB at T30)
        flat = torch.cat([s.reshape(-1), s2.reshape(-1), torch.tensor([float(count)], dtype=s.dtype, device=s.device)])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        n = s.numel()
        s.copy_(flat[:n].view_as(s))
        s2.copy_(flat[n:2 * n].view_as(s2))
        count = int(round(float(flat[-1].item())))
    return count


class Ensemble:
    """`total_members` members over the ranks of the default process group; this rank's block
    is one speedy context on `device`."""

    def __init__(self, pkg, total_members, device=0, trunc=30, sppt_on=1, seed=0, rank=0, world=1, precision=0, nsteps=0):
        self.rank, self.world, self.total = rank, world, total_members
        self.lo, self.hi = block_partition(total_members, world, rank)
        if self.hi == self.lo:
            raise ValueError("more ranks than members")
        self.ctx = pkg.Speedy(trunc=trunc, nmembers=self.hi - self.lo, device=device, sppt_on=sppt_on, seed=seed,
                              member_offset=self.lo, precision=precision, nsteps=nsteps)
        self.device = device

    def model_init(self, bc_path, *date):
        self.ctx.model_init(bc_path, *date)

    def run_steps(self, n):
        return self.ctx.run_steps(n)

    def local_moments(self):
        """device tensors (41, il, ix): sum and sum of squares over the resident members"""
        import ctypes
        import torch
        c = self.ctx
        dev = torch.device("cuda", self.device)
        s = torch.empty((N_OUT_LEVELS, c.il, c.ix), dtype=torch.float64, device=dev)
        s2 = torch.empty_like(s)
        torch.cuda.current_stream(dev).synchronize()
        rc = c.L.speedy_ensemble_sums_dev(c.h, ctypes.c_void_p(s.data_ptr()), ctypes.c_void_p(s2.data_ptr()))
        if rc:
            raise RuntimeError(c.L.speedy_last_error().decode())
        c.synchronize()
        return s, s2

    def mean_spread(self, group=None):
        """ensemble mean and spread of the 41 output levels over ALL ranks' members"""
        s, s2 = self.local_moments()
        n = allreduce_moments(s, s2, self.hi - self.lo, group)
        return moments_to_mean_spread(s, s2, n)

    def close(self):
        self.ctx.close()
