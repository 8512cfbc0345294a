"""Sharded MSM across the GPUs of one box: one process per GPU, torch.distributed for the plumbing.

The MSM sum_i s_i P_i shards by index range: rank r of R owns pairs
[r N / R, (r+1) N / R) -- its slice of the SRS stays resident on its GPU -- and produces ONE
partial G1 point (18 u64, un-normalised projective).  The only exchange on the data path is an
all-gather of R x 144 bytes (NCCL over NVLink / NVSwitch on GPUs, gloo in the CPU tests); every
rank then adds the R partials and normalises.  The message is latency-bound (~1 kB), so there is
nothing for a fused compute+collective kernel to overlap; NTTs do not shard ("replicas only":
independent polynomials of a round are dealt to different GPUs).

`compute_partial` and `sum_points` are injectable so the host-side logic (bounds, gather order,
result replication) is testable on CPU with world_size 2 over gloo.
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np


def shard_bounds(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """index range [lo, hi) of rank `rank`; ranges tile [0, n) in rank order, sizes differ by <= 1"""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    lo = n * rank // world_size
    hi = n * (rank + 1) // world_size
    return lo, hi


def slice_of_prefix(lo: int, hi: int, length: int) -> Tuple[int, int]:
    """the part of this rank's index range [lo, hi) that lies inside the first `length` indices, as
    (offset, count); count == 0 when the prefix ends before the rank's range starts"""
    if length <= lo:
        return lo, 0
    return lo, min(hi, length) - lo


def all_gather_points(partial, group=None):
    """gather one 18-limb point per rank -> tensor [world, 18] in rank order (int64 carrier dtype)"""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    partial = partial.reshape(18).contiguous()
    out = torch.empty((world, 18), dtype=partial.dtype, device=partial.device)
    dist.all_gather_into_tensor(out.view(-1), partial, group=group)
    return out


def sharded_msm(partial, sum_points: Callable, group=None):
    """partial: this rank's partial sum as an int64[18] tensor (bit pattern of the u64 limbs), on the
    device the process group communicates on.  Returns sum_points(gathered [world, 18])."""
    gathered = all_gather_points(partial, group)
    return sum_points(gathered)


class ShardedCommitter:
    """Per-rank state of a sharded KZG commitment: SRS slice on this GPU + reusable device buffers."""

    def __init__(self, pkg, ctx, n_total: int, tau: int, rank: int, world_size: int, group=None, precompute=0,
                 n_global=None):
        import torch

        self.pkg, self.ctx, self.group = pkg, ctx, group
        self.rank, self.world = rank, world_size
        self.n_total = n_total
        self.lo, self.hi = shard_bounds(n_total, world_size, rank)
        self.setup = pkg.Setup.generate_srs(self.hi - self.lo, tau, ctx, first=self.lo)
        if precompute is not None:  # window bits, 0 = auto
            self.precompute(precompute)
        self.device = torch.device("cuda", ctx.device)
        self.d_partial = torch.zeros(18, dtype=torch.int64, device=self.device)
        self.d_out = torch.zeros(18, dtype=torch.int64, device=self.device)
        self.d_gather = torch.zeros((world_size, 18), dtype=torch.int64, device=self.device)

    def precompute(self, window_bits: int = 0):
        """window levels next to this rank's SRS slice (Setup.precompute); 0 = the library's choice for the slice"""
        self.setup.precompute(window_bits)
        return self

    def commit_device(self, d_scalars) -> "torch.Tensor":
        """d_scalars: int64 tensor holding this rank's (hi - lo) x 4 u64 Montgomery limbs in HBM.
        Returns the normalised commitment (int64[18] on the device), identical on every rank."""
        import torch
        import torch.distributed as dist

        self.ctx.bind_torch_stream(torch)
        lib, h = self.ctx.lib, self.ctx.handle
        n = self.hi - self.lo
        if self.world == 1:
            self.ctx.check(lib.bpk_msm_g1_dev(h, self.setup.handle, 0, d_scalars.data_ptr(), n, 1,
                                              self.d_out.data_ptr()), "bpk_msm_g1_dev")
            return self.d_out
        self.ctx.check(lib.bpk_msm_g1_dev(h, self.setup.handle, 0, d_scalars.data_ptr(), n, 0,
                                          self.d_partial.data_ptr()), "bpk_msm_g1_dev")
        dist.all_gather_into_tensor(self.d_gather.view(-1), self.d_partial, group=self.group)
        self.ctx.check(lib.bpk_g1_sum_dev(h, self.d_gather.data_ptr(), self.world, self.d_out.data_ptr()),
                       "bpk_g1_sum_dev")
        return self.d_out

    def commit_prefix(self, d_coeffs, length: int) -> "torch.Tensor":
        """Commitment to a polynomial of `length` <= n_total coefficients that is REPLICATED in every
        rank's HBM (the device prover keeps identical state on all ranks): each rank multiplies only its own
        index range of the coefficients with its SRS slice, then the partial sums are gathered and added.
        d_coeffs: int64 tensor [>= length, 4].  Returns the normalised commitment, identical on every rank."""
        import torch
        import torch.distributed as dist

        self.ctx.bind_torch_stream(torch)
        lib, h = self.ctx.lib, self.ctx.handle
        first, count = slice_of_prefix(self.lo, self.hi, length)
        src = d_coeffs[first:first + count] if count else d_coeffs[0:1]
        final = 1 if self.world == 1 else 0
        dst = self.d_out if self.world == 1 else self.d_partial
        self.ctx.check(lib.bpk_msm_g1_dev(h, self.setup.handle, 0, src.data_ptr(), count, final, dst.data_ptr()),
                       "bpk_msm_g1_dev")
        if self.world > 1:
            dist.all_gather_into_tensor(self.d_gather.view(-1), self.d_partial, group=self.group)
            self.ctx.check(lib.bpk_g1_sum_dev(h, self.d_gather.data_ptr(), self.world, self.d_out.data_ptr()),
                           "bpk_g1_sum_dev")
        return self.d_out

    def commit_prefix_many(self, items) -> np.ndarray:
        """commit_prefix for several replicated polynomials at once: items = [(d_coeffs, length), ...].
        One bpk_msm_g1_dev_batch call (the MSMs overlap on separate streams), ONE all-gather of
        len(items) x 144 bytes per rank, then one sum per commitment.  Returns uint64[len(items), 18] on the host."""
        import ctypes

        import torch
        import torch.distributed as dist

        self.ctx.bind_torch_stream(torch)
        lib, h = self.ctx.lib, self.ctx.handle
        k = len(items)
        srcs, counts = [], []
        for d_coeffs, length in items:
            first, count = slice_of_prefix(self.lo, self.hi, length)
            srcs.append(d_coeffs[first:first + count] if count else d_coeffs[0:1])
            counts.append(count)
        ptrs = (ctypes.c_void_p * k)(*[t.data_ptr() for t in srcs])
        firsts = (ctypes.c_size_t * k)(*([0] * k))
        lens = (ctypes.c_size_t * k)(*counts)
        partial = torch.empty((k, 18), dtype=torch.int64, device=self.device)
        self.ctx.check(lib.bpk_msm_g1_dev_batch(h, self.setup.handle, k, ptrs, firsts, lens,
                                                1 if self.world == 1 else 0, partial.data_ptr()), "bpk_msm_g1_dev_batch")
        if self.world == 1:
            return partial.cpu().numpy().view(np.uint64)
        gathered = torch.empty((self.world, k, 18), dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(gathered.view(-1), partial.view(-1), group=self.group)
        per_commit = gathered.permute(1, 0, 2).contiguous()          # [k, world, 18]
        out = torch.empty((k, 18), dtype=torch.int64, device=self.device)
        for i in range(k):
            self.ctx.check(lib.bpk_g1_sum_dev(h, per_commit[i].data_ptr(), self.world, out[i].data_ptr()),
                           "bpk_g1_sum_dev")
        return out.cpu().numpy().view(np.uint64)

    def commit_host(self, scalars: np.ndarray, d_staging=None) -> np.ndarray:
        """end-to-end: this rank's scalars in (pinned) host memory -> sharded MSM -> commitment on the host.
        The upload is pipelined with the accumulation inside bpk_msm_g1_from_host (include/bpk.h)."""
        import torch
        import torch.distributed as dist

        self.ctx.bind_torch_stream(torch)
        lib, h = self.ctx.lib, self.ctx.handle
        sc = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        n = sc.shape[0]
        final = 1 if self.world == 1 else 0
        dst = self.d_out if self.world == 1 else self.d_partial
        self.ctx.check(lib.bpk_msm_g1_from_host(h, self.setup.handle, 0, sc.ctypes.data, n, final, dst.data_ptr()),
                       "bpk_msm_g1_from_host")
        if self.world > 1:
            dist.all_gather_into_tensor(self.d_gather.view(-1), self.d_partial, group=self.group)
            self.ctx.check(lib.bpk_g1_sum_dev(h, self.d_gather.data_ptr(), self.world, self.d_out.data_ptr()),
                           "bpk_g1_sum_dev")
        return self.d_out.cpu().numpy().view(np.uint64)
