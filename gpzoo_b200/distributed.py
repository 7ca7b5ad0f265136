"""Data-parallel plumbing for the ELBO step: spots are sharded across ranks, the small shared
parameters are replicated, and ONE all-reduce per step sums a flat buffer
    [ dZ | dsigma | dlengthscale | (dgdp) | dmu | dLu | dW | (dW_cf) | ELBO ]
Per-spot parameters (V, GaussianPrior.mean/scale) live on the rank that owns the spot and need no communication.
The KL term is replicated, so each rank weights it by 1/world_size (`model.elbo(..., kl_weight=1/world)`) and the
all-reduced sum counts it exactly once.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the transport.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, world: int, rank: int):
    """Contiguous block of spots owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owner_split(idx: torch.Tensor, n_total: int, world: int, rank: int):
    """Global minibatch indices (utilities.py:605 draws one global idx) -> positions owned by `rank`
    (mask into idx) and the local indices inside this rank's shard."""
    lo, hi = shard_range(n_total, world, rank)
    mask = (idx >= lo) & (idx < hi)
    return mask, idx[mask] - lo


class FlatGradReducer:
    """Packs the gradients of the shared parameters (+ one scalar) into a flat buffer, all-reduces it once and
    scatters the sums back into `.grad`."""

    def __init__(self, params, device=None, dtype=None, group=None):
        self.params = list(params)
        p0 = self.params[0]
        self.numels = [p.numel() for p in self.params]
        self.flat = torch.zeros(sum(self.numels) + 1, dtype=dtype or p0.dtype, device=device or p0.device)
        self.group = group

    @property
    def world(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def pack(self, scalar):
        o = 0
        for p, n in zip(self.params, self.numels):
            if p.grad is None:
                self.flat[o:o + n].zero_()
            else:
                self.flat[o:o + n].copy_(p.grad.reshape(-1))
            o += n
        self.flat[o] = scalar.detach() if torch.is_tensor(scalar) else float(scalar)

    def unpack(self):
        o = 0
        for p, n in zip(self.params, self.numels):
            if p.grad is None:
                p.grad = self.flat[o:o + n].view_as(p).clone()
            else:
                p.grad.copy_(self.flat[o:o + n].view_as(p.grad))
            o += n
        return self.flat[o]

    def all_reduce(self, scalar):
        """Sum gradients and `scalar` (the rank's ELBO share) over all ranks; returns the summed scalar."""
        if self.world == 1:
            return scalar.detach() if torch.is_tensor(scalar) else scalar
        self.pack(scalar)
        dist.all_reduce(self.flat, group=self.group)
        return self.unpack()
