"""Band sharding across GPUs: one process per GPU (torch.distributed, NCCL over NVLink).

The forward is independent per band (each band needs only the K maps, its wavelength window of the
OTF / templates and its own tables, and produces its own slice of y), so bands are dealt to ranks
with no data-path collective.  The adjoint's T^T C^T is linear and wavelength-local, so every rank
finishes its partial [K, N, N] map gradient locally and the only exchange per CG iteration is one
all-reduce(sum) of that array (12 MB at K=6, N=501 fp64); the CG scalars are then computed
redundantly on identical data, so they need no collective.  
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import numpy as np


def band_cost(n_wave: int, n_det: int, nb: int, n_pointing: int, n_slit: int, na: int, n_pix: int,
              local_a: int, local_b: int, bytes_per_real: int = 8) -> float:
    """Rough seconds-like cost of one forward+adjoint of a band: FFT/OTF streams (HBM-bound) plus the
    spectral contraction (FMA-bound).  Only ratios matter."""
    nf = n_pix * (n_pix // 2 + 1)
    stream_bytes = n_wave * (6 * nf * 2 * bytes_per_real + 6 * n_pix * n_pix * bytes_per_real
                             + 2 * n_pointing * local_a * local_b * bytes_per_real)
    flops = 4.0 * n_det * n_wave * nb * n_pointing * n_slit * na
    return stream_bytes / 3.0e12 + flops / 2.0e13


def partition_bands(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Longest-processing-time assignment of bands to ranks (deterministic)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world_size
    parts: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        parts[r].append(i)
        loads[r] += costs[i]
    return [sorted(p) for p in parts]


class Comm:
    """Thin wrapper over torch.distributed for the one collective the path needs."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world_size = dist.get_world_size(group)

    def allreduce_sum(self, tensor):
        if self.world_size > 1:
            self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM, group=self.group)
        return tensor

    def allreduce_max(self, tensor):
        if self.world_size > 1:
            self.dist.all_reduce(tensor, op=self.dist.ReduceOp.MAX, group=self.group)
        return tensor

    def barrier(self):
        if self.world_size > 1:
            self.dist.barrier(group=self.group)


def init_from_env(backend: Optional[str] = None) -> Optional[Comm]:
    """Initialise torch.distributed from torchrun's environment; None when single-process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend=backend)
    return Comm()


def local_band_indices(costs: Sequence[float], comm: Optional[Comm]) -> List[int]:
    if comm is None:
        return list(range(len(costs)))
    return partition_bands(costs, comm.world_size)[comm.rank]
