"""Wavelength sharding across GPUs: one process per GPU (torch.distributed, NCCL over NVLink).

The cube wavelength axis is cut into contiguous, cost-balanced ranges (`partition_lambda`).  The FFT,
OTF and slit stages are wavelength-local; a band whose window straddles a cut is computed as partial
sums (its spectral contraction is split along K = (lambda, beta)).  Exchanges per application:
  * between forward and adjoint, the detector blocks of the bands that are SHARED by several ranks are
    summed among exactly those ranks (`BandExchange`: one small all-reduce per shared band on a
    sub-communicator; 2-3 ranks and 6-22 MB each instead of 144 MB across all 8);
  * at the end, one all-reduce(sum) of the [K, N, N] map gradient (12 MB at K=6, N=501 fp64).
The CG scalars are then computed redundantly on identical data, so they need no collective.

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


def partition_lambda(cost_per_lambda: Sequence[float], world_size: int) -> List[tuple]:
    """Cut the cube wavelength axis into `world_size` contiguous ranges of (nearly) equal cumulative
    cost.  The FFT / OTF stages shard perfectly by wavelength; a band whose window straddles a cut is
    simply computed as two partial sums (its spectral contraction is split along K = (lambda, beta))
    that the all-reduce of the detector vector adds up."""
    c = np.asarray(cost_per_lambda, dtype=np.float64)
    n = len(c)
    total = float(c.sum())
    if total <= 0:
        raise ValueError("no wavelength carries any work")
    cum = np.concatenate([[0.0], np.cumsum(c)])
    cuts = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        k = int(np.searchsorted(cum, target, side="left"))
        k = min(max(k, cuts[-1] + 1), n - (world_size - r))
        cuts.append(k)
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(world_size)]


def lambda_costs(n_lambda: int, bands: Sequence[dict], n_pix: int, bytes_per_real: int = 8) -> np.ndarray:
    """Per-wavelength cost model in seconds, calibrated on a B200 with the C4 workload in fp64
    (profiles/r01_app_kernels.md); only the ratios matter for the partition.  `bands`: dicts with
    wave_start, n_wave, n_det, nb, n_pointing, n_slit, na, local_a, local_b and, optionally, hull_rows
    (cube rows the band's field of view touches; default: every row) and srf (default 7).  A covered wavelength costs
      * the two column passes of its 2-D FFT pair, the two template x OTF streams and the memset,
      * the two row passes, proportional to the row pairs in the hull of the bands covering it,
    plus, per band covering it, its share of the gather / scatter index streams (2.2 TB/s effective)
    and of the spectral contraction (56 TFLOP/s fp64-equivalent as an int8-sliced product on tcgen05, digit cutting
    included; 130 in fp32)."""
    scale = (n_pix / 501.0) ** 2 * (bytes_per_real / 8.0) ** 0.8
    # round 2 kernels (warp-per-transform chirp-z): 7.9 ns per 1-D transform -> 2 x 251 column transforms per
    # plane = 4.0 us, 7.9 ns per row pair and direction = 15.8 ns per pair; k1 + k1^T + hull zeroing = 1.5 us
    col_passes, streams, per_pair = 4.0e-6 * scale, 1.5e-6 * scale, 15.8e-9 * (n_pix / 501.0) * (bytes_per_real / 8.0) ** 0.8
    gemm_rate = 5.6e13 if bytes_per_real == 8 else 1.3e14
    cost = np.zeros(n_lambda)
    hull = np.zeros(n_lambda)
    for b in bands:
        sl = slice(b["wave_start"], b["wave_start"] + b["n_wave"])
        hull[sl] = np.maximum(hull[sl], min(n_pix, b.get("hull_rows", n_pix)))
        flops = 4.0 * b["n_det"] * b["nb"] * b["n_pointing"] * b["n_slit"] * b["na"]
        stream = 2.0 * b["n_pointing"] * (b["local_a"] * b["local_b"] + b["n_slit"] * b["na"] * b["nb"]) * bytes_per_real
        # the index streams are bound by LSU wavefronts, which grow with the taps per output (srf rows of 4)
        # faster than with the bytes: empirical exponent from the per-rank kernel times at 8 GPUs
        stream *= (b.get("srf", 7) / 7.0) ** 0.5
        cost[sl] += flops / gemm_rate + stream / 2.2e12
    covered = hull > 0
    cost[covered] += col_passes + streams + per_pair * (hull[covered] / 2.0 + 1.0)
    return cost


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


def band_rank_sets(band_windows: Sequence[tuple], ranges: Sequence[tuple]) -> List[List[int]]:
    """For every band (wavelength window [start, stop)), the ranks whose wavelength range overlaps it."""
    out = []
    for w0, w1 in band_windows:
        out.append([r for r, (l0, l1) in enumerate(ranges) if min(w1, l1) > max(w0, l0)])
    return out


class BandExchange:
    """Sums the detector block of every band among the ranks that computed a share of it.

    Built collectively (every rank of `comm` must construct it with the same arguments, in the same
    order: torch.distributed.new_group is a collective).  `blocks[b]` = (offset, size) of band b in the
    detector vector.  A rank only ends up with complete blocks for the bands it touches -- which is all
    its own adjoint reads."""

    def __init__(self, comm: Comm, band_windows: Sequence[tuple], blocks: Sequence[tuple],
                 my_range: Optional[tuple]):
        import torch
        self.comm = comm
        dist = comm.dist
        mine = torch.tensor(list(my_range) if my_range is not None else [0, 1 << 30], dtype=torch.int64)
        if torch.cuda.is_available() and dist.get_backend(comm.group) == "nccl":
            mine = mine.cuda()
        gathered = [torch.zeros_like(mine) for _ in range(comm.world_size)]
        dist.all_gather(gathered, mine, group=comm.group)
        self.ranges = [tuple(int(v) for v in g.cpu()) for g in gathered]
        self.rank_sets = band_rank_sets(band_windows, self.ranges)
        self.blocks = list(blocks)
        groups = {}
        self.plan = []  # (offset, size, group) for the shared bands this rank takes part in, band order
        for b, ranks in enumerate(self.rank_sets):
            if len(ranks) < 2:
                continue
            key = tuple(ranks)
            if key not in groups:  # every rank creates every group, in band order (collective call)
                groups[key] = None if len(ranks) == comm.world_size else dist.new_group(ranks=list(ranks))
            if comm.rank in ranks:
                self.plan.append((int(blocks[b][0]), int(blocks[b][1]), groups[key] if groups[key] is not None
                                  else comm.group))

    def owned_blocks(self):
        """(offset, size) of the detector blocks this rank accounts for in a sum over all bands: every band is
        owned by the lowest rank that holds a share of it -- after `reduce_shared` that rank's block is complete."""
        return [(int(off), int(size)) for (off, size), ranks in zip(self.blocks, self.rank_sets)
                if ranks and ranks[0] == self.comm.rank]

    def reduce_shared(self, y):
        """In place: y[block of b] = sum over the ranks sharing band b, for the bands of this rank."""
        dist = self.comm.dist
        for off, size, group in self.plan:
            dist.all_reduce(y[off: off + size], op=dist.ReduceOp.SUM, group=group)
        return y

    def bytes_per_application(self, itemsize: int) -> int:
        return sum(size for _, size, _ in self.plan) * itemsize


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
