"""Sharding of cells over GPUs and the one collective of the path: basin aggregates.

Cells (catchments / raster pixels) never read each other -- the only cross-cell operations in the
reference are the ``np.sum`` sites of the diagnostic integrals (reference ``bmi_topoflow_glacier.py:567-568``,
``:613-614``, ``:623-624``, ``:1486-1494``) and the driver-side ``* da_m2`` (``examples/run_topoflow_glacier.py:115``).
So every rank owns a contiguous block of cells with no halo, runs the fused kernel on it, and the per-basin,
area-weighted sums ``[T_chunk, n_basin, 3]`` are combined with ONE ``all_reduce(SUM, float64)`` per chunk.
``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU tests) carries it.
"""

from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

__all__ = ["shard_bounds", "shard_sizes", "BasinAggregates", "basin_sums_host", "dist_info"]

AGG_NAMES = ("runoff_m3s", "swe_m3", "iwe_m3")  # sum(M_total*da_m2), sum(h_swe*da_m2), sum(h_iwe*da_m2)


def shard_sizes(n_cells: int, world: int, align: int = 128) -> list[int]:
    """Cells per rank: contiguous blocks, every block but the last a multiple of ``align`` cells."""
    if world < 1 or n_cells < 0:
        raise ValueError("bad world size / cell count")
    per = -(-n_cells // world)
    per = -(-per // align) * align
    sizes, left = [], n_cells
    for _ in range(world):
        k = min(per, left)
        sizes.append(k)
        left -= k
    return sizes


def shard_bounds(n_cells: int, world: int, rank: int, align: int = 128) -> Tuple[int, int]:
    """Half-open cell range ``[lo, hi)`` owned by ``rank``."""
    sizes = shard_sizes(n_cells, world, align)
    lo = int(np.sum(sizes[:rank]))
    return lo, lo + sizes[rank]


def dist_info() -> Tuple[int, int]:
    """(rank, world) of the default process group, (0, 1) when not initialised."""
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:  # noqa: BLE001
        pass
    return 0, 1


def basin_sums_host(values: np.ndarray, da_m2: np.ndarray, basin_id: np.ndarray, n_basin: int) -> np.ndarray:
    """Host statement of one aggregate: ``out[b] = sum(values[i] * da_m2[i] for basin_id[i] == b)``."""
    return np.bincount(basin_id, weights=np.asarray(values, dtype=np.float64) * da_m2, minlength=n_basin)


class BasinAggregates:
    """Per-rank partial sums -> global sums.

    Default: ``buffer`` is the ``[T_chunk, n_basin, 3]`` float64 tensor the kernel accumulates into with
    floating-point atomics; ``reduce`` makes it global (in place) with one ``all_reduce(SUM)``.  Reproducible to
    rounding (~1e-15), not bit for bit: the order of the atomics is not fixed.

    ``exponents=(E0, E1, E2)`` (``MeltEngine.agg_exponents()``, which agrees them across ranks with one
    ``all_reduce(MAX)`` -- every rank must scale with the same exponents) selects ORDER-INDEPENDENT sums instead: the kernel
    splits every contribution into two fixed-point int64 words and adds them with integer atomics
    (``TFG_OPT_EXACT_AGG``); ``accumulator`` is that int64 tensor (pass it to ``MeltEngine.run(basin_agg=...)``),
    ``reduce`` all-reduces the integers -- exactly -- and decodes them into ``buffer``.  The result is bit-identical
    for any GPU count (shards are 128-cell aligned), launch split or scheduling order.  Works on CUDA tensors with
    NCCL and on CPU tensors with gloo.
    """

    def __init__(self, chunk_steps: int, n_basin: int, device=None, group=None, exponents=None):
        import torch

        self.torch = torch
        self.group = group
        self.n_basin = int(n_basin)
        self.chunk_steps = int(chunk_steps)
        self.buffer = torch.zeros(self.chunk_steps, self.n_basin, 3, dtype=torch.float64, device=device)
        self.exponents = None if exponents is None else tuple(int(e) for e in exponents)
        self.accumulator = None
        if self.exponents is not None:
            self.accumulator = torch.zeros(self.chunk_steps * self.n_basin * 3 * 2 + 1, dtype=torch.int64, device=device)
            e = torch.tensor(self.exponents, dtype=torch.float64, device=device)
            self._scale_hi, self._scale_lo = torch.exp2(e - 40.0), torch.exp2(e - 82.0)

    @property
    def target(self):
        """What ``MeltEngine.run(basin_agg=...)`` should accumulate into."""
        return self.buffer if self.accumulator is None else self.accumulator

    @property
    def n_left_out(self) -> int:
        """Exact mode: contributions the kernel left out (non-finite, or beyond 2^E_q); 0 for sane inputs."""
        return 0 if self.accumulator is None else int(self.accumulator[-1].item())

    def zero(self):
        self.buffer.zero_()
        if self.accumulator is not None:
            self.accumulator.zero_()
        return self.target

    def decode(self):
        """int64 {hi, lo} words -> float64 ``buffer`` (value = hi * 2^(E-40) + lo * 2^(E-82)); a pure function of
        the integer sums, hence as reproducible as they are."""
        w = self.accumulator[:-1].view(self.chunk_steps, self.n_basin, 3, 2).to(self.torch.float64)
        self.torch.add(w[..., 0] * self._scale_hi, w[..., 1] * self._scale_lo, out=self.buffer)
        return self.buffer

    def reduce(self, async_op: bool = False):
        import torch.distributed as dist

        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
        if self.accumulator is not None:
            if multi:
                dist.all_reduce(self.accumulator, op=dist.ReduceOp.SUM, group=self.group)
            self.decode()
            return None
        if multi:
            return dist.all_reduce(self.buffer, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        return None

    def basin_area(self, da_m2, basin_id) -> "object":
        """Static per-basin area (all ranks): ``sum(da_m2)`` by basin."""
        import torch.distributed as dist

        torch = self.torch
        area = torch.zeros(self.n_basin, dtype=torch.float64, device=da_m2.device)
        area.index_add_(0, basin_id.to(torch.int64), da_m2.to(torch.float64))
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(area, op=dist.ReduceOp.SUM, group=self.group)
        return area
