"""Host-side partitioning used by the multi-GPU paths.

* sort-first tile sharding: GPU tile `i` (row-major over the tile grid) belongs to rank `i % world` -- the rule
  `sr_context_set_tile_shard` applies on the device (csrc/sr_raster.cuh: `tile % shard_world == shard_rank`);
* frame batching (frames too small to shard, BASELINE.json config 5): frame `k` belongs to rank `k % world`
  (realtime_example/src/main.rs:90-93 turntable);
* timing: every multi-GPU number is the MAX over ranks of a device time.
"""
from __future__ import annotations

import numpy as np


def tile_grid(width: int, height: int, tile_w: int, tile_h: int):
    return (width + tile_w - 1) // tile_w, (height + tile_h - 1) // tile_h


def tile_owner(tile_index: int, world: int) -> int:
    return tile_index % world


def owned_tiles(width: int, height: int, tile_w: int, tile_h: int, rank: int, world: int) -> np.ndarray:
    ntx, nty = tile_grid(width, height, tile_w, tile_h)
    return np.arange(rank, ntx * nty, world, dtype=np.int64)


def ownership_mask(width: int, height: int, tile_w: int, tile_h: int, rank: int, world: int) -> np.ndarray:
    """bool [height, width]: pixels whose GPU tile this rank rasterises and writes back."""
    ntx, _ = tile_grid(width, height, tile_w, tile_h)
    ys, xs = np.mgrid[0:height, 0:width]
    tile = (ys // tile_h) * ntx + (xs // tile_w)
    return (tile % world) == rank


def frames_for_rank(nframes: int, rank: int, world: int):
    return list(range(rank, nframes, world))


def composite(parts, masks) -> np.ndarray:
    """Composite of disjoint per-rank tile sets: no depth merge is needed because ownership is disjoint."""
    out = np.zeros_like(parts[0])
    for part, mask in zip(parts, masks):
        out[mask] = part[mask]
    return out


def max_over_ranks(value: float) -> float:
    """MAX-reduce a timing over the ranks of the default process group (gloo on CPU, NCCL on GPUs)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
