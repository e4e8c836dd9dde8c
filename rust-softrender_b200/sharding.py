"""Host-side partitioning used by the multi-GPU paths.

* sort-first tile sharding: GPU tile `i` (row-major over the tile grid) belongs to rank `i % world` -- the rule
  `sr_context_set_tile_shard` applies on the device (csrc/sr_raster.cuh: `tile % shard_world == shard_rank`);
* frame batching (frames too small to shard, BASELINE.json config 5): frame `k` belongs to rank `k % world`
  (realtime_example/src/main.rs:90-93 turntable);
* range-sharded front end (sr_shard, csrc/sr_api.cu opaque_triangles_ranged): rank r rasterises the triangles
  [r*T/world, (r+1)*T/world) into its own key buffer; a key is (order-preserving depth bits << 32) | (primitive + 1) and
  the owner of a tile takes the per-pixel MAX over the ranks' keys -- the fragment with the largest (depth, submission
  index), which is what one GPU's atomicMax reduces (triangle.rs:120-126: `d >= dt`, later wins ties);
* weighted tile ownership of those frames (`owner_table`): rank 0 holds k0 of every k0 + k*(world-1) consecutive tiles;
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


def triangle_range(ntris: int, rank: int, world: int):
    """[begin, end) of the triangles rank `rank` rasterises in a range-sharded frame (same integer arithmetic as the library)."""
    return ntris * rank // world, ntris * (rank + 1) // world


def depth_keys(depth: np.ndarray, winner: np.ndarray) -> np.ndarray:
    """64-bit visibility keys of a rendered plane: order-preserving map of the f32 depth (csrc/sr_common.cuh sr_depth_key)
    in the high word, 1 + primitive in the low word (0 = nothing drawn: the key of the far depth f32::MIN)."""
    bits = np.ascontiguousarray(depth, np.float32).view(np.uint32).astype(np.uint64)
    neg = (bits & np.uint64(0x80000000)) != 0
    key = np.where(neg, (~bits) & np.uint64(0xFFFFFFFF), bits | np.uint64(0x80000000))
    return (key << np.uint64(32)) | winner.astype(np.uint64)


def merge_keys(keys) -> np.ndarray:
    """Per-pixel max over the ranks' key planes = the key a single GPU reduces with atomicMax."""
    out = keys[0].copy()
    for k in keys[1:]:
        np.maximum(out, k, out=out)
    return out


def owner_table(world: int, k0: int = 1, k: int = 1):
    """Tile -> rank pattern of range-sharded frames (csrc/sr_api.cu shard_set_shares): slot i of the period goes to the rank
    furthest behind its share (rank 0: k0 slots, every other rank: k)."""
    period = k0 + k * (world - 1)
    want = [k0] + [k] * (world - 1)
    got = [0] * world
    table = []
    for i in range(period):
        best, worst = 0, -1e30
        for r in range(world):
            if got[r] >= want[r]:
                continue
            deficit = want[r] * (i + 1) / period - got[r]
            if deficit > worst + 1e-12:
                worst, best = deficit, r
        table.append(best)
        got[best] += 1
    return table


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
