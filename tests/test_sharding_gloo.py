"""world_size-2 `gloo` tests (CPU) of the host-side multi-GPU logic: tile ownership, disjoint composite,
frame batching and the max-over-ranks timing reduction."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import softrender_b200 as sr
        from softrender_b200 import scenes, sharding
        import helpers as H
        import oracle_binding as ob

        w, h, tw, th = 200, 150, 64, 32
        # every rank holds the whole scene (geometry replicated) and renders the full frame with the oracle;
        # it contributes only the pixels of the tiles it owns
        rng = np.random.default_rng(5)
        n = 400
        verts = H.random_screen_triangles(rng, n, w, h, integer_depth=True)
        idx = np.arange(3 * n, dtype=np.uint32)
        u = scenes.suzanne_uniforms(w, h)
        ofb = ob.OracleFramebuffer(w, h)
        ofb.clear(H.CLEAR)
        od = ob.OracleDraw(sr.TRIANGLE, idx)
        od.set_vertices(verts, 1)
        od.fragment_run(ofb, sr.FS_FLAT, u)
        full = np.concatenate([ofb.color.reshape(h, w, 4), ofb.depth.reshape(h, w, 1)], axis=2).astype(np.float32)

        mask = sharding.ownership_mask(w, h, tw, th, rank, world)
        mine = np.where(mask[:, :, None], full, np.float32(0))
        parts = [torch.zeros(h, w, 5) for _ in range(world)] if rank == 0 else None
        dist.gather(torch.from_numpy(mine), parts, dst=0)
        masks = [torch.zeros(h, w, dtype=torch.uint8) for _ in range(world)] if rank == 0 else None
        dist.gather(torch.from_numpy(mask.astype(np.uint8)), masks, dst=0)

        # timing reduction: the reported time is the slowest rank's
        slowest = sharding.max_over_ranks(1.0 + rank)

        # frame batching: 64 turntable frames, k % world
        frames = sharding.frames_for_rank(64, rank, world)
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([len(frames)], dtype=torch.int64))

        if rank == 0:
            m = [x.numpy().astype(bool) for x in masks]
            total = sum(x.astype(np.int32) for x in m)
            assert (total == 1).all(), "tiles must partition the frame: every pixel owned exactly once"
            comp = sharding.composite([x.numpy() for x in parts], m)
            assert np.array_equal(comp.view(np.uint32), full.view(np.uint32)), "composite of disjoint tiles != full frame"
            assert slowest == float(world)
            assert sum(int(c.item()) for c in counts) == 64
            tiles = np.concatenate([sharding.owned_tiles(w, h, tw, th, r, world) for r in range(world)])
            ntx, nty = sharding.tile_grid(w, h, tw, th)
            assert sorted(tiles.tolist()) == list(range(ntx * nty))
            open(os.path.join(out_dir, "ok"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_tile_composite_and_timing(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok").exists()
