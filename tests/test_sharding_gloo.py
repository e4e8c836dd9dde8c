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


def _worker_ranged(rank: int, world: int, port: int, out_dir: str):
    """Range-sharded front end on the host side: every rank rasterises ITS triangle range (with the oracle, global
    primitive numbers), the ranks' (depth, winner) planes become keys, the per-pixel max over ranks must be the frame the
    oracle renders from the whole mesh -- depth ties between ranges included."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import softrender_b200 as sr
        from softrender_b200 import scenes, sharding
        import helpers as H
        import oracle_binding as ob

        w, h = 160, 120
        rng = np.random.default_rng(11)
        n = 900
        verts = H.random_screen_triangles(rng, n, w, h, integer_depth=True, max_size=25.0)
        idx = np.arange(3 * n, dtype=np.uint32)
        u = scenes.suzanne_uniforms(w, h)

        def render(i0, i1):
            ofb = ob.OracleFramebuffer(w, h)
            ofb.clear(H.CLEAR)
            od = ob.OracleDraw(sr.TRIANGLE, idx[3 * i0:3 * i1])
            od.set_vertices(verts, 1)
            od.fragment_run(ofb, sr.FS_FLAT, u)
            win = np.where(ofb.winner > 0, ofb.winner + np.uint32(i0), np.uint32(0))  # global primitive numbers
            return ofb.depth.copy(), win

        t0, t1 = sharding.triangle_range(n, rank, world)
        d, wv = render(t0, t1)
        mine = sharding.depth_keys(d, wv)
        parts = [torch.zeros(w * h, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(mine.view(np.int64)))
        merged = sharding.merge_keys([p.numpy().view(np.uint64) for p in parts])
        if rank == 0:
            fd, fw = render(0, n)
            assert np.array_equal(merged, sharding.depth_keys(fd, fw)), "max-merge of the ranks' keys != single-rank keys"
            ranges = [sharding.triangle_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            for k0, k in [(1, 1), (2, 1), (0, 1), (1, 3)]:
                t = sharding.owner_table(world, k0, k)
                assert len(t) == k0 + k * (world - 1) and t.count(0) == k0 and all(t.count(r) == k for r in range(1, world))
            open(os.path.join(out_dir, "ok_ranged"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_range_sharded_key_merge(tmp_path):
    world = 2
    mp.spawn(_worker_ranged, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok_ranged").exists()


def test_two_rank_tile_composite_and_timing(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok").exists()
