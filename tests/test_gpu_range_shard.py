"""Tile-sharded frames with a range-sharded front end (sr_shard, include/softrender_b200.h), exercised on ONE device:
the ranks of the group are contexts (= CUDA streams) of this process that exchange their keys through directly mapped
memory exactly as ranks on different GPUs do over NVLink (the kernels, the progress words and the merge are the same
code; only the pointers come from sr_shard_connect_local instead of CUDA IPC).  The composited frame must be
bit-identical to the single-context frame, which the other parity tests compare with the oracle
(the reference's tile-parallel loop: src/pipeline/stages/fragment.rs:240-253)."""
import os

import numpy as np
import pytest

import softrender_b200 as sr
from softrender_b200 import scenes

import helpers as H
import oracle_binding as ob

pytestmark = pytest.mark.gpu
# compute-sanitizer serialises kernels: the ranks of a single-process group cannot overlap, every cross-rank wait times out
# (SR_SHARD_TIMEOUT_MS keeps that short) and the frames are incomplete.  The run is then only a memory / race check of the
# kernels; the result assertions are skipped.
SERIALISED = bool(os.environ.get("SR_UNDER_SANITIZER"))


@pytest.fixture(scope="module")
def P():
    from softrender_b200 import pipeline
    return pipeline


def _scene(seed, w, h):
    """> 65536 triangles (the range path's threshold): mostly pixel-sized with exact depth ties between primitives of
    different ranks' ranges, plus mid-size and big ones that go through the per-tile lists (PHASE 1 sweep)."""
    rng = np.random.default_rng(seed)
    verts = np.concatenate([H.random_screen_triangles(rng, 90_000, w, h, max_size=2.5, integer_depth=True),
                            H.random_screen_triangles(rng, 600, w, h, max_size=60.0, integer_depth=True),
                            H.random_screen_triangles(rng, 40, w, h, integer_depth=True),
                            H.random_screen_triangles(rng, 30_000, w, h, max_size=2.5)])
    return verts, np.arange(verts.shape[0], dtype=np.uint32)


def _single(P, ctx, w, h, verts, idx, u):
    fb = P.RenderBuffer.with_dimensions(ctx, w, h)
    fb.clear(H.CLEAR)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    pipe.draw_from_vertices(sr.TRIANGLE, verts, idx, 1).run(sr.FS_FLAT)
    out = fb.download()
    pipe.destroy()
    fb.destroy()
    return out


@pytest.mark.parametrize("world,lanes", [(2, 1), (3, 2), (4, 1)])
def test_range_sharded_frame_is_bit_identical(P, ctx, world, lanes):
    w, h = 640, 360
    verts, idx = _scene(100 + world, w, h)
    u = scenes.suzanne_uniforms(w, h)
    expect = _single(P, ctx, w, h, verts, idx, u)
    # against the oracle as well (coverage, depth and -- flat shader -- colour bit-exact)
    ofb = ob.OracleFramebuffer(w, h)
    ofb.clear(H.CLEAR)
    od = ob.OracleDraw(sr.TRIANGLE, idx)
    od.set_vertices(verts, 1)
    od.fragment_run(ofb, sr.FS_FLAT, u)
    H.compare_framebuffers(expect, ofb, exact_color=True, what="single-context frame")

    ctxs = [[P.Context(0) for _ in range(lanes)] for _ in range(world)]  # [rank][lane]
    groups, fbs, pipes = [], [], []
    try:
        for r in range(world):
            for c in ctxs[r]:
                c.set_tile_shard(r, world)
            groups.append(P.ShardGroup(ctxs[r][0], w, h, lanes))
        for g in groups:
            g.connect_local(groups)
        for r in range(world):
            for lane, c in enumerate(ctxs[r]):
                groups[r].attach(c, lane)
        # one target per lane, owned by rank 0; the other ranks draw into it through an alias (NVLink peer pointer on a real box)
        targets = [P.RenderBuffer.with_dimensions(ctxs[0][lane], w, h) for lane in range(lanes)]
        fbs = [[targets[lane] if r == 0 else targets[lane].alias(ctxs[r][lane]) for lane in range(lanes)] for r in range(world)]
        pipes = [[P.Pipeline.from_framebuffer(fbs[r][lane], u) for lane in range(lanes)] for r in range(world)]
        before = ctxs[1][0].launch_count()
        nframes = 5
        for f in range(nframes):  # several frames per lane: the progress words carry the frame number
            lane = f % lanes
            for r in range(world):
                fbs[r][lane].clear(H.CLEAR)
                pipes[r][lane].draw_from_vertices(sr.TRIANGLE, verts, idx, 1).run(sr.FS_FLAT)
        for r in range(world):
            for c in ctxs[r]:
                c.synchronize()
        assert ctxs[1][0].launch_count() > before
        if not SERIALISED:
            assert [g.status() for g in groups] == [0] * world, "a rank gave up waiting for a peer"
            for lane in range(lanes):
                H.assert_bits_equal(targets[lane].download(), expect, f"range-sharded frame, world {world}, lane {lane}")
    finally:
        for r in range(world):
            for c in ctxs[r]:
                P.ShardGroup.detach(c)
        for row in pipes:
            for p_ in row:
                p_.destroy()
        for r in range(len(fbs) - 1, -1, -1):
            for fb in fbs[r]:
                fb.destroy()
        for g in groups:
            g.destroy()
        for row in ctxs:
            for c in row:
                c.close()


@pytest.mark.parametrize("world,fused,chunks", [(2, False, True), (3, False, True), (4, False, True), (2, False, False), (3, False, False), (2, True, False)])
def test_range_sharded_mesh_with_lazy_vertex_stage(P, ctx, world, fused, chunks, monkeypatch):
    """A real mesh through run_to_fragment: with a shard group attached the vertex stage is recorded and each rank shades only
    the vertices it needs.  `chunks` = the chunk-culled front end (a coherent mesh: every rank rasterises the 1024-triangle
    chunks that can reach ITS tile rows, all depth layers of a row on one rank, and resolves those rows); otherwise triangle
    ranges (vertex range of the rank's triangles + the vertices the winners of its tiles reference).  `fused` = the variant
    that merges the peers' keys inside the resolve.  Frame bit-identical to the single-context frame; a later consumer of the
    draw's vertices still sees the whole mesh."""
    w, h = 640, 360
    mesh = scenes.make_grid(200, 170, 4, seed=0x5EED0003)
    assert mesh.ntris >= 65536
    u = scenes.grid_uniforms(w, h)
    vp = scenes.Viewport.new(w, h, 0.1, 100.0)

    def draw(pipe, gmesh):
        return pipe.render_mesh(sr.TRIANGLE, gmesh).run_to_fragment(vp, sr.VS_SUZANNE)

    fb1 = P.RenderBuffer.with_dimensions(ctx, w, h)
    fb1.clear(H.CLEAR)
    p1, m1 = P.Pipeline.from_framebuffer(fb1, u), P.Mesh(ctx, mesh)
    d1 = draw(p1, m1)
    verts_expect = d1.download(0)
    d1.run(sr.FS_SUZANNE)
    expect = fb1.download()
    for x in (p1, m1, fb1):
        x.destroy()

    monkeypatch.setenv("SR_SHARD_SHARES", "1,1,16" if fused else "2,1,0")  # (read by sr_shard_create: tuning knob)
    if chunks:
        monkeypatch.setenv("SR_SHARD_CHUNKS", "1")  # opt-in mode (read at every draw)
    else:
        monkeypatch.delenv("SR_SHARD_CHUNKS", raising=False)
    ctxs = [P.Context(0) for _ in range(world)]
    groups = []
    for r, c in enumerate(ctxs):
        c.set_tile_shard(r, world)
        groups.append(P.ShardGroup(c, w, h, 1))
    for g in groups:
        g.connect_local(groups)
    for g, c in zip(groups, ctxs):
        g.attach(c, 0)
    target = P.RenderBuffer.with_dimensions(ctxs[0], w, h)
    fbs = [target] + [target.alias(c) for c in ctxs[1:]]
    pipes = [P.Pipeline.from_framebuffer(fb, u) for fb in fbs]
    meshes = [P.Mesh(c, mesh) for c in ctxs]
    last = None
    for f in range(3):
        stages = []
        for r in range(world):
            fbs[r].clear(H.CLEAR)
            stages.append(draw(pipes[r], meshes[r]))
        for st in stages:
            dup = st.duplicate()
            st.run(sr.FS_SUZANNE)
            last = dup
    for c in ctxs:
        c.synchronize()
    if not SERIALISED:
        assert [g.status() for g in groups] == [0] * world
        H.assert_bits_equal(target.download(), expect, f"range-sharded mesh frame, world {world}")
    # the duplicate of the last rank's draw (lazy vertex stage, partly shaded by its frame) yields the whole mesh on demand
    H.assert_bits_equal(last.download(0), verts_expect, "vertices after a lazy vertex stage")
    for c in ctxs:
        P.ShardGroup.detach(c)
    for x in pipes + meshes:
        x.destroy()
    for fb in fbs[::-1]:
        fb.destroy()
    for g in groups:
        g.destroy()
    for c in ctxs:
        c.close()


def test_range_shard_falls_back_for_ineligible_draws(P, ctx):
    """Draws the range path does not take (few triangles, draws onto existing contents, blending) use plain sort-first
    tile sharding on the same contexts and still composite to the single-context frame."""
    w, h, world = 320, 200, 2
    rng = np.random.default_rng(7)
    verts = H.random_screen_triangles(rng, 3000, w, h, max_size=30.0)
    idx = np.arange(verts.shape[0], dtype=np.uint32)
    u = scenes.suzanne_uniforms(w, h)
    expect = _single(P, ctx, w, h, verts, idx, u)
    ctxs = [P.Context(0) for _ in range(world)]
    groups = []
    for r, c in enumerate(ctxs):
        c.set_tile_shard(r, world)
        groups.append(P.ShardGroup(c, w, h, 1))
    for g in groups:
        g.connect_local(groups)
    for g, c in zip(groups, ctxs):
        g.attach(c, 0)
    target = P.RenderBuffer.with_dimensions(ctxs[0], w, h)
    fbs = [target, target.alias(ctxs[1])]
    pipes = [P.Pipeline.from_framebuffer(fb, u) for fb in fbs]
    for fb, p_ in zip(fbs, pipes):
        fb.clear(H.CLEAR)
        p_.draw_from_vertices(sr.TRIANGLE, verts, idx, 1).run(sr.FS_FLAT)
    for c in ctxs:
        c.synchronize()
    H.assert_bits_equal(target.download(), expect, "fallback frame")
    for c in ctxs:
        P.ShardGroup.detach(c)
    for p_ in pipes:
        p_.destroy()
    fbs[1].destroy()
    fbs[0].destroy()
    for g in groups:
        g.destroy()
    for c in ctxs:
        c.close()


def test_shard_group_argument_checks(P, ctx):
    with pytest.raises(Exception):
        P.ShardGroup(ctx, 64, 64, 1)  # no tile shard set on the context
    c = P.Context(0)
    c.set_tile_shard(1, 2)
    g = P.ShardGroup(c, 64, 64, 2)
    with pytest.raises(Exception):
        g.attach(c, 2)  # lane out of range
    with pytest.raises(Exception):
        g.connect([b"\0" * 64])  # wrong number of handles
    assert g.status() == 0
    g.destroy()
    c.close()
