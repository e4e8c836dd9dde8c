"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same inputs.

Bar (BASELINE.json north_star): tile bins, pixel coverage and depth-test outcomes bit-exact; depth
bit-exact (tolerance stated: 1e-6 relative, expected and asserted: 0); colour within 1/255 per channel
(shaders call powf, libm vs CUDA differ in the last ulps) and bit-exact for the power-free test shaders.
"""
import numpy as np
import pytest

import softrender_b200 as sr
from softrender_b200 import scenes

import helpers as H
import oracle_binding as ob

pytestmark = pytest.mark.gpu

COLOR_TOL = 1.0 / 255.0


@pytest.fixture(scope="module")
def P():
    from softrender_b200 import pipeline
    return pipeline


def make_fb(P, ctx, w, h, stencil=False, winner=True):
    fb = P.RenderBuffer.with_dimensions(ctx, w, h, stencil=stencil)
    if winner:
        fb.enable_winner(True)
    fb.clear(H.CLEAR)
    return fb


def oracle_fb(w, h, stencil=False):
    fb = ob.OracleFramebuffer(w, h, {False: 0, True: 8}.get(stencil, stencil))  # stencil: False, True (u8), 16 or 32
    fb.clear(H.CLEAR)
    return fb


# ------------------------------------------------------------------------------------------------------
# stage by stage on the Suzanne scene (config 1)
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("size", [256, 1024])
def test_suzanne_stages_and_frame(P, ctx, size):
    mesh = H.suzanne_mesh()
    u = scenes.suzanne_uniforms(size, size)
    vp = scenes.Viewport.new(size, size, 0.001, 1000.0)

    od = ob.OracleDraw(sr.TRIANGLE, mesh.indices)
    od.vertex_run(sr.VS_SUZANNE, u, mesh.vertices)
    o_clip_verts = od.data(0)

    fb = make_fb(P, ctx, size, size)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    gmesh = P.Mesh(ctx, mesh)
    gs = pipe.render_mesh(sr.TRIANGLE, gmesh).run(sr.VS_SUZANNE)
    H.assert_bits_equal(gs.download(0), o_clip_verts, "vertex stage")

    # clipper: the kept triangles are the oracle's literal sequence at the reported positions;
    # everything dropped is a triangle with two bit-identical vertices
    od.clip_primitives()
    o_tris = od.data(3).reshape(-1, 3, o_clip_verts.shape[1])
    gs = gs.clip_primitives()
    g_tris = gs.download(3).reshape(-1, 3, o_clip_verts.shape[1])
    seq = gs.download_sequence()
    assert len(o_tris) == 16 * mesh.ntris  # fully inside: 18-vertex polygon -> 16 triangles each
    assert len(g_tris) == len(seq) and np.all(np.diff(seq.astype(np.int64)) > 0)
    H.assert_bits_equal(g_tris, o_tris[seq], "clipped triangles")
    dropped = np.ones(len(o_tris), bool)
    dropped[seq] = False
    d = o_tris[dropped]
    same = (np.all(H.bits(d[:, 0]) == H.bits(d[:, 1]), axis=1) | np.all(H.bits(d[:, 0]) == H.bits(d[:, 2]), axis=1)
            | np.all(H.bits(d[:, 1]) == H.bits(d[:, 2]), axis=1))
    assert same.all()

    od.finish(vp)
    fs = gs.finish(vp)
    H.assert_bits_equal(fs.download(3).reshape(-1, 3, o_clip_verts.shape[1]), od.data(3).reshape(o_tris.shape)[seq], "finish")

    ofb = oracle_fb(size, size)
    od.fragment_run(ofb, sr.FS_SUZANNE, u)
    fs.run(sr.FS_SUZANNE)
    H.compare_framebuffers(fb.download(), ofb, color_tol=COLOR_TOL, what=f"suzanne {size}")
    assert np.array_equal(fb.download_winner(), ofb.winner), "coverage / winning primitive"

    # the no-clip path gives the identical image (mesh fully inside the frustum)
    fb2 = make_fb(P, ctx, size, size)
    pipe2 = P.Pipeline.from_framebuffer(fb2, u)
    pipe2.render_mesh(sr.TRIANGLE, gmesh).run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)
    a, b = fb.download(), fb2.download()
    H.assert_bits_equal(a, b, "clip vs run_to_fragment")
    for x in (pipe, pipe2, gmesh, fb, fb2):
        x.destroy()


def test_suzanne_matches_reference_golden_image(P, ctx):
    """Loose visual golden: examples/suzanne.png, box-downsampled 4x."""
    size = 2000
    mesh = H.suzanne_mesh()
    u = scenes.suzanne_uniforms(size, size)
    vp = scenes.Viewport.new(size, size, 0.001, 1000.0)
    fb = make_fb(P, ctx, size, size, winner=False)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    gmesh = P.Mesh(ctx, mesh)
    pipe.render_mesh(sr.TRIANGLE, gmesh).run(sr.VS_SUZANNE).clip_primitives().finish(vp).run(sr.FS_SUZANNE)
    img = np.clip(fb.download()[:, :3].reshape(size, size, 3), 0, 1)
    img = img.reshape(500, 4, 500, 4, 3).mean(axis=(1, 3))
    gold = np.load(H.GOLDEN + "/suzanne_gold_500.npz")["rgb"].astype(np.float32) / 255.0
    lit_o = np.abs(img - 0.01).max(-1) > 0.02
    lit_g = np.abs(gold - 0.01).max(-1) > 0.02

    def bbox(m):
        ys, xs = np.nonzero(m)
        return np.array([xs.min(), xs.max(), ys.min(), ys.max()])

    assert np.abs(bbox(lit_o) - bbox(lit_g)).max() <= 1
    assert np.abs(bbox(lit_g) - np.array([88, 477, 78, 440])).max() <= 1
    iou = (lit_o & lit_g).sum() / (lit_o | lit_g).sum()
    assert iou >= 0.94
    both = lit_o & lit_g
    assert np.median(np.abs(img - gold)[both]) * 255 <= 2.0
    for x in (pipe, gmesh, fb):
        x.destroy()


# ------------------------------------------------------------------------------------------------------
# rasteriser on injected screen-space triangles (bit-exact colour with the flat shader)
# ------------------------------------------------------------------------------------------------------
def run_both_screen(P, ctx, w, h, verts, indices, *, prim=sr.TRIANGLE, fs=sr.FS_FLAT, cull=None, blend=sr.BLEND_REPLACE,
                    stencil=False, stencil_cfg=None, stencil_value=None, aa=False, gen=None, draws=1, oracle_tile=None,
                    init=None):
    u = scenes.suzanne_uniforms(w, h)
    ofb = oracle_fb(w, h, stencil)
    fb = make_fb(P, ctx, w, h, stencil=stencil)
    if init is not None:
        col, dep, st = init
        ofb.color[:] = col
        ofb.depth[:] = dep
        if stencil:
            ofb.stencil[:] = st
        fb.upload_planes(col, dep, st if stencil else None)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    if stencil_cfg:
        pipe.set_stencil_config(*stencil_cfg)
    for _ in range(draws):
        od = ob.OracleDraw(prim, indices, stencil_value)
        od.set_vertices(verts, 1)
        od.cull = sr.CULL_NONE if cull is None else cull
        od.blend, od.aa, od.tile = blend, aa, oracle_tile
        fsd = pipe.draw_from_vertices(prim, verts, indices, 1, stencil_value)
        if gen:
            for which, gv in gen.items():
                od.set_generated(which, gv)
                fsd.set_generated(which, gv)
        od.fragment_run(ofb, fs, u, *(stencil_cfg or (0, 0)))
        fsd.cull_faces(cull).with_blend(blend).antialiased_lines(aa).run(fs)
    out = fb.download()
    win = fb.download_winner()
    st = fb.download_planes(stencil=True)[2] if stencil else None
    pipe.destroy()
    fb.destroy()
    return out, win, st, ofb


@pytest.mark.parametrize("w,h,n,seed", [(64, 64, 40, 1), (200, 150, 300, 2), (333, 257, 2000, 3), (1000, 700, 5000, 4)])
def test_random_triangles_opaque(P, ctx, w, h, n, seed):
    rng = np.random.default_rng(seed)
    verts = H.random_screen_triangles(rng, n, w, h)
    idx = np.arange(3 * n, dtype=np.uint32)
    out, win, _, ofb = run_both_screen(P, ctx, w, h, verts, idx)
    assert np.array_equal(win, ofb.winner)
    H.compare_framebuffers(out, ofb, exact_color=True, what="random opaque")


def test_depth_ties_later_primitive_wins(P, ctx):
    """`d >= dt`: equal depth goes to the later primitive (triangle.rs:126); shuffled submission order."""
    rng = np.random.default_rng(7)
    w, h, n = 96, 80, 150
    verts = H.random_screen_triangles(rng, n, w, h, integer_depth=True)
    idx = rng.permutation(n).astype(np.uint32)[:, None] * 3 + np.arange(3, dtype=np.uint32)[None, :]
    out, win, _, ofb = run_both_screen(P, ctx, w, h, verts, idx.reshape(-1))
    assert np.array_equal(win, ofb.winner)
    H.compare_framebuffers(out, ofb, exact_color=True, what="depth ties")
    # and the reference's default overlapping 128x128 tiling gives the same image (idempotent state)
    _, _, _, ofb128 = run_both_screen(P, ctx, w, h, verts, idx.reshape(-1), oracle_tile=(16, 16))
    H.assert_bits_equal(ofb128.color, ofb.color, "tiling independence")


def test_tiny_and_subpixel_triangles(P, ctx):
    rng = np.random.default_rng(11)
    w, h, n = 256, 192, 20000
    verts = H.random_screen_triangles(rng, n, w, h, max_size=1.5, margin=0.02)
    idx = np.arange(3 * n, dtype=np.uint32)
    out, win, _, ofb = run_both_screen(P, ctx, w, h, verts, idx)
    assert np.array_equal(win, ofb.winner)
    H.compare_framebuffers(out, ofb, exact_color=True, what="tiny triangles")


def test_inclusive_edges_half_integer_vertices(P, ctx):
    """Pixel centres exactly on an edge are covered; two triangles sharing an edge both cover them and the
    later one wins at equal depth."""
    w = h = 32
    z = -1.0

    def v(x, y, c):
        return [x, y, z, 1.0] + c

    red, green = [1, 0, 0, 1], [0, 1, 0, 1]
    verts = np.array([v(4.5, 4.5, red), v(20.5, 4.5, red), v(4.5, 20.5, red),
                      v(20.5, 4.5, green), v(20.5, 20.5, green), v(4.5, 20.5, green)], np.float32)
    idx = np.arange(6, dtype=np.uint32)
    out, win, _, ofb = run_both_screen(P, ctx, w, h, verts, idx)
    assert np.array_equal(win, ofb.winner)
    H.compare_framebuffers(out, ofb, exact_color=True, what="inclusive edges")
    W = win.reshape(h, w)
    assert W[4, 4] == 1 and W[4, 20] == 2 and W[20, 4] == 2  # corners: shared-edge end points go to the later triangle
    for k in range(5, 20):
        assert W[k, 24 - k] == 2  # the shared diagonal x + y = 24 (pixel centres on it) belongs to triangle 2


@pytest.mark.parametrize("cull", [sr.CLOCKWISE, sr.COUNTER_CLOCKWISE])
def test_cull_faces(P, ctx, cull):
    rng = np.random.default_rng(5)
    w, h, n = 128, 128, 400
    verts = H.random_screen_triangles(rng, n, w, h)
    # a zero-area triangle whose shoelace sum is -0.0 counts as clockwise
    verts[:3, :2] = [[3.0, 3.0], [3.0, 3.0], [3.0, 3.0]]
    idx = np.arange(3 * n, dtype=np.uint32)
    out, win, _, ofb = run_both_screen(P, ctx, w, h, verts, idx, cull=cull)
    assert np.array_equal(win, ofb.winner)
    H.compare_framebuffers(out, ofb, exact_color=True, what="cull")


def test_degenerate_offscreen_and_nonfinite(P, ctx):
    w, h = 64, 48
    rng = np.random.default_rng(9)
    verts = H.random_screen_triangles(rng, 60, w, h)
    t = verts.reshape(-1, 3, 8)
    t[0, :, :2] = [[10, 10], [20, 20], [30, 30]]       # det == 0
    t[1, :, 2] = 0.5                                    # z >= 0: covered but rejected
    t[2, :, 0] += 1e4                                   # entirely off screen (clamp artefact column)
    t[3, 0, 0] = np.inf                                 # +inf clamps to the last pixel
    t[4, 1, 1] = -np.inf
    t[5, 0, 0] = np.nan                                 # NaN: primitive skipped (reference would panic)
    t[6, :, 2] = [-1.0, np.nan, -2.0]                   # NaN depth: z < 0 is false
    idx = np.arange(len(verts), dtype=np.uint32)
    out, win, _, ofb = run_both_screen(P, ctx, w, h, verts, idx)
    assert np.array_equal(win, ofb.winner)
    H.compare_framebuffers(out, ofb, exact_color=True, what="degenerate")


def test_second_draw_uses_existing_depth(P, ctx):
    rng = np.random.default_rng(21)
    w, h, n = 160, 120, 500
    verts = H.random_screen_triangles(rng, n, w, h, integer_depth=True)
    idx = np.arange(3 * n, dtype=np.uint32)
    out, win, _, ofb = run_both_screen(P, ctx, w, h, verts, idx, draws=2)
    assert np.array_equal(win, ofb.winner)
    H.compare_framebuffers(out, ofb, exact_color=True, what="two draws")


@pytest.mark.parametrize("area,min_tris,precheck", [(0, 65536, 1), (16, 0, 1), (16, 0, 0), (16, 0, 2), (64, 0, 0), (64, 0, 3), (4096, 0, 1), (4096, 0, 0), (1, 0, 1)])
def test_opaque_path_split_is_invisible(P, ctx, area, min_tris, precheck):
    """The opaque path sends small triangles through the visibility buffer (k_micro) and the rest through per-tile
    lists; where the split lies (and whether a second draw re-initialises the keys from the stored depth) must not
    change a single bit.  Mixture of sub-pixel, small and large triangles with exact depth ties, drawn twice."""
    rng = np.random.default_rng(77)
    w, h = 333, 190
    tiny = H.random_screen_triangles(rng, 6000, w, h, max_size=2.5, margin=0.02, integer_depth=True)
    mid = H.random_screen_triangles(rng, 800, w, h, max_size=9.0, integer_depth=True)
    big = H.random_screen_triangles(rng, 60, w, h, integer_depth=True)
    verts = np.concatenate([tiny, mid, big])
    n = len(verts) // 3
    idx = (rng.permutation(n).astype(np.uint32)[:, None] * 3 + np.arange(3, dtype=np.uint32)[None, :]).reshape(-1)
    ctx.set_micro(area, min_tris, precheck)
    try:
        out, win, _, ofb = run_both_screen(P, ctx, w, h, verts, idx, draws=2)
    finally:
        ctx.set_micro()
    assert np.array_equal(win, ofb.winner)
    H.compare_framebuffers(out, ofb, exact_color=True, what=f"micro area {area}")


def test_auto_split_many_mid_size_triangles(P, ctx):
    """From 49152 triangles on the library lets the per-triangle front end walk boxes of up to 1024 pixels (the default
    `SR_MICRO_AREA_AUTO` split, sr_micro_area_for): 60k triangles of up to ~30x30 pixels with exact depth ties, drawn
    twice, must still match the oracle bit for bit (winner ids, depth, flat colour)."""
    rng = np.random.default_rng(4242)
    w, h, n = 640, 360, 60000
    verts = H.random_screen_triangles(rng, n, w, h, max_size=30.0, integer_depth=True)
    idx = np.arange(3 * n, dtype=np.uint32)
    ctx.set_micro()  # auto
    out, win, _, ofb = run_both_screen(P, ctx, w, h, verts, idx, draws=2)
    assert np.array_equal(win, ofb.winner)
    H.compare_framebuffers(out, ofb, exact_color=True, what="auto split")


def test_list_arena_overflow_is_replayed(P, ctx):
    """The opaque path enqueues a draw against the current per-tile list capacity without synchronising; when the
    lists do not fit, the tile pass skips itself on the device and is replayed with a larger arena.  Two draws back
    to back with a 4-entry arena must still give the oracle's frame."""
    rng = np.random.default_rng(79)
    w, h, n = 320, 200, 400
    verts = H.random_screen_triangles(rng, n, w, h, integer_depth=True)
    idx = np.arange(3 * n, dtype=np.uint32)
    ctx.set_list_capacity(4)
    try:
        out, win, _, ofb = run_both_screen(P, ctx, w, h, verts, idx, draws=2)
        assert ctx.list_capacity() > 4
    finally:
        ctx.set_list_capacity(1 << 20)
    assert np.array_equal(win, ofb.winner)
    H.compare_framebuffers(out, ofb, exact_color=True, what="replayed tile pass")


def test_wait_for_covers_a_replayed_pass(P, ctx):
    """sr_context_wait_for(waiter, other, 0) promises that the waiter starts after everything enqueued on `other` so far.
    A draw whose tile lists overflowed skipped itself on the device and is only re-enqueued by the next call that settles
    the context: wait_for must settle `other` before it records its event, or a consumer on the waiter's stream would read
    an incomplete frame.  The consumer here is a render-to-texture-style pass on a second context that draws into the
    SAME pixels through an alias: a second opaque draw of nearer triangles, whose result depends on the first frame being
    complete underneath it (it starts from the depth already stored)."""
    rng = np.random.default_rng(81)
    w, h, n = 320, 200, 400
    far = H.random_screen_triangles(rng, n, w, h, integer_depth=True)
    near = H.random_screen_triangles(rng, 60, w, h, integer_depth=True)
    near[:, 2] = -0.5  # in front of everything in `far` (depths -1 .. -5)
    idx_far, idx_near = np.arange(3 * n, dtype=np.uint32), np.arange(3 * 60, dtype=np.uint32)
    u = scenes.suzanne_uniforms(w, h)
    ofb = oracle_fb(w, h)
    for v, ix in ((far, idx_far), (near, idx_near)):
        od = ob.OracleDraw(sr.TRIANGLE, ix)
        od.set_vertices(v, 1)
        od.fragment_run(ofb, sr.FS_FLAT, u)
    other, waiter = P.Context(0), P.Context(0)
    fb = make_fb(P, other, w, h)
    view = fb.alias(waiter)
    p_other, p_waiter = P.Pipeline.from_framebuffer(fb, u), P.Pipeline.from_framebuffer(view, u)
    other.set_list_capacity(4)  # the first draw's lists cannot fit: its tile pass skips itself
    p_other.draw_from_vertices(sr.TRIANGLE, far, idx_far, 1).run(sr.FS_FLAT)
    waiter.wait_for(other, 0)   # must re-enqueue the skipped pass on `other` before the event
    p_waiter.draw_from_vertices(sr.TRIANGLE, near, idx_near, 1).run(sr.FS_FLAT)
    waiter.synchronize()
    other.synchronize()
    assert other.list_capacity() > 4
    H.compare_framebuffers(fb.download(), ofb, exact_color=True, what="consumer after wait_for")
    for x in (p_waiter, p_other, view, fb):
        x.destroy()
    waiter.close()
    other.close()


def test_handles_of_different_contexts_are_rejected(P, ctx):
    """Every object of a draw lives in one context (= one CUDA stream that orders their work and their memory reuse)."""
    other = P.Context(0)
    u = scenes.suzanne_uniforms(64, 64)
    fb = make_fb(P, ctx, 64, 64)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    mesh = P.Mesh(other, H.suzanne_mesh())
    with pytest.raises(Exception, match="different contexts"):
        pipe.render_mesh(sr.TRIANGLE, mesh)
    tex = P.Texture(other, scenes.checker_texture(16, 2))
    with pytest.raises(Exception, match="another context"):
        pipe.bind_texture(tex)
    fb2 = make_fb(P, other, 64, 64)
    with pytest.raises(Exception, match="another context"):
        pipe.with_framebuffer(fb2)
    for x in (tex, mesh, fb2, pipe, fb):
        x.destroy()
    other.close()


@pytest.mark.parametrize("world", [2, 3])
def test_tile_sharding_is_invisible(P, ctx, world):
    """Sort-first sharding: `world` contexts each rasterise the tiles with index % world == rank into ONE shared
    framebuffer; the union must be bit-identical to the unsharded frame."""
    rng = np.random.default_rng(78)
    w, h = 300, 200
    verts = np.concatenate([H.random_screen_triangles(rng, 5000, w, h, max_size=3.0, integer_depth=True),
                            H.random_screen_triangles(rng, 300, w, h, integer_depth=True)])
    n = len(verts) // 3
    idx = np.arange(3 * n, dtype=np.uint32)
    u = scenes.suzanne_uniforms(w, h)
    out1, win1, _, ofb = run_both_screen(P, ctx, w, h, verts, idx)
    # rank 0 owns the framebuffer, the other ranks open it through CUDA IPC-free aliasing (same device): every rank
    # draws with its own shard setting into the same pixels
    fb = make_fb(P, ctx, w, h)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    try:
        for rank in range(world):
            ctx.set_tile_shard(rank, world)
            fb.clear(H.CLEAR)  # lazy clear: every rank produces the clear colour of ITS tiles on chip
            pipe.draw_from_vertices(sr.TRIANGLE, verts, idx, 1).run(sr.FS_FLAT)
    finally:
        ctx.set_tile_shard(0, 1)
    out = fb.download()
    H.assert_bits_equal(out, out1, "sharded frame")
    pipe.destroy()
    fb.destroy()


def test_one_pixel_frames_draw_nothing(P, ctx):
    """1xN framebuffers have no tiles (fragment.rs:188-216): nothing is ever drawn."""
    verts = np.array([[-5, -5, -1, 1, 1, 0, 0, 1], [9, -5, -1, 1, 1, 0, 0, 1], [0, 9, -1, 1, 1, 0, 0, 1]], np.float32)
    out, win, _, ofb = run_both_screen(P, ctx, 1, 1, verts, np.arange(3, dtype=np.uint32))
    H.compare_framebuffers(out, ofb, exact_color=True, what="1x1")
    assert win.max() == 0


# ------------------------------------------------------------------------------------------------------
# bins
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cull", [sr.CULL_NONE, sr.CLOCKWISE])
def test_tile_bins_match_oracle(P, ctx, cull):
    rng = np.random.default_rng(31)
    w, h, n = 700, 450, 3000
    verts = H.random_screen_triangles(rng, n, w, h, max_size=90)
    verts.reshape(-1, 3, 8)[7, 0, 0] = np.nan
    idx = rng.permutation(3 * n).astype(np.uint32)
    u = scenes.suzanne_uniforms(w, h)
    fb = make_fb(P, ctx, w, h)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    fsd = pipe.draw_from_vertices(sr.TRIANGLE, verts, idx, 1).cull_faces(None if cull == 0 else cull)
    g_off, g_ids = fsd.bins()
    od = ob.OracleDraw(sr.TRIANGLE, idx)
    od.set_vertices(verts, 1)
    tw, th = P.tile_size()
    o_off, o_ids = od.bins(w, h, tw, th, cull)
    assert np.array_equal(g_off, o_off)
    assert np.array_equal(g_ids, o_ids)
    pipe.destroy()
    fb.destroy()


def _clamped_boxes(verts, idx, w, h):
    """Per triangle, in f32 as the kernels and triangle.rs:64-78 compute them: frame-clamped bounding box area, the
    candidate-tightening predicate and 'has a NaN coordinate'."""
    p = verts[idx].reshape(-1, 3, verts.shape[1]).astype(np.float32)
    x, y = p[:, :, 0], p[:, :, 1]
    det = (y[:, 1] - y[:, 2]) * (x[:, 0] - x[:, 2]) + (x[:, 2] - x[:, 1]) * (y[:, 0] - y[:, 2])
    xmin, xmax, ymin, ymax = x.min(1), x.max(1), y.min(1), y.max(1)
    tight = (xmax - xmin < np.float32(4.99)) & (ymax - ymin < np.float32(4.99)) & (np.abs(det) >= 1)

    def clamp(v, hi):
        return np.where(v < 0, 0, np.where(v > hi, hi, np.trunc(np.nan_to_num(v)))).astype(np.int64)
    area = (clamp(xmax, w - 1) - clamp(xmin, w - 1) + 1) * (clamp(ymax, h - 1) - clamp(ymin, h - 1) + 1)
    return area, tight, np.isnan(x).any(1) | np.isnan(y).any(1)


@pytest.mark.parametrize("n_small,n_big", [(90_000, 2500), (0, 3000)])
def test_opaque_fast_path_lists_match_oracle_bins(P, ctx, n_small, n_big):
    """SURVEY 8 a7 on the HEADLINE path: the per-tile lists k_micro + k_large_fill (big draws) and k_bin_small (draws of
    <= 8192 triangles) build equal the oracle's bins restricted to the triangles that path sends through lists: those
    whose frame-clamped bounding box exceeds the draw's split area and that the tightened small-triangle path does not
    take (every triangle for a small draw)."""
    rng = np.random.default_rng(33 + n_small)
    w, h = 700, 450
    parts = [H.random_screen_triangles(rng, n_big, w, h, max_size=90)]
    if n_small:
        parts.append(H.random_screen_triangles(rng, n_small, w, h, max_size=2.5))
    verts = np.concatenate(parts)
    n = len(verts) // 3
    idx = rng.permutation(3 * n).astype(np.uint32)
    u = scenes.suzanne_uniforms(w, h)
    fb = make_fb(P, ctx, w, h)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    pipe.draw_from_vertices(sr.TRIANGLE, verts, idx, 1).run(sr.FS_FLAT)
    tw, th = P.tile_size()
    ntiles = ((w + tw - 1) // tw) * ((h + th - 1) // th)
    g_off, g_ids, area_split = ctx.last_opaque_lists(ntiles)
    assert (area_split > 0) == (n > 8192)
    od = ob.OracleDraw(sr.TRIANGLE, idx)
    od.set_vertices(verts, 1)
    o_off, o_ids = od.bins(w, h, tw, th, 0)
    area, tight, nan = _clamped_boxes(verts, idx, w, h)
    listed = ~nan & ((area > area_split) & ~tight if area_split else np.ones(n, bool))
    keep = listed[o_ids]
    want_ids = o_ids[keep]
    tile_of_entry = np.repeat(np.arange(ntiles), np.diff(o_off.astype(np.int64)))
    want_off = np.concatenate([[0], np.cumsum(np.bincount(tile_of_entry[keep], minlength=ntiles))])
    assert np.array_equal(g_ids, want_ids)
    assert np.array_equal(g_off.astype(np.int64), want_off)
    assert len(want_ids) > 1000
    pipe.destroy()
    fb.destroy()


# ------------------------------------------------------------------------------------------------------
# ordered path: blend, stencil, discard, lines, points
# ------------------------------------------------------------------------------------------------------
def test_alpha_over_blend_in_order(P, ctx):
    rng = np.random.default_rng(41)
    w, h, n = 150, 110, 400
    verts = H.random_screen_triangles(rng, n, w, h)
    idx = np.arange(3 * n, dtype=np.uint32)
    out, win, _, ofb = run_both_screen(P, ctx, w, h, verts, idx, blend=sr.BLEND_ALPHA_OVER)
    assert np.array_equal(win, ofb.winner)
    H.compare_framebuffers(out, ofb, exact_color=True, what="alpha over")


@pytest.mark.parametrize("blend", [sr.BLEND_ALPHA_OVER, sr.BLEND_ADDITIVE])
@pytest.mark.parametrize("max_size,n,integer_depth,stencil", [(1.2, 300_000, False, False), (2.5, 100_000, True, False),
                                                               (2.5, 100_000, False, True), (5.0, 40_000, False, False)])
def test_ordered_blend_of_pixel_sized_triangles(P, ctx, blend, max_size, n, integer_depth, stencil):
    """Blended meshes of pixel-sized triangles take k_tile_ordered's lane-per-triangle sweep (one triangle per lane, bidding
    rounds for the pixels two triangles of a run share) and, batch by batch, interleaved or contiguous row ownership.  Random
    positions in random order, several fragments per covered pixel: the triangles of a run overlap all the time, every pixel sees several
    blended fragments, and f32 blending is not associative -- any fragment applied out of submission order shows.  With
    integer depths ties are everywhere (`d >= dt`, triangle.rs:126); a stencil attachment with the default config must not
    change the path's result.  Winner plane, depth and colours bit-exact (power-free shader)."""
    rng = np.random.default_rng(int(max_size * 10) + n + blend)
    w, h = 330, 200  # 6 x 7 tiles, partial tiles at the right and bottom edges
    verts = H.random_screen_triangles(rng, n, w, h, max_size=max_size, integer_depth=integer_depth, margin=0.02)
    idx = np.arange(3 * n, dtype=np.uint32)
    out, win, _, ofb = run_both_screen(P, ctx, w, h, verts, idx, blend=blend, stencil=stencil, stencil_value=1 if stencil else None)
    assert np.array_equal(win, ofb.winner)
    H.compare_framebuffers(out, ofb, exact_color=True, what="ordered blend of pixel-sized triangles")
    assert (ofb.winner != 0).mean() > 0.4


def test_ordered_list_arena_overflow_is_replayed(P, ctx):
    """The ordered path enqueues its bin fill and tile pass against the current capacity of its group-list arenas
    without a host synchronisation; with 2-entry arenas both skip themselves on the device and are replayed with larger
    arenas at the next call.  Blended triangles + lines + points of one draw, then a second blended draw on top."""
    rng = np.random.default_rng(83)
    w, h = 200, 150
    tri = H.random_screen_triangles(rng, 300, w, h)
    lines = H.random_screen_triangles(rng, 40, w, h)[:80]
    pts = H.random_screen_triangles(rng, 100, w, h)[:300]
    ctx.set_list_capacity(2)
    try:
        out, win, _, ofb = run_both_screen(P, ctx, w, h, tri, np.arange(len(tri), dtype=np.uint32), gen={2: lines, 1: pts},
                                           blend=sr.BLEND_ALPHA_OVER, draws=2)
        assert min(ctx.ordered_list_capacity()) > 2
    finally:
        ctx.set_list_capacity(1 << 20)
    assert np.array_equal(win, ofb.winner)
    H.compare_framebuffers(out, ofb, exact_color=True, what="replayed ordered pass")


def test_discarding_shader(P, ctx):
    rng = np.random.default_rng(43)
    w, h, n = 130, 90, 300
    verts = H.random_screen_triangles(rng, n, w, h)
    idx = np.arange(3 * n, dtype=np.uint32)
    out, win, _, ofb = run_both_screen(P, ctx, w, h, verts, idx, fs=sr.FS_DISCARD_CHECKER)
    assert np.array_equal(win, ofb.winner)
    H.compare_framebuffers(out, ofb, exact_color=True, what="discard")


@pytest.mark.parametrize("test,op", [(sr.STENCIL_ALWAYS, sr.STENCIL_INCREMENT_WRAP), (sr.STENCIL_EQUAL, sr.STENCIL_KEEP),
                                     (sr.STENCIL_GREATER_THAN, sr.STENCIL_REPLACE), (sr.STENCIL_NOT_EQUAL, sr.STENCIL_INVERT),
                                     (sr.STENCIL_LESS_THAN_EQ, sr.STENCIL_DECREMENT_SAT), (sr.STENCIL_NEVER, sr.STENCIL_ZERO),
                                     (sr.STENCIL_ALWAYS, sr.STENCIL_INCREMENT_SAT), (sr.STENCIL_ALWAYS, sr.STENCIL_DECREMENT_WRAP)])
def test_stencil(P, ctx, test, op):
    """Stencil test+op run for every pixel of the clamped bbox BEFORE the coverage test, also
    for zero-area triangles."""
    rng = np.random.default_rng(47)
    w, h, n = 100, 84, 120
    verts = H.random_screen_triangles(rng, n, w, h)
    verts.reshape(-1, 3, 8)[0, :, :2] = [[10, 10], [20, 20], [30, 30]]
    idx = np.arange(3 * n, dtype=np.uint32)
    init = (np.tile(np.float32(H.CLEAR), (w * h, 1)), np.full(w * h, np.float32(-3.4028235e38)),
            rng.integers(0, 4, w * h).astype(np.uint8))
    out, win, st, ofb = run_both_screen(P, ctx, w, h, verts, idx, stencil=True, stencil_cfg=(test, op), stencil_value=2, init=init)
    assert np.array_equal(st, ofb.stencil)
    assert np.array_equal(win, ofb.winner)
    H.compare_framebuffers(out, ofb, exact_color=True, what="stencil")


@pytest.mark.parametrize("bits", [16, 32])
@pytest.mark.parametrize("op", [sr.STENCIL_INVERT, sr.STENCIL_REPLACE, sr.STENCIL_INCREMENT_WRAP, sr.STENCIL_DECREMENT_WRAP,
                                sr.STENCIL_INCREMENT_SAT, sr.STENCIL_DECREMENT_SAT])
def test_stencil_wider_types(P, ctx, bits, op):
    """u16 and u32 stencil attachments (the Stencil trait covers every integer width, src/stencil.rs:9-60): initial values
    around 0 and the type's MAX so that wrapping and saturation happen at the type's own bounds; triangles, lines and
    points all run the stencil step."""
    rng = np.random.default_rng(4700 + bits + op)
    w, h, n = 100, 84, 90
    smax = (1 << bits) - 1
    verts = H.random_screen_triangles(rng, n, w, h)
    idx = np.arange(3 * n, dtype=np.uint32)
    lines = np.zeros((40, 8), np.float32)
    lines[:, 0], lines[:, 1] = rng.uniform(0, w, 40), rng.uniform(0, h, 40)
    lines[:, 2], lines[:, 3], lines[:, 4:] = -0.2, 1, rng.uniform(0, 1, (40, 4))
    pts = lines[:15].copy()
    pts[:, :2] = rng.uniform(0, 80, (15, 2))
    edge = np.array([0, 1, 2, smax - 2, smax - 1, smax, 300 & smax, 70000 & smax], np.uint64)
    init = (np.tile(np.float32(H.CLEAR), (w * h, 1)), np.full(w * h, np.float32(-3.4028235e38)),
            edge[rng.integers(0, len(edge), w * h)].astype(np.uint16 if bits == 16 else np.uint32))
    value = smax - 1  # the mesh's stencil value is of the buffer's type
    out, win, st, ofb = run_both_screen(P, ctx, w, h, verts, idx, stencil=bits, stencil_cfg=(sr.STENCIL_LESS_THAN_EQ, op),
                                        stencil_value=value, init=init, gen={2: lines, 1: pts}, draws=2)
    assert st.dtype == ofb.stencil.dtype and st.dtype.itemsize * 8 == bits
    assert np.array_equal(st, ofb.stencil)
    assert np.array_equal(win, ofb.winner)
    H.compare_framebuffers(out, ofb, exact_color=True, what=f"u{bits} stencil")
    assert (ofb.stencil != init[2]).any()


def test_texture_buffer_storage(P, ctx):
    """RGBAf32TextureBuffer storage (declare_texture_buffer!, src/framebuffer/texturebuffer.rs:72-210): colour and depth live in
    planes of their own.  The opaque path (big and small draws, a second draw onto existing contents), the ordered path
    (alpha_over + stencil) and the accessors must give exactly what the AoS RenderBuffer gives, i.e. the oracle's frame; the
    colour plane is then bound IN PLACE as the texture of a second pass (TextureBufferRef, texturebuffer.rs:12-47)."""
    rng = np.random.default_rng(909)
    w, h = 328, 200
    u = scenes.suzanne_uniforms(w, h)
    fb = P.RenderBuffer.with_dimensions(ctx, w, h, stencil=True, texture_buffer=True)
    fb.enable_winner(True)
    fb.clear(H.CLEAR)
    ofb = oracle_fb(w, h, True)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    draws = [(70_000, 3.0, sr.BLEND_REPLACE, (0, 0)), (500, 40.0, sr.BLEND_REPLACE, (0, 0)),
             (300, 40.0, sr.BLEND_ALPHA_OVER, (sr.STENCIL_LESS_THAN_EQ, sr.STENCIL_INCREMENT_WRAP))]
    for n, size, blend, st in draws:
        verts = H.random_screen_triangles(rng, n, w, h, max_size=size)
        idx = np.arange(3 * n, dtype=np.uint32)
        od = ob.OracleDraw(sr.TRIANGLE, idx, 2)
        od.set_vertices(verts, 1)
        od.blend = blend
        od.fragment_run(ofb, sr.FS_FLAT, u, *st)
        pipe.set_stencil_config(*st)
        pipe.draw_from_vertices(sr.TRIANGLE, verts, idx, 1, stencil=2).with_blend(blend).run(sr.FS_FLAT)
        assert np.array_equal(fb.download_winner(), ofb.winner)
        H.compare_framebuffers(fb.download(), ofb, exact_color=True, what=f"texture buffer, {n} triangles")  # the PixelRead view (AoS records)
        col, dep, st_plane = fb.download_planes(stencil=True)                                                # the planes themselves
        H.assert_bits_equal(col, ofb.color, "colour plane")
        H.assert_bits_equal(dep, ofb.depth, "depth plane")
        assert np.array_equal(st_plane, ofb.stencil)
    fb.set_pixel(7, 9, rgba=(0.5, 0.25, 0.125, 1.0), depth=-0.75)
    assert fb.pixel(7, 9)[:2] == ((0.5, 0.25, 0.125, 1.0), -0.75)
    assert np.array_equal(fb.download_rgba8().reshape(-1, 4)[9 * w + 7], [127, 63, 31, 255])
    # second pass: the colour plane sampled in place by a full-screen pass into an ordinary RenderBuffer
    src_color = fb.download_planes()[0].reshape(h, w, 4)
    fb2 = make_fb(P, ctx, w, h)
    p2 = P.Pipeline.from_framebuffer(fb2, u)
    p2.bind_framebuffer_texture(fb)
    p2.set_sampler(sr.FILTER_BILINEAR, sr.EDGE_WRAP)
    quad = np.array([[-1, -1, 0, 1, -0.3, 1.2], [1, -1, 0, 1, 1.4, 1.2], [1, 1, 0, 1, 1.4, -0.1], [-1, 1, 0, 1, -0.3, -0.1]], np.float32)
    qi = np.array([0, 1, 2, 0, 2, 3], np.uint32)
    vp = scenes.Viewport.new(w, h, 0.1, 10.0)
    qm = P.Mesh(ctx, vertices=quad, indices=qi)
    p2.render_mesh(sr.TRIANGLE, qm).run_to_fragment(vp, sr.VS_PASSTHROUGH).run(sr.FS_TEXTURE_UNLIT)
    ofb2 = oracle_fb(w, h)
    od2 = ob.OracleDraw(sr.TRIANGLE, qi)
    od2.vertex_run_to_fragment(vp, sr.VS_PASSTHROUGH, u, quad)
    od2.fragment_run(ofb2, sr.FS_TEXTURE_UNLIT, u, texture=src_color, sampler=(sr.FILTER_BILINEAR, sr.EDGE_WRAP, None))
    H.compare_framebuffers(fb2.download(), ofb2, exact_color=True, what="second pass over the texture buffer's colour plane")
    for x in (p2, qm, fb2, pipe, fb):
        x.destroy()


def test_texture_buffer_with_two_colour_planes(P, ctx):
    """declare_texture_buffer! takes one OR MORE named colours (src/framebuffer/texturebuffer.rs:72-110): the pixel colour is then the
    tuple of them (:129-133), the fragment shader returns the tuple and set_pixel_unchecked stores each into its own plane
    (:141-147); clear takes a tuple too (:181-197).  A G-buffer pass -- Suzanne's lit colour in plane 0, the interpolated normal in
    plane 1 -- through the clip path and the no-clip path, big and small draws, against the oracle; then the NORMAL plane is bound in
    place as the texture of a second pass (the named TextureBufferRef accessor, :110-117)."""
    mesh = H.suzanne_mesh()
    size = 320
    u = scenes.suzanne_uniforms(size, size)
    vp = scenes.Viewport.new(size, size, 0.001, 1000.0)
    clear0, clear1 = H.CLEAR, (0.5, 0.25, -1.0, 0.0)
    from softrender_b200._abi import SoftrenderError
    fb = P.RenderBuffer.with_dimensions(ctx, size, size, texture_buffer=2)
    fb.enable_winner(True)
    gmesh = P.Mesh(ctx, mesh)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    # freshly created: Color::empty() in both planes, Depth::far()
    assert not fb.download_attachment(0).any() and not fb.download_attachment(1).any()
    for clip in (False, True):
        fb.clear_attachment(0, clear0)
        fb.clear_attachment(1, clear1)
        ofb = ob.OracleFramebuffer(size, size, two_colors=True)
        ofb.clear(clear0, clear1)
        od = ob.OracleDraw(sr.TRIANGLE, mesh.indices)
        if clip:
            od.vertex_run(sr.VS_SUZANNE, u, mesh.vertices)
            od.clip_primitives()
            od.finish(vp)
            fs = pipe.render_mesh(sr.TRIANGLE, gmesh).run(sr.VS_SUZANNE).clip_primitives().finish(vp)
        else:
            od.vertex_run_to_fragment(vp, sr.VS_SUZANNE, u, mesh.vertices)
            fs = pipe.render_mesh(sr.TRIANGLE, gmesh).run_to_fragment(vp, sr.VS_SUZANNE)
        od.fragment_run(ofb, sr.FS_SUZANNE_GBUFFER, u)
        fs.run(sr.FS_SUZANNE_GBUFFER)
        assert np.array_equal(fb.download_winner(), ofb.winner), "coverage / winning primitive"
        c0, c1 = fb.download_attachment(0), fb.download_attachment(1)
        col, dep, _ = fb.download_planes()
        H.assert_bits_equal(col, c0, "download_planes' colour is plane 0")
        H.assert_bits_equal(dep, ofb.depth, "depth plane")
        assert np.abs(c0 - ofb.color).max() <= COLOR_TOL, "plane 0: lit colour"
        H.assert_bits_equal(c1, ofb.color1, "plane 1: interpolated normals (no shader arithmetic: bit-exact)")
        drawn = ofb.winner != 0
        assert drawn.any() and (~drawn).any()
        assert np.all(c1[~drawn] == np.float32(clear1)) and np.all(c0[~drawn] == np.float32(clear0))
    # a second draw onto the existing contents (no pending clear): big screen-space triangles, 8 attributes each
    rng = np.random.default_rng(77)
    n = 60
    verts = H.random_screen_triangles(rng, n, size, size, max_size=90.0, nk=8)
    idx = np.arange(3 * n, dtype=np.uint32)
    od = ob.OracleDraw(sr.TRIANGLE, idx)
    od.set_vertices(verts, 1)
    od.fragment_run(ofb, sr.FS_SUZANNE_GBUFFER, u)
    pipe.draw_from_vertices(sr.TRIANGLE, verts, idx, 1).run(sr.FS_SUZANNE_GBUFFER)
    assert np.array_equal(fb.download_winner(), ofb.winner)
    # (arbitrary "normals" make the lit colour a power of garbage: NaN where the oracle has NaN)
    c0, c1 = fb.download_attachment(0), fb.download_attachment(1)
    H.assert_bits_equal(c1, ofb.color1, "plane 1 after the second draw")
    ok = np.isfinite(ofb.color) & np.isfinite(c0)
    assert np.array_equal(np.isnan(c0), np.isnan(ofb.color)) and np.abs(c0 - ofb.color)[ok].max() <= COLOR_TOL
    # accessors see plane 0 and the depth
    px = 5 * size + 3
    dep = fb.download_planes()[1]
    got = fb.pixel(3, 5)
    assert np.array_equal(np.float32(got[0]), c0[px], equal_nan=True) and np.float32(got[1]) == dep[px]
    # type agreement between shader and target, both ways
    with pytest.raises(SoftrenderError):
        pipe.draw_from_vertices(sr.TRIANGLE, verts, idx, 1).run(sr.FS_SUZANNE)
    fb1 = make_fb(P, ctx, size, size)
    p1 = P.Pipeline.from_framebuffer(fb1, u)
    with pytest.raises(SoftrenderError):
        p1.draw_from_vertices(sr.TRIANGLE, verts, idx, 1).run(sr.FS_SUZANNE_GBUFFER)
    with pytest.raises(SoftrenderError):
        pipe.draw_from_vertices(sr.TRIANGLE, verts, idx, 1).with_blend(sr.BLEND_ALPHA_OVER).run(sr.FS_SUZANNE_GBUFFER)
    with pytest.raises(SoftrenderError):
        fb.download()  # no 20-byte pixel view of a tuple colour
    with pytest.raises(SoftrenderError):
        p1.bind_framebuffer_attachment(fb, 2)
    # second pass: the NORMAL plane sampled in place
    normals = fb.download_attachment(1).reshape(size, size, 4)
    p1.bind_framebuffer_attachment(fb, 1)
    p1.set_sampler(sr.FILTER_BILINEAR, sr.EDGE_CLAMP)
    quad = np.array([[-1, -1, 0, 1, 0, 1], [1, -1, 0, 1, 1, 1], [1, 1, 0, 1, 1, 0], [-1, 1, 0, 1, 0, 0]], np.float32)
    qi = np.array([0, 1, 2, 0, 2, 3], np.uint32)
    qm = P.Mesh(ctx, vertices=quad, indices=qi)
    p1.render_mesh(sr.TRIANGLE, qm).run_to_fragment(vp, sr.VS_PASSTHROUGH).run(sr.FS_TEXTURE_UNLIT)
    ofb2 = oracle_fb(size, size)
    od2 = ob.OracleDraw(sr.TRIANGLE, qi)
    od2.vertex_run_to_fragment(vp, sr.VS_PASSTHROUGH, u, quad)
    od2.fragment_run(ofb2, sr.FS_TEXTURE_UNLIT, u, texture=normals, sampler=(sr.FILTER_BILINEAR, sr.EDGE_CLAMP, None))
    H.compare_framebuffers(fb1.download(), ofb2, exact_color=True, what="second pass over the normal plane")
    for x in (p1, qm, fb1, pipe, gmesh, fb):
        x.destroy()


def test_user_blend_function_additive(P, ctx):
    """A third registered blend, the stand-in for a user's GenericBlend::new(|a, b| a + b) (src/color/blend.rs:57-76; recipe in
    INTEGRATION.md): strictly ordered path, triangles + antialiased lines + points, colours bit-exact (f32 addition is not
    associative, so submission order is observable)."""
    rng = np.random.default_rng(613)
    w, h, n = 150, 110, 400
    verts = H.random_screen_triangles(rng, n, w, h, max_size=40.0)
    idx = np.arange(3 * n, dtype=np.uint32)
    lines = np.zeros((60, 8), np.float32)
    lines[:, 0], lines[:, 1] = rng.uniform(0, w, 60), rng.uniform(0, h, 60)
    lines[:, 2], lines[:, 3], lines[:, 4:] = -0.05, 1, rng.uniform(0, 1, (60, 4))
    out, win, _, ofb = run_both_screen(P, ctx, w, h, verts, idx, blend=sr.BLEND_ADDITIVE, aa=True, gen={2: lines, 1: lines[:20]}, draws=2)
    assert np.array_equal(win, ofb.winner)
    H.compare_framebuffers(out, ofb, exact_color=True, what="additive blend")
    assert ofb.color.max() > 2.0  # contributions accumulated
    assert [r["name"] for r in P.registry(3)] == ["replace", "alpha_over", "additive"]


def test_clear_resets_the_stencil_plane_on_the_opaque_path(P, ctx):
    """RenderBuffer::clear resets stencil to its default (renderbuffer/mod.rs:126-133).  The clear is recorded and produced on
    chip by the next draw; when that draw takes the opaque path (stencil Always / Keep never touches the plane) the stencil
    values of the previous frame must still be gone."""
    rng = np.random.default_rng(17)
    w, h, n = 120, 90, 150
    verts = H.random_screen_triangles(rng, n, w, h)
    idx = np.arange(3 * n, dtype=np.uint32)
    u = scenes.suzanne_uniforms(w, h)
    fb = make_fb(P, ctx, w, h, stencil=True)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    pipe.set_stencil_config(sr.STENCIL_ALWAYS, sr.STENCIL_INCREMENT_WRAP)
    pipe.draw_from_vertices(sr.TRIANGLE, verts, idx, 1).run(sr.FS_FLAT)
    assert fb.download_planes(stencil=True)[2].max() > 0
    fb.clear(H.CLEAR)
    pipe.set_stencil_config(sr.STENCIL_ALWAYS, sr.STENCIL_KEEP)
    pipe.draw_from_vertices(sr.TRIANGLE, verts, idx, 1).run(sr.FS_FLAT)  # opaque path, lazy clear
    assert fb.download_planes(stencil=True)[2].max() == 0
    pipe.destroy()
    fb.destroy()


def test_pixel_write_accessors(P, ctx):
    """PixelWrite::pixel_mut / FramebufferAccessorMut::{set_depth, set_stencil} (src/pixels/mod.rs:77-98,
    src/framebuffer/accessor.rs:52-70) through sr_framebuffer_set_pixel: a depth written by hand is what the next draw's
    depth test sees, a stencil value written by hand is what its stencil test sees."""
    w, h = 40, 30
    fb = P.RenderBuffer.with_dimensions(ctx, w, h, stencil=16)
    fb.clear(H.CLEAR)
    fb.set_pixel(5, 6, rgba=(0.25, 0.5, 0.75, 1.0), depth=-0.125, stencil=40000)
    assert fb.pixel(5, 6) == ((0.25, 0.5, 0.75, 1.0), -0.125, 40000)
    assert fb.pixel(6, 6)[2] == 0 and fb.pixel(6, 6)[0] == tuple(np.float32(H.CLEAR).tolist())
    with pytest.raises(Exception):
        fb.set_pixel(w, 0, depth=-1.0)  # RenderError::InvalidPixelCoordinate
    with pytest.raises(Exception):
        fb.set_pixel(0, 0, stencil=70000)  # does not fit u16
    # a full-frame triangle at z = -1: the hand-written pixel (depth -0.125, nearer) must survive the depth test
    u = scenes.suzanne_uniforms(w, h)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    tri = np.array([[-50, -50, -1, 1, 1, 0, 0, 1], [200, -50, -1, 1, 1, 0, 0, 1], [-50, 200, -1, 1, 1, 0, 0, 1]], np.float32)
    pipe.draw_from_vertices(sr.TRIANGLE, tri, np.arange(3, dtype=np.uint32), 1).run(sr.FS_FLAT)
    assert fb.pixel(5, 6)[0] == (0.25, 0.5, 0.75, 1.0) and fb.pixel(5, 6)[1] == -0.125
    assert fb.pixel(6, 6)[0] == (1.0, 0.0, 0.0, 1.0) and fb.pixel(6, 6)[1] == -1.0
    # stencil Equal against 40000 with a nearer triangle: only the hand-written pixel passes
    pipe.set_stencil_config(sr.STENCIL_EQUAL, sr.STENCIL_KEEP)
    tri[:, 2] = -0.01
    tri[:, 4:8] = [0, 1, 0, 1]
    pipe.draw_from_vertices(sr.TRIANGLE, tri, np.arange(3, dtype=np.uint32), 1, stencil=40000).run(sr.FS_FLAT)
    assert fb.pixel(5, 6)[0] == (0.0, 1.0, 0.0, 1.0) and fb.pixel(6, 6)[0] == (1.0, 0.0, 0.0, 1.0)
    pipe.destroy()
    fb.destroy()


@pytest.mark.parametrize("aa", [False, True])
def test_lines(P, ctx, aa):
    rng = np.random.default_rng(53)
    w, h, n = 140, 100, 200
    v = np.zeros((2 * n, 8), np.float32)
    v[:, 0] = rng.uniform(-20, w + 20, 2 * n)
    v[:, 1] = rng.uniform(-20, h + 20, 2 * n)
    v[:, 2] = -rng.uniform(0.5, 5, 2 * n)
    v[:, 3] = 1
    v[:, 4:] = rng.uniform(0, 1, (2 * n, 4))
    v[0, :2], v[1, :2] = (5.5, 7.5), (5.5, 60.2)     # vertical
    v[2, :2], v[3, :2] = (3.2, 9.5), (120.7, 9.5)    # horizontal
    v[4, :2], v[5, :2] = (30.0, 30.0), (30.0, 30.0)  # zero length
    idx = np.arange(2 * n, dtype=np.uint32)
    blend = sr.BLEND_ALPHA_OVER if aa else sr.BLEND_REPLACE
    out, win, _, ofb = run_both_screen(P, ctx, w, h, v, idx, prim=sr.LINE, aa=aa, blend=blend)
    assert np.array_equal(win, ofb.winner)
    H.compare_framebuffers(out, ofb, exact_color=True, what="lines")


def test_points_half_open_bounds(P, ctx):
    rng = np.random.default_rng(59)
    w, h, n = 90, 70, 3000
    v = np.zeros((n, 8), np.float32)
    v[:, 0] = rng.uniform(-3, w + 3, n)
    v[:, 1] = rng.uniform(-3, h + 3, n)
    v[:, 2] = -rng.uniform(0.5, 5, n)
    v[:, 3] = 1
    v[:, 4:] = rng.uniform(0, 1, (n, 4))
    v[0, :2] = (w - 1 + 0.5, 10.5)  # last column: never drawn (point.rs:46)
    v[1, :2] = (10.5, h - 1 + 0.5)  # last row: never drawn
    idx = np.arange(n, dtype=np.uint32)
    out, win, _, ofb = run_both_screen(P, ctx, w, h, v, idx, prim=sr.POINT)
    assert np.array_equal(win, ofb.winner)
    H.compare_framebuffers(out, ofb, exact_color=True, what="points")
    W = win.reshape(h, w)
    assert W[:, w - 1].max() == 0 and W[h - 1, :].max() == 0


def test_mixed_generated_primitives_order(P, ctx):
    """Triangles, then lines, then points (fragment.rs:268-311), indexed before generated."""
    rng = np.random.default_rng(61)
    w, h = 120, 96
    tri = H.random_screen_triangles(rng, 80, w, h)
    gen_tri = H.random_screen_triangles(rng, 50, w, h)
    lines = H.random_screen_triangles(rng, 40, w, h)[:80]
    pts = H.random_screen_triangles(rng, 100, w, h)[:300]
    out, win, _, ofb = run_both_screen(P, ctx, w, h, tri, np.arange(len(tri), dtype=np.uint32),
                                       gen={3: gen_tri, 2: lines, 1: pts})
    assert np.array_equal(win, ofb.winner)
    H.compare_framebuffers(out, ofb, exact_color=True, what="mixed")


@pytest.mark.parametrize("world", [1, 3])
def test_opaque_lines_and_points_through_the_visibility_buffer(P, ctx, world):
    """With Blend=(), no stencil and Bresenham lines the whole draw is order independent: triangles, lines and points are
    reduced into one visibility buffer (k_micro / k_lines_vis / k_points_vis) and the resolve shades whichever won.
    Integer depths force ties between the three kinds (later kind / later primitive wins, fragment.rs:268-311); drawn
    twice (the second draw starts from the stored depth) and, for world > 1, tile-sharded into one framebuffer."""
    rng = np.random.default_rng(97)
    w, h = 260, 170
    tri = H.random_screen_triangles(rng, 400, w, h, integer_depth=True)
    lines = H.random_screen_triangles(rng, 300, w, h, integer_depth=True)[:600]
    lines[0, :2], lines[1, :2] = (-30.0, 20.5), (w + 40.0, 90.25)   # clipped on both sides
    lines[2, :2], lines[3, :2] = (40.5, 40.5), (40.5, 40.5)          # zero length: nothing is drawn
    pts = H.random_screen_triangles(rng, 400, w, h, integer_depth=True)[:1200]
    pts[0, :2] = (w - 1 + 0.5, 10.5)                                 # last column: never drawn (point.rs:46)
    idx = np.arange(len(tri), dtype=np.uint32)
    out1, win1, _, ofb = run_both_screen(P, ctx, w, h, tri, idx, gen={2: lines, 1: pts}, draws=2)
    assert np.array_equal(win1, ofb.winner)
    H.compare_framebuffers(out1, ofb, exact_color=True, what="opaque lines/points")
    kinds = np.digitize(ofb.winner[ofb.winner > 0] - 1, [len(tri) // 3, len(tri) // 3 + len(lines) // 2])
    assert all((kinds == k).sum() > 50 for k in (0, 1, 2)), "every primitive kind must win some pixels"
    if world > 1:
        fb = make_fb(P, ctx, w, h)
        pipe = P.Pipeline.from_framebuffer(fb, scenes.suzanne_uniforms(w, h))
        try:
            for rank in range(world):
                ctx.set_tile_shard(rank, world)
                fb.clear(H.CLEAR)
                for _ in range(2):
                    d = pipe.draw_from_vertices(sr.TRIANGLE, tri, idx, 1)
                    d.set_generated(2, lines)
                    d.set_generated(1, pts)
                    d.run(sr.FS_FLAT)
        finally:
            ctx.set_tile_shard(0, 1)
        H.assert_bits_equal(fb.download(), out1, "sharded opaque lines/points")
        pipe.destroy()
        fb.destroy()


# ------------------------------------------------------------------------------------------------------
# geometry stage
# ------------------------------------------------------------------------------------------------------
def _clip_space_triangles(rng, n, nk=8):
    v = np.zeros((3 * n, 4 + nk), np.float32)
    v[:, 3] = rng.uniform(0.2, 3.0, 3 * n)
    v[:, 0] = rng.uniform(-1.8, 1.8, 3 * n) * v[:, 3]
    v[:, 1] = rng.uniform(-1.8, 1.8, 3 * n) * v[:, 3]
    v[:, 2] = rng.uniform(-0.6, 1.6, 3 * n) * v[:, 3]
    neg = rng.uniform(0, 1, 3 * n) < 0.05
    v[neg, 3] *= -1  # behind the eye
    v[:, 4:] = rng.uniform(-1, 1, (3 * n, nk))
    return v


@pytest.mark.parametrize("prim", [sr.TRIANGLE, sr.LINE, sr.POINT])
def test_clip_primitives_random(P, ctx, prim):
    rng = np.random.default_rng(71)
    n = 600
    verts = _clip_space_triangles(rng, n)
    idx = rng.integers(0, len(verts), prim * 500).astype(np.uint32)
    u = scenes.suzanne_uniforms(64, 64)
    fb = make_fb(P, ctx, 64, 64)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    od = ob.OracleDraw(prim, idx)
    od.set_vertices(verts, 0).clip_primitives()
    gs = pipe.draw_from_vertices(prim, verts, idx, 0).clip_primitives()
    S = verts.shape[1]
    if prim == sr.TRIANGLE:
        o = od.data(3).reshape(-1, 3, S)
        g = gs.download(3).reshape(-1, 3, S)
        seq = gs.download_sequence()
        H.assert_bits_equal(g, o[seq], "clipped triangles")
        dropped = np.ones(len(o), bool)
        dropped[seq] = False
        d = o[dropped]
        same = (np.all(H.bits(d[:, 0]) == H.bits(d[:, 1]), axis=1) | np.all(H.bits(d[:, 0]) == H.bits(d[:, 2]), axis=1)
                | np.all(H.bits(d[:, 1]) == H.bits(d[:, 2]), axis=1))
        assert same.all()
    else:
        H.assert_bits_equal(gs.download(prim), od.data(prim), "clipped")
    pipe.destroy()
    fb.destroy()


def test_correct_clipper_opt_in(P, ctx):
    """SR_GS_CLIP_SH, the opt-in Sutherland-Hodgman clipper (the fix src/lib.rs asks for): bit-identical with the oracle's
    definition, every output vertex inside the frustum (up to rounding of the intersections), triangles entirely inside
    come through unchanged, and the rendered frame matches."""
    rng = np.random.default_rng(73)
    n = 900
    verts = _clip_space_triangles(rng, n)
    idx = rng.integers(0, len(verts), 3 * 700).astype(np.uint32)
    w, h = 96, 80
    u = scenes.suzanne_uniforms(w, h)
    vp = scenes.Viewport.new(w, h, 0.1, 100.0)
    fb = make_fb(P, ctx, w, h)
    ofb = oracle_fb(w, h)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    od = ob.OracleDraw(sr.TRIANGLE, idx)
    od.set_vertices(verts, 0).clip_primitives(correct=True)
    gs = pipe.draw_from_vertices(sr.TRIANGLE, verts, idx, 0).clip_primitives(correct=True)
    S = verts.shape[1]
    g = gs.download(3).reshape(-1, 3, S)
    o = od.data(3).reshape(-1, 3, S)
    H.assert_bits_equal(g, o, "Sutherland-Hodgman output")
    assert 0 < len(g) <= 7 * 700
    x, y, z, ww = (g[..., k].astype(np.float64) for k in range(4))
    eps = 1e-4 * np.maximum(1.0, np.abs(ww))
    assert (x >= -ww - eps).all() and (x <= ww + eps).all() and (y >= -ww - eps).all() and (y <= ww + eps).all()
    assert (z >= -eps).all() and (z <= ww + eps).all()
    tri = verts[idx].reshape(-1, 3, S)
    inside = ((np.abs(tri[..., 0]) <= tri[..., 3]) & (np.abs(tri[..., 1]) <= tri[..., 3]) & (tri[..., 2] >= 0) & (tri[..., 2] <= tri[..., 3])).all(axis=1)
    assert inside.sum() >= 2
    gset = {bytes(t.tobytes()) for t in g}
    assert all(bytes(t.tobytes()) in gset for t in tri[inside])
    od.finish(vp).fragment_run(ofb, sr.FS_FLAT, u)
    gs.finish(vp).run(sr.FS_FLAT)
    assert np.array_equal(fb.download_winner(), ofb.winner)
    H.compare_framebuffers(fb.download(), ofb, exact_color=True, what="correct clipper frame")
    pipe.destroy()
    fb.destroy()


@pytest.mark.parametrize("gs_id", [sr.GS_FACE_NORMALS, sr.GS_VERTEX_NORMALS])
def test_normal_visualisation_geometry_shaders(P, ctx, gs_id):
    size = 200
    mesh = H.suzanne_mesh()
    u = scenes.suzanne_uniforms(size, size)
    vp = scenes.Viewport.new(size, size, 0.001, 1000.0)
    fb = make_fb(P, ctx, size, size)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    gmesh = P.Mesh(ctx, mesh)
    od = ob.OracleDraw(sr.TRIANGLE, mesh.indices)
    od.vertex_run(sr.VS_SUZANNE, u, mesh.vertices).geometry_run(gs_id, u)
    gs = pipe.render_mesh(sr.TRIANGLE, gmesh).run(sr.VS_SUZANNE).run(gs_id)
    H.assert_bits_equal(gs.download(2), od.data(2), "normal lines")
    ofb = oracle_fb(size, size)
    od.finish(vp).fragment_run(ofb, sr.FS_GREEN, u)
    gs.finish(vp).run(sr.FS_GREEN)
    assert np.array_equal(fb.download_winner(), ofb.winner)
    H.compare_framebuffers(fb.download(), ofb, exact_color=True, what="normal lines frame")
    for x in (pipe, gmesh, fb):
        x.destroy()


# ------------------------------------------------------------------------------------------------------
# full_example scene parts (config 2) and the synthetic grid (configs 3/4, reduced)
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("textured,camera_distance", [(False, 2.0), (True, 2.0), (True, 0.9)])
def test_full_example_scene(P, ctx, textured, camera_distance):
    w, h = 480, 270
    mesh = H.suzanne_mesh(with_uv=True)
    tex = scenes.checker_texture(64, 8)
    vp = scenes.Viewport.new(w, h, 0.1, 1000.0)
    fb = make_fb(P, ctx, w, h)
    ofb = oracle_fb(w, h)
    gmesh = P.Mesh(ctx, mesh)
    gtex = P.Texture(ctx, tex)
    fs = sr.FS_FULL_EXAMPLE_TEXTURED if textured else sr.FS_FULL_EXAMPLE
    pipe = None
    for k, (rot, off) in enumerate([(45.0, -1.6), (165.0, 0.0), (285.0, 1.6)]):
        u = scenes.full_example_uniforms(w / h, np.deg2rad(75.0), camera_distance, np.deg2rad(rot), np.deg2rad(65.0), off)
        if pipe is None:
            pipe = P.Pipeline.from_framebuffer(fb, u)
            pipe.bind_texture(gtex)
            pipe.set_sampler(sr.FILTER_BILINEAR, sr.EDGE_CLAMP)  # config 2 samples Bilinear + Clamp
        else:
            pipe.set_uniforms(u)
        od = ob.OracleDraw(sr.TRIANGLE, mesh.indices)
        od.blend = sr.BLEND_ALPHA_OVER
        od.vertex_run(sr.VS_FULL_EXAMPLE, u, mesh.vertices)
        gs = pipe.render_mesh(sr.TRIANGLE, gmesh).run(sr.VS_FULL_EXAMPLE)
        H.assert_bits_equal(gs.download(0), od.data(0), "full_example vertex stage")
        if camera_distance < 1.0:
            od.clip_primitives()
            gs = gs.clip_primitives()
        od.finish(vp).fragment_run(ofb, fs, u, texture=tex, sampler=(sr.FILTER_BILINEAR, sr.EDGE_CLAMP, None) if tex is not None else None)
        gs.finish(vp).with_blend(sr.BLEND_ALPHA_OVER).run(fs)
    assert np.array_equal(fb.download_winner(), ofb.winner)
    H.compare_framebuffers(fb.download(), ofb, color_tol=COLOR_TOL, what="full_example")
    for x in (pipe, gmesh, gtex, fb):
        x.destroy()


@pytest.mark.parametrize("reverse", [False, True])
def test_grid_scene_reduced(P, ctx, reverse):
    """Config 3's generator at 1/100 of the triangle count on a 960x540 frame (same px^2 per triangle)."""
    w, h = 960, 540
    mesh = scenes.make_grid(125, 100, 4, seed=0x5EED0003, reverse=reverse)
    u = scenes.grid_uniforms(w, h)
    vp = scenes.Viewport.new(w, h, 0.1, 100.0)
    fb = make_fb(P, ctx, w, h)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    gmesh = P.Mesh(ctx, mesh)
    pipe.render_mesh(sr.TRIANGLE, gmesh).run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)
    ofb = oracle_fb(w, h)
    od = ob.OracleDraw(sr.TRIANGLE, mesh.indices)
    od.vertex_run_to_fragment(vp, sr.VS_SUZANNE, u, mesh.vertices).fragment_run(ofb, sr.FS_SUZANNE, u)
    assert np.array_equal(fb.download_winner(), ofb.winner)
    H.compare_framebuffers(fb.download(), ofb, color_tol=COLOR_TOL, what="grid")
    for x in (pipe, gmesh, fb):
        x.destroy()


@pytest.mark.parametrize("reverse", [False, True])
@pytest.mark.parametrize("cells", [(125, 100), (250, 200)])
def test_grid_scene_reduced_blended_in_order(P, ctx, reverse, cells):
    """The same generator drawn with the additive blend (strictly ordered path): a coherent mesh of pixel-sized triangles is
    what k_tile_ordered's lane-per-triangle sweep and per-batch interleaved row ownership are for.  The flat shader returns
    the interpolated world position, so colours are sums of f32 values in submission order: bit-exact, and back to front
    (`reverse`) every layer passes the depth test and contributes."""
    w, h = 960, 540
    mesh = scenes.make_grid(*cells, 4, seed=0x5EED0003, reverse=reverse)
    u = scenes.grid_uniforms(w, h)
    vp = scenes.Viewport.new(w, h, 0.1, 100.0)
    fb = make_fb(P, ctx, w, h)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    gmesh = P.Mesh(ctx, mesh)
    pipe.render_mesh(sr.TRIANGLE, gmesh).run_to_fragment(vp, sr.VS_SUZANNE).with_blend(sr.BLEND_ADDITIVE).run(sr.FS_FLAT)
    ofb = oracle_fb(w, h)
    od = ob.OracleDraw(sr.TRIANGLE, mesh.indices)
    od.blend = sr.BLEND_ADDITIVE
    od.vertex_run_to_fragment(vp, sr.VS_SUZANNE, u, mesh.vertices).fragment_run(ofb, sr.FS_FLAT, u)
    assert np.array_equal(fb.download_winner(), ofb.winner)
    H.compare_framebuffers(fb.download(), ofb, exact_color=True, what="blended grid")
    for x in (pipe, gmesh, fb):
        x.destroy()


def test_turntable_frames(P, ctx):
    """Config 5: turntable frames k use model rotation 3 deg * (k+1) about y (realtime_example/src/main.rs:90-93);
    every frame is a full clear + draw on the same pipeline (set_uniforms between frames)."""
    size = 256
    mesh = H.suzanne_mesh()
    vp = scenes.Viewport.new(size, size, 0.001, 1000.0)
    fb = make_fb(P, ctx, size, size)
    pipe = P.Pipeline.from_framebuffer(fb, scenes.suzanne_uniforms(size, size))
    gmesh = P.Mesh(ctx, mesh)
    for k in (0, 7, 31, 63):
        u = scenes.suzanne_uniforms(size, size, rotation_y=np.deg2rad(3.0 * (k + 1)))
        pipe.set_uniforms(u)
        fb.clear(H.CLEAR)
        pipe.render_mesh(sr.TRIANGLE, gmesh).run(sr.VS_SUZANNE).clip_primitives().finish(vp).run(sr.FS_SUZANNE)
        ofb = oracle_fb(size, size)
        od = ob.OracleDraw(sr.TRIANGLE, mesh.indices)
        od.vertex_run(sr.VS_SUZANNE, u, mesh.vertices).clip_primitives().finish(vp).fragment_run(ofb, sr.FS_SUZANNE, u)
        assert np.array_equal(fb.download_winner(), ofb.winner), f"turntable frame {k}"
        H.compare_framebuffers(fb.download(), ofb, color_tol=COLOR_TOL, what=f"turntable frame {k}")
    for x in (pipe, gmesh, fb):
        x.destroy()


# ------------------------------------------------------------------------------------------------------
# BASELINE.json's full sizes through size-independent properties of the CUDA path (idempotence, split and shard
# invariance); the full-size comparison with the oracle itself is tests/test_gpu_sizes_full.py
# ------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def grid10m():
    return scenes.make_grid(1250, 1000, 4, seed=0x5EED0003)


def _draw_grid(P, ctx, gmesh, w, h, vp, u, draws=1, winner=True):
    fb = make_fb(P, ctx, w, h, winner=winner)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    for _ in range(draws):
        pipe.render_mesh(sr.TRIANGLE, gmesh).run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)
    out = fb.download()
    win = fb.download_winner() if winner else None
    pipe.destroy()
    fb.destroy()
    return out, win


def test_full_size_grid10m_properties(P, ctx, grid10m):
    """Config 3 at full size (10M triangles, 3840x2160).  Properties that hold for the reference's algorithm:
    (1) idempotence: drawing the same mesh a second time over the result changes nothing (`d >= dt` lets exactly the
        stored winner pass again, triangle.rs:126) -- this also runs the keys-from-stored-depth initialisation;
    (2) the split between the per-triangle and the per-tile path is invisible (every triangle through tile lists);
    (3) sort-first sharding is invisible (3 shards into one framebuffer);
    (4) coverage sanity: the mesh covers a centred rectangle, every covered pixel carries z < 0 and a valid id."""
    w, h = 3840, 2160
    u = scenes.grid_uniforms(w, h)
    vp = scenes.Viewport.new(w, h, 0.1, 100.0)
    gmesh = P.Mesh(ctx, grid10m)
    ref, win = _draw_grid(P, ctx, gmesh, w, h, vp, u)
    # (4)
    covered = win > 0
    assert covered.sum() > 3_000_000
    assert win.max() <= grid10m.ntris
    assert (ref[covered, 4] < 0).all()
    assert np.array_equal(ref[~covered, :4], np.broadcast_to(np.array(H.CLEAR, np.float32), ((~covered).sum(), 4)))
    # (1)
    twice, win2 = _draw_grid(P, ctx, gmesh, w, h, vp, u, draws=2)
    H.assert_bits_equal(twice, ref, "second identical draw")
    assert np.array_equal(win2, win)
    # (2)
    ctx.set_micro(0, 65536, 0)
    try:
        lists_only, win3 = _draw_grid(P, ctx, gmesh, w, h, vp, u)
    finally:
        ctx.set_micro()
    H.assert_bits_equal(lists_only, ref, "all triangles through per-tile lists")
    assert np.array_equal(win3, win)
    # (3)
    fb = make_fb(P, ctx, w, h, winner=False)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    try:
        for rank in range(3):
            ctx.set_tile_shard(rank, 3)
            fb.clear(H.CLEAR)
            pipe.render_mesh(sr.TRIANGLE, gmesh).run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)
    finally:
        ctx.set_tile_shard(0, 1)
    H.assert_bits_equal(fb.download(), ref, "3-way sharded frame")
    for x in (pipe, fb, gmesh):
        x.destroy()


def test_full_size_grid10m_reverse_order(P, ctx):
    """Back-to-front submission (maximal overdraw) of config 3 gives the same depth plane as front-to-back: the
    visible surface does not depend on submission order when no two layers tie in depth."""
    w, h = 3840, 2160
    u = scenes.grid_uniforms(w, h)
    vp = scenes.Viewport.new(w, h, 0.1, 100.0)
    depth = []
    for reverse in (False, True):
        gmesh = P.Mesh(ctx, scenes.make_grid(1250, 1000, 4, seed=0x5EED0003, reverse=reverse))
        out, _ = _draw_grid(P, ctx, gmesh, w, h, vp, u, winner=False)
        depth.append(out[:, 4].copy())
        gmesh.destroy()
    same = depth[0].view(np.uint32) == depth[1].view(np.uint32)
    # within a layer, pixel centres exactly on shared edges tie and go to the later triangle: the depth is the same
    # value either way (both triangles interpolate the shared edge), so the planes must agree bit for bit
    assert same.mean() > 0.9999, f"{(~same).sum()} depth values differ between submission orders"


def test_presentation_readback_rgba8(P, ctx, tmp_path):
    """sr_framebuffer_download_rgba8 = the loop of realtime_example/src/main.rs:100-116, `(c * 255.0) as u8` per channel
    (truncating, saturating, NaN -> 0), both byte orders; checked on hostile colour values and on a rendered frame,
    and the PNG written from it decodes back to the same bytes."""
    w, h = 37, 19  # odd size: exercises the non-vectorised tail
    rng = np.random.default_rng(5)
    color = rng.uniform(-0.5, 1.5, (w * h, 4)).astype(np.float32)
    color[:12, 0] = [0.0, -0.0, 1.0, 0.999999, 1.0 / 255.0, 254.999 / 255.0, np.nan, np.inf, -np.inf, 1e30, -1e30, 0.5]
    fb = P.RenderBuffer.with_dimensions(ctx, w, h)
    fb.upload_planes(color=color, depth=np.zeros(w * h, np.float32))
    with np.errstate(invalid="ignore", over="ignore"):
        scaled = color * np.float32(255.0)
        expect = np.where(np.isnan(scaled), 0.0, np.clip(np.trunc(scaled), 0.0, 255.0)).astype(np.uint8).reshape(h, w, 4)
    assert np.array_equal(fb.download_rgba8(), expect)
    assert np.array_equal(fb.download_rgba8(abgr=True), expect[..., ::-1])
    fb.clear(H.CLEAR)  # a pending (lazy) clear must be visible too
    assert np.array_equal(fb.download_rgba8()[0, 0], (np.asarray(H.CLEAR, np.float32) * np.float32(255)).astype(np.uint8))
    fb.destroy()
    # a rendered frame, 4-pixel vector path (width % 4 == 0), and the PNG round trip
    size = 64
    mesh = H.suzanne_mesh()
    from softrender_b200 import scenes
    import softrender_b200 as sr
    fb = P.RenderBuffer.with_dimensions(ctx, size, size)
    fb.clear(H.CLEAR)
    pipe = P.Pipeline.from_framebuffer(fb, scenes.suzanne_uniforms(size, size))
    gm = P.Mesh(ctx, mesh)
    pipe.render_mesh(sr.TRIANGLE, gm).run_to_fragment(scenes.Viewport.new(size, size, 0.001, 1000.0), sr.VS_SUZANNE).run(sr.FS_SUZANNE)
    f32 = fb.download()[:, :4].reshape(size, size, 4)
    img = fb.copy_to_image()
    assert np.array_equal(img, np.clip(np.trunc(f32 * np.float32(255.0)), 0, 255).astype(np.uint8))
    assert (img[..., :3].max(axis=-1) > 60).sum() > 200  # the model is there
    path = tmp_path / "frame.png"
    fb.save_png(str(path))
    import struct, zlib
    blob = path.read_bytes()
    assert blob[:8] == b"\x89PNG\r\n\x1a\n" and struct.unpack(">II", blob[16:24]) == (size, size)
    pos, idat = 8, b""
    while pos < len(blob):
        n, tag = struct.unpack(">I", blob[pos:pos + 4])[0], blob[pos + 4:pos + 8]
        assert zlib.crc32(blob[pos + 4:pos + 8 + n]) & 0xFFFFFFFF == struct.unpack(">I", blob[pos + 8 + n:pos + 12 + n])[0]
        if tag == b"IDAT":
            idat += blob[pos + 8:pos + 8 + n]
        pos += 12 + n
    rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(size, 1 + size * 4)
    assert np.array_equal(rows[:, 1:].reshape(size, size, 4), img) and not rows[:, 0].any()
    pipe.destroy(); gm.destroy(); fb.destroy()


def test_exact_division_shortcut(P, ctx):
    """The coverage path replaces `n / det` by a reciprocal + two FMA corrections; it must be the IEEE quotient,
    bit for bit, over its whole validity range (4e9 random operand pairs incl. all-ones/sparse mantissas)."""
    for seed in (1, 0xDEADBEEF):
        assert ctx.selftest_division(seed, 2_000_000_000) == 0


def test_randomised_scenarios(P, ctx):
    """250 random scenarios (tests/fuzz_scenarios.py): frame sizes, triangle / line / point mixes, blend, stencil, cull,
    antialiased lines, one or two draws, every split setting -- each must equal the oracle bit for bit.  The seeds
    include the ones with which the campaign (profiles/scripts/fuzz.py) found the too-narrow tile rectangle of Wu lines."""
    from fuzz_scenarios import run_scenario
    failures = [m for m in (run_scenario(P, ctx, run_both_screen, seed) for seed in list(range(1000, 1240)) + [1291, 1558, 2010, 2052]
                            + list(range(5000, 5006))) if m]
    assert not failures, "\n".join(failures)


def test_randomised_pipeline_scenarios(P, ctx):
    """80 random scenes through the whole builder chain (tests/fuzz_scenarios.py::run_pipeline_scenario): Suzanne, subdivided
    Suzanne or a displaced grid; random camera distance (meshes cross the frustum planes when close), rotation, frame size;
    both example shader sets, textured or not (8-bit image or a render target sampled in place, any Filter / Edge);
    run_to_fragment / literal clipper / Sutherland-Hodgman clipper; blend, cull, one or two draws, optionally tile-sharded.
    Winner and depth bit-exact, colour within 1/255."""
    from fuzz_scenarios import run_pipeline_scenario
    # the four seeds past the range are regressions: Nearest sampling is a step function of the interpolated uv, which the lit
    # shaders' contracted (FMA) attribute interpolation moved by an ulp across a texel boundary -- the uv plane is exact now
    seeds = list(range(100, 180)) + [1005969, 1009630, 1012157, 1012189]
    failures = [m for m in (run_pipeline_scenario(P, ctx, ob, scenes, seed) for seed in seeds) if m]
    assert not failures, "\n".join(failures)


def test_randomised_geometry_scenarios(P, ctx):
    """60 random scenes of point / line meshes and the normal-visualisation geometry shaders (Bresenham or Wu, blended or
    not, clipped or not, alone or on top of the shaded mesh): tests/fuzz_scenarios.py::run_geometry_scenario."""
    from fuzz_scenarios import run_geometry_scenario
    failures = [m for m in (run_geometry_scenario(P, ctx, ob, scenes, seed) for seed in range(100, 160)) if m]
    assert not failures, "\n".join(failures)


def test_antialiased_lines_across_tile_and_band_boundaries(P, ctx):
    """Wu plots the rows trunc(yend) and trunc(yend) + 1 with yend up to half a pixel past the clipped end point, so a line
    reaches up to two rows / columns beyond the truncated end-point box: lines ending just before tile rows (multiples of 32),
    tile columns (64) and band rows (4) must not lose those pixels (found by the randomised campaign)."""
    w, h = 200, 140
    rows = []
    for k, edge in enumerate([31.95, 63.6, 95.99, 3.9, 7.51, 127.75]):
        for j, slope in enumerate([0.0, 0.45, -0.45, 0.95]):
            x0 = 5.3 + 7.1 * j + 3.3 * k
            rows.append(((x0, edge - 20 * slope), (x0 + 20.0, edge)))          # non-steep, ends next to a tile / band row
            rows.append(((edge + 0.2, 8.2 + 5 * j), (edge + 0.2 + 9 * slope, 40.7 + 5 * j)))  # steep, next to a tile column
    v = np.zeros((2 * len(rows), 8), np.float32)
    for i, (a, b) in enumerate(rows):
        v[2 * i, :2], v[2 * i + 1, :2] = a, b
    v[:, 2] = -1.0 - 0.01 * np.arange(len(v))
    v[:, 3] = 1.0
    v[:, 4:] = np.random.default_rng(7).uniform(0.2, 1.0, (len(v), 4))
    idx = np.arange(len(v), dtype=np.uint32)
    for blend in (sr.BLEND_ALPHA_OVER, sr.BLEND_REPLACE):
        out, win, _, ofb = run_both_screen(P, ctx, w, h, v, idx, prim=sr.LINE, aa=True, blend=blend)
        assert np.array_equal(win, ofb.winner)
        H.compare_framebuffers(out, ofb, exact_color=True, what="antialiased lines at tile boundaries")


def test_children_may_outlive_their_context(P):
    """A garbage collector destroys handles in arbitrary order: children released after sr_context_destroy must still find a
    valid context (every device buffer keeps it alive), and a second destroy of the context is an error, not a crash."""
    from softrender_b200._abi import SoftrenderError
    c2 = P.Context(0)
    fb = P.RenderBuffer.with_dimensions(c2, 64, 48)
    mesh = P.Mesh(c2, H.suzanne_mesh())
    pipe = P.Pipeline.from_framebuffer(fb, scenes.suzanne_uniforms(64, 48))
    fb.clear(H.CLEAR)
    pipe.render_mesh(sr.TRIANGLE, mesh).run_to_fragment(scenes.Viewport.new(64, 48, 0.001, 1000.0), sr.VS_SUZANNE).run(sr.FS_SUZANNE)
    st = pipe.render_mesh(sr.TRIANGLE, mesh).run(sr.VS_SUZANNE)  # a draw left unfinished
    assert (fb.download()[:, :3].max(axis=1) > 0.05).sum() > 50
    c2.close()
    for x in (st, pipe, mesh, fb):
        x.destroy()


def test_duplicate_and_with_framebuffer(P, ctx):
    """`duplicate()` (vertex.rs:44, geometry.rs:43, fragment.rs:122) re-uses one stage's output for several passes; both
    copies must stay usable and independent (finish() normalises a shared position buffer copy-on-write).
    `Pipeline::with_framebuffer` (mod.rs:126-140) swaps the target and RESETS the stencil configuration."""
    w, h = 120, 90
    mesh = H.suzanne_mesh()
    u = scenes.suzanne_uniforms(w, h)
    vp = scenes.Viewport.new(w, h, 0.001, 1000.0)
    vp2 = scenes.Viewport.new(w // 2, h // 2, 0.001, 1000.0)  # second pass into the top-left quarter
    fb = make_fb(P, ctx, w, h)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    gmesh = P.Mesh(ctx, mesh)
    gs = pipe.render_mesh(sr.TRIANGLE, gmesh).run(sr.VS_SUZANNE)
    dup = gs.duplicate()
    gs.clip_primitives().finish(vp).run(sr.FS_SUZANNE)
    dup.finish(vp2).with_blend(sr.BLEND_ALPHA_OVER).run(sr.FS_SUZANNE)
    ofb = oracle_fb(w, h)
    for view, clip, blend in ((vp, True, sr.BLEND_REPLACE), (vp2, False, sr.BLEND_ALPHA_OVER)):
        od = ob.OracleDraw(sr.TRIANGLE, mesh.indices)
        od.blend = blend
        od.vertex_run(sr.VS_SUZANNE, u, mesh.vertices)
        if clip:
            od.clip_primitives()
        od.finish(view).fragment_run(ofb, sr.FS_SUZANNE, u)
    assert np.array_equal(fb.download_winner(), ofb.winner)
    H.compare_framebuffers(fb.download(), ofb, color_tol=COLOR_TOL, what="duplicate")
    # with_framebuffer: a configuration that would reject everything is reset to Always / Keep
    fbs = make_fb(P, ctx, w, h, stencil=True)
    pipe.set_stencil_config(sr.STENCIL_NEVER, sr.STENCIL_ZERO)
    pipe.with_framebuffer(fbs)
    pipe.render_mesh(sr.TRIANGLE, gmesh, stencil=3).run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)
    ofs = oracle_fb(w, h, stencil=True)
    od = ob.OracleDraw(sr.TRIANGLE, mesh.indices, 3)
    od.vertex_run_to_fragment(vp, sr.VS_SUZANNE, u, mesh.vertices).fragment_run(ofs, sr.FS_SUZANNE, u, sr.STENCIL_ALWAYS, sr.STENCIL_KEEP)
    assert np.array_equal(fbs.download_winner(), ofs.winner) and (ofs.winner > 0).sum() > 100
    H.compare_framebuffers(fbs.download(), ofs, color_tol=COLOR_TOL, what="with_framebuffer")
    for x in (pipe, gmesh, fb, fbs):
        x.destroy()


def test_error_behaviour(P, ctx):
    from softrender_b200._abi import SoftrenderError
    u = scenes.suzanne_uniforms(8, 8)
    fb0 = P.RenderBuffer.with_dimensions(ctx, 0, 4)
    with pytest.raises(SoftrenderError):  # assert!(width > 0) (src/pipeline/mod.rs:129)
        P.Pipeline.from_framebuffer(fb0, u)
    fb = make_fb(P, ctx, 8, 8)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    mesh = P.Mesh(ctx, vertices=np.zeros((4, 6), np.float32), indices=np.array([0, 1, 2, 3], np.uint32))
    with pytest.raises(SoftrenderError):  # assert_eq!(indices.len() % num_vertices, 0) (mod.rs:148)
        pipe.render_mesh(sr.TRIANGLE, mesh)
    with pytest.raises(SoftrenderError) as e:  # RenderError::InvalidPixelCoordinate (src/error.rs:9)
        fb.pixel(8, 0)
    assert e.value.status == sr.ERR_INVALID_PIXEL_COORDINATE
    rgba, depth, _ = fb.pixel(7, 7)
    assert np.float32(depth).view(np.uint32) == sr.DEPTH_FAR_BITS and np.allclose(rgba, H.CLEAR)
    for x in (pipe, mesh, fb, fb0):
        x.destroy()


# ------------------------------------------------------------------------------------------------------
# render-to-texture: a framebuffer's colour sampled in place by a later pass
# (TextureBufferRef, src/framebuffer/texturebuffer.rs:12-58) with every Filter / Edge of src/texture.rs:21-45
# ------------------------------------------------------------------------------------------------------
def _uv_triangles(rng, n, w, h, lo=-0.8, hi=1.9):
    """n random screen-space triangles whose K is a texture coordinate reaching well outside [0, 1]."""
    v = H.random_screen_triangles(rng, n, w, h, nk=2, max_size=0.6 * min(w, h))
    v[:, 4:6] = rng.uniform(lo, hi, (3 * n, 2)).astype(np.float32)
    return v


@pytest.mark.parametrize("blend", [sr.BLEND_REPLACE, sr.BLEND_ALPHA_OVER])
@pytest.mark.parametrize("edge", [sr.EDGE_CLAMP, sr.EDGE_WRAP, sr.EDGE_BORDER])
@pytest.mark.parametrize("filt", [sr.FILTER_NEAREST, sr.FILTER_BILINEAR])
def test_render_to_texture(P, ctx, filt, edge, blend):
    """Pass 1 renders flat triangles into A; pass 2 draws uv-mapped triangles into B with the texture_unlit shader
    sampling A's colour attachment in place.  f32 texels, no transcendental: colours are bit-exact."""
    rng = np.random.default_rng(1000 + 10 * filt + edge)
    wa, ha, wb, hb = 97, 61, 230, 170
    border = (0.25, 0.5, 0.75, 0.5)
    # pass 1 (alpha of the flat colours is random: pass 2's alpha_over then really blends)
    va = H.random_screen_triangles(rng, 120, wa, ha)
    ia = np.arange(va.shape[0], dtype=np.uint32)
    out_a, _, _, ofa = run_both_screen_keep(P, ctx, wa, ha, va, ia)
    fba, pipe_a = out_a
    H.compare_framebuffers(fba.download(), ofa, exact_color=True, what="pass 1")
    # pass 2
    vb = _uv_triangles(rng, 60, wb, hb)
    ib = np.arange(vb.shape[0], dtype=np.uint32)
    u = scenes.suzanne_uniforms(wb, hb)
    ofb = oracle_fb(wb, hb)
    od = ob.OracleDraw(sr.TRIANGLE, ib, None)
    od.set_vertices(vb, 1)
    od.cull, od.blend, od.aa, od.tile = sr.CULL_NONE, blend, False, None
    od.fragment_run(ofb, sr.FS_TEXTURE_UNLIT, u, texture=ofa.color.reshape(ha, wa, 4), sampler=(filt, edge, border))
    fbb = make_fb(P, ctx, wb, hb)
    pipe = P.Pipeline.from_framebuffer(fbb, u)
    pipe.bind_framebuffer_texture(fba)
    pipe.set_sampler(filt, edge, border)
    pipe.draw_from_vertices(sr.TRIANGLE, vb, ib, 1).with_blend(blend).run(sr.FS_TEXTURE_UNLIT)
    assert np.array_equal(fbb.download_winner(), ofb.winner)
    H.compare_framebuffers(fbb.download(), ofb, exact_color=True, what=f"render-to-texture filter {filt} edge {edge} blend {blend}")
    # the source is untouched by being sampled
    H.compare_framebuffers(fba.download(), ofa, exact_color=True, what="source after sampling")
    for x in (pipe, fbb, pipe_a, fba):
        x.destroy()


def run_both_screen_keep(P, ctx, w, h, verts, indices, fs=sr.FS_FLAT):
    """One opaque screen-space draw on both sides; returns the live GPU framebuffer + pipeline and the oracle's."""
    u = scenes.suzanne_uniforms(w, h)
    ofb = oracle_fb(w, h)
    fb = make_fb(P, ctx, w, h)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    od = ob.OracleDraw(sr.TRIANGLE, indices, None)
    od.set_vertices(verts, 1)
    od.cull, od.blend, od.aa, od.tile = sr.CULL_NONE, sr.BLEND_REPLACE, False, None
    od.fragment_run(ofb, fs, u)
    pipe.draw_from_vertices(sr.TRIANGLE, verts, indices, 1).run(fs)
    return (fb, pipe), None, None, ofb


@pytest.mark.parametrize("edge", [sr.EDGE_CLAMP, sr.EDGE_WRAP, sr.EDGE_BORDER])
@pytest.mark.parametrize("filt", [sr.FILTER_NEAREST, sr.FILTER_BILINEAR])
def test_image_texture_sampler_modes(P, ctx, filt, edge):
    """The same Filter / Edge set on an 8-bit image texture (texel / 255, decode_gamma: full_example/src/texture.rs:33-40,84);
    coverage and depth bit-exact, colour within 1/255 (powf)."""
    rng = np.random.default_rng(2000 + 10 * filt + edge)
    w, h = 210, 140
    img = rng.integers(0, 256, (23, 37, 4), dtype=np.uint8)
    border = (0.9, 0.1, 0.3, 1.0)
    vb = _uv_triangles(rng, 50, w, h)
    ib = np.arange(vb.shape[0], dtype=np.uint32)
    u = scenes.suzanne_uniforms(w, h)
    ofb = oracle_fb(w, h)
    od = ob.OracleDraw(sr.TRIANGLE, ib, None)
    od.set_vertices(vb, 1)
    od.cull, od.blend, od.aa, od.tile = sr.CULL_NONE, sr.BLEND_REPLACE, False, None
    od.fragment_run(ofb, sr.FS_TEXTURE_UNLIT, u, texture=img, sampler=(filt, edge, border))
    fb = make_fb(P, ctx, w, h)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    tex = P.Texture(ctx, img)
    pipe.bind_texture(tex)
    pipe.set_sampler(filt, edge, border)
    pipe.draw_from_vertices(sr.TRIANGLE, vb, ib, 1).run(sr.FS_TEXTURE_UNLIT)
    assert np.array_equal(fb.download_winner(), ofb.winner)
    out = fb.download()
    H.assert_bits_equal(out[:, 4], ofb.depth, "image sampler depth")
    # Bilinear + Wrap extrapolates for negative coordinates (fract keeps the sign, the ratio goes negative): colours leave
    # [0, 1] (NaN where decode_gamma meets a negative value), so the 1/255 bar is taken relative to the magnitude there
    assert np.array_equal(np.isnan(out[:, :4]), np.isnan(ofb.color))
    err = np.nan_to_num(np.abs(out[:, :4].astype(np.float64) - ofb.color), nan=0.0)
    assert np.all(err <= COLOR_TOL * np.maximum(1.0, np.nan_to_num(np.abs(ofb.color), nan=0.0))), f"image sampler filter {filt} edge {edge}: {err.max()}"
    for x in (pipe, tex, fb):
        x.destroy()


def test_render_to_texture_lit_scene_and_state(P, ctx):
    """The shipped textured shader (full_example/src/shaders.rs:108-162) fed from a render target instead of an image;
    a source whose clear is still only recorded; binding rules."""
    from softrender_b200._abi import SoftrenderError
    w, h = 320, 200
    mesh = H.suzanne_mesh(with_uv=True)
    u = scenes.full_example_uniforms(w / h, np.deg2rad(75.0), 2.0, np.deg2rad(45.0), np.deg2rad(65.0), 0.0)
    # source: a cleared-only framebuffer (the clear is lazy: it must be materialised before it is sampled) ...
    src = P.RenderBuffer.with_dimensions(ctx, 16, 9)
    src.clear((0.2, 0.4, 0.6, 0.8))
    osrc = ob.OracleFramebuffer(16, 9)
    osrc.clear((0.2, 0.4, 0.6, 0.8))
    rng = np.random.default_rng(77)
    vb = _uv_triangles(rng, 30, w, h)
    ib = np.arange(vb.shape[0], dtype=np.uint32)
    us = scenes.suzanne_uniforms(w, h)
    ofb = oracle_fb(w, h)
    od = ob.OracleDraw(sr.TRIANGLE, ib, None)
    od.set_vertices(vb, 1)
    od.cull, od.blend, od.aa, od.tile = sr.CULL_NONE, sr.BLEND_REPLACE, False, None
    od.fragment_run(ofb, sr.FS_TEXTURE_UNLIT, us, texture=osrc.color.reshape(9, 16, 4), sampler=(sr.FILTER_BILINEAR, sr.EDGE_WRAP, None))
    fb = make_fb(P, ctx, w, h)
    pipe = P.Pipeline.from_framebuffer(fb, us)
    with pytest.raises(SoftrenderError):  # no texture bound
        pipe.draw_from_vertices(sr.TRIANGLE, vb, ib, 1).run(sr.FS_TEXTURE_UNLIT)
    pipe.bind_framebuffer_texture(fb)
    with pytest.raises(SoftrenderError) as e:  # a target cannot be its own texture (&mut vs & borrow in the reference)
        pipe.draw_from_vertices(sr.TRIANGLE, vb, ib, 1).run(sr.FS_TEXTURE_UNLIT)
    assert e.value.status == sr.ERR_INVALID_STATE
    with pytest.raises(SoftrenderError):
        pipe.set_sampler(2, 0)
    with pytest.raises(SoftrenderError):
        pipe.set_sampler(0, 3)
    pipe.bind_framebuffer_texture(src)
    pipe.set_sampler(sr.FILTER_BILINEAR, sr.EDGE_WRAP)
    pipe.draw_from_vertices(sr.TRIANGLE, vb, ib, 1).run(sr.FS_TEXTURE_UNLIT)
    H.compare_framebuffers(fb.download(), ofb, exact_color=True, what="cleared-only source")
    # ... then the lit, textured full_example shader sampling a rendered target (uv of the mesh), Nearest + Clamp
    rs = np.random.default_rng(5)
    va = H.random_screen_triangles(rs, 80, 64, 48)
    ia = np.arange(va.shape[0], dtype=np.uint32)
    (fba, pipe_a), _, _, ofa = run_both_screen_keep(P, ctx, 64, 48, va, ia)
    vp = scenes.Viewport.new(w, h, 0.1, 1000.0)
    ofb2 = oracle_fb(w, h)
    od2 = ob.OracleDraw(sr.TRIANGLE, mesh.indices)
    od2.vertex_run_to_fragment(vp, sr.VS_FULL_EXAMPLE, u, mesh.vertices)
    od2.fragment_run(ofb2, sr.FS_FULL_EXAMPLE_TEXTURED, u, texture=ofa.color.reshape(48, 64, 4), sampler=(sr.FILTER_NEAREST, sr.EDGE_CLAMP, None))
    fb2 = make_fb(P, ctx, w, h)
    pipe2 = P.Pipeline.from_framebuffer(fb2, u)
    pipe2.bind_framebuffer_texture(fba)
    pipe2.set_sampler(sr.FILTER_NEAREST, sr.EDGE_CLAMP)
    gm = P.Mesh(ctx, mesh)
    pipe2.render_mesh(sr.TRIANGLE, gm).run_to_fragment(vp, sr.VS_FULL_EXAMPLE).run(sr.FS_FULL_EXAMPLE_TEXTURED)
    assert np.array_equal(fb2.download_winner(), ofb2.winner)
    H.compare_framebuffers(fb2.download(), ofb2, color_tol=COLOR_TOL, what="lit scene textured from a render target")
    for x in (pipe2, gm, fb2, pipe_a, fba):
        x.destroy()
    for x in (pipe, fb, src):
        x.destroy()


def test_render_to_texture_full_screen_pass_through_the_builder_chain(P, ctx):
    """The post-processing shape of the feature (what bench.py times): pass 1 renders Suzanne, pass 2 is a clip-space quad
    through the passthrough vertex shader (Vin = clip xyzw + uv) and run_to_fragment, texture_unlit sampling pass 1 in place."""
    w, h = 300, 180
    mesh = H.suzanne_mesh()
    u = scenes.suzanne_uniforms(w, h)
    vp = scenes.Viewport.new(w, h, 0.001, 1000.0)
    ofa, fba = oracle_fb(w, h), make_fb(P, ctx, w, h)
    ob.OracleDraw(sr.TRIANGLE, mesh.indices).vertex_run_to_fragment(vp, sr.VS_SUZANNE, u, mesh.vertices).fragment_run(ofa, sr.FS_SUZANNE, u)
    pa = P.Pipeline.from_framebuffer(fba, u)
    gm = P.Mesh(ctx, mesh)
    pa.render_mesh(sr.TRIANGLE, gm).run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)
    quad = np.array([[-1, -1, 0, 1, 0, 1], [1, -1, 0, 1, 1, 1], [1, 1, 0, 1, 1, 0], [-1, 1, 0, 1, 0, 0]], np.float32)
    qi = np.array([0, 1, 2, 0, 2, 3], np.uint32)
    # the oracle samples ITS pass-1 image (lit colours agree within 1/255 only, so pass 2 inherits that tolerance);
    # sampling the GPU's own pass-1 pixels instead makes pass 2 bit-exact
    src = fba.download()[:, :4].reshape(h, w, 4).copy()
    for filt in (sr.FILTER_NEAREST, sr.FILTER_BILINEAR):
        ofb = oracle_fb(w, h)
        ob.OracleDraw(sr.TRIANGLE, qi).vertex_run_to_fragment(vp, sr.VS_PASSTHROUGH, u, quad).fragment_run(
            ofb, sr.FS_TEXTURE_UNLIT, u, texture=src, sampler=(filt, sr.EDGE_CLAMP, None))
        fbb = make_fb(P, ctx, w, h)
        pb = P.Pipeline.from_framebuffer(fbb, u)
        pb.bind_framebuffer_texture(fba)
        pb.set_sampler(filt, sr.EDGE_CLAMP)
        qm = P.Mesh(ctx, vertices=quad, indices=qi)
        pb.render_mesh(sr.TRIANGLE, qm).run_to_fragment(vp, sr.VS_PASSTHROUGH).run(sr.FS_TEXTURE_UNLIT)
        out = fbb.download()
        assert np.array_equal(fbb.download_winner(), ofb.winner) and int((ofb.winner > 0).sum()) == w * h  # every pixel, once
        H.compare_framebuffers(out, ofb, exact_color=True, what=f"full-screen second pass filter {filt}")
        for x in (pb, qm, fbb):
            x.destroy()
    H.compare_framebuffers(fba.download(), ofa, color_tol=COLOR_TOL, what="pass 1")
    for x in (pa, gm, fba):
        x.destroy()


# ------------------------------------------------------------------------------------------------------
# draws of a handful of triangles onto a plain RenderBuffer WITHOUT the winner plane: k_tile_few (one launch, no lists)
# ------------------------------------------------------------------------------------------------------
def _few_fb(P, ctx, w, h):
    fb = P.RenderBuffer.with_dimensions(ctx, w, h)  # (no enable_winner: the introspection plane selects the general path)
    fb.clear(H.CLEAR)
    return fb


@pytest.mark.parametrize("w,h", [(300, 180), (256, 128), (67, 45)])
@pytest.mark.parametrize("cull", [sr.CULL_NONE, sr.CLOCKWISE])
def test_few_triangles_direct_path(P, ctx, w, h, cull):
    """Up to eight triangles of any size (k_tile_few): big ones that cover whole tiles, slivers, triangles that stick out of
    the frame, exact depth ties between them (the later one wins), a vertex at z >= 0, NaN coordinates; drawn twice -- onto a
    recorded clear and then, with other triangles, onto the existing contents (stored depth = the bar to pass).  Frame sizes
    with partial tiles and widths that are not a multiple of four (no 16-byte stores).  Bit-exact against the oracle."""
    rng = np.random.default_rng(w * 7 + h + cull)
    u = scenes.suzanne_uniforms(w, h)
    fb = _few_fb(P, ctx, w, h)
    ofb = oracle_fb(w, h)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    for draw in range(3):
        n = int(rng.integers(1, 9))
        verts = H.random_screen_triangles(rng, n, w, h, max_size=0.9 * max(w, h), integer_depth=(draw == 1), margin=0.3)
        if draw == 2 and n >= 3:
            verts[0:3, 2] = [0.5, -1.0, -2.0]     # a triangle crossing z = 0: fragments with z >= 0 are dropped (triangle.rs:120)
            verts[3, 0] = np.nan                  # the reference panics on NaN; defined as "skipped" 
        idx = np.arange(3 * n, dtype=np.uint32)
        od = ob.OracleDraw(sr.TRIANGLE, idx)
        od.set_vertices(verts, 1)
        od.cull = cull
        od.fragment_run(ofb, sr.FS_FLAT, u)
        pipe.draw_from_vertices(sr.TRIANGLE, verts, idx, 1).cull_faces(cull).run(sr.FS_FLAT)
        H.compare_framebuffers(fb.download(), ofb, exact_color=True, what=f"few triangles, draw {draw}")
    for x in (pipe, fb):
        x.destroy()


@pytest.mark.parametrize("filt", [sr.FILTER_NEAREST, sr.FILTER_BILINEAR])
@pytest.mark.parametrize("edge", [sr.EDGE_CLAMP, sr.EDGE_WRAP, sr.EDGE_BORDER])
def test_full_screen_pass_direct_path(P, ctx, filt, edge):
    """The render-to-texture second pass as an application issues it (no winner plane): a quad of two triangles with texture
    coordinates beyond [0, 1], texture_unlit sampling a rendered target in place, every Filter x Edge -- through k_tile_few."""
    rng = np.random.default_rng(31 + filt * 3 + edge)
    w, h = 328, 200
    u = scenes.suzanne_uniforms(w, h)
    vp = scenes.Viewport.new(w, h, 0.1, 10.0)
    src_fb = make_fb(P, ctx, w, h)
    p1 = P.Pipeline.from_framebuffer(src_fb, u)
    verts = H.random_screen_triangles(rng, 300, w, h, max_size=40.0)
    p1.draw_from_vertices(sr.TRIANGLE, verts, np.arange(900, dtype=np.uint32), 1).run(sr.FS_FLAT)
    src = src_fb.download()[:, :4].reshape(h, w, 4).copy()
    quad = np.array([[-1, -1, 0, 1, -0.3, 1.2], [1, -1, 0, 1, 1.4, 1.2], [1, 1, 0, 1, 1.4, -0.1], [-1, 1, 0, 1, -0.3, -0.1]], np.float32)
    qi = np.array([0, 1, 2, 0, 2, 3], np.uint32)
    border = (0.25, 0.5, 0.75, 1.0)
    fb = _few_fb(P, ctx, w, h)
    p2 = P.Pipeline.from_framebuffer(fb, u)
    p2.bind_framebuffer_texture(src_fb)
    p2.set_sampler(filt, edge, border)
    qm = P.Mesh(ctx, vertices=quad, indices=qi)
    p2.render_mesh(sr.TRIANGLE, qm).run_to_fragment(vp, sr.VS_PASSTHROUGH).run(sr.FS_TEXTURE_UNLIT)
    ofb = oracle_fb(w, h)
    od = ob.OracleDraw(sr.TRIANGLE, qi)
    od.vertex_run_to_fragment(vp, sr.VS_PASSTHROUGH, u, quad)
    od.fragment_run(ofb, sr.FS_TEXTURE_UNLIT, u, texture=src, sampler=(filt, edge, border))
    H.compare_framebuffers(fb.download(), ofb, exact_color=True, what="full-screen pass, direct path")
    for x in (p2, qm, fb, p1, src_fb):
        x.destroy()


@pytest.mark.parametrize("fs,nk", [(sr.FS_SUZANNE, 8), (sr.FS_FULL_EXAMPLE, 8), (sr.FS_FULL_EXAMPLE_TEXTURED, 10), (sr.FS_GREEN, 4)])
def test_few_triangles_direct_path_lit_shaders(P, ctx, fs, nk):
    """k_tile_few with the registered lit shaders (attribute planes carried in the per-CTA records: 2 planes, 3 with texture
    coordinates) and an image texture: depth bit-exact, colour within 1/255 where the oracle's colour is a number (random
    "normals" can drive a power's base negative -- NaN in both)."""
    rng = np.random.default_rng(500 + fs)
    w, h = 200, 136
    u = scenes.full_example_uniforms(w / h, np.deg2rad(75.0), 2.0, np.deg2rad(45.0), np.deg2rad(65.0), 0.0) if fs in (
        sr.FS_FULL_EXAMPLE, sr.FS_FULL_EXAMPLE_TEXTURED) else scenes.suzanne_uniforms(w, h)
    fb = _few_fb(P, ctx, w, h)
    ofb = oracle_fb(w, h)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    tex_img = scenes.checker_texture(64, 8)
    tex = P.Texture(ctx, tex_img)
    if fs == sr.FS_FULL_EXAMPLE_TEXTURED:
        pipe.bind_texture(tex)
        pipe.set_sampler(sr.FILTER_BILINEAR, sr.EDGE_CLAMP)
    n = 8
    verts = H.random_screen_triangles(rng, n, w, h, nk=nk, max_size=1.5 * w, margin=0.1)
    verts[:, 8:11] = verts[:, 8:11] * 2.0 - 1.0 if nk >= 8 else verts[:, 8:11]  # normals with both signs
    idx = np.arange(3 * n, dtype=np.uint32)
    od = ob.OracleDraw(sr.TRIANGLE, idx)
    od.set_vertices(verts, 1)
    kw = dict(texture=tex_img, sampler=(sr.FILTER_BILINEAR, sr.EDGE_CLAMP, None)) if fs == sr.FS_FULL_EXAMPLE_TEXTURED else {}
    od.fragment_run(ofb, fs, u, **kw)
    pipe.draw_from_vertices(sr.TRIANGLE, verts, idx, 1).run(fs)
    out = fb.download()
    H.assert_bits_equal(out[:, 4], ofb.depth, "depth")
    col = out[:, :4]
    assert np.array_equal(np.isnan(col), np.isnan(ofb.color))
    ok = np.isfinite(ofb.color) & np.isfinite(col)
    assert np.abs(col - ofb.color)[ok].max() <= COLOR_TOL
    assert (ofb.depth > np.float32(-3e38)).mean() > 0.02
    for x in (pipe, tex, fb):
        x.destroy()
