"""CPU tests: pin the oracle against the reference's own test (pixel index layout) and the known-answer
vectors derived line by line from the reference source, and against the reference's
golden image examples/suzanne.png (committed as a 4x box-downsampled fixture)."""
import numpy as np
import pytest

import softrender_b200 as sr
from softrender_b200 import scenes

import helpers as H
import oracle_binding as ob

L = ob.lib()


def test_coordinate_index_reference_test():
    """src/geometry/coordinate.rs:71-85 (`coordinate_index`): row-major x + y*width on a 10x20 frame."""
    for y in range(20):
        for x in range(10):
            assert L.so_coordinate_index(x, y, 10) == x + 10 * y


def test_depth_far_is_f32_min():
    assert np.float32(L.so_depth_far()).view(np.uint32) == 0xFF7FFFFF  # depth.rs:31


@pytest.mark.parametrize("w,h,count,last", [(1024, 1024, 64, (896, 896, 1023, 1023)), (2000, 2000, 256, None),
                                            (1920, 1080, 135, None), (3840, 2160, 510, (3712, 2048, 3839, 2159)),
                                            (7680, 4320, 2040, None), (800, 600, 35, None), (129, 129, 1, (0, 0, 128, 128)),
                                            (1, 1, 0, None)])
def test_tile_lists(w, h, count, last):
    """fragment.rs:188-216 at the default 128x128 tile size."""
    t = ob.tiles(w, h)
    assert len(t) == count
    if last:
        assert tuple(t[-1]) == last
    if count:
        covered = np.zeros((h, w), bool)
        for x0, y0, x1, y1 in t:
            covered[y0:y1 + 1, x0:x1 + 1] = True
        assert covered.all()


def test_stencil_truth_table():
    """src/stencil.rs:112-123 (test compares MESH value `mask` against BUFFER value) and :147-158."""
    for value in (0, 1, 2, 254, 255):
        for mask in (0, 1, 2, 255):
            exp = {sr.STENCIL_ALWAYS: True, sr.STENCIL_NEVER: False, sr.STENCIL_LESS_THAN: mask < value,
                   sr.STENCIL_LESS_THAN_EQ: mask <= value, sr.STENCIL_GREATER_THAN: mask > value,
                   sr.STENCIL_GREATER_THAN_EQ: mask >= value, sr.STENCIL_EQUAL: mask == value,
                   sr.STENCIL_NOT_EQUAL: mask != value}
            for t, e in exp.items():
                assert bool(L.so_stencil_test(t, value, mask)) == e
            ops = {sr.STENCIL_KEEP: value, sr.STENCIL_INVERT: (~value) & 255, sr.STENCIL_ZERO: 0, sr.STENCIL_REPLACE: mask,
                   sr.STENCIL_INCREMENT_WRAP: (value + 1) & 255, sr.STENCIL_DECREMENT_WRAP: (value - 1) & 255,
                   sr.STENCIL_INCREMENT_SAT: min(value + 1, 255), sr.STENCIL_DECREMENT_SAT: max(value - 1, 0)}
            for o, e in ops.items():
                assert L.so_stencil_op(o, value, mask) == e


def _clip(a, b, c):
    v = np.zeros((3, 8), np.float32)
    v[0, :4], v[1, :4], v[2, :4] = a, b, c
    v[:, 4:] = np.arange(12, dtype=np.float32).reshape(3, 4)
    d = ob.OracleDraw(sr.TRIANGLE, np.arange(3, dtype=np.uint32))
    d.set_vertices(v, 0).clip_primitives()
    t = d.data(3).reshape(-1, 3, 8)

    def nondeg(tri):
        b_ = [x.tobytes() for x in tri]
        return len(set(b_)) == 3
    return v, t, [i for i in range(len(t)) if nondeg(t[i])]


def test_clipper_known_answers():
    """geometry.rs:261-299 literally: the clipper amplifies instead of clipping."""
    A, B, C = (-0.5, -0.5, 0.5, 1.0), (0.5, -0.5, 0.5, 1.0), (0.0, 0.5, 0.5, 1.0)
    v, t, nd = _clip(A, B, C)  # fully inside: polygon of 18 -> 16 triangles, one real = (a,b,c) at fan index 5
    assert len(t) == 16 and nd == [5]
    assert np.array_equal(t[5], v)
    v, t, nd = _clip((-2.0, -0.5, 0.5, 1.0), B, C)  # a outside Left only: 19 -> 17, three non-degenerate
    assert len(t) == 17 and len(nd) == 3
    assert any(np.array_equal(t[i], v) for i in nd)  # the unclipped (a,b,c) survives
    assert all(np.array_equal(t[i][0], v[0]) for i in nd)  # fan pivot is the OUTSIDE vertex a
    _, t, _ = _clip((-2.0, -0.5, 0.5, 1.0), (-3.0, 0.2, 0.5, 1.0), C)  # a,b outside Left: 18 -> 16
    assert len(t) == 16
    v, t, nd = _clip((-2.0, -0.5, 0.5, 1.0), (-3.0, 0.2, 0.5, 1.0), (-2.5, 0.5, 0.5, 1.0))  # all outside Left: 15 -> 13
    assert len(t) == 13 and any(np.array_equal(t[i], v) for i in nd)
    _, t, nd = _clip((-2.0, -2.0, 0.5, 1.0), B, C)  # a outside Left+Top: 20 -> 18, five non-degenerate
    assert len(t) == 18 and len(nd) == 5
    _, t, nd = _clip((0.1, 0.1, -0.5, -1.0), B, C)  # a with w<0 (outside all six planes): 24 -> 22, seventeen non-degenerate
    assert len(t) == 22 and len(nd) >= 16  # the non-degenerate count depends on coincident intersections of the chosen vertex


def _raster(verts, idx, w=32, h=32, prim=sr.TRIANGLE, **kw):
    fb = ob.OracleFramebuffer(w, h, 8 if kw.get("stencil") else 0)
    fb.clear(H.CLEAR)
    d = ob.OracleDraw(prim, np.asarray(idx, np.uint32), kw.get("stencil_value"))
    d.set_vertices(np.asarray(verts, np.float32), 1)
    d.cull = kw.get("cull", 0)
    d.tile = kw.get("tile")
    d.fragment_run(fb, sr.FS_FLAT, scenes.suzanne_uniforms(w, h), kw.get("stencil_test", 0), kw.get("stencil_op", 0))
    return fb


def test_inclusive_edges_and_tie_break():
    z = -1.0
    red, green = [1, 0, 0, 1], [0, 1, 0, 1]
    tri = lambda pts, c: [[x, y, z, 1.0] + c for x, y in pts]  # noqa: E731
    verts = tri([(4.5, 4.5), (20.5, 4.5), (4.5, 20.5)], red) + tri([(20.5, 4.5), (20.5, 20.5), (4.5, 20.5)], green)
    fb = _raster(verts, range(6))
    W = fb.winner.reshape(32, 32)
    assert W[4, 4] == 1 and W[4, 19] == 1  # centres on triangle 1's own edges are covered
    assert all(W[k, 24 - k] == 2 for k in range(4, 21))  # shared diagonal: both cover, later wins at equal depth
    assert W[3, 4] == 0 and W[21, 21] == 0
    # single triangle: edge pixels covered on all three sides
    fb1 = _raster(verts[:3], range(3))
    W1 = fb1.winner.reshape(32, 32)
    assert all(W1[k, 24 - k] == 1 for k in range(4, 21)) and W1[4, 4:21].all() and W1[4:21, 4].all()


def test_cull_negative_zero_is_clockwise():
    p = [3.0, 3.0, -1.0, 1.0, 1, 1, 1, 1]
    big = [[2.5, 2.5, -1, 1, 1, 0, 0, 1], [12.5, 2.5, -1, 1, 1, 0, 0, 1], [2.5, 12.5, -1, 1, 1, 0, 0, 1]]
    # shoelace of identical points: 9+9+9-9-9-9 = +0.0 -> CounterClockwise; flipping the sign bit needs a real -0.0
    x1, y1, x2, y2, x3, y3 = -1.0, 0.0, 0.0, 0.0, 1.0, 0.0
    a = np.float32(x1 * y2 + x2 * y3 + x3 * y1 - x2 * y1 - x3 * y2 - x1 * y3)
    assert not np.signbit(a)
    fb = _raster(big, range(3), cull=sr.COUNTER_CLOCKWISE)
    ccw_drawn = fb.winner.max()
    fb = _raster(big, range(3), cull=sr.CLOCKWISE)
    cw_drawn = fb.winner.max()
    assert {int(ccw_drawn), int(cw_drawn)} == {0, 1}  # culled under exactly one winding
    del p


def test_det_zero_draws_nothing_but_applies_stencil():
    verts = [[10, 10, -1, 1, 1, 1, 1, 1], [20, 20, -1, 1, 1, 1, 1, 1], [30, 30, -1, 1, 1, 1, 1, 1]]
    fb = _raster(verts, range(3), w=40, h=40, stencil=True, stencil_value=1, stencil_test=sr.STENCIL_ALWAYS,
                 stencil_op=sr.STENCIL_INCREMENT_WRAP)
    assert fb.winner.max() == 0
    S = fb.stencil.reshape(40, 40)
    assert S[10:31, 10:31].min() == 1 and S.sum() == 21 * 21


def test_nonnegative_z_rejected_and_points_last_row_column():
    verts = [[2.5, 2.5, 0.0, 1, 1, 1, 1, 1], [12.5, 2.5, 0.0, 1, 1, 1, 1, 1], [2.5, 12.5, 0.0, 1, 1, 1, 1, 1]]
    assert _raster(verts, range(3)).winner.max() == 0
    pts = [[31.5, 5.5, -1, 1, 1, 1, 1, 1], [5.5, 31.5, -1, 1, 1, 1, 1, 1], [30.9, 30.9, -1, 1, 1, 1, 1, 1]]
    W = _raster(pts, range(3), prim=sr.POINT).winner.reshape(32, 32)
    assert W[5, 31] == 0 and W[31, 5] == 0 and W[30, 30] == 3
    assert _raster(verts, range(3), w=1, h=1).winner.max() == 0  # 1x1: no tiles at all


def test_reference_tiling_equals_one_tile_and_literal_equals_binned():
    """(i) overlapping 16x16 tiles == one frame tile for idempotent state; (ii) shuffled order with
    exact depth ties; (iii) bins derived from the clamped bbox cover every pixel that was drawn."""
    rng = np.random.default_rng(3)
    w, h, n = 64, 64, 120
    verts = H.random_screen_triangles(rng, n, w, h, integer_depth=True)
    idx = (rng.permutation(n).astype(np.uint32)[:, None] * 3 + np.arange(3, dtype=np.uint32)).reshape(-1)
    one = _raster(verts, idx, w, h)
    tiled = _raster(verts, idx, w, h, tile=(16, 16))
    assert np.array_equal(one.winner, tiled.winner)
    H.assert_bits_equal(one.color, tiled.color)
    H.assert_bits_equal(one.depth, tiled.depth)
    d = ob.OracleDraw(sr.TRIANGLE, idx)
    d.set_vertices(verts, 1)
    off, ids = d.bins(w, h, 16, 16)
    for tile in range(16):
        tx, ty = tile % 4, tile // 4
        winners = set(np.unique(one.winner.reshape(h, w)[ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16])) - {0}
        assert {x - 1 for x in winners} <= set(ids[off[tile]:off[tile + 1]].tolist())
        assert np.all(np.diff(ids[off[tile]:off[tile + 1]].astype(np.int64)) > 0)


def test_threaded_reference_structure_matches_canonical():
    """The timing mode (thread pool + atomic cursors + 128x128 tiles) renders the same image."""
    size = 192
    mesh = H.suzanne_mesh()
    u = scenes.suzanne_uniforms(size, size)
    vp = scenes.Viewport.new(size, size, 0.001, 1000.0)
    frames = []
    for nthreads, tile in ((1, None), (4, (128, 128))):
        fb = ob.OracleFramebuffer(size, size)
        fb.clear(H.CLEAR)
        d = ob.OracleDraw(sr.TRIANGLE, mesh.indices)
        d.tile = tile
        d.vertex_run(sr.VS_SUZANNE, u, mesh.vertices, nthreads).clip_primitives(1).finish(vp, nthreads)
        d.fragment_run(fb, sr.FS_SUZANNE, u, nthreads=nthreads)
        frames.append(fb)
    H.assert_bits_equal(frames[0].color, frames[1].color)
    H.assert_bits_equal(frames[0].depth, frames[1].depth)


def test_suzanne_golden_image():
    """examples/suzanne.png (reference golden, older revision): silhouette bbox equal after 4x box downsample
    (x 88..477, y 78..440 at 500^2), lit-coverage IoU >= 0.94, median colour error <= 2/255."""
    size = 500
    mesh = H.suzanne_mesh()
    u = scenes.suzanne_uniforms(size, size)
    vp = scenes.Viewport.new(size, size, 0.001, 1000.0)
    fb = ob.OracleFramebuffer(size, size)
    fb.clear(H.CLEAR)
    d = ob.OracleDraw(sr.TRIANGLE, mesh.indices)
    d.vertex_run(sr.VS_SUZANNE, u, mesh.vertices, 4).clip_primitives().finish(vp)
    assert len(d.data(3)) == 3 * 16 * 968
    d.tile = (128, 128)
    d.fragment_run(fb, sr.FS_SUZANNE, u, nthreads=8)
    img = np.clip(fb.color[:, :3].reshape(size, size, 3), 0, 1)
    gold = np.load(H.GOLDEN + "/suzanne_gold_500.npz")["rgb"].astype(np.float32) / 255.0
    lit_o = np.abs(img - 0.01).max(-1) > 0.02
    lit_g = np.abs(gold - 0.01).max(-1) > 0.02

    def bbox(m):
        ys, xs = np.nonzero(m)
        return np.array([xs.min(), xs.max(), ys.min(), ys.max()])

    assert np.abs(bbox(lit_g) - np.array([88, 477, 78, 440])).max() == 0
    assert np.abs(bbox(lit_o) - bbox(lit_g)).max() <= 1  # the golden is anti-aliased by the downsample
    assert (lit_o & lit_g).sum() / (lit_o | lit_g).sum() >= 0.94
    assert np.median(np.abs(img - gold)[lit_o & lit_g]) * 255 <= 2.0


def test_mesh_fixture_matches_tobj_indexing():
    m = H.suzanne_mesh()
    assert m.vertices.shape == (1966, 6) and m.ntris == 968


def test_grid_generator_counts():
    g = scenes.make_grid(10, 8, 4)
    assert g.ntris == 2 * 10 * 8 * 4 and len(g.vertices) == 4 * 11 * 9
    r = scenes.make_grid(10, 8, 4, reverse=True)
    assert np.array_equal(np.sort(g.indices.reshape(-1, 3), axis=0), np.sort(r.indices.reshape(-1, 3), axis=0))
    # config 3 / 4 sizes without building them
    assert 4 * 2 * 1250 * 1000 == 10_000_000 and 4 * 1251 * 1001 == 5_009_004
    assert 4 * 2 * 5000 * 2500 == 100_000_000 and 4 * 5001 * 2501 == 50_030_004


def test_texture_sampler_known_answers():
    """texture(t, coord, filter, edge) (src/texture.rs:14-45) with the arithmetic of full_example/src/texture.rs:25-84,
    worked by hand on a 3x2 texture: row 0 = 0, 10, 20 (red), row 1 = 100, 110, 120; green = 1, blue = 2, alpha = 3."""
    t = np.zeros((2, 3, 4), np.float32)
    t[..., 0] = np.array([[0, 10, 20], [100, 110, 120]], np.float32)
    t[..., 1], t[..., 2], t[..., 3] = 1, 2, 3
    N, B = 0, 1
    CLAMP, WRAP, BORDER = 0, 1, 2
    s = lambda u, v, f, e, b=None: ob.texture_sample(t, u, v, (f, e, b))
    # Nearest: x = round(u * 2), y = round(v * 1); round is half away from zero (texture.rs:53-54)
    assert s(0.0, 0.0, N, CLAMP)[0] == 0 and s(0.24, 0.0, N, CLAMP)[0] == 0 and s(0.25, 0.0, N, CLAMP)[0] == 10
    assert s(0.75, 0.49, N, CLAMP)[0] == 20 and s(0.74, 0.5, N, CLAMP)[0] == 110 and s(1.0, 1.0, N, CLAMP)[0] == 120
    assert list(s(0.5, 0.0, N, CLAMP)) == [10, 1, 2, 3]  # f32 colours come back as stored: no /255, no gamma
    # Clamp: (u.min(1).max(0), v.min(1).max(0)) (texture.rs:28); NaN.min(1) = 1 in Rust
    assert s(-3.0, 7.0, N, CLAMP)[0] == 100 and s(9.0, -1.0, N, CLAMP)[0] == 20 and s(float("nan"), 0.0, N, CLAMP)[0] == 20
    # Wrap: fract keeps the sign (texture.rs:29): 1.25 -> 0.25, -0.75 -> -0.75 -> round(-1.5) = -2 -> `as u32` saturates to 0
    assert s(1.25, 2.0, N, WRAP)[0] == 10 and s(3.5, 1.75, N, WRAP)[0] == 110 and s(-0.75, 0.0, N, WRAP)[0] == 0
    # Border(C): outside [0,1]^2 or NaN -> C, inside as Clamp (src/texture.rs:43-44)
    c = [0.5, 0.25, 0.125, 1.0]
    assert list(s(1.0001, 0.5, N, BORDER, c)) == c and list(s(0.5, -1e-9, B, BORDER, c)) == c and list(s(float("nan"), 0.5, B, BORDER, c)) == c
    assert s(1.0, 1.0, N, BORDER, c)[0] == 120 and s(0.0, 0.0, B, BORDER, c)[0] == s(0.0, 0.0, B, CLAMP)[0]
    # Bilinear (texture.rs:58-82): uu = u*2 + 0.5, x = floor(uu), ratio = uu - x; vv = v*1 + 0.5
    # u = 0.5, v = 0.25: uu = 1.5 -> x = 1, ur = 0.5; vv = 0.75 -> y = 0, vr = 0.75
    #   (10*0.5 + 20*0.5) * 0.25 + (110*0.5 + 120*0.5) * 0.75 = 3.75 + 86.25 = 90
    assert s(0.5, 0.25, B, CLAMP)[0] == 90.0
    # u = 0, v = 0: uu = 0.5 -> x = 0, ur = 0.5; vv = 0.5 -> y = 0, vr = 0.5: (0*.5 + 10*.5)*.5 + (100*.5 + 110*.5)*.5 = 55
    assert s(0.0, 0.0, B, CLAMP)[0] == 55.0
    # u = 1, v = 1: uu = 2.5 -> x = 2, x+1 = 3 clamped to the last column (the reference would index out of bounds):
    #   vv = 1.5 -> y = 1, y+1 clamped: every tap is texel (2,1) = 120
    assert s(1.0, 1.0, B, CLAMP)[0] == 120.0
    # Bilinear + Wrap below zero extrapolates: u = -0.5 -> fract = -0.5, uu = -0.5, floor = -1 -> `as u32` = 0, ratio = -0.5,
    # opposite = 1.5: row 0: 0*1.5 + 10*(-0.5) = -5; row 1: 100*1.5 + 110*(-0.5) = 95; v = 0 -> vr = 0.5: (-5 + 95) / 2 = 45
    assert s(-0.5, 0.0, B, WRAP)[0] == 45.0
    # image texels: u8 / 255 then decode_gamma on r, g, b only (texture.rs:33-40,84)
    img = np.zeros((1, 1, 4), np.uint8)
    img[0, 0] = (255, 128, 0, 64)
    r = ob.texture_sample(img, 0.3, 0.6, (N, CLAMP, None))
    exp = np.array([1.0, np.float32(128 / 255) ** np.float32(2.2), 0.0, np.float32(64) / np.float32(255)], np.float32)
    assert np.allclose(r, exp, rtol=1e-6, atol=0) and r[3] == exp[3]
    # default sampler state = `impl Default for Filter` / `for Edge`: Nearest + Clamp (src/texture.rs:27-31,43-45)
    chk = scenes.checker_texture(16, 4)
    for u, v in [(0.1, 0.9), (0.5, 0.5), (1.0, 0.0), (0.333, 0.777)]:
        assert np.array_equal(ob.texture_sample(chk, u, v, (N, CLAMP, None)), ob.texture_sample(chk, u, v, None))


def test_stencil_ops_on_wider_types():
    """StencilOp::op over the Stencil trait's wrapping / saturating / not (src/stencil.rs:9-60,147-158) at the bounds of u8,
    u16 and u32, worked by hand."""
    import softrender_b200 as sr
    f = ob.lib().so_stencil_op_wide
    for bits in (8, 16, 32):
        m = (1 << bits) - 1
        assert f(sr.STENCIL_INCREMENT_WRAP, m, 0, bits) == 0 and f(sr.STENCIL_DECREMENT_WRAP, 0, 0, bits) == m
        assert f(sr.STENCIL_INCREMENT_SAT, m, 0, bits) == m and f(sr.STENCIL_DECREMENT_SAT, 0, 0, bits) == 0
        assert f(sr.STENCIL_INCREMENT_SAT, m - 1, 0, bits) == m and f(sr.STENCIL_DECREMENT_SAT, 1, 0, bits) == 0
        assert f(sr.STENCIL_INVERT, 0, 0, bits) == m and f(sr.STENCIL_INVERT, 0x5A, 0, bits) == m ^ 0x5A
        assert f(sr.STENCIL_REPLACE, 3, m - 7, bits) == m - 7 and f(sr.STENCIL_ZERO, m, 1, bits) == 0 and f(sr.STENCIL_KEEP, m, 1, bits) == m
    assert f(sr.STENCIL_INCREMENT_WRAP, 255, 0, 16) == 256 and f(sr.STENCIL_INCREMENT_WRAP, 65535, 0, 32) == 65536


def test_two_colour_planes_in_the_oracle():
    """A texture buffer declared with two colour planes (declare_texture_buffer!, src/framebuffer/texturebuffer.rs:72-147): the
    fragment shader returns the tuple, set_pixel_unchecked stores each colour into its own plane (:141-147), clear takes the
    tuple (:181-197).  Known answers: one flat-depth triangle whose three vertices carry the same normal -- plane 1 holds that
    normal on every covered pixel and its clear colour elsewhere; plane 0 is what the one-plane Suzanne shader produces."""
    import softrender_b200 as sr
    from softrender_b200 import scenes
    w, h = 16, 12
    u = scenes.suzanne_uniforms(w, h)
    normal = np.float32([0.0, 0.6, 0.8, 0.0])
    verts = np.zeros((3, 12), np.float32)
    verts[:, :4] = [[2.0, 2.0, -1.0, 1.0], [13.0, 3.0, -1.0, 1.0], [6.0, 10.0, -1.0, 1.0]]
    verts[:, 4:8] = [[0.1, 0.2, 0.3, 1.0], [0.4, 0.1, 0.2, 1.0], [0.3, 0.5, 0.1, 1.0]]  # world positions
    verts[:, 8:12] = normal
    idx = np.arange(3, dtype=np.uint32)
    clear0, clear1 = (0.1, 0.2, 0.3, 1.0), (9.0, 8.0, 7.0, 6.0)
    two = ob.OracleFramebuffer(w, h, two_colors=True)
    two.clear(clear0, clear1)
    one = ob.OracleFramebuffer(w, h)
    one.clear(clear0)
    for fb, fs in ((two, sr.FS_SUZANNE_GBUFFER), (one, sr.FS_SUZANNE)):
        od = ob.OracleDraw(sr.TRIANGLE, idx)
        od.set_vertices(verts, 1)
        od.fragment_run(fb, fs, u)
    covered = two.winner != 0
    assert 20 < covered.sum() < w * h and np.array_equal(covered, one.winner != 0)
    assert np.array_equal(two.color, one.color) and np.array_equal(two.depth, one.depth)  # plane 0 and depth: the one-plane result
    # u + v + w is 1 only up to rounding, so the interpolated constant normal is within an ulp or two of the constant
    assert np.allclose(two.color1[covered], normal, rtol=0, atol=3e-7)
    assert np.array_equal(two.color1[~covered], np.tile(np.float32(clear1), ((~covered).sum(), 1)))
    # the shader's return type and the target's colour type must agree (a type error in the reference)
    od = ob.OracleDraw(sr.TRIANGLE, idx)
    od.set_vertices(verts, 1)
    with pytest.raises(Exception):
        od.fragment_run(one, sr.FS_SUZANNE_GBUFFER, u)
    with pytest.raises(Exception):
        od.fragment_run(two, sr.FS_SUZANNE, u)
