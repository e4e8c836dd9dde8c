"""RGBAu8Color render targets (src/color/predefined.rs:26; SURVEY.md section 8 f rank 1): an 8-byte AoS pixel
{r, g, b, a as u8, f32 depth} through clear, both tile kernels, read-back and the sharded composite, against the oracle's
u8 colour attachment.  A registered shader's f32 colour is stored as `(c * 255.0) as u8`; lines multiply the alpha channel
with the integer rule of src/color/helper.rs:36-42 (coverage cast to u8 first -- the reference's quirk, kept)."""
import numpy as np
import pytest

import softrender_b200 as sr
from softrender_b200 import scenes

import helpers as H
import oracle_binding as ob

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    from softrender_b200 import pipeline
    return pipeline


def _pair(P, ctx, w, h, stencil=False):
    fb = P.RenderBuffer.with_dimensions(ctx, w, h, stencil=stencil, u8_color=True)
    fb.enable_winner(True)
    fb.clear(H.CLEAR)
    ofb = ob.OracleFramebuffer(w, h, 8 if stencil else 0, u8_color=True)
    ofb.clear(H.CLEAR)
    return fb, ofb


def _compare(fb, ofb, what, tol=0):
    px = fb.download()
    assert px.dtype == fb.U8_PIXEL and px.nbytes == fb.width * fb.height * 8
    H.assert_bits_equal(px["depth"], ofb.depth, what + " depth")
    diff = np.abs(px["rgba"].astype(np.int16) - ofb.color.astype(np.int16))
    assert diff.max() <= tol, f"{what}: {int((diff > tol).sum())} channel values differ by more than {tol} (max {diff.max()})"
    col, dep, _ = fb.download_planes()
    assert col.dtype == np.uint8 and np.array_equal(col, px["rgba"]) and np.array_equal(dep.view(np.uint32), px["depth"].view(np.uint32))
    assert np.array_equal(fb.download_rgba8().reshape(-1, 4), px["rgba"])
    assert np.array_equal(fb.download_rgba8(abgr=True).reshape(-1, 4), px["rgba"][:, ::-1])


@pytest.mark.parametrize("n,max_size", [(400, None), (70_000, 3.0)])
def test_opaque_triangles_on_u8_target(P, ctx, n, max_size):
    """Small draw (k_bin_small lists) and big draw (k_micro + resolve); a second draw lands on existing contents."""
    rng = np.random.default_rng(5 + n)
    w, h = 330, 210
    fb, ofb = _pair(P, ctx, w, h)
    u = scenes.suzanne_uniforms(w, h)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    for k in range(2):
        verts = H.random_screen_triangles(rng, n, w, h, max_size=max_size, integer_depth=(k == 1))
        verts[:, 4:] = rng.uniform(-0.2, 1.3, (len(verts), 4))  # outside [0,1] too: the cast saturates
        idx = np.arange(3 * n, dtype=np.uint32)
        od = ob.OracleDraw(sr.TRIANGLE, idx)
        od.set_vertices(verts, 1)
        od.fragment_run(ofb, sr.FS_FLAT, u)
        pipe.draw_from_vertices(sr.TRIANGLE, verts, idx, 1).run(sr.FS_FLAT)
        assert np.array_equal(fb.download_winner(), ofb.winner)
        _compare(fb, ofb, f"u8 opaque draw {k}")
    assert len(np.unique(ofb.color)) > 100
    pipe.destroy()
    fb.destroy()


def test_lit_scene_on_u8_target(P, ctx):
    """Suzanne through the real shaders: coverage and depth bit-exact; the u8 colour may differ by one level where the
    f32 colours differ within the 1/255 parity tolerance (powf: libm vs SFU)."""
    size = 256
    mesh = H.suzanne_mesh()
    u = scenes.suzanne_uniforms(size, size)
    vp = scenes.Viewport.new(size, size, 0.001, 1000.0)
    fb, ofb = _pair(P, ctx, size, size)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    gm = P.Mesh(ctx, mesh)
    pipe.render_mesh(sr.TRIANGLE, gm).run(sr.VS_SUZANNE).clip_primitives().finish(vp).run(sr.FS_SUZANNE)
    od = ob.OracleDraw(sr.TRIANGLE, mesh.indices)
    od.vertex_run(sr.VS_SUZANNE, u, mesh.vertices).clip_primitives().finish(vp).fragment_run(ofb, sr.FS_SUZANNE, u)
    assert np.array_equal(fb.download_winner(), ofb.winner)
    _compare(fb, ofb, "suzanne on u8", tol=1)
    assert (ofb.winner > 0).sum() > 5000
    for x in (pipe, gm, fb):
        x.destroy()


@pytest.mark.parametrize("aa", [False, True])
def test_lines_and_points_on_u8_target(P, ctx, aa):
    """Lines: c.mul_alpha(ColorAlpha::from_scalar(alpha)) with u8 arithmetic (line.rs:100, color/mod.rs:26-33,
    helper.rs:36-42): the coverage truncates to 0 or 1, the alpha channel becomes (a8 * (cov / 255)) as u8.  Bresenham lines
    and points ride the opaque path, Wu lines the ordered one."""
    rng = np.random.default_rng(77 + aa)
    w, h, n = 160, 120, 150
    fb, ofb = _pair(P, ctx, w, h)
    u = scenes.suzanne_uniforms(w, h)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    tris = H.random_screen_triangles(rng, 50, w, h)
    lines = np.zeros((2 * n, 8), np.float32)
    lines[:, 0], lines[:, 1] = rng.uniform(-10, w + 10, 2 * n), rng.uniform(-10, h + 10, 2 * n)
    lines[:, 2], lines[:, 3], lines[:, 4:] = -rng.uniform(0.01, 0.09, 2 * n), 1, rng.uniform(0, 1, (2 * n, 4))
    pts = lines[:60].copy()
    idx = np.arange(150, dtype=np.uint32)
    od = ob.OracleDraw(sr.TRIANGLE, idx)
    od.set_vertices(tris, 1)
    od.set_generated(2, lines).set_generated(1, pts)
    od.aa = aa
    od.fragment_run(ofb, sr.FS_FLAT, u)
    fs = pipe.draw_from_vertices(sr.TRIANGLE, tris, idx, 1)
    fs.set_generated(2, lines)
    fs.set_generated(1, pts)
    fs.antialiased_lines(aa).run(sr.FS_FLAT)
    assert np.array_equal(fb.download_winner(), ofb.winner)
    _compare(fb, ofb, "u8 lines + points")
    is_line = (ofb.winner > 50) & (ofb.winner <= 50 + n)  # canonical order: 50 triangles, then the lines, then the points
    assert is_line.sum() > 500 and set(np.unique(ofb.color[is_line, 3]).tolist()) <= {0, 1}  # every line fragment's alpha is 0 or 1
    pipe.destroy()
    fb.destroy()


def test_stencil_and_discard_on_u8_target(P, ctx):
    """The strictly ordered kernel on a u8 target (active stencil, then a discarding shader)."""
    rng = np.random.default_rng(91)
    w, h, n = 140, 100, 200
    fb, ofb = _pair(P, ctx, w, h, stencil=True)
    u = scenes.suzanne_uniforms(w, h)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    pipe.set_stencil_config(sr.STENCIL_LESS_THAN_EQ, sr.STENCIL_INCREMENT_WRAP)
    for k, fsid in enumerate((sr.FS_FLAT, sr.FS_DISCARD_CHECKER)):
        verts = H.random_screen_triangles(rng, n, w, h)
        idx = np.arange(3 * n, dtype=np.uint32)
        od = ob.OracleDraw(sr.TRIANGLE, idx, 3)
        od.set_vertices(verts, 1)
        od.fragment_run(ofb, fsid, u, sr.STENCIL_LESS_THAN_EQ, sr.STENCIL_INCREMENT_WRAP)
        pipe.draw_from_vertices(sr.TRIANGLE, verts, idx, 1, stencil=3).run(fsid)
        assert np.array_equal(fb.download_winner(), ofb.winner)
        _compare(fb, ofb, f"u8 ordered draw {k}")
        assert np.array_equal(fb.download_planes(stencil=True)[2], ofb.stencil)
    pipe.destroy()
    fb.destroy()


def test_u8_target_accessors_and_limits(P, ctx):
    w, h = 32, 20
    fb = P.RenderBuffer.with_dimensions(ctx, w, h, u8_color=True)
    fb.clear((0.5, 1.0, 0.0, 2.0))
    assert fb.pixel(3, 4)[0] == (127.0, 255.0, 0.0, 255.0)  # (c * 255.0) as u8: truncation and saturation
    fb.set_pixel(3, 4, rgba=(1, 2, 3, 4), depth=-0.5)
    assert fb.pixel(3, 4)[:2] == ((1.0, 2.0, 3.0, 4.0), -0.5)
    with pytest.raises(Exception):
        fb.set_pixel(0, 0, rgba=(0.5, 0, 0, 0))  # not a u8 value
    u = scenes.suzanne_uniforms(w, h)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    tri = np.array([[-50, -50, -1, 1, 1, 0, 0, 1], [200, -50, -1, 1, 1, 0, 0, 1], [-50, 200, -1, 1, 1, 0, 0, 1]], np.float32)
    with pytest.raises(Exception, match="Blend = \\(\\) only"):
        pipe.draw_from_vertices(sr.TRIANGLE, tri, np.arange(3, dtype=np.uint32), 1).with_blend(sr.BLEND_ALPHA_OVER).run(sr.FS_FLAT)
    other = P.RenderBuffer.with_dimensions(ctx, w, h)
    p2 = P.Pipeline.from_framebuffer(other, u)
    p2.bind_framebuffer_texture(fb)
    quad = np.array([[-1, -1, 0, 1, 0, 1], [1, -1, 0, 1, 1, 1], [1, 1, 0, 1, 1, 0]], np.float32)
    qm = P.Mesh(ctx, vertices=quad, indices=np.arange(3, dtype=np.uint32))
    with pytest.raises(Exception, match="texture source"):
        p2.render_mesh(sr.TRIANGLE, qm).run_to_fragment(scenes.Viewport.new(w, h, 0.1, 10.0), sr.VS_PASSTHROUGH).run(sr.FS_TEXTURE_UNLIT)
    p2.bind_framebuffer_texture(None)
    for x in (p2, qm, other, pipe, fb):
        x.destroy()


@pytest.mark.parametrize("ranged", [False, True])
def test_sharded_composite_into_u8_target(P, ctx, ranged):
    """Two ranks (contexts of this process) composite their tiles into one RGBAu8Color target: plain sort-first sharding and
    the range-sharded front end (sr_shard); identical to the single-context frame."""
    rng = np.random.default_rng(123)
    w, h, world = 400, 260, 2
    n = 80_000 if ranged else 3000
    verts = H.random_screen_triangles(rng, n, w, h, max_size=3.0 if ranged else 30.0, integer_depth=True)
    idx = np.arange(3 * n, dtype=np.uint32)
    u = scenes.suzanne_uniforms(w, h)
    fb1 = P.RenderBuffer.with_dimensions(ctx, w, h, u8_color=True)
    fb1.clear(H.CLEAR)
    p1 = P.Pipeline.from_framebuffer(fb1, u)
    p1.draw_from_vertices(sr.TRIANGLE, verts, idx, 1).run(sr.FS_FLAT)
    expect = fb1.download()
    p1.destroy()
    fb1.destroy()
    ctxs = [P.Context(0) for _ in range(world)]
    groups = []
    for r, c in enumerate(ctxs):
        c.set_tile_shard(r, world)
        if ranged:
            groups.append(P.ShardGroup(c, w, h, 1))
    for g in groups:
        g.connect_local(groups)
    for g, c in zip(groups, ctxs):
        g.attach(c, 0)
    target = P.RenderBuffer.with_dimensions(ctxs[0], w, h, u8_color=True)
    fbs = [target, target.alias(ctxs[1])]
    pipes = [P.Pipeline.from_framebuffer(fb, u) for fb in fbs]
    for _ in range(2):
        for fb, p_ in zip(fbs, pipes):
            fb.clear(H.CLEAR)
            p_.draw_from_vertices(sr.TRIANGLE, verts, idx, 1).run(sr.FS_FLAT)
    for c in ctxs:
        c.synchronize()
    assert all(g.status() == 0 for g in groups)
    got = target.download()
    assert np.array_equal(got["rgba"], expect["rgba"]) and np.array_equal(got["depth"].view(np.uint32), expect["depth"].view(np.uint32))
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
