"""Randomised differential scenarios against the oracle through the screen-space injection entry (the harness of
tests/test_gpu_parity.py::run_both_screen): random frame sizes, primitive mixes, triangle sizes, blend / stencil / cull /
line antialiasing state, draw counts, split settings.  Used by test_gpu_parity.py::test_randomised_scenarios and by
profiles/scripts/fuzz.py (long campaigns)."""
import numpy as np

import softrender_b200 as sr
import helpers as H


def run_scenario(P, ctx, run_both_screen, seed: int, big: bool = False) -> str:
    """Runs one scenario; returns "" when the GPU frame equals the oracle's, a description of the mismatch otherwise.
    big: frames of up to 2200x1400 with 60k-300k triangles (the wide per-draw split, many tiles, long lists)."""
    rng = np.random.default_rng(seed)
    if big:
        w, h = int(rng.integers(600, 2200)), int(rng.integers(400, 1400))
        ntri = int(rng.choice([60000, 150000, 300000]))
        max_size = float(rng.choice([1.5, 3.0, 8.0, 30.0]))
    else:
        w, h = int(rng.integers(2, 400)), int(rng.integers(2, 300))
        ntri = int(rng.choice([0, 1, 7, 100, 900, 5000, 60000], p=[.05, .05, .1, .25, .3, .2, .05]))
        max_size = float(rng.choice([1.5, 3.0, 8.0, 30.0, 0.35 * min(w, h) + 1]))
    integer_depth = bool(rng.integers(0, 2))
    tri = H.random_screen_triangles(rng, max(ntri, 1), w, h, max_size=max_size, integer_depth=integer_depth)[:3 * ntri]
    gen = {}
    if rng.random() < 0.4:
        gen[2] = H.random_screen_triangles(rng, 200, w, h, integer_depth=integer_depth)[:2 * int(rng.integers(1, 300))]
    if rng.random() < 0.4:
        gen[1] = H.random_screen_triangles(rng, 200, w, h, integer_depth=integer_depth)[:int(rng.integers(1, 600))]
    if rng.random() < 0.2:
        gen[3] = H.random_screen_triangles(rng, int(rng.integers(1, 200)), w, h, max_size=max_size, integer_depth=integer_depth)
    blend = sr.BLEND_ALPHA_OVER if rng.random() < 0.3 else sr.BLEND_REPLACE
    aa = bool(rng.random() < 0.3) and 2 in gen
    if aa:
        blend = sr.BLEND_ALPHA_OVER if rng.random() < 0.7 else blend
    fs = sr.FS_DISCARD_CHECKER if rng.random() < 0.1 else sr.FS_FLAT
    cull = [None, sr.CLOCKWISE, sr.COUNTER_CLOCKWISE][int(rng.integers(0, 3))]
    stencil = rng.random() < 0.15
    kw = {}
    if stencil:
        kw = dict(stencil=True, stencil_cfg=(int(rng.integers(0, 8)), int(rng.integers(0, 8))), stencil_value=int(rng.integers(0, 4)),
                  init=(np.tile(np.float32(H.CLEAR), (w * h, 1)), np.full(w * h, np.float32(-3.4028235e38)),
                        rng.integers(0, 4, w * h).astype(np.uint8)))
    draws = int(rng.integers(1, 3))
    mode = int(rng.integers(0, 4))
    if mode == 0:
        ctx.set_micro()
    elif mode == 1:
        ctx.set_micro(int(rng.choice([0, 1, 16, 64, 1024, 4096])), 0, int(rng.integers(0, 4)))
    elif mode == 2:
        ctx.set_micro(16, 65536, int(rng.integers(0, 4)))
    else:
        ctx.set_micro(1024, 0, 0)
    idx = np.arange(len(tri), dtype=np.uint32)
    if ntri and rng.random() < 0.3:
        idx = (rng.permutation(ntri).astype(np.uint32)[:, None] * 3 + np.arange(3, dtype=np.uint32)[None, :]).reshape(-1)
    tiny_arena = rng.random() < 0.12  # per-tile list arenas too small: the passes skip themselves and are replayed
    if tiny_arena:
        ctx.set_list_capacity(int(rng.integers(1, 48)))
    what = (f"seed {seed}: {w}x{h} ntri {ntri} size {max_size:.1f} gen {[(k, len(v)) for k, v in gen.items()]} blend {blend} aa {aa} "
            f"fs {fs} cull {cull} stencil {stencil} draws {draws} split mode {mode} tiny arena {tiny_arena}")
    try:
        out, win, st, ofb = run_both_screen(P, ctx, w, h, tri if ntri else np.zeros((0, 8), np.float32), idx, fs=fs, cull=cull, blend=blend,
                                            aa=aa, gen=gen or None, draws=draws, **kw)
        if not np.array_equal(win, ofb.winner):
            return what + ": winner plane differs"
        if stencil and not np.array_equal(st, ofb.stencil):
            return what + ": stencil plane differs"
        H.compare_framebuffers(out, ofb, exact_color=True, what=what)
    except AssertionError as e:
        return what + ": " + str(e)[:200]
    finally:
        ctx.set_micro()
        if tiny_arena:
            ctx.set_list_capacity(1 << 20)
    return ""


def run_pipeline_scenario(P, ctx, ob, scenes, seed: int) -> str:
    """The whole builder chain on a random scene: Suzanne / subdivided Suzanne / a small displaced grid, random camera
    distance (meshes cross the frustum planes when close), rotation and frame size, both example shader sets, with or
    without the geometry stage (literal or Sutherland-Hodgman clipper), blend, cull, optionally tile-sharded.
    Winner plane and depth must be bit-exact, colour within 1/255."""
    rng = np.random.default_rng(seed)
    w, h = int(rng.integers(16, 500)), int(rng.integers(16, 400))
    kind = int(rng.integers(0, 4))
    full = kind >= 2  # full_example shader set (needs uv) vs suzanne shader set
    if kind == 3:
        mesh = scenes.make_grid(int(rng.integers(4, 40)), int(rng.integers(4, 30)), int(rng.integers(1, 4)), seed=seed)
        full = False
    else:
        mesh = H.suzanne_mesh(with_uv=full)
        if rng.random() < 0.25:
            mesh = scenes.subdivide(mesh, 1)
    dist = float(rng.choice([0.6, 0.9, 1.3, 2.0, 3.5]))
    rot = float(rng.uniform(0, 2 * np.pi))
    if kind == 3:
        u = scenes.grid_uniforms(w, h)
        near, far = 0.1, 100.0
    elif full:
        u = scenes.full_example_uniforms(w / h, np.deg2rad(75.0), dist, rot, np.deg2rad(65.0), float(rng.uniform(-1.5, 1.5)))
        near, far = 0.1, 1000.0
    else:
        u = scenes.suzanne_uniforms(w, h, rotation_y=rot)
        near, far = 0.001, 1000.0
    vp = scenes.Viewport.new(w, h, near, far)
    vs = sr.VS_FULL_EXAMPLE if full else sr.VS_SUZANNE
    textured = full and rng.random() < 0.5
    fs = (sr.FS_FULL_EXAMPLE_TEXTURED if textured else sr.FS_FULL_EXAMPLE) if full else sr.FS_SUZANNE
    clip = int(rng.integers(0, 3))  # 0: run_to_fragment, 1: literal clipper, 2: Sutherland-Hodgman
    blend = sr.BLEND_ALPHA_OVER if rng.random() < 0.35 else sr.BLEND_REPLACE
    cull = [None, sr.CLOCKWISE, sr.COUNTER_CLOCKWISE][int(rng.integers(0, 3))]
    # (single-process emulation of sharding: every rank draws into the same framebuffer object, so each rank must own a tile --
    # a rank without tiles would leave its lazily recorded clear pending for the whole frame)
    world = min(int(rng.choice([1, 1, 2, 3])), ((w + 63) // 64) * ((h + 31) // 32))
    draws = int(rng.integers(1, 3))
    what = (f"pipeline seed {seed}: {w}x{h} mesh kind {kind} ({mesh.ntris} tris) dist {dist} clip {clip} fs {fs} blend {blend} cull {cull} "
            f"world {world} draws {draws}")
    tex = None
    if textured:  # any size, including 1x1 and non-square: the bilinear taps at the last row / column are clamped
        tex = (scenes.checker_texture(64, 8) if rng.random() < 0.3 else
               rng.integers(0, 256, (int(rng.integers(1, 70)), int(rng.integers(1, 70)), 4), dtype=np.uint8))
    # sampler state and texel format from a stream of their own (the scenarios of earlier campaigns keep their seeds):
    # Filter x Edge of src/texture.rs:21-45; the texture is an 8-bit image or a render target sampled in place
    sampler, rtt = None, False
    if textured:
        srng = np.random.default_rng(seed ^ 0x7E57)
        if srng.random() < 0.6:
            sampler = (int(srng.integers(0, 2)), int(srng.integers(0, 3)), tuple(float(x) for x in srng.uniform(0, 1, 4)))
        rtt = srng.random() < 0.4
        if rtt:
            tex = srng.uniform(0, 1, (tex.shape[0], tex.shape[1], 4)).astype(np.float32)
        what += f" sampler {sampler} rtt {rtt}"
    ofb = ob.OracleFramebuffer(w, h)
    ofb.clear(H.CLEAR)
    for _ in range(draws):
        od = ob.OracleDraw(sr.TRIANGLE, mesh.indices)
        od.cull = sr.CULL_NONE if cull is None else cull
        od.blend = blend
        if clip == 0:
            od.vertex_run_to_fragment(vp, vs, u, mesh.vertices)
        else:
            od.vertex_run(vs, u, mesh.vertices).clip_primitives(correct=(clip == 2)).finish(vp)
        od.fragment_run(ofb, fs, u, texture=tex, sampler=sampler)
    fb = P.RenderBuffer.with_dimensions(ctx, w, h)
    fb.enable_winner(True)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    gmesh = P.Mesh(ctx, mesh)
    gtex = None
    if textured and rtt:
        gtex = P.RenderBuffer.with_dimensions(ctx, tex.shape[1], tex.shape[0])
        gtex.upload_planes(tex.reshape(-1, 4), np.zeros(tex.shape[0] * tex.shape[1], np.float32), None)
        pipe.bind_framebuffer_texture(gtex)
    elif textured:
        gtex = P.Texture(ctx, tex)
        pipe.bind_texture(gtex)
    if sampler is not None:
        pipe.set_sampler(*sampler)
    try:
        for rank in range(world):
            ctx.set_tile_shard(rank, world)
            fb.clear(H.CLEAR)
            for _ in range(draws):
                st = pipe.render_mesh(sr.TRIANGLE, gmesh)
                st = st.run_to_fragment(vp, vs) if clip == 0 else st.run(vs).clip_primitives(correct=(clip == 2)).finish(vp)
                st.cull_faces(cull).with_blend(blend).run(fs)
        ctx.set_tile_shard(0, 1)
        win = fb.download_winner()
        out = fb.download()
        if world == 1 and not np.array_equal(win, ofb.winner):  # (the winner plane is per draw call and rank: compared unsharded only)
            return what + f": winner plane differs at {int((win != ofb.winner).sum())} pixels"
        H.compare_framebuffers(out, ofb, color_tol=1.0 / 255.0, what=what)
    except AssertionError as e:
        return what + ": " + str(e)[:200]
    finally:
        ctx.set_tile_shard(0, 1)
        for x in (pipe, gmesh, fb) + ((gtex,) if gtex is not None else ()):
            x.destroy()
    return ""


def run_geometry_scenario(P, ctx, ob, scenes, seed: int) -> str:
    """Line / point meshes and the normal-visualisation geometry shaders through the builder chain: Suzanne's vertices drawn
    as points or as lines over random index pairs (clipped or not), or face / vertex normals (full_example/src/shaders.rs:35-89)
    with Bresenham or Wu lines, on top of the shaded mesh or alone.  Winner and depth bit-exact, colour within 1/255."""
    rng = np.random.default_rng(seed)
    w, h = int(rng.integers(16, 420)), int(rng.integers(16, 320))
    mesh = H.suzanne_mesh(with_uv=True)
    dist = float(rng.choice([0.7, 1.0, 2.0, 3.0]))
    u = scenes.full_example_uniforms(w / h, np.deg2rad(75.0), dist, float(rng.uniform(0, 6.28)), np.deg2rad(65.0), float(rng.uniform(-1.0, 1.0)))
    vp = scenes.Viewport.new(w, h, 0.1, 1000.0)
    mode = int(rng.integers(0, 4))  # 0 points, 1 lines, 2 face normals, 3 vertex normals
    aa = bool(rng.integers(0, 2)) and mode != 0
    blend = sr.BLEND_ALPHA_OVER if (aa or rng.random() < 0.3) else sr.BLEND_REPLACE
    under = mode >= 2 and rng.random() < 0.5  # draw the shaded mesh first, normals on top (config 2's two passes)
    clip = bool(rng.integers(0, 2))
    what = f"geometry seed {seed}: {w}x{h} mode {mode} aa {aa} blend {blend} clip {clip} mesh-under {under} dist {dist}"
    nv = len(mesh.vertices)
    if mode == 0:
        prim, idx = sr.POINT, rng.integers(0, nv, int(rng.integers(1, 3000))).astype(np.uint32)
    elif mode == 1:
        prim, idx = sr.LINE, rng.integers(0, nv, 2 * int(rng.integers(1, 800))).astype(np.uint32)
    else:
        prim, idx = sr.TRIANGLE, mesh.indices
    gs = [None, None, sr.GS_FACE_NORMALS, sr.GS_VERTEX_NORMALS][mode]
    fs_top = sr.FS_GREEN if mode >= 2 else sr.FS_FULL_EXAMPLE
    ofb = ob.OracleFramebuffer(w, h)
    ofb.clear(H.CLEAR)
    fb = P.RenderBuffer.with_dimensions(ctx, w, h)
    fb.enable_winner(True)
    fb.clear(H.CLEAR)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    gmesh = P.Mesh(ctx, vertices=mesh.vertices, indices=idx)
    gfull = P.Mesh(ctx, mesh) if under else None
    try:
        if under:
            od = ob.OracleDraw(sr.TRIANGLE, mesh.indices)
            od.vertex_run_to_fragment(vp, sr.VS_FULL_EXAMPLE, u, mesh.vertices).fragment_run(ofb, sr.FS_FULL_EXAMPLE, u)
            pipe.render_mesh(sr.TRIANGLE, gfull).run_to_fragment(vp, sr.VS_FULL_EXAMPLE).run(sr.FS_FULL_EXAMPLE)
        od = ob.OracleDraw(prim, idx)
        od.blend, od.aa = blend, aa
        od.vertex_run(sr.VS_FULL_EXAMPLE, u, mesh.vertices)
        st = pipe.render_mesh(prim, gmesh).run(sr.VS_FULL_EXAMPLE)
        if gs is not None:
            od.geometry_run(gs, u)
            st = st.run(gs)
        if clip:
            od.clip_primitives()
            st = st.clip_primitives()
        od.finish(vp).fragment_run(ofb, fs_top, u)
        st.finish(vp).with_blend(blend).antialiased_lines(aa).run(fs_top)
        if not np.array_equal(fb.download_winner(), ofb.winner):
            return what + ": winner plane differs"
        H.compare_framebuffers(fb.download(), ofb, color_tol=1.0 / 255.0, what=what)
    except AssertionError as e:
        return what + ": " + str(e)[:200]
    finally:
        for x in (pipe, gmesh, fb) + ((gfull,) if gfull is not None else ()):
            x.destroy()
    return ""
