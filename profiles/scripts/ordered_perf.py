"""Throughput of the paths beside config 3: full_example scene (config 2), ordered (blend / stencil) path, lines, points."""
import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import softrender_b200 as sr
from softrender_b200 import pipeline as P, scenes
import helpers as H
ctx = P.Context(0); ctx.set_stage_timing(os.environ.get("SR_STAGES", "1") == "1")

def timeit(fn, n=10):
    for _ in range(3): fn()
    ctx.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    ctx.synchronize(); return (time.perf_counter() - t0) / n * 1e3

# ---- config 2: full_example scene at 1920x1080, 3 instanced meshes, alpha_over blend, textured; +subdivided mesh ----
w, h = 1920, 1080
vp = scenes.Viewport.new(w, h, 0.1, 1000.0)
tex = P.Texture(ctx, scenes.checker_texture(512, 8))
fb = P.RenderBuffer.with_dimensions(ctx, w, h)
for name, mesh in (("suzanne", H.suzanne_mesh(with_uv=True)), ("suzanne x2 subdivided", scenes.subdivide(H.suzanne_mesh(with_uv=True), 2))):
    gm = P.Mesh(ctx, mesh)
    us = [scenes.full_example_uniforms(w / h, np.deg2rad(75.0), 2.0, np.deg2rad(rot), np.deg2rad(65.0), off) for rot, off in [(45.0, -1.6), (165.0, 0.0), (285.0, 1.6)]]
    pipe = P.Pipeline.from_framebuffer(fb, us[0]); pipe.bind_texture(tex); pipe.set_sampler(sr.FILTER_BILINEAR, sr.EDGE_CLAMP)
    for blend in (None, sr.BLEND_ALPHA_OVER):
        def frame():
            fb.clear(H.CLEAR)
            for u in us:
                pipe.set_uniforms(u)
                st = pipe.render_mesh(sr.TRIANGLE, gm).run(sr.VS_FULL_EXAMPLE).finish(vp)
                if blend is not None: st = st.with_blend(blend)
                st.run(sr.FS_FULL_EXAMPLE_TEXTURED)
        ms = timeit(frame)
        print(f"full_example 1080p {name} ({mesh.ntris} tris x3) blend={'alpha_over' if blend is not None else '()'}: {ms:.3f} ms/frame, {ctx.stage_times()}")
    pipe.destroy(); gm.destroy()
fb.destroy()

# ---- ordered path at scale: grid meshes at 4K with alpha_over (strict order), and stencil ----
w, h = 3840, 2160
vp = scenes.Viewport.new(w, h, 0.1, 100.0)
u = scenes.grid_uniforms(w, h)
fb = P.RenderBuffer.with_dimensions(ctx, w, h)
fbs = P.RenderBuffer.with_dimensions(ctx, w, h, stencil=True) if hasattr(P.RenderBuffer, "with_dimensions") else None
for nx, ny in ((125, 100), (395, 316), (1250, 1000)):
    mesh = scenes.make_grid(nx, ny, 4)
    gm = P.Mesh(ctx, mesh)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    for label, blend in (("opaque", None), ("alpha_over (ordered)", sr.BLEND_ALPHA_OVER)):
        def frame():
            fb.clear(H.CLEAR)
            st = pipe.render_mesh(sr.TRIANGLE, gm).run_to_fragment(vp, sr.VS_SUZANNE)
            if blend is not None: st = st.with_blend(blend)
            st.run(sr.FS_SUZANNE)
        ms = timeit(frame, 5)
        print(f"grid {mesh.ntris} tris 4K {label}: {ms:.3f} ms/frame = {mesh.ntris / ms / 1e3:.1f} Mtris/s, {ctx.stage_times()}")
    pipe.destroy(); gm.destroy()
fb.destroy()
