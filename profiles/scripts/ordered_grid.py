"""Ordered path at scale: config-3 grid with alpha_over blending (strict submission order)."""
import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import softrender_b200 as sr
from softrender_b200 import pipeline as P, scenes
import helpers as H
ctx = P.Context(0)
w, h = 3840, 2160
vp = scenes.Viewport.new(w, h, 0.1, 100.0); u = scenes.grid_uniforms(w, h)
fb = P.RenderBuffer.with_dimensions(ctx, w, h)
nx, ny = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1250x1000").split("x"))
mesh = scenes.make_grid(nx, ny, 4); gm = P.Mesh(ctx, mesh); pipe = P.Pipeline.from_framebuffer(fb, u)
def frame():
    fb.clear(H.CLEAR)
    pipe.render_mesh(sr.TRIANGLE, gm).run_to_fragment(vp, sr.VS_SUZANNE).with_blend(sr.BLEND_ALPHA_OVER).run(sr.FS_SUZANNE)
for _ in range(2): frame()
ctx.synchronize(); t0 = time.perf_counter()
for _ in range(3): frame()
ctx.synchronize(); ms = (time.perf_counter() - t0) / 3 * 1e3
print(f"grid {mesh.ntris} tris 4K alpha_over (ordered): {ms:.3f} ms/frame = {mesh.ntris / ms / 1e3:.1f} Mtris/s")
