"""Sweep of the k_micro / tile-list split (micro_area) over triangle sizes."""
import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import softrender_b200 as sr
from softrender_b200 import pipeline as P, scenes
import helpers as H
ctx = P.Context(0); ctx.set_stage_timing(os.environ.get("SR_STAGES", "1") == "1")
def timeit(fn, n=8):
    for _ in range(3): fn()
    ctx.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    ctx.synchronize(); return (time.perf_counter() - t0) / n * 1e3
areas = [int(a) for a in (sys.argv[1].split(",") if len(sys.argv) > 1 else "16,32,64,128,256,1024,4096".split(","))]
w, h = 3840, 2160
vp = scenes.Viewport.new(w, h, 0.1, 100.0); u = scenes.grid_uniforms(w, h)
fb = P.RenderBuffer.with_dimensions(ctx, w, h)
for nx, ny in [tuple(int(v) for v in g.split('x')) for g in (sys.argv[2].split(',') if len(sys.argv) > 2 else '40x32,125x100,395x316,700x560,1250x1000'.split(','))]:
    mesh = scenes.make_grid(nx, ny, 4); gm = P.Mesh(ctx, mesh); pipe = P.Pipeline.from_framebuffer(fb, u)
    def frame():
        fb.clear(H.CLEAR); pipe.render_mesh(sr.TRIANGLE, gm).run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)
    res = []
    for a in areas:
        ctx.set_micro(a, 65536, 0)
        ms = timeit(frame); st = ctx.stage_times()
        res.append(f"{a}: {ms:.3f} (micro {st['micro_ms']:.3f} raster {st['raster_ms']:.3f})")
    print(f"grid {mesh.ntris} tris (~{2550*1428/ (mesh.ntris/4):.0f} px^2 each):", " | ".join(res))
    pipe.destroy(); gm.destroy()
fb.destroy()
# Suzanne 1024^2 and full_example-like 1080p
for size, label in ((1024, "suzanne 1024^2"), (2000, "suzanne 2000^2")):
    mesh = H.suzanne_mesh(); gm = P.Mesh(ctx, mesh)
    fb = P.RenderBuffer.with_dimensions(ctx, size, size); vp = scenes.Viewport.new(size, size, 0.001, 1000.0)
    pipe = P.Pipeline.from_framebuffer(fb, scenes.suzanne_uniforms(size, size))
    def frame():
        fb.clear(H.CLEAR); pipe.render_mesh(sr.TRIANGLE, gm).run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)
    res = []
    for a in areas:
        ctx.set_micro(a, 0, 0)
        ms = timeit(frame, 20); st = ctx.stage_times()
        res.append(f"{a}: {ms:.3f} (micro {st['micro_ms']:.3f} raster {st['raster_ms']:.3f})")
    print(label, " | ".join(res))
    pipe.destroy(); gm.destroy(); fb.destroy()
