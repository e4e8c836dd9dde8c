import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import softrender_b200 as sr
from softrender_b200 import pipeline as P, scenes
import helpers as H
size = 1024
mesh = H.suzanne_mesh()
vp = scenes.Viewport.new(size, size, 0.001, 1000.0)
ctx = P.Context(0); ctx.set_stage_timing(os.environ.get("SR_STAGES", "1") == "1")
fb = P.RenderBuffer.with_dimensions(ctx, size, size)
us = [scenes.suzanne_uniforms(size, size, rotation_y=np.deg2rad(3.0 * (k + 1))) for k in range(64)]
pipe = P.Pipeline.from_framebuffer(fb, us[0])
gm = P.Mesh(ctx, mesh)
def frames(clip):
    for u in us:
        pipe.set_uniforms(u); fb.clear(H.CLEAR)
        d = pipe.render_mesh(sr.TRIANGLE, gm)
        if clip: d.run(sr.VS_SUZANNE).clip_primitives().finish(vp).run(sr.FS_SUZANNE)
        else: d.run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)
    ctx.synchronize()
for clip in (True, False):
    frames(clip); t0 = time.perf_counter(); frames(clip); dt = time.perf_counter() - t0
    print("clip" if clip else "noclip", "64 frames in %.2f ms -> %.0f frames/s, %.1f us/frame" % (dt * 1e3, 64 / dt, dt / 64 * 1e6), ctx.stage_times())
# host-side cost only: time to ENQUEUE 64 frames (no sync inside)
for clip in (True, False):
    ctx.synchronize(); t0 = time.perf_counter()
    for u in us:
        pipe.set_uniforms(u); fb.clear(H.CLEAR)
        d = pipe.render_mesh(sr.TRIANGLE, gm)
        if clip: d.run(sr.VS_SUZANNE).clip_primitives().finish(vp).run(sr.FS_SUZANNE)
        else: d.run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)
    t1 = time.perf_counter(); ctx.synchronize(); t2 = time.perf_counter()
    print("clip" if clip else "noclip", "enqueue %.1f us/frame, drain %.1f us/frame, launches/frame %d" % ((t1 - t0) / 64 * 1e6, (t2 - t1) / 64 * 1e6, 0))
l0 = ctx.launch_count(); frames(True); print("launches per clip frame", (ctx.launch_count() - l0) / 64)
l0 = ctx.launch_count(); frames(False); print("launches per noclip frame", (ctx.launch_count() - l0) / 64)
