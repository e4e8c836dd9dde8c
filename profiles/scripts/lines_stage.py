import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import softrender_b200 as sr
from softrender_b200 import pipeline as P, scenes
import helpers as H
ctx = P.Context(0); ctx.set_stage_timing(True)
w, h = 1920, 1080
vp = scenes.Viewport.new(w, h, 0.1, 1000.0)
fb = P.RenderBuffer.with_dimensions(ctx, w, h)
mesh = scenes.subdivide(H.suzanne_mesh(with_uv=True), 2)
gm = P.Mesh(ctx, mesh)
u = scenes.full_example_uniforms(w / h, np.deg2rad(75.0), 2.0, np.deg2rad(165.0), np.deg2rad(65.0), 0.0)
pipe = P.Pipeline.from_framebuffer(fb, u)
for it in range(3):
    fb.clear(H.CLEAR)
    t0 = time.perf_counter()
    a = pipe.render_mesh(sr.TRIANGLE, gm).run(sr.VS_FULL_EXAMPLE); ctx.synchronize(); t1 = time.perf_counter()
    b = a.run(sr.GS_FACE_NORMALS); ctx.synchronize(); t2 = time.perf_counter()
    c = b.finish(vp); ctx.synchronize(); t3 = time.perf_counter()
    if len(sys.argv) > 1 and sys.argv[1] == 'aa': c = c.antialiased_lines(True).with_blend(sr.BLEND_ALPHA_OVER)
    c.run(sr.FS_GREEN); ctx.synchronize(); t4 = time.perf_counter()
    print("vertex %.3f  gs %.3f  finish %.3f  fragment %.3f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3), ctx.stage_times())
