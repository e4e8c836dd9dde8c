import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import softrender_b200 as sr
from softrender_b200 import pipeline as P, scenes
import helpers as H
ctx = P.Context(0)
w, h = 1920, 1080
vp = scenes.Viewport.new(w, h, 0.1, 1000.0)
tex = P.Texture(ctx, scenes.checker_texture(512, 8))
fb = P.RenderBuffer.with_dimensions(ctx, w, h)
mesh = H.suzanne_mesh(with_uv=True); gm = P.Mesh(ctx, mesh)
us = [scenes.full_example_uniforms(w / h, np.deg2rad(75.0), 2.0, np.deg2rad(rot), np.deg2rad(65.0), off) for rot, off in [(45.0, -1.6), (165.0, 0.0), (285.0, 1.6)]]
pipe = P.Pipeline.from_framebuffer(fb, us[0]); pipe.bind_texture(tex); pipe.set_sampler(sr.FILTER_BILINEAR, sr.EDGE_CLAMP)
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 10):
    fb.clear(H.CLEAR)
    for u in us:
        pipe.set_uniforms(u)
        pipe.render_mesh(sr.TRIANGLE, gm).run(sr.VS_FULL_EXAMPLE).finish(vp).with_blend(sr.BLEND_ALPHA_OVER).run(sr.FS_FULL_EXAMPLE_TEXTURED)
ctx.synchronize()
