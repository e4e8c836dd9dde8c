"""One render-to-texture second pass in a loop (for ncu): full-screen quad at 3840x2160, texture_unlit, sampler from argv
(filter edge), source = a rendered 3840x2160 target bound in place."""
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import softrender_b200 as sr
from softrender_b200 import pipeline as P, scenes
import helpers as H
filt, edge, n = (int(a) for a in (sys.argv[1:4] + ["1", "0", "6"][len(sys.argv) - 1:]))
ctx = P.Context(0)
w, h = 3840, 2160
u = scenes.grid_uniforms(w, h); vp = scenes.Viewport.new(w, h, 0.1, 100.0)
src = P.RenderBuffer.with_dimensions(ctx, w, h); src.clear(H.CLEAR)
gm = P.Mesh(ctx, scenes.make_grid(395, 316, 4))
P.Pipeline.from_framebuffer(src, u).render_mesh(sr.TRIANGLE, gm).run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)
dst = P.RenderBuffer.with_dimensions(ctx, w, h)
quad = np.zeros((4, 6), np.float32)
for i, (x, y) in enumerate([(0, 0), (w, 0), (w, h), (0, h)]):
    quad[i] = (x, y, -1.0, 1.0, x / w, y / h)
idx = np.array([0, 1, 2, 0, 2, 3], np.uint32)
p2 = P.Pipeline.from_framebuffer(dst, u)
p2.bind_framebuffer_texture(src); p2.set_sampler(filt, edge, (0, 0, 0, 1))
for _ in range(n):
    dst.clear(H.CLEAR); p2.draw_from_vertices(sr.TRIANGLE, quad, idx, 1).run(sr.FS_TEXTURE_UNLIT)
ctx.synchronize()
print("done")
