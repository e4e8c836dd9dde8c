import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import softrender_b200 as sr
from softrender_b200 import pipeline as P, scenes
import helpers as H
torch.cuda.init()
def used(): f, t = torch.cuda.mem_get_info(); return (t - f) / 2**20
ctx = P.Context(0)
size = 512
mesh = H.suzanne_mesh(with_uv=True)
vp = scenes.Viewport.new(size, size, 0.1, 1000.0)
fb = P.RenderBuffer.with_dimensions(ctx, size, size)
tex = P.Texture(ctx, scenes.checker_texture(64, 8))
u = scenes.full_example_uniforms(1.0, np.deg2rad(75.0), 1.0, 0.3, np.deg2rad(65.0), 0.0)
pipe = P.Pipeline.from_framebuffer(fb, u); pipe.bind_texture(tex); pipe.set_sampler(sr.FILTER_BILINEAR, sr.EDGE_CLAMP)
marks = []
for it in range(3001):
    gm = P.Mesh(ctx, mesh)  # upload + destroy every frame (the e2e pattern)
    fb.clear(H.CLEAR)
    st = pipe.render_mesh(sr.TRIANGLE, gm).run(sr.VS_FULL_EXAMPLE)
    dup = st.duplicate()
    st.clip_primitives(correct=bool(it & 1)).finish(vp).with_blend(sr.BLEND_ALPHA_OVER if it % 3 == 0 else sr.BLEND_REPLACE).run(sr.FS_FULL_EXAMPLE_TEXTURED)
    dup.run(sr.GS_FACE_NORMALS).finish(vp).antialiased_lines(bool(it % 5 == 0)).run(sr.FS_GREEN)
    if it % 7 == 0: fb.download_rgba8()
    gm.destroy()
    if it % 500 == 0:
        ctx.synchronize(); marks.append(round(used(), 1))
print("device MiB in use at frames 0,500,...:", marks)
assert marks[-1] - marks[1] < 8, "device memory keeps growing"
print("no growth")
