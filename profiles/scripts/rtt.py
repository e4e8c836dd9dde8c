"""Render-to-texture second pass: a full-screen quad (2 triangles) at 3840x2160 with the texture_unlit shader sampling a
3840x2160 render target in place, for every Filter / Edge, plus the same pass from an 8-bit image texture.
Algorithmic bytes of the pass: 20 B/pixel written + 16 B/pixel of source colour read (Nearest, 1:1 mapping)."""
import sys, os, time, gc
gc.disable()  # a full collection in the middle of one mode's timed frames showed up as +50..85 us on that line
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import softrender_b200 as sr
from softrender_b200 import pipeline as P, scenes
import helpers as H
ctx = P.Context(0)
def timeit(fn, n=20):
    for _ in range(5): fn()
    ctx.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    ctx.synchronize(); return (time.perf_counter() - t0) / n * 1e3
w, h = 3840, 2160
u = scenes.grid_uniforms(w, h); vp = scenes.Viewport.new(w, h, 0.1, 100.0)
src = P.RenderBuffer.with_dimensions(ctx, w, h); src.clear(H.CLEAR)
mesh = scenes.make_grid(395, 316, 4); gm = P.Mesh(ctx, mesh)
p1 = P.Pipeline.from_framebuffer(src, u)
p1.render_mesh(sr.TRIANGLE, gm).run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)  # pass 1: something to sample
dst = P.RenderBuffer.with_dimensions(ctx, w, h)
quad = np.zeros((4, 6), np.float32)  # x, y, z, 1/w, u, v
for i, (x, y) in enumerate([(0, 0), (w, 0), (w, h), (0, h)]):
    quad[i] = (x, y, -1.0, 1.0, x / w, y / h)
idx = np.array([0, 1, 2, 0, 2, 3], np.uint32)
p2 = P.Pipeline.from_framebuffer(dst, u)
img = P.Texture(ctx, scenes.checker_texture(1024, 16))
names = {0: "nearest", 1: "bilinear"}, {0: "clamp", 1: "wrap", 2: "border"}
# the reference's own render-to-texture source: a texture buffer, whose colour plane is re-used as the texture (texturebuffer.rs:63-66)
tsrc = P.RenderBuffer.with_dimensions(ctx, w, h, texture_buffer=True); tsrc.clear(H.CLEAR)
pt = P.Pipeline.from_framebuffer(tsrc, u)
pt.render_mesh(sr.TRIANGLE, gm).run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)
for label, bind in (("render target f32 (in place)", lambda: p2.bind_framebuffer_texture(src)),
                    ("texture buffer f32 plane     ", lambda: p2.bind_framebuffer_texture(tsrc)),
                    ("image rgba8 1024^2", lambda: p2.bind_texture(img))):
    bind()
    modes = [(f, e) for f in (0, 1) for e in (0, 1, 2)]
    if os.environ.get("RTT_REVERSE") == "1": modes.reverse()  # (order check: a mode's time must not depend on its position)
    for filt, edge in modes:
        p2.set_sampler(filt, edge, (0, 0, 0, 1))
        def frame():
            dst.clear(H.CLEAR); p2.draw_from_vertices(sr.TRIANGLE, quad, idx, 1).run(sr.FS_TEXTURE_UNLIT)
        ms = timeit(frame)
        gbs = w * h * (20 + (16 if "f32" in label else 0)) / ms / 1e6
        print(f"{label:30s} {names[0][filt]:8s} {names[1][edge]:6s}: {ms*1e3:7.1f} us/pass  ({gbs:6.0f} GB/s algorithmic)")
# where the time of the pass goes: library stage timers, and the same quad with the flat shader (no sampling) for comparison
ctx.set_stage_timing(True)
p2.bind_framebuffer_texture(src); p2.set_sampler(0, 0)
quad4 = np.zeros((4, 8), np.float32); quad4[:, :4] = quad[:, :4]; quad4[:, 4:] = 0.5
for label, fn in (("texture_unlit nearest/clamp", lambda: p2.draw_from_vertices(sr.TRIANGLE, quad, idx, 1).run(sr.FS_TEXTURE_UNLIT)),
                  ("flat shader, same quad", lambda: p2.draw_from_vertices(sr.TRIANGLE, quad4, idx, 1).run(sr.FS_FLAT))):
    def frame():
        dst.clear(H.CLEAR); fn()
    ms = timeit(frame)
    print(f"{label}: {ms*1e3:.1f} us/pass; stages {ctx.stage_times()}")
