import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import softrender_b200 as sr
from softrender_b200 import pipeline as P, scenes
import helpers as H
ctx = P.Context(0)
def timeit(fn, n=5):
    for _ in range(2): fn()
    ctx.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    ctx.synchronize(); return (time.perf_counter() - t0) / n * 1e3
# config 2 second pass: face-normal lines of the three instances at 1080p (green shader, Bresenham and Wu)
w, h = 1920, 1080
vp = scenes.Viewport.new(w, h, 0.1, 1000.0)
fb = P.RenderBuffer.with_dimensions(ctx, w, h)
for name, mesh in (("suzanne", H.suzanne_mesh(with_uv=True)), ("suzanne x2 subdivided", scenes.subdivide(H.suzanne_mesh(with_uv=True), 2))):
    gm = P.Mesh(ctx, mesh)
    us = [scenes.full_example_uniforms(w / h, np.deg2rad(75.0), 2.0, np.deg2rad(rot), np.deg2rad(65.0), off) for rot, off in [(45.0, -1.6), (165.0, 0.0), (285.0, 1.6)]]
    pipe = P.Pipeline.from_framebuffer(fb, us[0])
    for aa in (False, True):
        def frame():
            fb.clear(H.CLEAR)
            for u in us:
                pipe.set_uniforms(u)
                st = pipe.render_mesh(sr.TRIANGLE, gm).run(sr.VS_FULL_EXAMPLE).run(sr.GS_FACE_NORMALS).finish(vp)
                if aa: st = st.antialiased_lines(True).with_blend(sr.BLEND_ALPHA_OVER)
                st.run(sr.FS_GREEN)
        print(f"face-normal lines 1080p {name} ({mesh.ntris} lines x3) aa={aa}: {timeit(frame):.3f} ms/frame")
    pipe.destroy(); gm.destroy()
fb.destroy()
# lines at scale: random segments at 4K
w, h = 3840, 2160
fb = P.RenderBuffer.with_dimensions(ctx, w, h)
pipe = P.Pipeline.from_framebuffer(fb, scenes.grid_uniforms(w, h))
rng = np.random.default_rng(1)
for n, length in ((100000, 30.0), (1000000, 8.0)):
    v = np.zeros((2 * n, 8), np.float32)
    v[0::2, 0] = rng.uniform(0, w, n); v[0::2, 1] = rng.uniform(0, h, n)
    ang = rng.uniform(0, 2 * np.pi, n)
    v[1::2, 0] = v[0::2, 0] + length * np.cos(ang); v[1::2, 1] = v[0::2, 1] + length * np.sin(ang)
    v[:, 2] = -rng.uniform(0.5, 5, 2 * n); v[:, 3] = 1; v[:, 4:] = rng.uniform(0, 1, (2 * n, 4))
    idx = np.arange(2 * n, dtype=np.uint32)
    def frame():
        fb.clear(H.CLEAR)
        pipe.draw_from_vertices(sr.LINE, v, idx, 1).run(sr.FS_FLAT)
    try:
        print(f"{n} random lines of {length} px at 4K: {timeit(frame, 3):.3f} ms/frame")
    except Exception as e:
        print("lines at scale skipped:", e)
