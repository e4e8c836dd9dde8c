"""The config-3 frame (or any CONFIGS entry of bench.py) in a plain loop, for ncu: clear + vertex + raster + resolve, one stream."""
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import softrender_b200 as sr
from softrender_b200 import pipeline as P
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "grid10m"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
w, h, mesh, u, vp = bench.build_scene(name)
ctx = P.Context(0)
fb = P.RenderBuffer.with_dimensions(ctx, w, h)
pipe = P.Pipeline.from_framebuffer(fb, u)
gm = P.Mesh(ctx, mesh)
for _ in range(n):
    fb.clear(bench.CLEAR)
    pipe.render_mesh(sr.TRIANGLE, gm).run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)
ctx.synchronize()
print("done", name, mesh.ntris, "triangles")
