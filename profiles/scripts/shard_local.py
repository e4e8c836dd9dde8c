"""Two ranks of a shard group as contexts of ONE device (sr_shard_connect_local), config 3: the range-sharded frame's kernels under ncu.
The peers' key buffers are local memory here, so the merge kernel's NVLink side is not represented -- instruction counts, occupancy and
the local memory traffic are.  Under ncu kernels are serialised: every cross-rank wait runs into its timeout (SR_SHARD_TIMEOUT_MS)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import softrender_b200 as sr  # noqa: E402
from softrender_b200 import pipeline as P, scenes  # noqa: E402

CLEAR = (0.01, 0.01, 0.01, 1.0)
w, h, world = 3840, 2160, 2
mesh = scenes.make_grid(1250, 1000, 4, seed=0x5EED0003)
u = scenes.grid_uniforms(w, h)
vp = scenes.Viewport.new(w, h, 0.1, 100.0)
ctxs = [P.Context(0) for _ in range(world)]
groups = []
for r, c in enumerate(ctxs):
    c.set_tile_shard(r, world)
    groups.append(P.ShardGroup(c, w, h, 1))
for g in groups:
    g.connect_local(groups)
for g, c in zip(groups, ctxs):
    g.attach(c, 0)
target = P.RenderBuffer.with_dimensions(ctxs[0], w, h)
fbs = [target, target.alias(ctxs[1])]
pipes = [P.Pipeline.from_framebuffer(fb, u) for fb in fbs]
meshes = [P.Mesh(c, mesh) for c in ctxs]
for f in range(int(os.environ.get("FRAMES", "3"))):
    for r in range(world):
        fbs[r].clear(CLEAR)
        pipes[r].render_mesh(sr.TRIANGLE, meshes[r]).run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)
for c in ctxs:
    c.synchronize()
print("status", [g.status() for g in groups], "covered", int((target.download()[:, 4] > np.float32(-3e38)).sum()))
