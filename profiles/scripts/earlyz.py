import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import softrender_b200 as sr
from softrender_b200 import pipeline as P, scenes
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "grid10m"
w, h, nx, ny, layers, seed, near, far = bench.CONFIGS[name]
u = scenes.grid_uniforms(w, h); vp = scenes.Viewport.new(w, h, near, far)
for rev in (False, True):
    mesh = scenes.make_grid(nx, ny, layers, seed=seed, reverse=rev)
    cx = P.Context(0); cx.set_stage_timing(True); fb = P.RenderBuffer.with_dimensions(cx, w, h); pp = P.Pipeline.from_framebuffer(fb, u); gm = P.Mesh(cx, mesh)
    ref = None
    for pc in (2, 0):
        cx.set_micro(16, 65536, pc)
        acc = {}
        for i in range(8):
            fb.clear(bench.CLEAR); pp.render_mesh(sr.TRIANGLE, gm).run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)
            cx.synchronize()
            st = cx.stage_times()
            if i >= 3:
                for k, v in st.items(): acc[k] = acc.get(k, 0) + v / 5
        out = fb.download()
        if ref is None: ref = out
        same = all(np.array_equal(a, b) for a, b in zip(ref, out)) if isinstance(out, tuple) else np.array_equal(ref, out)
        print("reverse", rev, "precheck", pc, {k: round(v, 4) for k, v in acc.items()}, "identical", same)
    pp.destroy(); gm.destroy(); fb.destroy(); cx.close()
