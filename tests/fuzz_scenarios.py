"""Randomised differential scenarios against the oracle through the screen-space injection entry (the harness of
tests/test_gpu_parity.py::run_both_screen): random frame sizes, primitive mixes, triangle sizes, blend / stencil / cull /
line antialiasing state, draw counts, split settings.  Used by test_gpu_parity.py::test_randomised_scenarios and by
profiles/scripts/fuzz.py (long campaigns)."""
import numpy as np

import softrender_b200 as sr
import helpers as H


def run_scenario(P, ctx, run_both_screen, seed: int) -> str:
    """Runs one scenario; returns "" when the GPU frame equals the oracle's, a description of the mismatch otherwise."""
    rng = np.random.default_rng(seed)
    w, h = int(rng.integers(2, 400)), int(rng.integers(2, 300))
    ntri = int(rng.choice([0, 1, 7, 100, 900, 5000, 60000], p=[.05, .05, .1, .25, .3, .2, .05]))
    max_size = float(rng.choice([1.5, 3.0, 8.0, 30.0, 0.35 * min(w, h) + 1]))
    integer_depth = bool(rng.integers(0, 2))
    tri = H.random_screen_triangles(rng, max(ntri, 1), w, h, max_size=max_size, integer_depth=integer_depth)[:3 * ntri]
    gen = {}
    if rng.random() < 0.4:
        gen[2] = H.random_screen_triangles(rng, 200, w, h, integer_depth=integer_depth)[:2 * int(rng.integers(1, 300))]
    if rng.random() < 0.4:
        gen[1] = H.random_screen_triangles(rng, 200, w, h, integer_depth=integer_depth)[:int(rng.integers(1, 600))]
    if rng.random() < 0.2:
        gen[3] = H.random_screen_triangles(rng, int(rng.integers(1, 200)), w, h, max_size=max_size, integer_depth=integer_depth)
    blend = sr.BLEND_ALPHA_OVER if rng.random() < 0.3 else sr.BLEND_REPLACE
    aa = bool(rng.random() < 0.3) and 2 in gen
    if aa:
        blend = sr.BLEND_ALPHA_OVER if rng.random() < 0.7 else blend
    fs = sr.FS_DISCARD_CHECKER if rng.random() < 0.1 else sr.FS_FLAT
    cull = [None, sr.CLOCKWISE, sr.COUNTER_CLOCKWISE][int(rng.integers(0, 3))]
    stencil = rng.random() < 0.15
    kw = {}
    if stencil:
        kw = dict(stencil=True, stencil_cfg=(int(rng.integers(0, 8)), int(rng.integers(0, 8))), stencil_value=int(rng.integers(0, 4)),
                  init=(np.tile(np.float32(H.CLEAR), (w * h, 1)), np.full(w * h, np.float32(-3.4028235e38)),
                        rng.integers(0, 4, w * h).astype(np.uint8)))
    draws = int(rng.integers(1, 3))
    mode = int(rng.integers(0, 4))
    if mode == 0:
        ctx.set_micro()
    elif mode == 1:
        ctx.set_micro(int(rng.choice([0, 1, 16, 64, 1024, 4096])), 0, int(rng.integers(0, 4)))
    elif mode == 2:
        ctx.set_micro(16, 65536, int(rng.integers(0, 4)))
    else:
        ctx.set_micro(1024, 0, 0)
    idx = np.arange(len(tri), dtype=np.uint32)
    if ntri and rng.random() < 0.3:
        idx = (rng.permutation(ntri).astype(np.uint32)[:, None] * 3 + np.arange(3, dtype=np.uint32)[None, :]).reshape(-1)
    what = (f"seed {seed}: {w}x{h} ntri {ntri} size {max_size:.1f} gen {[(k, len(v)) for k, v in gen.items()]} blend {blend} aa {aa} "
            f"fs {fs} cull {cull} stencil {stencil} draws {draws} split mode {mode}")
    try:
        out, win, st, ofb = run_both_screen(P, ctx, w, h, tri if ntri else np.zeros((0, 8), np.float32), idx, fs=fs, cull=cull, blend=blend,
                                            aa=aa, gen=gen or None, draws=draws, **kw)
        if not np.array_equal(win, ofb.winner):
            return what + ": winner plane differs"
        if stencil and not np.array_equal(st, ofb.stencil):
            return what + ": stencil plane differs"
        H.compare_framebuffers(out, ofb, exact_color=True, what=what)
    except AssertionError as e:
        return what + ": " + str(e)[:200]
    finally:
        ctx.set_micro()
    return ""
