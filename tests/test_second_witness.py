"""The oracle against a second, independent restatement (tests/second_witness.py: numpy float32, written from the Rust
sources) -- bit for bit on small scenes.  CPU only.  The oracle is what every GPU parity test trusts; the reference itself
cannot be run here (no Rust toolchain) and ships no test for these functions, so two independent readings that agree to
the last bit are the strongest pin available."""
import numpy as np
import pytest

import softrender_b200 as sr
from softrender_b200 import scenes

import helpers as H
import oracle_binding as ob
import second_witness as W2


def _clip_space_records(rng, n, nk=4, spread=1.6):
    """n clip-space vertices: most inside the frustum, many outside one or two planes, a few behind the eye (w < 0)."""
    v = np.zeros((n, 4 + nk), np.float32)
    w = rng.uniform(0.2, 3.0, n).astype(np.float32)
    w[rng.random(n) < 0.08] *= -1.0
    v[:, 3] = w
    v[:, 0] = rng.uniform(-spread, spread, n) * np.abs(w)
    v[:, 1] = rng.uniform(-spread, spread, n) * np.abs(w)
    v[:, 2] = rng.uniform(-0.3, 1.3, n) * np.abs(w)
    v[:, 4:] = rng.uniform(0, 1, (n, nk))
    return v


def test_normalize_matches_second_witness():
    rng = np.random.default_rng(1)
    for vpt in [(0, 0, 64, 48, 0.1, 100.0), (3, 5, 57, 33, 0.001, 1000.0), (0, 0, 1920, 1080, 0.5, 7.0)]:
        v = _clip_space_records(rng, 500)
        vp = scenes.Viewport(*[float(x) for x in vpt])
        od = ob.OracleDraw(sr.TRIANGLE, np.arange(498, dtype=np.uint32))
        od.set_vertices(v, 0).finish(vp)
        H.assert_bits_equal(od.data(0), W2.normalize(v, vpt), f"normalize {vpt}")


@pytest.mark.parametrize("seed", [2, 3, 4])
def test_clip_primitives_match_second_witness(seed):
    rng = np.random.default_rng(seed)
    n = 120
    v = _clip_space_records(rng, 3 * n)
    v[:9] = _clip_space_records(rng, 9, spread=0.9)  # three triangles fully inside: 18-vertex polygon, 16 triangles each
    v[:9, 3] = np.abs(v[:9, 3])
    v[:9, 2] = np.abs(v[:9, 2]) * 0.5
    od = ob.OracleDraw(sr.TRIANGLE, np.arange(3 * n, dtype=np.uint32))
    od.set_vertices(v, 0).clip_primitives()
    want = []
    for t in range(n):
        want += W2.clip_triangle(v[3 * t], v[3 * t + 1], v[3 * t + 2])
    got = od.data(3)
    assert len(got) == 3 * len(want) and len(want) > n
    H.assert_bits_equal(got, np.concatenate(want), "clipped triangles")
    # lines and points through the same planes
    ol = ob.OracleDraw(sr.LINE, np.arange(2 * n, dtype=np.uint32))
    ol.set_vertices(v[:2 * n], 0).clip_primitives()
    wl = [x for x in (W2.clip_line(v[2 * i], v[2 * i + 1]) for i in range(n)) if x is not None]
    H.assert_bits_equal(ol.data(2), np.concatenate(wl), "clipped lines")
    op = ob.OracleDraw(sr.POINT, np.arange(3 * n, dtype=np.uint32))
    op.set_vertices(v, 0).clip_primitives()
    wp = [x for x in (W2.clip_point(v[i]) for i in range(3 * n)) if x is not None]
    H.assert_bits_equal(op.data(1), np.stack(wp), "clipped points")
    assert 0 < len(wl) < n and 0 < len(wp) < 3 * n


@pytest.mark.parametrize("w,h,n,seed,cull", [(64, 64, 60, 5, 0), (61, 37, 150, 6, 0), (48, 64, 80, 7, sr.CLOCKWISE),
                                             (33, 29, 80, 8, sr.COUNTER_CLOCKWISE), (129, 129, 40, 9, 0)])
def test_rasterize_triangle_matches_second_witness(w, h, n, seed, cull):
    """Coverage (winner plane), depth and the interpolated colour of the flat shader: bit for bit.  Scenes include exact
    depth ties (later primitive wins), pixel centres exactly on edges (half-integer vertices), zero-area and z >= 0
    triangles."""
    rng = np.random.default_rng(seed)
    verts = np.concatenate([H.random_screen_triangles(rng, n // 2, w, h), H.random_screen_triangles(rng, n - n // 2, w, h, integer_depth=True)])
    t = verts.reshape(-1, 3, 8)
    t[0, :, :2] = [[4.5, 4.5], [20.5, 4.5], [4.5, 20.5]]      # centres on the edges
    t[1, :, :2] = [[20.5, 4.5], [20.5, 20.5], [4.5, 20.5]]    # shares the diagonal with triangle 0
    t[2, :, :2] = [[10, 10], [20, 20], [30, 30]]              # det = 0
    t[3, :, 2] = [0.5, 0.25, 0.125]                           # z >= 0: covered but rejected
    t[4, :, 2] = [-1.0, 0.5, -1.0]                            # z crosses zero inside the triangle
    idx = np.arange(3 * n, dtype=np.uint32)
    u = scenes.suzanne_uniforms(w, h)
    ofb = ob.OracleFramebuffer(w, h)
    ofb.clear(H.CLEAR)
    od = ob.OracleDraw(sr.TRIANGLE, idx)
    od.set_vertices(verts, 1)
    od.cull = cull
    od.fragment_run(ofb, sr.FS_FLAT, u)
    color = np.tile(np.float32(H.CLEAR), (h, w, 1))
    depth = np.full((h, w), np.float32(-3.4028235e38))
    winner = np.zeros((h, w), np.uint32)
    for i in range(n):
        W2.rasterize_triangle(color, depth, winner, t[i, 0], t[i, 1], t[i, 2], i, cull)
    assert np.array_equal(winner.reshape(-1), ofb.winner)
    H.assert_bits_equal(depth.reshape(-1), ofb.depth, "depth")
    H.assert_bits_equal(color.reshape(-1, 4), ofb.color, "colour")
    assert (winner > 0).sum() > 0.1 * w * h


def test_clip_then_raster_pipeline_matches_second_witness():
    """The three functions chained the way examples/suzanne.rs chains them: clip_primitives -> finish -> fragment run."""
    rng = np.random.default_rng(10)
    w, h, n = 64, 48, 40
    v = _clip_space_records(rng, 3 * n, spread=1.3)
    v[:, 2] = np.abs(v[:, 2])
    vpt = (0, 0, w, h, 0.1, 50.0)
    vp = scenes.Viewport(*[float(x) for x in vpt])
    u = scenes.suzanne_uniforms(w, h)
    ofb = ob.OracleFramebuffer(w, h)
    ofb.clear(H.CLEAR)
    od = ob.OracleDraw(sr.TRIANGLE, np.arange(3 * n, dtype=np.uint32))
    od.set_vertices(v, 0).clip_primitives().finish(vp).fragment_run(ofb, sr.FS_FLAT, u)
    tris = []
    for t in range(n):
        tris += W2.clip_triangle(v[3 * t], v[3 * t + 1], v[3 * t + 2])
    screen = W2.normalize(np.concatenate(tris), vpt).reshape(-1, 3, 8)
    color = np.tile(np.float32(H.CLEAR), (h, w, 1))
    depth = np.full((h, w), np.float32(-3.4028235e38))
    winner = np.zeros((h, w), np.uint32)
    for i, tri in enumerate(screen):
        if not np.isfinite(tri[:, :2]).all():
            continue  # (the reference panics on NaN coordinates, triangle.rs:70; the oracle skips them)
        W2.rasterize_triangle(color, depth, winner, tri[0], tri[1], tri[2], i)
    assert np.array_equal(winner.reshape(-1), ofb.winner)
    H.assert_bits_equal(depth.reshape(-1), ofb.depth, "depth")
    H.assert_bits_equal(color.reshape(-1, 4), ofb.color, "colour")
    assert (winner > 0).sum() > 100
