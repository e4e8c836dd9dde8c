"""A SECOND, independent restatement of three reference functions in numpy float32 -- test infrastructure only.

Written from the Rust sources (cited per function), NOT from oracle/sr_oracle.cpp: a different author-path through the same
lines, in a different language and with different control flow (vectorised over pixels), so that a misreading of the
reference in the oracle would have to be made twice, identically, to go unnoticed.  tests/test_second_witness.py compares
the two bit for bit on small scenes.  Every operation below is a single IEEE f32 operation (numpy ufuncs never contract
a*b+c into an FMA), in the order the Rust expression trees evaluate.

  normalize            src/geometry/clipvertex.rs:89-127   (+ nalgebra 0.12 Matrix4 * Vector4: k-ascending dot products)
  clip_primitives      src/pipeline/stages/geometry.rs:261-336, src/geometry/clip.rs:22-63, src/numeric/interpolate.rs:56-57
  rasterize_triangle   src/pipeline/stages/rasterization/triangle.rs:48-143, src/numeric/interpolate.rs:49-51
"""
from __future__ import annotations

import numpy as np

F = np.float32
PLANES = ("left", "right", "top", "bottom", "near", "far")  # ALL_CLIPPING_PLANES, clip.rs:22-29


def normalize(v: np.ndarray, viewport) -> np.ndarray:
    """ClipVertex::normalize for records [n, 4+nk]; viewport = (x, y, width, height, near, far)."""
    left, bottom, width, height, near, far = (F(t) for t in viewport)
    right, top = left + width, bottom + height
    two, mtwo, zero, one = F(2.0), F(-2.0), F(0.0), F(1.0)
    m = [[(right - left) / two, zero, zero, (right + left) / two],
         [zero, (top - bottom) / mtwo, zero, (top + bottom) / two],
         [zero, zero, (far - near) / mtwo, (far + near) / mtwo],
         [zero, zero, zero, one]]
    out = np.array(v, dtype=F, copy=True)
    x, y, z, w = (out[:, i].copy() for i in range(4))
    ndc = [x / w, y / w, z / w, np.full_like(x, one)]
    for r in range(3):  # screen.w is overwritten below
        acc = np.zeros_like(x)
        for k in range(4):  # nalgebra: accumulator starts at zero, k ascending
            acc = acc + m[r][k] * ndc[k]
        out[:, r] = acc
    out[:, 3] = one / w
    return out


def _inside(plane: str, p: np.ndarray) -> bool:  # ClippingPlane::has_inside, clip.rs:33-44
    x, y, z, w = p[0], p[1], p[2], p[3]
    return {"left": x >= -w, "right": x <= w, "top": y >= -w, "bottom": y <= w, "near": z >= F(0.0), "far": z <= w}[plane]


def _intersect(plane: str, v1: np.ndarray, v2: np.ndarray) -> np.ndarray:  # ClippingPlane::intersect, clip.rs:47-63
    x1, y1, z1, w1 = v1[:4]
    x2, y2, z2, w2 = v2[:4]
    a, b = {"left": (w1 + x1, w2 + x2), "right": (w1 - x1, w2 - x2), "top": (w1 + y1, w2 + y2),
            "bottom": (w1 - y1, w2 - y2), "near": (z1, z2), "far": (w1 - z1, w2 - z2)}[plane]
    with np.errstate(all="ignore"):
        t = F(a) / (F(a) - F(b))
        return ((F(1.0) - t) * v1 + t * v2).astype(F)  # linear_interpolate on every component, interpolate.rs:56-57


def clip_triangle(a: np.ndarray, b: np.ndarray, c: np.ndarray):
    """geometry.rs:264-298: the polygon the loop builds, then the triangles it emits (list of [3, 4+nk] arrays)."""
    polygon = []
    for s, p in ((a, b), (b, c), (c, a)):
        for plane in PLANES:
            s_in, p_in = _inside(plane, s), _inside(plane, p)
            if s_in != p_in:
                polygon.append(_intersect(plane, s, p))
            if p_in:
                polygon.append(p.copy())
    if len(polygon) == 3:
        return [np.stack(polygon)]
    out = []
    if len(polygon) > 3:
        last = polygon[-1]
        for i in range(len(polygon) - 2):
            out.append(np.stack([last, polygon[i], polygon[i + 1]]))
    return out


def clip_line(start: np.ndarray, end: np.ndarray):
    """geometry.rs:300-327; returns [2, 4+nk] or None."""
    start, end = start.copy(), end.copy()
    intersections = 0
    for plane in PLANES:
        s_in, p_in = _inside(plane, start), _inside(plane, end)
        if s_in != p_in:
            x = _intersect(plane, start, end)
            if s_in:
                end = x
            elif p_in:
                start = x
            intersections += 1
        elif not s_in:
            return None
        if intersections > 2:
            break
    return np.stack([start, end])


def clip_point(p: np.ndarray):
    return p.copy() if all(_inside(pl, p) for pl in PLANES) else None  # geometry.rs:329-333


def _clamp_as_int(value, lo: int, hi: int) -> int:  # triangle.rs:66-72
    if value < F(lo):
        return lo
    if value > F(hi):
        return hi
    return int(np.trunc(value))


def rasterize_triangle(color, depth, winner, a, b, c, prim: int, cull: int = 0, blend=None):
    """triangle.rs:48-143 on planes color [h, w, 4], depth [h, w], winner [h, w] with one frame-sized tile
    ((0,0),(w-1,h-1)) -- fragment.rs:188-216 for frames up to 129 pixels a side -- and the flat test shader
    (colour = the interpolated K[0..4)).  cull: 0 None, 1 Clockwise, 2 CounterClockwise.  Vectorised over the bounding box."""
    h, w = depth.shape
    x1, y1 = F(a[0]), F(a[1])
    x2, y2 = F(b[0]), F(b[1])
    x3, y3 = F(c[0]), F(c[1])
    if cull:
        area = x1 * y2 + x2 * y3 + x3 * y1 - x2 * y1 - x3 * y2 - x1 * y3  # left to right
        if cull == (1 if np.signbit(area) else 2):
            return
    det = (y2 - y3) * (x1 - x3) + (x3 - x2) * (y1 - y3)
    minx = _clamp_as_int(min(min(x1, x2), x3), 0, w - 1)
    miny = _clamp_as_int(min(min(y1, y2), y3), 0, h - 1)
    maxx = _clamp_as_int(max(max(x1, x2), x3), 0, w - 1)
    maxy = _clamp_as_int(max(max(y1, y2), y3), 0, h - 1)
    if maxx < minx or maxy < miny:
        return
    ys, xs = np.mgrid[miny:maxy + 1, minx:maxx + 1]
    px, py = xs.astype(F) + F(0.5), ys.astype(F) + F(0.5)
    with np.errstate(all="ignore"):
        u = ((y2 - y3) * (px - x3) + (x3 - x2) * (py - y3)) / det
        v = ((y3 - y1) * (px - x3) + (x1 - x3) * (py - y3)) / det
        wgt = F(1.0) - u - v
        inside = ~((u < 0) | (v < 0) | (wgt < 0))

        def bary(ka, kb, kc):  # interpolate.rs:49-51 per component
            return (F(ka) * u + F(kb) * v) + F(kc) * wgt
        z = bary(a[2], b[2], c[2])
        sub_d = depth[miny:maxy + 1, minx:maxx + 1]
        passed = inside & (z < 0) & (z >= sub_d)
        src = np.stack([bary(a[4 + i], b[4 + i], c[4 + i]) for i in range(4)], axis=-1)
    sub_c = color[miny:maxy + 1, minx:maxx + 1]
    sub_w = winner[miny:maxy + 1, minx:maxx + 1]
    if blend is not None:
        src = blend(src, sub_c)
    sub_c[passed] = src[passed]
    sub_d[passed] = z[passed]
    sub_w[passed] = prim + 1
