"""CPU check of the rasteriser's "candidate tightening" rule (rust-softrender_b200/csrc/sr_raster.cuh,
sr_tightening_applies / sr_tight_lo / sr_tight_hi; proof in DESIGN.md).

The reference tests every pixel of a triangle's integer bounding box with its f32 inside test
(src/pipeline/stages/rasterization/triangle.rs:74-113).  The CUDA path skips pixels whose centre lies more than 1/16
outside the f32 bounding box of a small, well-conditioned triangle.  This test restates both in numpy float32 (IEEE,
no FMA -- the reference's arithmetic) and asserts on adversarial random triangles that no skipped pixel would have
passed the reference's test.
"""
import numpy as np

F = np.float32


def reference_inside(x, y, px, py):
    """triangle.rs:104-113 for pixel (px, py) of triangles with vertex arrays x[3], y[3] (float32, broadcast)."""
    x1, x2, x3 = x
    y1, y2, y3 = y
    cx = px.astype(F) + F(0.5)
    cy = py.astype(F) + F(0.5)
    det = (y2 - y3) * (x1 - x3) + (x3 - x2) * (y1 - y3)
    with np.errstate(all="ignore"):
        u = ((y2 - y3) * (cx - x3) + (x3 - x2) * (cy - y3)) / det
        v = ((y3 - y1) * (cx - x3) + (x1 - x3) * (cy - y3)) / det
        w = F(1.0) - u - v
    return ~((u < 0) | (v < 0) | (w < 0)), det


def tightened_range(lo_f, hi_f, lo_i, hi_i):
    """sr_tight_lo / sr_tight_hi for one axis: keep p iff ceil(lo - 0.5625) <= p <= floor(hi - 0.4375)."""
    a = np.maximum(lo_i, np.ceil(lo_f - F(0.5625)).astype(np.int64))
    b = np.minimum(hi_i, np.floor(hi_f - F(0.4375)).astype(np.int64))
    return a, b


def check(x, y, width=4096, height=4096):
    x = [c.astype(F) for c in x]
    y = [c.astype(F) for c in y]
    xmin, xmax = np.minimum(np.minimum(x[0], x[1]), x[2]), np.maximum(np.maximum(x[0], x[1]), x[2])
    ymin, ymax = np.minimum(np.minimum(y[0], y[1]), y[2]), np.maximum(np.maximum(y[0], y[1]), y[2])

    def clamp_as_int(v, hi):
        return np.where(v < 0, 0, np.where(v > hi, hi, np.trunc(np.clip(v, -1e9, 1e9)))).astype(np.int64)

    minx, maxx = clamp_as_int(xmin, width - 1), clamp_as_int(xmax, width - 1)
    miny, maxy = clamp_as_int(ymin, height - 1), clamp_as_int(ymax, height - 1)
    _, det = reference_inside(x, y, minx, miny)
    applies = (xmax - xmin < F(4.99)) & (ymax - ymin < F(4.99)) & (np.abs(det) >= F(1.0))
    lx, hx = tightened_range(xmin, xmax, minx, maxx)
    ly, hy = tightened_range(ymin, ymax, miny, maxy)
    violations = 0
    tested = 0
    skipped = 0
    for oy in range(7):
        for ox in range(7):
            px, py = minx + ox, miny + oy
            in_box = (px <= maxx) & (py <= maxy)
            inside, _ = reference_inside(x, y, px, py)
            kept = (px >= lx) & (px <= hx) & (py >= ly) & (py <= hy)
            dropped = in_box & applies & ~kept
            violations += int((dropped & inside).sum())
            tested += int((in_box & applies).sum())
            skipped += int(dropped.sum())
    return violations, tested, skipped, int(applies.sum())


def test_no_skipped_pixel_passes_the_reference_test():
    rng = np.random.default_rng(20261017)
    n = 400_000
    total_tested = total_skipped = total_applies = 0
    for mode in range(6):
        base_x = rng.uniform(-3, 4090, n)
        base_y = rng.uniform(-3, 4090, n)
        if mode == 1:  # vertices hugging pixel centres and pixel edges
            base_x = np.floor(base_x) + rng.choice([0.0, 0.5, 0.4375, 0.5625], n) + rng.normal(0, 2e-4, n)
            base_y = np.floor(base_y) + rng.choice([0.0, 0.5, 0.4375, 0.5625], n) + rng.normal(0, 2e-4, n)
        size = rng.uniform(0.05, 4.9, n) if mode != 2 else rng.uniform(1.0, 2.5, n)
        x = [base_x + rng.uniform(0, 1, n) * size for _ in range(3)]
        y = [base_y + rng.uniform(0, 1, n) * size for _ in range(3)]
        if mode == 3:  # slivers: third vertex almost on the edge 1-2, det near the threshold of 1
            t = rng.uniform(-0.2, 1.2, n)
            off = rng.choice([-1.0, 1.0], n) * rng.uniform(0.9, 1.3, n)
            ex, ey = x[1] - x[0], y[1] - y[0]
            ln = np.hypot(ex, ey) + 1e-9
            x[2] = x[0] + t * ex - off * ey / (ln * ln)
            y[2] = y[0] + t * ey + off * ex / (ln * ln)
        if mode == 4:  # axis-aligned edges exactly on the margin positions
            x[1] = x[0]
            y[2] = y[0]
        if mode == 5:  # near the upper coordinate range the ABI allows (16384 x 8192 frames)
            x = [c + 12000 for c in x]
            y = [c + 4000 for c in y]
        v, tested, skipped, applies = check(x, y, 16384, 8192)
        assert v == 0, f"mode {mode}: {v} pixels skipped by the tightening rule pass the reference inside test"
        total_tested += tested
        total_skipped += skipped
        total_applies += applies
    # the rule must actually bite (otherwise this test proves nothing)
    assert total_applies > 500_000
    assert total_skipped > 0.3 * total_tested
