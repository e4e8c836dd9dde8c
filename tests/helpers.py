"""Shared helpers of the parity tests: run the same scene through the oracle and through the CUDA
library (via the C ABI) and compare."""
from __future__ import annotations

import os

import numpy as np

import softrender_b200 as sr
from softrender_b200 import scenes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CLEAR = (0.01, 0.01, 0.01, 1.0)  # examples/suzanne.rs:75


def suzanne_mesh(with_uv: bool = False) -> scenes.MeshData:
    z = np.load(os.path.join(GOLDEN, "suzanne_mesh.npz"))
    return scenes.MeshData(z["vertices_uv"] if with_uv else z["vertices"], z["indices_uv"] if with_uv else z["indices"])


def bits(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bits_equal(a, b, what=""):
    a, b = bits(a), bits(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    bad = np.nonzero(a != b)
    assert bad[0].size == 0, f"{what}: {bad[0].size} of {a.size} values differ, first at {tuple(x[0] for x in bad)}"


def compare_framebuffers(gpu_aos: np.ndarray, ofb, *, color_tol=1.0 / 255.0, exact_color=False, what=""):
    """gpu_aos: [n,5] from RenderBuffer.download(); ofb: OracleFramebuffer."""
    assert_bits_equal(gpu_aos[:, 4], ofb.depth, what + " depth")
    if exact_color:
        assert_bits_equal(gpu_aos[:, :4], ofb.color, what + " colour")
    else:
        err = np.abs(gpu_aos[:, :4].astype(np.float64) - ofb.color.astype(np.float64))
        assert np.all(np.isfinite(gpu_aos[:, :4]) == np.isfinite(ofb.color)), what + " colour finiteness"
        err = np.where(np.isfinite(err), err, 0.0)
        assert err.max() <= color_tol, f"{what}: max colour error {err.max()} > {color_tol}"


def random_screen_triangles(rng, n, width, height, nk=4, *, integer_depth=False, max_size=None, margin=0.25):
    """n random screen-space triangles as records [3n, 4+nk] (x, y, z<0, 1/w, rgba...)."""
    max_size = max_size or 0.35 * min(width, height)
    cx = rng.uniform(-margin * width, (1 + margin) * width, n)
    cy = rng.uniform(-margin * height, (1 + margin) * height, n)
    size = rng.uniform(0.3, max_size, n)
    v = np.zeros((n, 3, 4 + nk), np.float32)
    for k in range(3):
        v[:, k, 0] = cx + rng.uniform(-1, 1, n) * size
        v[:, k, 1] = cy + rng.uniform(-1, 1, n) * size
        if integer_depth:
            v[:, k, 2] = -rng.integers(1, 6, n)[:, None][:, 0]
        else:
            v[:, k, 2] = -rng.uniform(0.1, 10.0, n)
        v[:, k, 3] = 1.0
        v[:, k, 4:] = rng.uniform(0, 1, (n, nk))
    if integer_depth:  # whole triangle at one integer depth: exact ties between primitives
        v[:, 1, 2] = v[:, 0, 2]
        v[:, 2, 2] = v[:, 0, 2]
    return v.reshape(3 * n, 4 + nk)
