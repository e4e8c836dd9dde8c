"""BASELINE.json's configs at their FULL sizes, CUDA path (through the C ABI) against the oracle.

The oracle's canonical mode is one frame-sized tile (tests/oracle_binding.py: `tile = None`): every primitive visits the
pixels of its own frame-clamped bounding box once, cost O(triangles + pixels) -- about 1.5 s for the 10 M triangles of
config 3 and half a minute for the 100 M of config 4 on one host core.  (The reference's default 128x128 tiling, where
every tile visits every primitive, is what the CPU *baseline* times; results are identical for Blend = () / stencil Keep,
SURVEY.md section 8 a8, which tests/test_gpu_parity.py::test_depth_ties_later_primitive_wins re-establishes.)

Bar: winner plane (which primitive owns each pixel = coverage mask + depth-test outcome) and depth bit-exact, colour
within 1/255 per channel (north_star; rasterization/triangle.rs:23-155, fragment.rs:268-311).
"""
import json
import os
import time

import numpy as np
import pytest

import softrender_b200 as sr
from softrender_b200 import scenes

import helpers as H
import oracle_binding as ob

pytestmark = pytest.mark.gpu

COLOR_TOL = 1.0 / 255.0
NTHREADS = os.cpu_count() or 1


@pytest.fixture(scope="module")
def P():
    from softrender_b200 import pipeline
    return pipeline


def _fb(P, ctx, w, h):
    fb = P.RenderBuffer.with_dimensions(ctx, w, h)
    fb.enable_winner(True)
    fb.clear(H.CLEAR)
    return fb


def _ofb(w, h):
    fb = ob.OracleFramebuffer(w, h)
    fb.clear(H.CLEAR)
    return fb


def _grid_vs_oracle(P, ctx, name, w, h, nx, ny, seed, reverse):
    mesh = scenes.make_grid(nx, ny, 4, seed=seed, reverse=reverse)
    u = scenes.grid_uniforms(w, h)
    vp = scenes.Viewport.new(w, h, 0.1, 100.0)
    fb = _fb(P, ctx, w, h)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    gmesh = P.Mesh(ctx, mesh)
    pipe.render_mesh(sr.TRIANGLE, gmesh).run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)
    gpu, gwin = fb.download(), fb.download_winner()
    for x in (pipe, gmesh, fb):
        x.destroy()
    t0 = time.perf_counter()
    ofb = _ofb(w, h)
    od = ob.OracleDraw(sr.TRIANGLE, mesh.indices)
    od.vertex_run_to_fragment(vp, sr.VS_SUZANNE, u, mesh.vertices, nthreads=NTHREADS).fragment_run(ofb, sr.FS_SUZANNE, u)
    oracle_s = time.perf_counter() - t0
    covered = int((ofb.winner > 0).sum())
    assert covered > 0.4 * w * h  # the mesh fills the centre of the frame
    assert np.array_equal(gwin, ofb.winner), f"{name}: {(gwin != ofb.winner).sum()} winner ids differ"
    H.compare_framebuffers(gpu, ofb, color_tol=COLOR_TOL, what=name)
    err = np.abs(gpu[:, :4].astype(np.float64) - ofb.color.astype(np.float64)).max()
    return {"config": name, "width": w, "height": h, "triangles": mesh.ntris, "reverse": reverse, "covered_pixels": covered,
            "winner_plane_equal": True, "depth_bit_exact": True, "max_colour_error": float(err), "colour_tolerance": COLOR_TOL,
            "oracle_seconds_one_tile": round(oracle_s, 2)}


@pytest.mark.parametrize("reverse", [False, True])
def test_config3_grid10m_full_size_vs_oracle(P, ctx, reverse):
    """Config 3 exactly as BASELINE.json names it: 10 M triangles, 3840x2160, front-to-back and back-to-front."""
    _grid_vs_oracle(P, ctx, "grid10m", 3840, 2160, 1250, 1000, 0x5EED0003, reverse)


@pytest.mark.slow
def test_config4_grid100m_full_size_vs_oracle(P, ctx):
    """Config 4: 100 M sub-pixel triangles at 7680x4320 (needs ~12 GB of host memory and ~1 min of oracle time; marked
    slow -- deselect with -m "gpu and not slow").  The result line is written to gpurun_out/ for profiles/."""
    rec = _grid_vs_oracle(P, ctx, "grid100m", 7680, 4320, 5000, 2500, 0x5EED0004, False)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "config4_full_size_parity.json"), "w") as fh:
        json.dump(rec, fh)


@pytest.mark.parametrize("camera_distance", [2.0, 0.9])
def test_config2_full_example_1080p_vs_oracle(P, ctx, camera_distance):
    """Config 2 as composed in SURVEY.md section 8d at its full size: 1920x1080, three instances of the twice-subdivided
    Suzanne (3 x 15,488 triangles), textured 4-light shader (512x512 checker, Bilinear + Clamp), alpha_over blend; the
    0.9 camera distance pushes the meshes through the left/right/near planes and the path includes clip_primitives;
    then the face-normal line pass (green shader, Bresenham) over the same frame."""
    w, h = 1920, 1080
    mesh = scenes.subdivide(H.suzanne_mesh(with_uv=True), 2)
    assert mesh.ntris == 15488
    tex = scenes.checker_texture(512, 8)
    vp = scenes.Viewport.new(w, h, 0.1, 1000.0)
    fb, ofb = _fb(P, ctx, w, h), _ofb(w, h)
    gmesh, gtex = P.Mesh(ctx, mesh), P.Texture(ctx, tex)
    us = [scenes.full_example_uniforms(w / h, np.deg2rad(75.0), camera_distance, np.deg2rad(rot), np.deg2rad(65.0), off)
          for rot, off in [(45.0, -1.6), (165.0, 0.0), (285.0, 1.6)]]
    pipe = P.Pipeline.from_framebuffer(fb, us[0])
    pipe.bind_texture(gtex)
    pipe.set_sampler(sr.FILTER_BILINEAR, sr.EDGE_CLAMP)
    clip = camera_distance < 1.0
    for u in us:
        pipe.set_uniforms(u)
        od = ob.OracleDraw(sr.TRIANGLE, mesh.indices)
        od.blend = sr.BLEND_ALPHA_OVER
        od.vertex_run(sr.VS_FULL_EXAMPLE, u, mesh.vertices, nthreads=NTHREADS)
        gs = pipe.render_mesh(sr.TRIANGLE, gmesh).run(sr.VS_FULL_EXAMPLE)
        if clip:
            od.clip_primitives()
            gs = gs.clip_primitives()
        od.finish(vp).fragment_run(ofb, sr.FS_FULL_EXAMPLE_TEXTURED, u, texture=tex, sampler=(sr.FILTER_BILINEAR, sr.EDGE_CLAMP, None))
        gs.finish(vp).with_blend(sr.BLEND_ALPHA_OVER).run(sr.FS_FULL_EXAMPLE_TEXTURED)
    assert np.array_equal(fb.download_winner(), ofb.winner), "config 2 triangle pass: winner ids of the last instance"
    H.compare_framebuffers(fb.download(), ofb, color_tol=COLOR_TOL, what="config 2 triangle pass")
    assert (ofb.depth > np.float32(-3e38)).sum() > 50_000
    for u in us:  # second pass: one face-normal line per triangle (full_example/src/shaders.rs:63-89)
        pipe.set_uniforms(u)
        od = ob.OracleDraw(sr.TRIANGLE, mesh.indices)
        od.vertex_run(sr.VS_FULL_EXAMPLE, u, mesh.vertices, nthreads=NTHREADS).geometry_run(sr.GS_FACE_NORMALS, u)
        gs = pipe.render_mesh(sr.TRIANGLE, gmesh).run(sr.VS_FULL_EXAMPLE).run(sr.GS_FACE_NORMALS)
        od.finish(vp).fragment_run(ofb, sr.FS_GREEN, u)
        gs.finish(vp).run(sr.FS_GREEN)
    assert np.array_equal(fb.download_winner(), ofb.winner), "config 2 line pass: winner ids of the last instance"
    H.compare_framebuffers(fb.download(), ofb, color_tol=COLOR_TOL, what="config 2 after the line pass")
    for x in (pipe, gmesh, gtex, fb):
        x.destroy()


def test_config1_suzanne_1024_and_2000_vs_oracle(P, ctx):
    """Config 1 at both sizes SURVEY.md section 8d names (1024^2 = the config, 2000^2 = the golden image), clip path."""
    mesh = H.suzanne_mesh()
    for size in (1024, 2000):
        u = scenes.suzanne_uniforms(size, size)
        vp = scenes.Viewport.new(size, size, 0.001, 1000.0)
        fb, ofb = _fb(P, ctx, size, size), _ofb(size, size)
        pipe = P.Pipeline.from_framebuffer(fb, u)
        gmesh = P.Mesh(ctx, mesh)
        pipe.render_mesh(sr.TRIANGLE, gmesh).run(sr.VS_SUZANNE).clip_primitives().finish(vp).run(sr.FS_SUZANNE)
        od = ob.OracleDraw(sr.TRIANGLE, mesh.indices)
        od.vertex_run(sr.VS_SUZANNE, u, mesh.vertices).clip_primitives().finish(vp).fragment_run(ofb, sr.FS_SUZANNE, u)
        assert np.array_equal(fb.download_winner(), ofb.winner), f"suzanne {size}"
        H.compare_framebuffers(fb.download(), ofb, color_tol=COLOR_TOL, what=f"suzanne {size}")
        for x in (pipe, gmesh, fb):
            x.destroy()
