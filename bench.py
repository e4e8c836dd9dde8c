#!/usr/bin/env python
"""bench.py -- headline benchmark of the draw hot path (BASELINE.json: Mtris/s and frames/s at 3840x2160).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config grid10m|...]

A "step" is one frame: clear + vertex stage + binning + tile rasterisation of the synthetic 10M-triangle
shaded mesh at 3840x2160 (config 3) with the mesh already resident in HBM.
N>1 (torchrun, one rank per GPU): `value` is frame batching (every GPU renders whole frames, no data-path
collective; "weak" scaling).  The north_star's split of ONE frame -- tiles sharded over the GPUs, every finished
tile stored straight into rank 0's framebuffer over NVLink peer memory, and the per-triangle front end sharded by
triangle range with the keys merged over NVLink inside the tile kernel -- is measured beside it for config 3
(`sharded`) and config 4 (`sharded_config4`) and compared bit for bit with the single-GPU frame.  NCCL carries
only rendezvous, barriers between batches and the timing reduction.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

CLEAR = (0.01, 0.01, 0.01, 1.0)
CONFIGS = {
    # name: (width, height, nx, ny, layers, seed, near, far)
    "grid10m": (3840, 2160, 1250, 1000, 4, 0x5EED0003, 0.1, 100.0),   # config 3 (the metric's config)
    "grid100m": (7680, 4320, 5000, 2500, 4, 0x5EED0004, 0.1, 100.0),  # config 4
    "grid1m": (3840, 2160, 395, 316, 4, 0x5EED0003, 0.1, 100.0),      # ~14 px^2 triangles: all through the per-tile lists
    "grid100k": (3840, 2160, 125, 100, 4, 0x5EED0003, 0.1, 100.0),    # ~145 px^2 triangles
}


def algorithmic_bytes(nverts: int, ntris: int, w: int, h: int, vin_bytes: int = 24) -> int:
    """B_alg = V*sizeof(Vin) + 3*T*4 + W*H*20 (SURVEY.md section 8d / BASELINE.md section 3)."""
    return nverts * vin_bytes + 3 * ntris * 4 + w * h * 20


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def build_scene(name: str):
    from softrender_b200 import scenes
    w, h, nx, ny, layers, seed, near, far = CONFIGS[name]
    mesh = scenes.make_grid(nx, ny, layers, seed=seed)
    return w, h, mesh, scenes.grid_uniforms(w, h), scenes.Viewport.new(w, h, near, far)


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the restated reference CPU path (oracle) on the host cores
# ------------------------------------------------------------------------------------------------------------
def native_oracle():
    """Builds oracle/libsr_oracle_native.so (-O3 -march=native -ffp-contract=off) ON the host that
    times it and points the binding at it; falls back to the portable -O3 build.  Returns the flag string for the record."""
    try:
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "native"], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL, timeout=300)
        os.environ["SR_ORACLE_SO"] = os.path.join(ROOT, "oracle", "libsr_oracle_native.so")
        return "-O3 -march=native -ffp-contract=off"
    except Exception:
        return "-O3 -ffp-contract=off (portable build; native build failed)"


def workload_config(name: str, in_flight: int, world: int) -> dict:
    """The `config` object of the JSON line; both arms print the same one (the CPU arm times this workload)."""
    w, h, nx, ny, layers = CONFIGS[name][:5]
    nverts = layers * (nx + 1) * (ny + 1)
    return {"workload": name, "width": w, "height": h, "triangles": 2 * nx * ny * layers, "vertices": nverts,
            "shader": "suzanne Blinn-Phong", "depth_test": True, "frames_in_flight": in_flight,
            "parallelism": "1 GPU" if world == 1 else
            f"frame batching: {world} GPUs each render whole frames of the workload (no data-path collective); the "
            f"tile-sharded single frame (sort-last front end, sort-first resolve, NVLink merge + composite) is under 'sharded'",
            "l2": "working set per frame (mesh + shaded vertices + framebuffer) exceeds the 126 MB L2; no flush needed"}


class CpuFrame:
    """One frame of a grid config on the restated reference CPU path (the oracle's threaded mode: thread pool, atomic
    cursors, every 128x128 reference tile visits EVERY triangle -- fragment.rs:29,240-311), cut into `nslices` slices
    of the reference's tile list: slice i = tiles i, i+nslices, ...  All slices together are exactly one full frame
    of the full-size workload; the vertex stage (all vertices) and the clear run in slice 0."""

    def __init__(self, name: str, nslices: int):
        import softrender_b200 as sr
        from softrender_b200 import scenes
        import oracle_binding as ob
        self.sr, self.ob = sr, ob
        self.w, self.h, nx, ny, layers, seed, near, far = CONFIGS[name]
        self.mesh = scenes.make_grid(nx, ny, layers, seed=seed)
        self.u = scenes.grid_uniforms(self.w, self.h)
        self.vp = scenes.Viewport.new(self.w, self.h, near, far)
        self.cores = os.cpu_count() or 1
        self.fb = ob.OracleFramebuffer(self.w, self.h)
        self.nslices = nslices
        self.ntiles = len(ob.tiles(self.w, self.h, 128, 128))
        self.draw = None

    def slice(self, i: int) -> float:
        sr, ob = self.sr, self.ob
        t0 = time.perf_counter()
        if i == 0 or self.draw is None:
            self.fb.clear(CLEAR)
            self.draw = ob.OracleDraw(sr.TRIANGLE, self.mesh.indices)
            self.draw.tile = (128, 128)
            self.draw.vertex_run_to_fragment(self.vp, sr.VS_SUZANNE, self.u, self.mesh.vertices, nthreads=self.cores)
        self.draw.fragment_run(self.fb, sr.FS_SUZANNE, self.u, nthreads=self.cores, tile_slice=(i, self.nslices), keep_winner=True)
        return time.perf_counter() - t0

    def tiles_in(self, i: int) -> int:
        return len(range(i, self.ntiles, self.nslices))


def cpu_baseline_sample(name: str, flags: str, stride: int = 3):
    """`cpu_baseline` of our arm: one slice (every `stride`-th reference tile, all triangles, plus the whole vertex stage) of
    the full-size frame on all host threads, scaled by the tile fraction -- a bounded sample of about 10 s."""
    f = CpuFrame(name, stride)
    t = f.slice(0)
    frac = f.tiles_in(0) / f.ntiles
    return {"value": f.mesh.ntris * frac / t / 1e6, "unit": "Mtris/s", "cores": f.cores, "kind": "port",
            "sample": f"{name} at full size ({f.mesh.ntris} triangles, {f.w}x{f.h}): vertex stage + {f.tiles_in(0)} of the "
                      f"{f.ntiles} reference tiles (every {stride}th, each visiting every triangle), {t:.2f} s on {f.cores} "
                      f"threads; value = triangles x tile fraction / time; oracle built {flags}"}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the box's host cores.  The Rust crate
    cannot be built in this image (no rustc/cargo, un-vendored dependencies), so this is the oracle's threaded mode --
    the restated reference CPU path with the reference's structure.  The K timed steps are the K tile slices of ONE
    full frame of the full-size workload (CpuFrame), so `value` = triangles / (sum of the K step times) is a whole-frame,
    same-config measurement while each step stays bounded; the W warm-up steps are slices of a frame that is discarded."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    flags = native_oracle()
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    f = CpuFrame(args.config, steps)
    for i in range(warmup):
        f.slice(i % steps)
    f.draw = None
    times = [f.slice(i) for i in range(steps)]
    total = float(sum(times))
    value = f.mesh.ntris / total / 1e6
    base = {"value": value, "unit": "Mtris/s", "cores": f.cores, "kind": "port",
            "sample": f"one full {args.config} frame ({f.mesh.ntris} triangles, {f.w}x{f.h}, all {f.ntiles} reference tiles of "
                      f"128x128, each visiting every triangle) timed as {steps} tile slices: {total:.2f} s on {f.cores} threads; "
                      f"oracle built {flags}"}
    line = {
        "impl": "reference", "metric": f"Mtris/s at {f.w}x{f.h}", "value": value, "unit": "Mtris/s",
        "frames_per_s": 1.0 / total,
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": total / steps * 1e3, "frame_s": total,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.config, args.in_flight, args.gpus),
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "Mtris/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "restated reference CPU path (the Rust reference cannot be built here); `config` describes the workload, "
                "its GPU-side keys (frames_in_flight, parallelism, l2) do not apply to this arm",
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def kernel_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(path))
    except Exception:
        return {}


def bind_to_gpu_numa_node(local_rank: int):
    """Best effort: run this rank's host threads (and so first-touch its pinned buffers) on the NUMA node its GPU hangs
    off, so the end-to-end copies do not cross the socket interconnect.  Returns a description for the JSON line."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local_rank), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())

        def parse_cpulist(text):
            cpus = set()
            for part in text.strip().split(","):
                lo, _, hi = part.strip().partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            return cpus

        if node < 0:
            # sysfs carries no node for the device (virtualised PCI topology): ask the driver -- the "CPU Affinity" column of
            # `nvidia-smi topo -m` for this GPU's row
            try:
                import re
                text = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
                rows = re.sub(r"\x1b\[[0-9;]*m", "", text).splitlines()  # (the header cells are underlined with escape codes)
                head = next(r for r in rows if "CPU Affinity" in r)
                col = [c.strip() for c in head.split("\t")].index("CPU Affinity")
                row = next(r for r in rows if r.startswith(f"GPU{local_rank}\t") or r.startswith(f"GPU{local_rank} "))
                cell = [c.strip() for c in row.split("\t")][col]
                cpus = parse_cpulist(cell) & os.sched_getaffinity(0)
                if cpus and cpus != os.sched_getaffinity(0):
                    os.sched_setaffinity(0, cpus)
                    return f"nvidia-smi topo cpu affinity {cell} ({len(cpus)} cpus)"
                return f"numa node unknown (sysfs -1; nvidia-smi topo cpu affinity '{cell}' = every allowed cpu: one node)"
            except Exception as e:
                return f"numa node unknown (sysfs -1; nvidia-smi topo: {type(e).__name__})"
        cpus = parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"numa node {node}: no allowed cpus"
        os.sched_setaffinity(0, cpus)
        return f"numa node {node} ({len(cpus)} cpus)"
    except Exception as e:  # containers without sysfs topology, single-socket boxes, ...
        return f"unbound ({type(e).__name__})"


def other_configs(ctx, headline=None):
    """Single-stream frame times of the other BASELINE configs on the same box (reported beside the headline, not part of it):
    config 1 (Suzanne 1024x1024, examples/suzanne.rs path) and config 2 as composed in SURVEY.md 8d (three Suzanne instances
    at 1920x1080, textured 4-light shader, alpha_over blend, then the face-normal line pass with the green shader)."""
    import softrender_b200 as sr
    from softrender_b200 import pipeline as P, scenes
    import helpers as H

    def per_frame_us(frame, n):
        for _ in range(5):
            frame()
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            frame()
        ctx.synchronize()
        return (time.perf_counter() - t0) / n * 1e6

    out = {}
    size = 1024
    mesh = H.suzanne_mesh()
    fb = P.RenderBuffer.with_dimensions(ctx, size, size)
    pipe = P.Pipeline.from_framebuffer(fb, scenes.suzanne_uniforms(size, size))
    gm = P.Mesh(ctx, mesh)
    vp = scenes.Viewport.new(size, size, 0.001, 1000.0)

    def suzanne():
        fb.clear(CLEAR)
        pipe.render_mesh(sr.TRIANGLE, gm).run(sr.VS_SUZANNE).clip_primitives().finish(vp).run(sr.FS_SUZANNE)

    out["config1_suzanne_1024_us_per_frame"] = per_frame_us(suzanne, 64)
    for x in (pipe, gm, fb):
        x.destroy()

    w, h = 1920, 1080
    mesh = H.suzanne_mesh(with_uv=True)
    fb = P.RenderBuffer.with_dimensions(ctx, w, h)
    tex = P.Texture(ctx, scenes.checker_texture(512, 8))
    gm = P.Mesh(ctx, mesh)
    us = [scenes.full_example_uniforms(w / h, np.deg2rad(75.0), 2.0, np.deg2rad(rot), np.deg2rad(65.0), off)
          for rot, off in [(45.0, -1.6), (165.0, 0.0), (285.0, 1.6)]]
    pipe = P.Pipeline.from_framebuffer(fb, us[0])
    pipe.bind_texture(tex)
    pipe.set_sampler(sr.FILTER_BILINEAR, sr.EDGE_CLAMP)
    vp = scenes.Viewport.new(w, h, 0.1, 1000.0)

    def full_example():
        fb.clear(CLEAR)
        for u in us:
            pipe.set_uniforms(u)
            pipe.render_mesh(sr.TRIANGLE, gm).run(sr.VS_FULL_EXAMPLE).finish(vp).with_blend(sr.BLEND_ALPHA_OVER).run(sr.FS_FULL_EXAMPLE_TEXTURED)
        for u in us:
            pipe.set_uniforms(u)
            pipe.render_mesh(sr.TRIANGLE, gm).run(sr.VS_FULL_EXAMPLE).run(sr.GS_FACE_NORMALS).finish(vp).run(sr.FS_GREEN)

    out["config2_full_example_1080p_us_per_frame"] = per_frame_us(full_example, 32)
    out["config2_composition"] = "3 x 968 triangles, textured 4-light shader, alpha_over; then 3 x 968 face-normal lines, green shader"
    # the same frame with the twice-subdivided instances BASELINE's config 2 names (3 x 15,488 triangles; tests/test_gpu_sizes_full.py
    # compares this frame with the oracle)
    gm_hi = P.Mesh(ctx, scenes.subdivide(mesh, 2))
    gm_lo, gm = gm, gm_hi
    out["config2_subdivided_1080p_us_per_frame"] = per_frame_us(full_example, 16)
    out["config2_subdivided_composition"] = "3 x 15,488 triangles, textured 4-light shader, alpha_over; then 3 x 15,488 face-normal lines"
    gm = gm_lo
    gm_hi.destroy()

    # render-to-texture: the config-2 frame above sampled in place by a full-screen second pass
    # (2 triangles through the passthrough vertex shader, texture_unlit, Bilinear + Clamp) into a second 1920x1080 target
    fb2 = P.RenderBuffer.with_dimensions(ctx, w, h)
    pipe2 = P.Pipeline.from_framebuffer(fb2, us[0])
    pipe2.bind_framebuffer_texture(fb)
    pipe2.set_sampler(sr.FILTER_BILINEAR, sr.EDGE_CLAMP)
    quad = np.array([[-1, -1, 0, 1, 0, 1], [1, -1, 0, 1, 1, 1], [1, 1, 0, 1, 1, 0], [-1, 1, 0, 1, 0, 0]], np.float32)  # clip xyzw + uv
    qm = P.Mesh(ctx, vertices=quad, indices=np.array([0, 1, 2, 0, 2, 3], np.uint32))

    def second_pass():
        fb2.clear(CLEAR)
        pipe2.render_mesh(sr.TRIANGLE, qm).run_to_fragment(vp, sr.VS_PASSTHROUGH).run(sr.FS_TEXTURE_UNLIT)

    out["render_to_texture_second_pass_1080p_us"] = per_frame_us(second_pass, 32)
    covered = int((fb2.download()[:, 4] > np.float32(-3e38)).sum())
    out["render_to_texture_second_pass_pixels"] = covered
    for x in (pipe2, qm, fb2):
        x.destroy()
    for x in (pipe, gm, tex, fb):
        x.destroy()

    # config 5: the 64-frame Suzanne turntable at 1024x1024 on this GPU (clip path), device resident, one stream
    size = 1024
    mesh = H.suzanne_mesh()
    fb = P.RenderBuffer.with_dimensions(ctx, size, size)
    gm = P.Mesh(ctx, mesh)
    vp = scenes.Viewport.new(size, size, 0.001, 1000.0)
    tt_uniforms = [scenes.suzanne_uniforms(size, size, rotation_y=np.deg2rad(3.0 * (k + 1))) for k in range(64)]
    pipe = P.Pipeline.from_framebuffer(fb, tt_uniforms[0])

    def turntable():
        for uu in tt_uniforms:
            pipe.set_uniforms(uu)
            fb.clear(CLEAR)
            pipe.render_mesh(sr.TRIANGLE, gm).run(sr.VS_SUZANNE).clip_primitives().finish(vp).run(sr.FS_SUZANNE)

    out["config5_turntable_64_frames_fps_one_stream"] = 64e6 / per_frame_us(turntable, 3)
    out["config5_note"] = "bench.py --config turntable measures the same batch with four frames in flight, with read-back, and across GPUs"
    for x in (pipe, gm, fb):
        x.destroy()

    # the strictly ordered path at scale: the headline frame (config 3: 10 M triangles at 3840x2160) drawn with alpha_over
    if headline is not None:
        hfb, hpipe, hmesh, hvp, ntris = headline

        def ordered():
            hfb.clear(CLEAR)
            hpipe.render_mesh(sr.TRIANGLE, hmesh).run_to_fragment(hvp, sr.VS_SUZANNE).with_blend(sr.BLEND_ALPHA_OVER).run(sr.FS_SUZANNE)

        us = per_frame_us(ordered, 8)
        out["config3_alpha_over_ordered_ms_per_frame"] = us / 1e3
        out["config3_alpha_over_ordered_Mtris_per_s"] = ntris / us

    # config 4 on ONE GPU (the yardstick of the tile-sharded runs at N > 1): 100 M sub-pixel triangles at 7680x4320
    try:
        if os.environ.get("SR_NO_CONFIG4"):
            raise RuntimeError("skipped (SR_NO_CONFIG4)")
        w4, h4, mesh4, u4, vp4 = build_scene("grid100m")
        fb = P.RenderBuffer.with_dimensions(ctx, w4, h4)
        pipe = P.Pipeline.from_framebuffer(fb, u4)
        gm = P.Mesh(ctx, mesh4)

        def grid100m():
            fb.clear(CLEAR)
            pipe.render_mesh(sr.TRIANGLE, gm).run_to_fragment(vp4, sr.VS_SUZANNE).run(sr.FS_SUZANNE)

        us = per_frame_us(grid100m, 10)
        out["config4_grid100m_8k_ms_per_frame"] = us / 1e3
        out["config4_grid100m_8k_Mtris_per_s"] = mesh4.ntris / us
        b4 = algorithmic_bytes(len(mesh4.vertices), mesh4.ntris, w4, h4)
        out["config4_roofline_frac"] = b4 / (us * 1e-6) / 1e9 / measured_peak_gbs()[0]
        for x in (pipe, gm, fb):
            x.destroy()
    except Exception as e:
        out["config4_error"] = f"{type(e).__name__}: {e}"[:200]
    return out


def nvlink_counters(index: int):
    """Sum of the NVLink data counters of one GPU (nvidia-smi nvlink -gt d), bytes: (tx, rx) or None."""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True, timeout=10).stdout
        tx = rx = 0
        seen = False
        for ln in out.splitlines():
            ln = ln.strip()
            if "Data Tx" in ln or "Data Rx" in ln:
                kib = float(ln.split(":")[-1].strip().split()[0])
                seen = True
                if "Data Tx" in ln:
                    tx += int(kib * 1024)
                else:
                    rx += int(kib * 1024)
        return (tx, rx) if seen else None
    except Exception:
        return None


def run_sharded(name, world, rank, local_rank, P, sr, dist, barrier, max_over_ranks, nframes):
    """ONE frame split over the GPUs the way the north_star asks: the framebuffer is sharded by screen tiles (tile i belongs
    to rank i % N, the reference's tile-parallel loop fragment.rs:240-253 across GPUs), every finished tile is stored
    straight into rank 0's framebuffer over NVLink, and -- new in round 2 -- the per-triangle front end is sharded by
    triangle range: the tile owner merges the ranks' keys over NVLink inside its tile kernel (sr_shard).  Timed with CUDA
    events on the lanes' streams over `nframes` pipelined frames (two frames in flight = two lanes, no host barrier inside
    the batch), max over ranks; the composited frames are compared bit for bit with the single-GPU frame."""
    import torch
    w, h, mesh, u, vp = build_scene(name)
    ntris = mesh.ntris
    lanes = 2
    ctxs = [P.Context(local_rank) for _ in range(lanes)]
    for c in ctxs:
        c.set_tile_shard(rank, world)
    single_frame, single_ms = None, None
    if rank == 0:  # the unsharded frame on one GPU: the bit-exact yardstick and the time the speed-up is quoted against
        c1 = P.Context(local_rank)
        fb1 = P.RenderBuffer.with_dimensions(c1, w, h)
        p1 = P.Pipeline.from_framebuffer(fb1, u)
        m1 = P.Mesh(c1, mesh)
        st1 = torch.cuda.ExternalStream(c1.stream, device=torch.device("cuda", local_rank))

        def one():
            fb1.clear(CLEAR)
            p1.render_mesh(sr.TRIANGLE, m1).run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)
        for _ in range(3):
            one()
        c1.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st1)
        for _ in range(10):
            one()
        e1.record(st1)
        c1.synchronize()
        single_ms = e0.elapsed_time(e1) / 10
        single_frame = fb1.download()
        for x in (p1, m1, fb1):
            x.destroy()
        c1.close()
    group = P.ShardGroup(ctxs[0], w, h, lanes)
    handles = [None] * world
    dist.all_gather_object(handles, group.export())
    group.connect(handles)
    for lane, c in enumerate(ctxs):
        group.attach(c, lane)
    targets = [P.RenderBuffer.with_dimensions(ctxs[lane], w, h) for lane in range(lanes)] if rank == 0 else None
    th = [[t.ipc_export() for t in targets] if rank == 0 else None]
    dist.broadcast_object_list(th, src=0)
    fbs = targets if rank == 0 else [P.RenderBuffer.ipc_open(ctxs[lane], th[0][lane], w, h) for lane in range(lanes)]
    pipes = [P.Pipeline.from_framebuffer(fbs[lane], u) for lane in range(lanes)]
    meshes = [P.Mesh(ctxs[lane], mesh) for lane in range(lanes)]
    streams = [torch.cuda.ExternalStream(c.stream, device=torch.device("cuda", local_rank)) for c in ctxs]

    def frame(lane):
        fbs[lane].clear(CLEAR)
        pipes[lane].render_mesh(sr.TRIANGLE, meshes[lane]).run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)

    def batch(nlanes, n):
        for c in ctxs:
            c.synchronize()
        barrier()
        torch.cuda.synchronize()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(nlanes)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(nlanes)]
        for lane in range(nlanes):
            starts[lane].record(streams[lane])
        for f in range(n):
            frame(f % nlanes)
        for lane in range(nlanes):
            ends[lane].record(streams[lane])
        for c in ctxs:
            c.synchronize()
        torch.cuda.synchronize()
        barrier()
        return max_over_ranks(max(starts[0].elapsed_time(e) for e in ends)) / n

    for f in range(2 * lanes):
        frame(f % lanes)
    nv0 = nvlink_counters(local_rank) if rank == 0 else None
    ms_pipelined = batch(lanes, nframes)
    nv1 = nvlink_counters(local_rank) if rank == 0 else None
    ms_latency = batch(1, max(nframes // 2, 4))
    stage_ms = None
    if os.environ.get("SR_SHARD_STAGES"):  # diagnosis: per-stage device times of this rank's part of a sharded frame
        acc = {}
        ctxs[0].set_stage_timing(True)
        for _ in range(4):
            frame(0)
            for k, v in ctxs[0].stage_times().items():
                acc[k] = acc.get(k, 0.0) + v / 4
            barrier()
        ctxs[0].set_stage_timing(False)
        allst = [None] * world
        dist.all_gather_object(allst, {k: round(v, 4) for k, v in acc.items()})
        stage_ms = allst
    status = group.status()
    st = torch.tensor([status], device="cuda", dtype=torch.int64)
    dist.all_reduce(st, op=dist.ReduceOp.MAX)
    same = None
    if rank == 0:
        same = all(bool(np.array_equal(t.download().view(np.uint32), single_frame.view(np.uint32))) for t in targets)
    barrier()
    for c in ctxs:
        P.ShardGroup.detach(c)
    for x in pipes + meshes:
        x.destroy()
    if rank != 0:
        for fb in fbs:
            fb.destroy()
    barrier()
    if rank == 0:
        for fb in fbs:
            fb.destroy()
    group.destroy()
    for c in ctxs:
        c.close()
    if rank != 0:
        return None
    ingest = w * h * 20 * (world - 1) / world
    rec = {"config": name, "width": w, "height": h, "triangles": ntris, "n_gpus": world,
           "ms_per_frame": ms_pipelined, "Mtris_per_s": ntris / (ms_pipelined * 1e-3) / 1e6, "frames_per_s": 1e3 / ms_pipelined,
           "frames_in_flight": lanes, "frames_timed": nframes,
           "ms_per_frame_one_in_flight": ms_latency,
           "single_gpu_ms_per_frame": single_ms, "speedup_vs_single_gpu": single_ms / ms_pipelined,
           "speedup_vs_single_gpu_one_in_flight": single_ms / ms_latency,
           "identical_to_single_gpu_frame": same, "peer_wait_timeouts": int(st.item()),
           "rank0_ingest_bytes_per_frame": int(ingest),
           "rank0_ingest_bound_ms_at_770GBps": ingest / 770e9 * 1e3,
           "timing": "CUDA events on the lanes' streams around the whole batch of pipelined frames, max over ranks; one barrier per batch",
           "split": "tiles i % N (sort-first resolve, peer-store composite into rank 0) + triangle ranges r*T/N..(r+1)*T/N "
                    "(sort-last front end, keys merged over NVLink inside the tile kernel); vertex stage replicated"}
    if stage_ms:
        rec["stage_ms_per_rank"] = stage_ms
    if nv0 and nv1:
        rec["nvlink_rank0_rx_bytes_per_frame"] = (nv1[1] - nv0[1]) / nframes
        rec["nvlink_rank0_tx_bytes_per_frame"] = (nv1[0] - nv0[0]) / nframes
    return rec


def run_ours(args):
    import torch
    import torch.distributed as dist
    import softrender_b200 as sr
    from softrender_b200 import pipeline as P

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if os.environ.get("SR_NO_NUMA", "0") == "0" else "off"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    w, h, mesh, u, vp = build_scene(args.config)
    ntris, nverts = mesh.ntris, len(mesh.vertices)
    ctx = P.Context(local_rank)
    if os.environ.get("SR_MICRO"):  # tuning experiments only: "area,min_triangles,precheck"
        a, m, pc = (int(x) for x in os.environ["SR_MICRO"].split(","))
        ctx.set_micro(a, m, int(pc))
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    # every rank owns a framebuffer and the whole mesh (geometry replicated)
    fb = P.RenderBuffer.with_dimensions(ctx, w, h)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    gmesh = P.Mesh(ctx, mesh)

    def frame(pipeline=pipe, target=fb, m=gmesh):
        target.clear(CLEAR)
        pipeline.render_mesh(sr.TRIANGLE, m).run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)

    # Throughput is measured with independent frames in flight on `args.in_flight` contexts (= CUDA streams, each with its
    # own framebuffer): the HBM-bound vertex stage of one frame overlaps the issue-bound rasteriser of another and kernel
    # tails are filled.  Every frame is still a complete clear + vertex + raster + resolve of the whole workload.
    flights = [(ctx, fb, pipe, gmesh, stream)]
    for _ in range(1, max(1, args.in_flight)):
        cx = P.Context(local_rank)
        ffb = P.RenderBuffer.with_dimensions(cx, w, h)
        flights.append((cx, ffb, P.Pipeline.from_framebuffer(ffb, u), P.Mesh(cx, mesh),
                        torch.cuda.ExternalStream(cx.stream, device=torch.device("cuda", local_rank))))
    for lane in flights:  # untimed set-up: every stream allocates its scratch memory and warms its caches once
        for _ in range(3):
            frame(lane[2], lane[1], lane[3])
        lane[0].synchronize()

    def timed(lanes, steps, warmup, min_seconds=0.5, max_blocks=300):
        """W warm-up steps, then blocks of exactly K steps, each bracketed by barrier + synchronize and timed with CUDA
        events on the launching streams; blocks are repeated until the timed region reaches `min_seconds` (a 20-step block
        is only ~6 ms) and the MEDIAN block (max over ranks per block) is reported."""
        def go(i):
            cx, ffb, fp, fm, _ = lanes[i % len(lanes)]
            frame(fp, ffb, fm)

        def block():
            for lane in lanes:
                lane[0].synchronize()
            barrier()
            torch.cuda.synchronize()
            starts = [torch.cuda.Event(enable_timing=True) for _ in lanes]
            ends = [torch.cuda.Event(enable_timing=True) for _ in lanes]
            for lane, e in zip(lanes, starts):
                e.record(lane[4])
            for i in range(steps):
                go(i)
            for lane, e in zip(lanes, ends):
                e.record(lane[4])
            for lane in lanes:
                lane[0].synchronize()
            torch.cuda.synchronize()
            barrier()
            return max(starts[0].elapsed_time(e) for e in ends)

        for i in range(warmup):
            go(i)
        first = max_over_ranks(block())
        nblocks = int(min(max_blocks, max(3, np.ceil(min_seconds * 1e3 / max(first, 1e-3)))))
        times = [block() for _ in range(nblocks)]
        if world > 1:
            t = torch.tensor(times, device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times = t.tolist()
        return float(np.median(times)) / steps, nblocks, float(np.sum(times)) * 1e-3

    # ---- headline: whole frames per GPU (N>1: frame batching, independent frame streams per rank) ----
    sampler = ClockSampler(local_rank)
    if not os.environ.get("SR_NO_SAMPLER"):
        sampler.start()
    launches0 = sum(f[0].launch_count() for f in flights)
    ms_per_step, nblocks, region_s = timed(flights, args.steps, args.warmup, min_seconds=0.0 if args.quick else 0.5)
    launches_per_step = (sum(f[0].launch_count() for f in flights) - launches0) / (args.warmup + (nblocks + 1) * args.steps)
    single_ms, _, _ = timed(flights[:1], args.steps, args.warmup, min_seconds=0.0 if args.quick else 0.5) if len(flights) > 1 else (ms_per_step, 0, 0)
    in_flight_used = len(flights)
    if single_ms < ms_per_step:  # frames too large to profit from overlap (config 4): report the single-stream number
        ms_per_step, in_flight_used = single_ms, 1
    sampler.stop_flag = True
    if sampler.is_alive():
        sampler.join(timeout=2)
    for cx, ffb, fp, fm, _ in flights[1:]:
        fp.destroy()
        fm.destroy()
        ffb.destroy()
        cx.close()

    # per-stage device times (CUDA events recorded by the library on its stream), averaged over a few more frames
    stage_acc, nstage = {}, 5
    ctx.set_stage_timing(True)
    for _ in range(nstage):
        frame()
        for k, v in ctx.stage_times().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v / nstage
    ctx.set_stage_timing(False)

    # ---- N>1: ONE frame sharded over the GPUs (tiles + triangle ranges), configs 3 and 4 ----
    sharded, sharded4 = None, None
    if world > 1:
        sharded = run_sharded(args.config, world, rank, local_rank, P, sr, dist, barrier, max_over_ranks, max(20, args.steps))
        if args.config == "grid10m" and not os.environ.get("SR_NO_CONFIG4") and not (args.quick and os.environ.get("SR_QUICK_CONFIG3_ONLY")):
            try:
                sharded4 = run_sharded("grid100m", world, rank, local_rank, P, sr, dist, barrier, max_over_ranks, 20)
            except Exception as e:
                sharded4 = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---- end to end through the C ABI with HOST buffers: mesh upload + draw + framebuffer read-back each step ----
    # Three frames are in flight (three contexts = three CUDA streams, one host thread each, the C ABI releases the GIL): the
    # read-back of frame k (D2H) overlaps the mesh upload of frame k+1 (H2D) on the full-duplex PCIe link.  Every step
    # still copies its own inputs from pinned host memory and reads its own result back inside the timed region.
    host_v = torch.from_numpy(mesh.vertices).pin_memory().numpy()
    host_i = torch.from_numpy(mesh.indices.astype(np.uint32)).pin_memory().numpy()
    depth = 3
    lanes = []
    for k in range(depth):
        cx = ctx if k == 0 else P.Context(local_rank)
        lfb = fb if k == 0 else P.RenderBuffer.with_dimensions(cx, w, h)
        lp = pipe if k == 0 else P.Pipeline.from_framebuffer(lfb, u)
        lanes.append((cx, lfb, lp, torch.empty((w * h, 5), dtype=torch.float32).pin_memory().numpy()))
    e2e_steps = max(depth, (args.steps + depth - 1) // depth * depth)

    def e2e_frame(lane):
        cx, lfb, lp, host_fb = lane
        m = P.Mesh(cx, vertices=host_v, indices=host_i)  # H2D of this step's inputs (pinned host memory)
        lfb.clear(CLEAR)
        lp.render_mesh(sr.TRIANGLE, m).run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)
        lfb.download(host_fb)  # D2H of the step's result (synchronises)
        m.destroy()

    def e2e_run(nsteps):
        start = threading.Barrier(depth + 1)

        def work(lane):
            start.wait()
            for _ in range(nsteps // depth):
                e2e_frame(lane)

        threads = [threading.Thread(target=work, args=(lane,)) for lane in lanes]
        for t in threads:
            t.start()
        barrier()
        start.wait()
        t0 = time.perf_counter()
        for t in threads:
            t.join()
        return time.perf_counter() - t0

    if args.quick:
        e2e_steps = depth
    e2e_run(depth)  # warm-up (allocations, first touches)
    e2e_s = max_over_ranks(e2e_run(e2e_steps) / e2e_steps)
    barrier()
    # the last frame read back must be the frame the device-resident path produces
    ref_fb = fb.download()
    e2e_ok = bool(np.array_equal(lanes[-1][3].view(np.uint32), ref_fb.view(np.uint32)))
    e2e = {"value": world * ntris / e2e_s / 1e6, "unit": "Mtris/s", "frames_per_s": world / e2e_s,
           "h2d_bytes_per_step": int(host_v.nbytes + host_i.nbytes + 576) * world, "d2h_bytes_per_step": int(w * h * 20) * world,
           "ms_per_step": e2e_s * 1e3, "steps": e2e_steps, "frames_in_flight": depth, "host_affinity": numa, "result_matches_resident_path": e2e_ok,
           "note": "PCIe-bound: upload of frame k+1 overlaps read-back of frame k (three contexts/streams)"}
    for cx, lfb, lp, _ in lanes[1:]:
        lp.destroy()
        lfb.destroy()
        cx.close()

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        b_alg = algorithmic_bytes(nverts, ntris, w, h)
        achieved = b_alg / (ms_per_step * 1e-3) / 1e9
        # Per-kernel figures.  `pipeline_bytes` = what this implementation's launch has to move by design (inputs +
        # intermediates: shaded vertices, keys), `alg_bytes` = the launch's share of SURVEY 8d's algorithmic bytes only
        # (input vertices for k_vertex, indices for k_micro, the 20 B/pixel frame for the resolve).
        kernels = {
            "k_vertex": {"ms": stage_acc.get("vertex_ms", 0.0), "pipeline_bytes": nverts * (24 + 48), "alg_bytes": nverts * 24},
            "k_micro": {"ms": stage_acc.get("micro_ms", 0.0), "pipeline_bytes": 3 * ntris * 4 + nverts * 16, "alg_bytes": 3 * ntris * 4},
            "k_tile_opaque(+offsets)": {"ms": stage_acc.get("raster_ms", 0.0), "pipeline_bytes": w * h * (8 + 8 + 20), "alg_bytes": w * h * 20},
        }
        for k in kernels.values():
            for which in ("pipeline", "alg"):
                gbs = k[f"{which}_bytes"] / (k["ms"] * 1e-3) / 1e9 if k["ms"] > 0 else None
                k[f"{which}_GBps"] = gbs
                k[f"{which}_frac"] = gbs / peak if gbs else None
        dom = max(kernels, key=lambda n: kernels[n]["ms"])
        traffic = kernel_traffic().get(args.config, {})  # per config; absent -> null
        line = {
            "metric": f"Mtris/s at {w}x{h}", "value": world * ntris / (ms_per_step * 1e-3) / 1e6, "unit": "Mtris/s",
            "value_definition": f"throughput with {in_flight_used} independent frame(s) in flight per GPU; one full draw on one stream is "
                                f"single_stream_ms_per_frame (SURVEY.md 8d)",
            "frames_per_s": world * 1e3 / ms_per_step,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "timed_blocks": nblocks, "timed_region_s": region_s,
            "single_stream_ms_per_frame": single_ms, "single_stream_Mtris_per_s": ntris / (single_ms * 1e-3) / 1e6,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(args.config, args.in_flight, world),
            "frames_in_flight_used": in_flight_used,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_single_stream": b_alg / (single_ms * 1e-3) / 1e9 / peak,
                         "traffic": traffic.get("frame_dram_bytes"), "peak_source": peak_src, "algorithmic_bytes_per_frame": b_alg,
                         "scope": "whole frame: B_alg / frame time (SURVEY.md 8d); per-kernel figures under 'kernels'",
                         "dominant_kernel": {"name": dom, "achieved": kernels[dom]["alg_GBps"], "frac": kernels[dom]["alg_frac"],
                                             "alg_bytes_per_launch": kernels[dom]["alg_bytes"], "pipeline_bytes_per_launch": kernels[dom]["pipeline_bytes"],
                                             "pipeline_frac": kernels[dom]["pipeline_frac"], "ms_per_launch": kernels[dom]["ms"],
                                             "traffic": traffic.get(dom),
                                             "note": "issue-slot bound, not HBM bound (DESIGN.md section 4): see profiles/"},
                         "kernels": kernels,
                         "stage_ms_avg": stage_acc},
            "e2e": e2e,
            "gpu_launches": int(round(launches_per_step * args.steps)),
            "clocks": sampler.summary(),
        }
        if sharded:
            line["sharded"] = sharded
        if sharded4:
            line["sharded_config4"] = sharded4
        if world == 1 and args.config == "grid10m" and not args.quick:
            try:
                line["other_configs"] = other_configs(ctx, (fb, pipe, gmesh, vp, mesh.ntris))
            except Exception as e:  # never let the side measurements take the headline line down
                line["other_configs"] = {"error": f"{type(e).__name__}: {e}"[:200]}
        if world == 1 and not args.no_cpu_baseline and not args.quick:
            line["cpu_baseline"] = cpu_baseline_sample(args.config, native_oracle())
        print(json.dumps(line))
    barrier()
    for x in (pipe, gmesh, fb):
        x.destroy()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------
# config 5: 64-frame turntable of the Suzanne scene (frames/s; frames k % N per GPU)
# ------------------------------------------------------------------------------------------------------------
def run_turntable(args):
    import torch
    import torch.distributed as dist
    import softrender_b200 as sr
    from softrender_b200 import pipeline as P, scenes, sharding
    import helpers as H

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    size, nframes = 1024, 64
    mesh = H.suzanne_mesh()
    vp = scenes.Viewport.new(size, size, 0.001, 1000.0)
    uniforms = [scenes.suzanne_uniforms(size, size, rotation_y=np.deg2rad(3.0 * (k + 1))) for k in range(nframes)]
    mine = sharding.frames_for_rank(nframes, rank, world)

    if args.impl == "reference":
        if rank != 0:
            return
        import oracle_binding as ob
        cores = os.cpu_count() or 1
        fbo = ob.OracleFramebuffer(size, size)
        t0 = time.perf_counter()
        for u in uniforms:
            fbo.clear(CLEAR)
            d = ob.OracleDraw(sr.TRIANGLE, mesh.indices)
            d.tile = (128, 128)
            d.vertex_run(sr.VS_SUZANNE, u, mesh.vertices, nthreads=cores).clip_primitives().finish(vp)
            d.fragment_run(fbo, sr.FS_SUZANNE, u, nthreads=cores)
        dt = time.perf_counter() - t0
        print(json.dumps({"impl": "reference", "metric": "frames/s, 64-frame Suzanne turntable at 1024x1024", "value": nframes / dt,
                          "unit": "frames/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": dt * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": "turntable", "frames": nframes, "width": size, "height": size, "triangles": mesh.ntris},
                          "cpu_baseline": {"value": nframes / dt, "unit": "frames/s", "cores": cores, "kind": "port",
                                           "sample": "all 64 frames, restated reference CPU path"},
                          "e2e": {"value": nframes / dt, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # Frames are independent, so `depth` of them are in flight per GPU: one context (= CUDA stream, framebuffer, host
    # thread) each; the C ABI releases the GIL.  Every frame is still a full clear + draw (+ read-back for e2e).
    depth = max(1, args.in_flight)
    lanes = []
    for k in range(depth):
        cx = P.Context(local_rank)
        lfb = P.RenderBuffer.with_dimensions(cx, size, size)
        lanes.append((cx, lfb, P.Pipeline.from_framebuffer(lfb, uniforms[0]), P.Mesh(cx, mesh),
                      torch.empty((size * size, 5), dtype=torch.float32).pin_memory().numpy()))
    ctx = lanes[0][0]

    def lane_frames(lane, frames, readback):
        cx, lfb, lp, lm, host_fb = lane
        host_fb8 = host_fb.reshape(-1).view(np.uint8)[:size * size * 4].reshape(size, size, 4)
        for k in frames:
            lp.set_uniforms(uniforms[k])
            lfb.clear(CLEAR)
            lp.render_mesh(sr.TRIANGLE, lm).run(sr.VS_SUZANNE).clip_primitives().finish(vp).run(sr.FS_SUZANNE)
            if readback == 1:
                lfb.download(host_fb)
            elif readback == 2:  # presentation read-back: (c * 255) as u8 on the device, 4 B/pixel (realtime_example/src/main.rs:100-116)
                lfb.download_rgba8(host_fb8, abgr=True)
        cx.synchronize()

    def batch(readback):
        threads = [threading.Thread(target=lane_frames, args=(lane, mine[i::depth], readback)) for i, lane in enumerate(lanes)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()

    def timed(readback):
        for _ in range(max(args.warmup, 1)):
            batch(readback)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            batch(readback)
        if world > 1:
            dist.barrier()
        return sharding.max_over_ranks((time.perf_counter() - t0) / args.steps)

    def launch_total():
        return sum(lane[0].launch_count() for lane in lanes)

    launches0 = launch_total()
    t_res = timed(0)
    launches = (launch_total() - launches0) // (args.steps + max(args.warmup, 1)) * args.steps
    t_e2e = timed(1)
    t_present = timed(2)
    if rank == 0:
        print(json.dumps({"metric": "frames/s, 64-frame Suzanne turntable at 1024x1024", "value": nframes / t_res, "unit": "frames/s",
                          "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": t_res * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": "turntable", "frames": nframes, "width": size, "height": size, "triangles": mesh.ntris,
                                     "path": "render_mesh -> vertex run -> clip_primitives -> finish -> fragment run (examples/suzanne.rs:121-147)",
                                     "parallelism": f"frames k % {world} per GPU", "frames_in_flight": depth},
                          "e2e": {"value": nframes / t_e2e, "unit": "frames/s", "h2d_bytes_per_step": 576 * len(mine) * world,
                                  "d2h_bytes_per_step": size * size * 20 * nframes, "note": "uniform upload + framebuffer read-back per frame"},
                          "e2e_present": {"value": nframes / t_present, "unit": "frames/s", "d2h_bytes_per_step": size * size * 4 * nframes,
                                          "note": "read-back as RGBA8 converted on the device (sr_framebuffer_download_rgba8), the realtime_example presentation path"},
                          "gpu_launches": int(launches)}))
    for cx, lfb, lp, lm, _ in lanes:
        lp.destroy()
        lm.destroy()
        lfb.destroy()
        cx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="grid10m", choices=sorted(CONFIGS) + ["turntable"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="development runs: short timed region, no side measurements (not a bench line)")
    ap.add_argument("--in-flight", type=int, default=3, help="independent frames in flight per GPU (CUDA streams)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # no cyclic-garbage collections inside timed regions (a full collection is a millisecond-scale host pause, several frames
    # of this workload; seen as a +50..85 us/frame outlier in profiles/scripts/rtt.py before it did the same)
    import gc
    gc.disable()
    if args.config == "turntable":
        run_turntable(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
