#!/usr/bin/env python
"""bench.py -- headline benchmark of the draw hot path (BASELINE.json: Mtris/s and frames/s at 3840x2160).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config grid10m|...]

A "step" is one frame: clear + vertex stage + binning + tile rasterisation of the synthetic 10M-triangle
shaded mesh at 3840x2160 (config 3) with the mesh already resident in HBM.
N>1 (torchrun, one rank per GPU): sort-first tile sharding -- geometry replicated, GPU tiles interleaved over
the ranks, every rank's tile rasteriser stores its finished tiles straight into rank 0's framebuffer over
NVLink peer memory (no staging copy, no separate collective); NCCL carries only the barrier and the timing
reduction.  Prints ONE JSON line (rank 0).
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
    "grid1m": (3840, 2160, 395, 316, 4, 0x5EED0003, 0.1, 100.0),      # reduced, for quick checks only
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
            time.sleep(0.2)

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
def cpu_reference_sample(name: str, steps: int = 1, warmup: int = 0):
    """Times the oracle's threaded mode (thread pool + atomic cursors, every 128x128 tile visits every primitive:
    the reference's structure) on a BOUNDED sample of the workload: same frame size, camera and shader, the mesh
    generator at 1/16 of the triangle count.  Cost is O(tiles x triangles), so Mtris/s is size-independent."""
    import softrender_b200 as sr
    from softrender_b200 import scenes
    import oracle_binding as ob
    w, h, nx, ny, layers, seed, near, far = CONFIGS[name]
    mesh = scenes.make_grid(max(nx // 4, 1), max(ny // 4, 1), layers, seed=seed)
    u = scenes.grid_uniforms(w, h)
    vp = scenes.Viewport.new(w, h, near, far)
    cores = os.cpu_count() or 1
    fb = ob.OracleFramebuffer(w, h)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        fb.clear(CLEAR)
        d = ob.OracleDraw(sr.TRIANGLE, mesh.indices)
        d.tile = (128, 128)
        d.vertex_run_to_fragment(vp, sr.VS_SUZANNE, u, mesh.vertices, nthreads=cores)
        d.fragment_run(fb, sr.FS_SUZANNE, u, nthreads=cores)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    t = float(np.mean(times))
    return {"value": mesh.ntris / t / 1e6, "unit": "Mtris/s", "cores": cores, "kind": "port",
            "sample": f"{name} frame {w}x{h}, generator at 1/16 triangles ({mesh.ntris} tris), 128x128 reference tiles, "
                      f"{t:.2f} s/frame on {cores} threads"}, t, mesh.ntris


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w, h = CONFIGS[args.config][:2]
    base, t, ntris = cpu_reference_sample(args.config, steps=max(1, min(args.steps, 3)), warmup=0)
    line = {
        "impl": "reference", "metric": f"Mtris/s at {w}x{h}", "value": base["value"], "unit": "Mtris/s",
        "frames_per_s": base["value"] * 1e6 / (CONFIGS[args.config][2] * CONFIGS[args.config][3] * CONFIGS[args.config][4] * 2),
        "n_gpus": args.gpus, "steps": max(1, min(args.steps, 3)), "warmup": 0, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.config, "width": w, "height": h, "note": "restated reference CPU path (the Rust reference cannot be built here)"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "Mtris/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import softrender_b200 as sr
    from softrender_b200 import pipeline as P

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    w, h, mesh, u, vp = build_scene(args.config)
    ntris, nverts = mesh.ntris, len(mesh.vertices)
    ctx = P.Context(local_rank)
    if os.environ.get("SR_MICRO"):  # tuning experiments only: "area,min_triangles,precheck"
        a, m, pc = (int(x) for x in os.environ["SR_MICRO"].split(","))
        ctx.set_micro(a, m, bool(pc))
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    # framebuffer: rank 0 owns it; other ranks map it through CUDA IPC and store their tiles into it over NVLink
    if rank == 0:
        fb = P.RenderBuffer.with_dimensions(ctx, w, h)
        fb.clear(CLEAR)
        handle = [fb.ipc_export()] if world > 1 else [None]
    else:
        handle = [None]
    if world > 1:
        dist.broadcast_object_list(handle, src=0)
        ctx.set_tile_shard(rank, world)
        if rank != 0:
            fb = P.RenderBuffer.ipc_open(ctx, handle[0], w, h)
    pipe = P.Pipeline.from_framebuffer(fb, u)
    gmesh = P.Mesh(ctx, mesh)

    def frame():
        fb.clear(CLEAR)
        pipe.render_mesh(sr.TRIANGLE, gmesh).run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)

    for _ in range(args.warmup):
        frame()
    ctx.synchronize()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(args.steps):
        frame()
    ev1.record(stream)
    ctx.synchronize()
    torch.cuda.synchronize()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - launches0
    stage = ctx.stage_times()  # CUDA events of the last timed step, on the launching stream
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps

    # ---- end to end through the C ABI with HOST buffers: mesh upload + draw + framebuffer read-back each step ----
    e2e = None
    host_v = torch.from_numpy(mesh.vertices).pin_memory().numpy()
    host_i = torch.from_numpy(mesh.indices.astype(np.uint32)).pin_memory().numpy()
    host_fb = torch.empty((w * h, 5), dtype=torch.float32).pin_memory().numpy() if rank == 0 else None
    e2e_steps = max(1, min(args.steps, 5))

    def e2e_frame():
        m = P.Mesh(ctx, vertices=host_v, indices=host_i)  # H2D of this step's inputs (pinned host memory)
        fb.clear(CLEAR)
        pipe.render_mesh(sr.TRIANGLE, m).run_to_fragment(vp, sr.VS_SUZANNE).run(sr.FS_SUZANNE)
        ctx.synchronize()
        barrier()
        if rank == 0:
            fb.download(host_fb)  # D2H of the step's result
        m.destroy()

    e2e_frame()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_frame()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": ntris / e2e_s / 1e6, "unit": "Mtris/s", "frames_per_s": 1.0 / e2e_s,
           "h2d_bytes_per_step": int(host_v.nbytes + host_i.nbytes + 576), "d2h_bytes_per_step": int(w * h * 20),
           "ms_per_step": e2e_s * 1e3, "steps": e2e_steps}

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        b_alg = algorithmic_bytes(nverts, ntris, w, h)
        achieved = b_alg / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": f"Mtris/s at {w}x{h}", "value": ntris / (ms_per_step * 1e-3) / 1e6, "unit": "Mtris/s",
            "frames_per_s": 1e3 / ms_per_step,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.config, "width": w, "height": h, "triangles": ntris, "vertices": nverts,
                       "shader": "suzanne Blinn-Phong", "depth_test": True,
                       "parallelism": "1 GPU" if world == 1 else f"sort-first tile sharding over {world} GPUs, peer-store composite to rank 0",
                       "l2": "working set per frame (mesh 240 MB + shaded vertices 240 MB + framebuffer 166 MB) exceeds the 126 MB L2; no flush needed"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_frame": b_alg,
                         "scope": "whole frame (vertex + binning + tile raster kernels): B_alg / frame time",
                         "stage_ms_last_step": stage},
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], _, _ = cpu_reference_sample(args.config)
        print(json.dumps(line))
    barrier()
    for x in (pipe, gmesh):
        x.destroy()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="grid10m", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
