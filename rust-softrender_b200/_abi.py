"""ctypes binding of libsoftrender_b200.so (the C ABI of include/softrender_b200.h).

There is no CPU fallback: if the CUDA library has not been built, importing this module raises."""
from __future__ import annotations

import ctypes
import os

from .scenes import Uniforms, Viewport

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SOFTRENDER_B200_LIB") or os.path.join(_HERE, "csrc", "libsoftrender_b200.so")  # env override: tuning builds only

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()' "
        "or make -C rust-softrender_b200/csrc). softrender_b200 has no CPU fallback.")

# Range-sharded frames keep a spinning wait kernel on the stream; a lazily loaded kernel's first launch could stall behind it.
# The library pre-loads what such frames launch (sr_shard_create); eager loading covers the rest when CUDA is not up yet.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
lib = ctypes.CDLL(LIB_PATH)

c_void_p, c_int, c_u32, c_u64, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_size_t
f32p = ctypes.POINTER(ctypes.c_float)
u32p = ctypes.POINTER(ctypes.c_uint32)
u64p = ctypes.POINTER(ctypes.c_uint64)
u8p = ctypes.POINTER(ctypes.c_uint8)
pp = ctypes.POINTER(c_void_p)


class StageTimes(ctypes.Structure):
    _fields_ = [("vertex_ms", ctypes.c_float), ("geometry_ms", ctypes.c_float), ("bin_ms", ctypes.c_float),
                ("vis_init_ms", ctypes.c_float), ("micro_ms", ctypes.c_float), ("raster_ms", ctypes.c_float), ("total_ms", ctypes.c_float)]


# every symbol include/softrender_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "sr_last_error": (ctypes.c_char_p, []),
    "sr_version": (c_int, []),
    "sr_tile_size": (c_int, [u32p, u32p]),
    "sr_context_create": (c_int, [c_int, pp]),
    "sr_registry_entry": (c_int, [c_u32, c_u32, c_void_p]),
    "sr_context_destroy": (c_int, [c_void_p]),
    "sr_context_synchronize": (c_int, [c_void_p]),
    "sr_context_stream": (c_void_p, [c_void_p]),
    "sr_context_set_tile_shard": (c_int, [c_void_p, c_u32, c_u32]),
    "sr_context_set_micro": (c_int, [c_void_p, c_u32, c_u32, c_u32]),
    "sr_context_set_stage_timing": (c_int, [c_void_p, c_int]),
    "sr_context_stage_timestamps": (c_int, [c_void_p, c_void_p, f32p]),
    "sr_context_wait_for": (c_int, [c_void_p, c_void_p, c_u32]),
    "sr_context_set_list_capacity": (c_int, [c_void_p, c_u32]),
    "sr_context_list_capacity": (c_int, [c_void_p, u32p]),
    "sr_context_ordered_list_capacity": (c_int, [c_void_p, u32p]),
    "sr_context_launch_count": (c_int, [c_void_p, u64p]),
    "sr_context_stage_times": (c_int, [c_void_p, ctypes.POINTER(StageTimes)]),
    "sr_framebuffer_create": (c_int, [c_void_p, c_u32, c_u32, c_u32, pp]),
    "sr_framebuffer_destroy": (c_int, [c_void_p]),
    "sr_framebuffer_clear": (c_int, [c_void_p, f32p]),
    "sr_framebuffer_dimensions": (c_int, [c_void_p, u32p, u32p]),
    "sr_framebuffer_download": (c_int, [c_void_p, c_void_p, c_size_t]),
    "sr_framebuffer_download_rgba8": (c_int, [c_void_p, u8p, c_size_t, c_u32]),
    "sr_framebuffer_download_planes": (c_int, [c_void_p, c_void_p, f32p, c_void_p]),
    "sr_framebuffer_upload_planes": (c_int, [c_void_p, c_void_p, f32p, c_void_p]),
    "sr_framebuffer_get_pixel": (c_int, [c_void_p, c_u32, c_u32, f32p, f32p, u32p]),
    "sr_framebuffer_set_pixel": (c_int, [c_void_p, c_u32, c_u32, f32p, f32p, u32p]),
    "sr_framebuffer_clear_attachment": (c_int, [c_void_p, c_u32, f32p]),
    "sr_framebuffer_download_attachment": (c_int, [c_void_p, c_u32, f32p]),
    "sr_framebuffer_enable_winner": (c_int, [c_void_p, c_int]),
    "sr_framebuffer_download_winner": (c_int, [c_void_p, u32p]),
    "sr_framebuffer_device_ptr": (c_void_p, [c_void_p]),
    "sr_framebuffer_ipc_export": (c_int, [c_void_p, c_void_p]),
    "sr_framebuffer_ipc_open": (c_int, [c_void_p, c_void_p, c_u32, c_u32, c_u32, pp]),
    "sr_framebuffer_alias": (c_int, [c_void_p, c_void_p, pp]),
    "sr_shard_create": (c_int, [c_void_p, c_u32, c_u32, c_u32, pp]),
    "sr_shard_export": (c_int, [c_void_p, c_void_p]),
    "sr_shard_connect": (c_int, [c_void_p, c_void_p, c_u32]),
    "sr_shard_connect_local": (c_int, [c_void_p, pp, c_u32]),
    "sr_context_attach_shard": (c_int, [c_void_p, c_void_p, c_u32]),
    "sr_shard_status": (c_int, [c_void_p, u32p]),
    "sr_shard_destroy": (c_int, [c_void_p]),
    "sr_mesh_create": (c_int, [c_void_p, c_void_p, c_u64, c_u32, c_void_p, c_u64, c_u32, pp]),
    "sr_mesh_destroy": (c_int, [c_void_p]),
    "sr_texture_create": (c_int, [c_void_p, u8p, c_u32, c_u32, pp]),
    "sr_texture_destroy": (c_int, [c_void_p]),
    "sr_pipeline_create": (c_int, [c_void_p, c_void_p, ctypes.POINTER(Uniforms), pp]),
    "sr_pipeline_destroy": (c_int, [c_void_p]),
    "sr_pipeline_set_uniforms": (c_int, [c_void_p, ctypes.POINTER(Uniforms)]),
    "sr_pipeline_set_framebuffer": (c_int, [c_void_p, c_void_p]),
    "sr_pipeline_set_stencil_config": (c_int, [c_void_p, c_u32, c_u32]),
    "sr_pipeline_bind_texture": (c_int, [c_void_p, c_void_p]),
    "sr_pipeline_bind_framebuffer_texture": (c_int, [c_void_p, c_void_p]),
    "sr_pipeline_bind_framebuffer_attachment": (c_int, [c_void_p, c_void_p, c_u32]),
    "sr_pipeline_set_sampler": (c_int, [c_void_p, c_u32, c_u32, ctypes.POINTER(ctypes.c_float)]),
    "sr_render_mesh": (c_int, [c_void_p, c_void_p, c_u32, c_int, c_u32, pp]),
    "sr_vertex_run": (c_int, [c_void_p, c_u32]),
    "sr_vertex_run_to_fragment": (c_int, [c_void_p, ctypes.POINTER(Viewport), c_u32]),
    "sr_geometry_run": (c_int, [c_void_p, c_u32]),
    "sr_geometry_clip_primitives": (c_int, [c_void_p]),
    "sr_geometry_finish": (c_int, [c_void_p, ctypes.POINTER(Viewport)]),
    "sr_draw_duplicate": (c_int, [c_void_p, pp]),
    "sr_fragment_set_cull_faces": (c_int, [c_void_p, c_u32]),
    "sr_fragment_set_antialiased_lines": (c_int, [c_void_p, c_int]),
    "sr_fragment_set_tile_size": (c_int, [c_void_p, c_u32, c_u32]),
    "sr_fragment_set_blend": (c_int, [c_void_p, c_u32]),
    "sr_fragment_run": (c_int, [c_void_p, c_u32]),
    "sr_draw_destroy": (c_int, [c_void_p]),
    "sr_draw_from_vertices": (c_int, [c_void_p, c_u32, f32p, c_u64, c_u32, c_int, u32p, c_u64, c_int, c_u32, pp]),
    "sr_draw_set_generated": (c_int, [c_void_p, c_int, f32p, c_u64, c_u32]),
    "sr_draw_count": (c_int, [c_void_p, c_int, u64p, u32p]),
    "sr_draw_download": (c_int, [c_void_p, c_int, f32p, c_u64]),
    "sr_draw_download_sequence": (c_int, [c_void_p, u32p, c_u64]),
    "sr_context_last_opaque_lists": (c_int, [c_void_p, u64p, u32p, c_u64, u64p, u32p]),
    "sr_draw_bins": (c_int, [c_void_p, u64p, u32p, c_u64, u64p]),
    "sr_selftest_division": (c_int, [c_void_p, c_u64, c_u64, u64p]),
}

for _name, (_res, _args) in SYMBOLS.items():
    _fn = getattr(lib, _name)  # AttributeError here = the library does not export a declared symbol
    _fn.restype = _res
    _fn.argtypes = _args


class SoftrenderError(RuntimeError):
    """A non-zero sr_status; `.status` holds the code (the reference panics or returns RenderError here)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"softrender_b200 status {status}: {message}")
        self.status = status


def check(status: int) -> None:
    if status != 0:
        raise SoftrenderError(status, (lib.sr_last_error() or b"").decode("utf-8", "replace"))
