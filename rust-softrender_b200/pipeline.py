"""Host-side mirror of the reference's builder API over the C ABI.

Same verbs, argument meaning and error behaviour as the Rust types (panics become SoftrenderError):

    Pipeline.from_framebuffer(fb, uniforms)              src/pipeline/mod.rs:120
      .render_mesh(TRIANGLE, mesh, stencil=None)         src/pipeline/mod.rs:146            -> VertexShader
         .run(VS_*)                                      src/pipeline/stages/vertex.rs:87   -> GeometryShader
            .run(GS_*) / .clip_primitives()              src/pipeline/stages/geometry.rs:132,261
            .finish(viewport)                            src/pipeline/stages/geometry.rs:60 -> FragmentShader
         .run_to_fragment(viewport, VS_*)                src/pipeline/stages/vertex.rs:123  -> FragmentShader
               .with_blend / cull_faces / tile_size ...  src/pipeline/stages/fragment.rs:82-160
               .run(FS_*)                                src/pipeline/stages/fragment.rs:168

Shaders are ids of the registered CUDA device functions (softrender_b200.constants) instead of closures.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _abi
from ._abi import check, lib
from .constants import *  # noqa: F401,F403
from .scenes import MeshData, Uniforms, Viewport


def _vp(p=None):
    return ctypes.c_void_p(p)


class Context:
    """Owns the device, the CUDA stream and scratch memory (replaces the thread pool of Pipeline::new)."""

    def __init__(self, device: int = 0):
        h = ctypes.c_void_p()
        check(lib.sr_context_create(device, ctypes.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            lib.sr_context_destroy(self.h)
            self.h = None

    def synchronize(self):
        check(lib.sr_context_synchronize(self.h))

    @property
    def stream(self) -> int:
        return lib.sr_context_stream(self.h) or 0

    def set_tile_shard(self, rank: int, world: int):
        check(lib.sr_context_set_tile_shard(self.h, rank, world))

    def set_micro(self, area: int = 0xFFFFFFFF, min_triangles: int = 65536, precheck: int = 0):
        """Tuning of the opaque triangle path (results never depend on it): see sr_context_set_micro.
        precheck: bit 0 = per-fragment key pre-check, bit 1 = early depth rejection off (library default 0)."""
        check(lib.sr_context_set_micro(self.h, area, min_triangles, int(precheck) & 3))

    def set_stage_timing(self, enable: bool = True):
        """Record the per-stage CUDA events `stage_times` / `stage_timestamps` read (off by default)."""
        check(lib.sr_context_set_stage_timing(self.h, 1 if enable else 0))

    def stage_timestamps(self, base_event: int):
        """ms from the caller's cudaEvent_t `base_event` to the stage events of the latest draw (see the C header)."""
        out = (ctypes.c_float * 8)()
        check(lib.sr_context_stage_timestamps(self.h, ctypes.c_void_p(base_event), out))
        return list(out)

    def wait_for(self, other: "Context", point: int = 0):
        """Work enqueued on this context from now on starts only when `other` has reached `point`
        (0: everything enqueued so far, 1: the raster front end of its latest opaque draw)."""
        check(lib.sr_context_wait_for(self.h, other.h, point))

    def last_opaque_lists(self, ntiles: int):
        """(offsets[ntiles + 1], ids, micro_area): the per-tile triangle lists the opaque fast path built for its latest draw."""
        total, area = ctypes.c_uint64(), ctypes.c_uint32()
        check(lib.sr_context_last_opaque_lists(self.h, None, None, 0, ctypes.byref(total), ctypes.byref(area)))
        n = total.value
        ids, off = np.zeros(max(n, 1), np.uint32), np.zeros(ntiles + 1, np.uint64)
        check(lib.sr_context_last_opaque_lists(self.h, off.ctypes.data_as(_abi.u64p), ids.ctypes.data_as(_abi.u32p), n,
                                               ctypes.byref(total), ctypes.byref(area)))
        return off, ids[:n], area.value

    def set_list_capacity(self, entries: int):
        check(lib.sr_context_set_list_capacity(self.h, entries))

    def ordered_list_capacity(self):
        """(points, lines, triangles) capacities of the ordered path's group-list arenas."""
        n = (ctypes.c_uint32 * 3)()
        check(lib.sr_context_ordered_list_capacity(self.h, n))
        return tuple(n)

    def list_capacity(self) -> int:
        n = ctypes.c_uint32()
        check(lib.sr_context_list_capacity(self.h, ctypes.byref(n)))
        return n.value

    def launch_count(self) -> int:
        n = ctypes.c_uint64()
        check(lib.sr_context_launch_count(self.h, ctypes.byref(n)))
        return n.value

    def selftest_division(self, seed: int, count: int) -> int:
        bad = ctypes.c_uint64()
        check(lib.sr_selftest_division(self.h, seed, count, ctypes.byref(bad)))
        return bad.value

    def stage_times(self) -> dict:
        t = _abi.StageTimes()
        check(lib.sr_context_stage_times(self.h, ctypes.byref(t)))
        return {k: getattr(t, k) for k, _ in t._fields_}


class ShaderInfo(ctypes.Structure):  # sr_shader_info
    _fields_ = [("id", ctypes.c_uint32), ("vin_floats", ctypes.c_uint32), ("nk", ctypes.c_uint32), ("discards", ctypes.c_uint32),
                ("needs_texture", ctypes.c_uint32), ("name", ctypes.c_char * 40), ("reference", ctypes.c_char * 72)]


def registry(kind: int):
    """The registered device functions of one kind (0 vertex, 1 geometry, 2 fragment, 3 blend) that stand in for the
    reference's closures: list of dicts with id, name, expected layouts and the reference closure each one mirrors."""
    out, i = [], 0
    while True:
        info = ShaderInfo()
        if lib.sr_registry_entry(kind, i, ctypes.byref(info)) != 0:
            return out
        out.append({"id": info.id, "name": info.name.decode(), "vin_floats": info.vin_floats, "nk": info.nk,
                    "discards": bool(info.discards), "needs_texture": bool(info.needs_texture), "reference": info.reference.decode()})
        i += 1


def write_png(path: str, rgba: np.ndarray):
    """Minimal RGBA8 PNG writer (filter type 0 on every row)."""
    import struct
    import zlib
    h, w, ch = rgba.shape
    assert ch == 4 and rgba.dtype == np.uint8
    raw = np.concatenate([np.zeros((h, 1), np.uint8), np.ascontiguousarray(rgba).reshape(h, w * 4)], axis=1).tobytes()

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def tile_size():
    w, h = ctypes.c_uint32(), ctypes.c_uint32()
    check(lib.sr_tile_size(ctypes.byref(w), ctypes.byref(h)))
    return w.value, h.value


class RenderBuffer:
    """RenderBuffer<ColorDepth[Stencil]Attachments<RGBAf32Color, f32[, u8]>> (src/framebuffer/renderbuffer/mod.rs)."""

    def __init__(self, ctx: Context, handle, width, height, fmt):
        self.ctx, self.h, self.width, self.height, self.format = ctx, handle, width, height, fmt

    _STENCIL_DTYPE = {FB_RGBAF32_DF32_S8: np.uint8, FB_RGBAF32_DF32_S16: np.uint16, FB_RGBAF32_DF32_S32: np.uint32,
                      FB_RGBAU8_DF32_S8: np.uint8, FB_TEXTURE_RGBAF32_DF32_S8: np.uint8}
    U8_PIXEL = np.dtype([("rgba", np.uint8, 4), ("depth", np.float32)])  # the 8-byte AoS pixel of an RGBAu8Color target

    @property
    def u8_color(self) -> bool:
        return self.format in (FB_RGBAU8_DF32, FB_RGBAU8_DF32_S8)

    @staticmethod
    def with_dimensions(ctx: Context, width: int, height: int, stencil=False, u8_color: bool = False,
                        texture_buffer: bool = False) -> "RenderBuffer":
        """stencil: False = stencil type `()`, True / 8 = u8, 16 = u16, 32 = u32 (the Stencil trait, src/stencil.rs:9-60).
        u8_color: colour attachment RGBAu8Color instead of RGBAf32Color (src/color/predefined.rs:17,26).
        texture_buffer: RGBAf32TextureBuffer storage -- colour and depth in planes of their own, the colour plane re-usable
        as a texture without copying (src/framebuffer/texturebuffer.rs:63-66,200-210)."""
        if texture_buffer == 2:  # declare_texture_buffer! with two colour planes (texturebuffer.rs:72-110)
            if stencil:
                raise ValueError("two-plane texture buffers have stencil type ()")
            fmt = FB_TEXTURE_2xRGBAF32_DF32
        elif texture_buffer:
            fmt = {False: FB_TEXTURE_RGBAF32_DF32, True: FB_TEXTURE_RGBAF32_DF32_S8, 8: FB_TEXTURE_RGBAF32_DF32_S8}[stencil]
        elif u8_color:
            fmt = {False: FB_RGBAU8_DF32, True: FB_RGBAU8_DF32_S8, 8: FB_RGBAU8_DF32_S8}[stencil]
        else:
            fmt = {False: FB_RGBAF32_DF32, True: FB_RGBAF32_DF32_S8, 8: FB_RGBAF32_DF32_S8, 16: FB_RGBAF32_DF32_S16,
                   32: FB_RGBAF32_DF32_S32}[stencil]
        h = ctypes.c_void_p()
        check(lib.sr_framebuffer_create(ctx.h, width, height, fmt, ctypes.byref(h)))
        return RenderBuffer(ctx, h, width, height, fmt)

    def destroy(self):
        if self.h:
            lib.sr_framebuffer_destroy(self.h)
            self.h = None

    def dimensions(self):
        return self.width, self.height

    def clear(self, color):
        c = (ctypes.c_float * 4)(*[float(x) for x in color])
        check(lib.sr_framebuffer_clear(self.h, c))

    def clear_attachment(self, index: int, color):
        """The colour of ONE plane of Framebuffer::clear's tuple (texturebuffer.rs:181-197); records a clear of the whole buffer."""
        c = (ctypes.c_float * 4)(*[float(x) for x in color])
        check(lib.sr_framebuffer_clear_attachment(self.h, index, c))

    def download_attachment(self, index: int) -> np.ndarray:
        """Colour plane `index` of a texture buffer: float32 [height*width, 4]."""
        out = np.empty((self.width * self.height, 4), np.float32)
        check(lib.sr_framebuffer_download_attachment(self.h, index, out.ctypes.data_as(_abi.f32p)))
        return out

    def download(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """AoS read-back: float32 [height*width, 5] = {r,g,b,a,depth}, index = x + y*width (an RGBAu8Color target: a
        structured array of U8_PIXEL = {rgba: 4 x u8, depth: f32})."""
        n = self.width * self.height
        if out is None:
            out = np.empty(n, self.U8_PIXEL) if self.u8_color else np.empty((n, 5), np.float32)
        check(lib.sr_framebuffer_download(self.h, out.ctypes.data_as(ctypes.c_void_p), out.nbytes))
        return out

    def download_rgba8(self, out: Optional[np.ndarray] = None, abgr: bool = False) -> np.ndarray:
        """Presentation read-back, uint8 [height, width, 4]: `(c * 255.0) as u8` per channel, converted on the device
        (realtime_example/src/main.rs:100-116).  abgr=True gives the byte order the example writes for SDL."""
        if out is None:
            out = np.empty((self.height, self.width, 4), np.uint8)
        check(lib.sr_framebuffer_download_rgba8(self.h, out.ctypes.data_as(_abi.u8p), out.nbytes, 1 if abgr else 0))
        return out

    def copy_to_image(self) -> np.ndarray:
        """RenderBuffer::copy_to_image of the reference's `image_compat` feature (examples/suzanne.rs:186-193):
        an RGBA8 image, row-major, top row first."""
        return self.download_rgba8()

    def save_png(self, path: str):
        """`image.save(path)` of examples/suzanne.rs:191 (8-bit RGBA PNG, zlib from the Python standard library)."""
        write_png(path, self.copy_to_image())

    def download_planes(self, stencil: bool = False):
        n = self.width * self.height
        color, depth = np.empty((n, 4), np.uint8 if self.u8_color else np.float32), np.empty(n, np.float32)
        st = np.empty(n, self._STENCIL_DTYPE[self.format]) if stencil else None
        check(lib.sr_framebuffer_download_planes(self.h, color.ctypes.data_as(ctypes.c_void_p), depth.ctypes.data_as(_abi.f32p),
                                                 st.ctypes.data_as(ctypes.c_void_p) if st is not None else None))
        return color, depth, st

    def upload_planes(self, color=None, depth=None, stencil=None):
        c = np.ascontiguousarray(color, np.uint8 if self.u8_color else np.float32) if color is not None else None
        d = np.ascontiguousarray(depth, np.float32) if depth is not None else None
        s = np.ascontiguousarray(stencil, self._STENCIL_DTYPE.get(self.format, np.uint8)) if stencil is not None else None
        check(lib.sr_framebuffer_upload_planes(self.h, c.ctypes.data_as(ctypes.c_void_p) if c is not None else None,
                                               d.ctypes.data_as(_abi.f32p) if d is not None else None,
                                               s.ctypes.data_as(ctypes.c_void_p) if s is not None else None))

    def pixel(self, x: int, y: int):
        """Checked accessor (PixelRead::pixel_ref): raises SoftrenderError(ERR_INVALID_PIXEL_COORDINATE) out of range."""
        rgba = (ctypes.c_float * 4)()
        d, s = ctypes.c_float(), ctypes.c_uint32()
        check(lib.sr_framebuffer_get_pixel(self.h, x, y, rgba, ctypes.byref(d), ctypes.byref(s)))
        return tuple(rgba), d.value, s.value

    def set_pixel(self, x: int, y: int, rgba=None, depth=None, stencil=None):
        """Checked write accessor (PixelWrite::pixel_mut, FramebufferAccessorMut::set_depth / set_stencil)."""
        c = (ctypes.c_float * 4)(*[float(v) for v in rgba]) if rgba is not None else None
        d = ctypes.byref(ctypes.c_float(depth)) if depth is not None else None
        s = ctypes.byref(ctypes.c_uint32(stencil)) if stencil is not None else None
        check(lib.sr_framebuffer_set_pixel(self.h, x, y, c, d, s))

    def enable_winner(self, enable: bool = True):
        check(lib.sr_framebuffer_enable_winner(self.h, 1 if enable else 0))

    def download_winner(self) -> np.ndarray:
        w = np.empty(self.width * self.height, np.uint32)
        check(lib.sr_framebuffer_download_winner(self.h, w.ctypes.data_as(_abi.u32p)))
        return w

    def device_ptr(self) -> int:
        return lib.sr_framebuffer_device_ptr(self.h) or 0

    def ipc_export(self) -> bytes:
        buf = ctypes.create_string_buffer(64)
        check(lib.sr_framebuffer_ipc_export(self.h, buf))
        return buf.raw

    def alias(self, ctx: Context) -> "RenderBuffer":
        """A second handle on the same pixels for another context of this process (sr_framebuffer_alias)."""
        h = ctypes.c_void_p()
        check(lib.sr_framebuffer_alias(ctx.h, self.h, ctypes.byref(h)))
        return RenderBuffer(ctx, h, self.width, self.height, self.format)

    @staticmethod
    def ipc_open(ctx: Context, handle: bytes, width: int, height: int) -> "RenderBuffer":
        h = ctypes.c_void_p()
        check(lib.sr_framebuffer_ipc_open(ctx.h, ctypes.create_string_buffer(handle, 64), width, height, FB_RGBAF32_DF32, ctypes.byref(h)))
        return RenderBuffer(ctx, h, width, height, FB_RGBAF32_DF32)


class ShardGroup:
    """This rank's end of a shard group (sr_shard): tile-sharded frames whose per-triangle front end is sharded by
    triangle range as well; the tile owner merges the ranks' keys over NVLink inside its tile kernel
    (include/softrender_b200.h; the reference's tile-parallel loop is src/pipeline/stages/fragment.rs:240-253).
    The context must already carry its tile shard (Context.set_tile_shard)."""

    def __init__(self, ctx: Context, width: int, height: int, lanes: int = 1):
        h = ctypes.c_void_p()
        check(lib.sr_shard_create(ctx.h, width, height, lanes, ctypes.byref(h)))
        self.ctx, self.h, self.lanes = ctx, h, lanes

    def export(self) -> bytes:
        buf = ctypes.create_string_buffer(64)
        check(lib.sr_shard_export(self.h, buf))
        return buf.raw

    def connect(self, handles):
        """handles: the exported handles of all ranks in rank order (one process per GPU)."""
        blob = ctypes.create_string_buffer(b"".join(bytes(x).ljust(64, b"\0") for x in handles), 64 * len(handles))
        check(lib.sr_shard_connect(self.h, blob, len(handles)))

    def connect_local(self, groups):
        """groups: the ShardGroup of every rank in rank order, all living in this process (tests)."""
        arr = (ctypes.c_void_p * len(groups))(*[g.h for g in groups])
        check(lib.sr_shard_connect_local(self.h, arr, len(groups)))

    def attach(self, ctx: Context, lane: int = 0):
        check(lib.sr_context_attach_shard(ctx.h, self.h, lane))

    @staticmethod
    def detach(ctx: Context):
        check(lib.sr_context_attach_shard(ctx.h, None, 0))

    def status(self) -> int:
        n = ctypes.c_uint32()
        check(lib.sr_shard_status(self.h, ctypes.byref(n)))
        return n.value

    def destroy(self):
        if self.h:
            lib.sr_shard_destroy(self.h)
            self.h = None


class Mesh:
    """Arc<Mesh<V>> resident in HBM (src/mesh.rs:12-20)."""

    def __init__(self, ctx: Context, data: MeshData = None, vertices=None, indices=None):
        if data is not None:
            vertices, indices = data.vertices, data.indices
        v = np.ascontiguousarray(vertices, np.float32)
        ix = np.ascontiguousarray(indices)
        if ix.dtype not in (np.uint32, np.uint64):
            ix = ix.astype(np.uint32)
        h = ctypes.c_void_p()
        check(lib.sr_mesh_create(ctx.h, v.ctypes.data_as(ctypes.c_void_p), v.shape[0], v.shape[1] if v.ndim == 2 else 0,
                                 ix.ctypes.data_as(ctypes.c_void_p), ix.size, ix.dtype.itemsize, ctypes.byref(h)))
        self.h, self.nverts, self.nindices = h, v.shape[0], ix.size

    def destroy(self):
        if self.h:
            lib.sr_mesh_destroy(self.h)
            self.h = None


class Texture:
    def __init__(self, ctx: Context, rgba: np.ndarray):
        t = np.ascontiguousarray(rgba, np.uint8)
        h = ctypes.c_void_p()
        check(lib.sr_texture_create(ctx.h, t.ctypes.data_as(_abi.u8p), t.shape[1], t.shape[0], ctypes.byref(h)))
        self.h = h

    def destroy(self):
        if self.h:
            lib.sr_texture_destroy(self.h)
            self.h = None


class _Stage:
    def __init__(self, pipeline: "Pipeline", handle):
        self.pipeline, self.h = pipeline, handle

    def _take(self):
        h, self.h = self.h, None
        if h is None:
            raise _abi.SoftrenderError(ERR_INVALID_STATE, "stage object already consumed (stage transitions take `self`)")
        return h

    def destroy(self):
        if self.h:
            lib.sr_draw_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    # parity-test introspection
    def download(self, which: int) -> np.ndarray:
        n, nk = ctypes.c_uint64(), ctypes.c_uint32()
        check(lib.sr_draw_count(self.h, which, ctypes.byref(n), ctypes.byref(nk)))
        out = np.zeros((n.value, 4 + nk.value), np.float32)
        if n.value:
            check(lib.sr_draw_download(self.h, which, out.ctypes.data_as(_abi.f32p), out.size))
        return out

    def download_sequence(self) -> np.ndarray:
        n, nk = ctypes.c_uint64(), ctypes.c_uint32()
        check(lib.sr_draw_count(self.h, 3, ctypes.byref(n), ctypes.byref(nk)))
        out = np.zeros(n.value // 3, np.uint32)
        if out.size:
            check(lib.sr_draw_download_sequence(self.h, out.ctypes.data_as(_abi.u32p), out.size))
        return out

    def set_generated(self, which: int, verts: np.ndarray):
        v = np.ascontiguousarray(verts, np.float32)
        check(lib.sr_draw_set_generated(self.h, which, v.ctypes.data_as(_abi.f32p), v.shape[0], v.shape[1] - 4))
        return self


class VertexShader(_Stage):
    def duplicate(self) -> "VertexShader":
        h = ctypes.c_void_p()
        check(lib.sr_draw_duplicate(self.h, ctypes.byref(h)))
        return VertexShader(self.pipeline, h)

    def run(self, vertex_shader: int) -> "GeometryShader":
        check(lib.sr_vertex_run(self.h, vertex_shader))
        return GeometryShader(self.pipeline, self._take())

    def run_to_fragment(self, viewport: Viewport, vertex_shader: int) -> "FragmentShader":
        check(lib.sr_vertex_run_to_fragment(self.h, ctypes.byref(viewport), vertex_shader))
        return FragmentShader(self.pipeline, self._take())


class GeometryShader(_Stage):
    def duplicate(self) -> "GeometryShader":
        h = ctypes.c_void_p()
        check(lib.sr_draw_duplicate(self.h, ctypes.byref(h)))
        return GeometryShader(self.pipeline, h)

    def run(self, geometry_shader: int) -> "GeometryShader":
        check(lib.sr_geometry_run(self.h, geometry_shader))
        return GeometryShader(self.pipeline, self._take())

    def clip_primitives(self, correct: bool = False) -> "GeometryShader":
        """GeometryShader::clip_primitives (geometry.rs:261-336), literally.  correct=True is the opt-in Sutherland-Hodgman
        clipper (SR_GS_CLIP_SH) the reference's author asks for (src/lib.rs, "Glaring Problems: Clipping")."""
        if correct:
            check(lib.sr_geometry_run(self.h, 3))
        else:
            check(lib.sr_geometry_clip_primitives(self.h))
        return GeometryShader(self.pipeline, self._take())

    def finish(self, viewport: Viewport) -> "FragmentShader":
        check(lib.sr_geometry_finish(self.h, ctypes.byref(viewport)))
        return FragmentShader(self.pipeline, self._take())


class FragmentShader(_Stage):
    def duplicate(self) -> "FragmentShader":
        h = ctypes.c_void_p()
        check(lib.sr_draw_duplicate(self.h, ctypes.byref(h)))
        return FragmentShader(self.pipeline, h)

    def cull_faces(self, winding: Optional[int]) -> "FragmentShader":
        check(lib.sr_fragment_set_cull_faces(self.h, CULL_NONE if winding is None else winding))
        return self

    def antialiased_lines(self, enable: bool) -> "FragmentShader":
        check(lib.sr_fragment_set_antialiased_lines(self.h, 1 if enable else 0))
        return self

    def tile_size(self, width: int, height: int) -> "FragmentShader":
        check(lib.sr_fragment_set_tile_size(self.h, width, height))
        return self

    def with_blend(self, blend: int) -> "FragmentShader":
        check(lib.sr_fragment_set_blend(self.h, blend))
        return self

    def bins(self):
        ntx = (self.pipeline.fb.width + tile_size()[0] - 1) // tile_size()[0]
        nty = (self.pipeline.fb.height + tile_size()[1] - 1) // tile_size()[1]
        offsets = np.zeros(ntx * nty + 1, np.uint64)
        total = ctypes.c_uint64()
        check(lib.sr_draw_bins(self.h, offsets.ctypes.data_as(_abi.u64p), None, 0, ctypes.byref(total)))
        ids = np.zeros(max(total.value, 1), np.uint32)
        check(lib.sr_draw_bins(self.h, offsets.ctypes.data_as(_abi.u64p), ids.ctypes.data_as(_abi.u32p), ids.size, ctypes.byref(total)))
        return offsets, ids[:total.value]

    def run(self, fragment_shader: int) -> None:
        """Draws (consumes the stage, like `FragmentShader::run(self, ..)`)."""
        check(lib.sr_fragment_run(self.h, fragment_shader))
        self.destroy()

    def run_keep(self, fragment_shader: int) -> "FragmentShader":
        """Draw without consuming (what `duplicate().run(..)` does in the reference)."""
        check(lib.sr_fragment_run(self.h, fragment_shader))
        return self


class Pipeline:
    """Pipeline<U, F, S> (src/pipeline/mod.rs:60-157)."""

    def __init__(self, ctx: Context, fb: RenderBuffer, uniforms: Uniforms):
        h = ctypes.c_void_p()
        check(lib.sr_pipeline_create(ctx.h, fb.h, ctypes.byref(uniforms), ctypes.byref(h)))
        self.ctx, self.fb, self.h = ctx, fb, h
        self._uniforms = uniforms.copy()

    @staticmethod
    def from_framebuffer(fb: RenderBuffer, uniforms: Uniforms) -> "Pipeline":
        return Pipeline(fb.ctx, fb, uniforms)

    def destroy(self):
        if self.h:
            lib.sr_pipeline_destroy(self.h)
            self.h = None

    def framebuffer(self) -> RenderBuffer:
        return self.fb

    def with_framebuffer(self, fb: RenderBuffer) -> "Pipeline":
        check(lib.sr_pipeline_set_framebuffer(self.h, fb.h))
        self.fb = fb
        return self

    def uniforms(self) -> Uniforms:
        return self._uniforms

    def set_uniforms(self, uniforms: Uniforms):
        """`*pipeline.uniforms_mut() = uniforms`."""
        check(lib.sr_pipeline_set_uniforms(self.h, ctypes.byref(uniforms)))
        self._uniforms = uniforms.copy()

    def set_stencil_config(self, test: int, op: int):
        check(lib.sr_pipeline_set_stencil_config(self.h, test, op))

    def bind_texture(self, tex: Optional[Texture]):
        check(lib.sr_pipeline_bind_texture(self.h, tex.h if tex else None))

    def bind_framebuffer_texture(self, src: Optional["RenderBuffer"]):
        """Render-to-texture: `src`'s colour attachment becomes the texture in place, no copy
        (TextureBufferRef, src/framebuffer/texturebuffer.rs:12-58)."""
        check(lib.sr_pipeline_bind_framebuffer_texture(self.h, src.h if src else None))

    def bind_framebuffer_attachment(self, src: "RenderBuffer", index: int):
        """Colour plane `index` of a texture buffer as the texture, in place (the named accessor, texturebuffer.rs:110-117)."""
        check(lib.sr_pipeline_bind_framebuffer_attachment(self.h, src.h, index))

    def set_sampler(self, filter: int, edge: int, border=None):
        """Filter / Edge of texture(t, coord, filter, edge) (src/texture.rs:14-45); `border` = Edge::Border's colour."""
        b = (ctypes.c_float * 4)(*[float(x) for x in border]) if border is not None else None
        check(lib.sr_pipeline_set_sampler(self.h, filter, edge, b))

    def render_mesh(self, primitive: int, mesh: Mesh, stencil: Optional[int] = None) -> VertexShader:
        h = ctypes.c_void_p()
        check(lib.sr_render_mesh(self.h, mesh.h, primitive, 0 if stencil is None else 1, stencil or 0, ctypes.byref(h)))
        return VertexShader(self, h)

    # parity-test injection: the state a vertex/geometry stage would have produced
    def draw_from_vertices(self, primitive: int, verts: np.ndarray, indices: np.ndarray, space: int,
                           stencil: Optional[int] = None):
        v = np.ascontiguousarray(verts, np.float32)
        ix = np.ascontiguousarray(indices, np.uint32)
        h = ctypes.c_void_p()
        check(lib.sr_draw_from_vertices(self.h, primitive, v.ctypes.data_as(_abi.f32p), v.shape[0], v.shape[1] - 4, space,
                                        ix.ctypes.data_as(_abi.u32p), ix.size, 0 if stencil is None else 1, stencil or 0,
                                        ctypes.byref(h)))
        return (FragmentShader if space else GeometryShader)(self, h)
