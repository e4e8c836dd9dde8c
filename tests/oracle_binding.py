"""ctypes binding of the TEST ORACLE (oracle/libsr_oracle.so).  Test infrastructure only."""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libsr_oracle.so")

import softrender_b200 as sr  # noqa: E402  (shim at the repo root; scene types only)
from softrender_b200.scenes import Uniforms, Viewport  # noqa: E402

u32p = ctypes.POINTER(ctypes.c_uint32)
u64p = ctypes.POINTER(ctypes.c_uint64)
f32p = ctypes.POINTER(ctypes.c_float)
u8p = ctypes.POINTER(ctypes.c_uint8)


class SoFramebuffer(ctypes.Structure):
    _fields_ = [("width", ctypes.c_uint32), ("height", ctypes.c_uint32), ("color", f32p), ("depth", f32p),
                ("stencil", u8p), ("winner", u32p), ("stencil_bytes", ctypes.c_uint32), ("color_u8", u8p),
                ("color1", f32p)]


class SoTexture(ctypes.Structure):
    _fields_ = [("width", ctypes.c_uint32), ("height", ctypes.c_uint32), ("rgba", u8p), ("texels_f32", f32p),
                ("stride", ctypes.c_uint32), ("filter", ctypes.c_uint32), ("edge", ctypes.c_uint32), ("border", ctypes.c_float * 4)]


def make_texture(texture, sampler=None):
    """so_texture of a (h, w, 4) array: uint8 = image texels, float32 = a framebuffer's colour sampled in place
    (texturebuffer.rs:12-58).  sampler = (filter, edge, border) or None for the reference's defaults (Nearest, Clamp: src/texture.rs:27-45).
    Returns (struct, array kept alive)."""
    filt, edge, border = sampler if sampler is not None else (0, 0, None)  # Filter::default() = Nearest, Edge::default() = Clamp
    b = (ctypes.c_float * 4)(*([float(x) for x in border] if border is not None else [0.0] * 4))
    if np.asarray(texture).dtype == np.float32:
        t = np.ascontiguousarray(texture, np.float32)
        return SoTexture(t.shape[1], t.shape[0], None, t.ctypes.data_as(f32p), 4, filt, edge, b), t
    t = np.ascontiguousarray(texture, np.uint8)
    return SoTexture(t.shape[1], t.shape[0], t.ctypes.data_as(u8p), None, 0, filt, edge, b), t


def texture_sample(texture, u, v, sampler=None) -> np.ndarray:
    """texture(t, (u, v), filter, edge) of the oracle for one coordinate."""
    tex, _keep = make_texture(texture, sampler)
    out = (ctypes.c_float * 4)()
    lib().so_texture_sample(ctypes.byref(tex), ctypes.c_float(u), ctypes.c_float(v), out)
    return np.array(out[:], np.float32)


class SoRasterState(ctypes.Structure):
    _fields_ = [("cull_faces", ctypes.c_uint32), ("blend", ctypes.c_uint32), ("antialiased_lines", ctypes.c_uint32),
                ("tile_width", ctypes.c_uint32), ("tile_height", ctypes.c_uint32),
                ("stencil_test", ctypes.c_uint32), ("stencil_op", ctypes.c_uint32)]


def build_oracle() -> str:
    # bench.py's CPU arm points SR_ORACLE_SO at the -march=native build it made on the box it times (oracle/Makefile)
    override = os.environ.get("SR_ORACLE_SO")
    if override and os.path.exists(override):
        return override
    src = os.path.join(ORACLE_DIR, "sr_oracle.cpp")
    if (not os.path.exists(ORACLE_SO)) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)
    return ORACLE_SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build_oracle())
        L.so_draw_create.restype = ctypes.c_void_p
        L.so_draw_create.argtypes = [ctypes.c_int, u32p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32]
        L.so_draw_destroy.argtypes = [ctypes.c_void_p]
        L.so_draw_vertex_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(Uniforms), f32p, ctypes.c_uint64,
                                         ctypes.c_uint32, ctypes.c_int]
        L.so_draw_vertex_run_to_fragment.argtypes = [ctypes.c_void_p, ctypes.POINTER(Viewport), ctypes.c_int,
                                                     ctypes.POINTER(Uniforms), f32p, ctypes.c_uint64, ctypes.c_uint32,
                                                     ctypes.c_int]
        L.so_draw_set_vertices.argtypes = [ctypes.c_void_p, f32p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int]
        L.so_draw_set_generated.argtypes = [ctypes.c_void_p, ctypes.c_int, f32p, ctypes.c_uint64, ctypes.c_uint32]
        L.so_draw_geometry_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(Uniforms), ctypes.c_int]
        L.so_draw_finish.argtypes = [ctypes.c_void_p, ctypes.POINTER(Viewport), ctypes.c_int]
        L.so_draw_fragment_run.argtypes = [ctypes.c_void_p, ctypes.POINTER(SoFramebuffer), ctypes.POINTER(SoRasterState),
                                           ctypes.c_int, ctypes.POINTER(Uniforms), ctypes.POINTER(SoTexture), ctypes.c_int]
        L.so_draw_fragment_run_tiles.argtypes = L.so_draw_fragment_run.argtypes + [ctypes.c_uint64, ctypes.c_uint64]
        L.so_texture_sample.restype = None
        L.so_texture_sample.argtypes = [ctypes.POINTER(SoTexture), ctypes.c_float, ctypes.c_float, ctypes.POINTER(ctypes.c_float)]
        L.so_draw_count.restype = ctypes.c_uint64
        L.so_draw_count.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.so_draw_data.restype = f32p
        L.so_draw_data.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.so_draw_nk.restype = ctypes.c_uint32
        L.so_draw_nk.argtypes = [ctypes.c_void_p]
        L.so_tiles.restype = ctypes.c_uint64
        L.so_tiles.argtypes = [ctypes.c_uint32] * 4 + [u32p, ctypes.c_uint64]
        L.so_draw_bins.restype = ctypes.c_uint64
        L.so_draw_bins.argtypes = [ctypes.c_void_p] + [ctypes.c_uint32] * 5 + [u64p, u32p]
        L.so_coordinate_index.restype = ctypes.c_uint64
        L.so_coordinate_index.argtypes = [ctypes.c_uint32] * 3
        L.so_stencil_test.restype = ctypes.c_int
        L.so_stencil_test.argtypes = [ctypes.c_uint32, ctypes.c_uint8, ctypes.c_uint8]
        L.so_stencil_op.restype = ctypes.c_uint8
        L.so_stencil_op.argtypes = [ctypes.c_uint32, ctypes.c_uint8, ctypes.c_uint8]
        L.so_stencil_op_wide.restype = ctypes.c_uint32
        L.so_stencil_op_wide.argtypes = [ctypes.c_uint32] * 4
        L.so_depth_far.restype = ctypes.c_float
        L.so_framebuffer_clear.argtypes = [ctypes.POINTER(SoFramebuffer), f32p]
        L.so_framebuffer_clear2.argtypes = [ctypes.POINTER(SoFramebuffer), f32p, f32p]
        _lib = L
    return _lib


def _fp(a: np.ndarray):
    return a.ctypes.data_as(f32p)


@dataclass
class OracleFramebuffer:
    """RenderBuffer<ColorDepth[Stencil]Attachments<RGBAf32Color, f32[, u8]>> held in numpy planes."""
    width: int
    height: int
    stencil_bits: int = 0
    u8_color: bool = False  # colour attachment RGBAu8Color (src/color/predefined.rs:26) instead of RGBAf32Color
    two_colors: bool = False  # a texture buffer declared with two colour planes (texturebuffer.rs:72-147): `color` and `color1`
    color: np.ndarray = field(init=False)
    depth: np.ndarray = field(init=False)
    stencil: np.ndarray | None = field(init=False)
    winner: np.ndarray = field(init=False)

    def __post_init__(self):
        n = self.width * self.height
        self.color = np.zeros((n, 4), np.uint8 if self.u8_color else np.float32)
        self.color1 = np.zeros((n, 4), np.float32) if self.two_colors else None
        self.depth = np.full(n, lib().so_depth_far(), np.float32)
        self.stencil = np.zeros(n, {16: np.uint16, 32: np.uint32}.get(self.stencil_bits, np.uint8)) if self.stencil_bits else None
        self.winner = np.zeros(n, np.uint32)

    def struct(self) -> SoFramebuffer:
        return SoFramebuffer(self.width, self.height, None if self.u8_color else _fp(self.color), _fp(self.depth),
                             self.stencil.ctypes.data_as(u8p) if self.stencil is not None else None,
                             self.winner.ctypes.data_as(u32p), self.stencil.itemsize if self.stencil is not None else 0,
                             self.color.ctypes.data_as(u8p) if self.u8_color else None,
                             _fp(self.color1) if self.color1 is not None else None)

    def clear(self, color, color1=None):
        c = np.asarray(color, np.float32)
        s = self.struct()
        if color1 is not None:
            lib().so_framebuffer_clear2(ctypes.byref(s), _fp(c), _fp(np.asarray(color1, np.float32)))
        else:
            lib().so_framebuffer_clear(ctypes.byref(s), _fp(c))


class OracleDraw:
    """The reference's VertexShader -> GeometryShader -> FragmentShader chain for one render_mesh call."""

    def __init__(self, primitive: int, indices: np.ndarray, stencil_value=None):
        self.indices = np.ascontiguousarray(indices, np.uint32)
        self.h = lib().so_draw_create(primitive, self.indices.ctypes.data_as(u32p), len(self.indices),
                                      0 if stencil_value is None else 1, 0 if stencil_value is None else stencil_value)
        if not self.h:
            raise ValueError("render_mesh: indices.len() % num_vertices != 0")
        self.cull = sr.CULL_NONE
        self.blend = sr.BLEND_REPLACE
        self.aa = False
        self.tile = None  # None -> one frame-sized tile (canonical); (128,128) is the reference default

    def __del__(self):
        if getattr(self, "h", None):
            lib().so_draw_destroy(self.h)
            self.h = None

    @staticmethod
    def _ck(rc):
        if rc != 0:
            raise RuntimeError(f"oracle status {rc}")

    def vertex_run(self, vs, uniforms, vertices: np.ndarray, nthreads=1):
        v = np.ascontiguousarray(vertices, np.float32)
        self._ck(lib().so_draw_vertex_run(self.h, vs, ctypes.byref(uniforms), _fp(v), v.shape[0], v.shape[1], nthreads))
        return self

    def vertex_run_to_fragment(self, viewport, vs, uniforms, vertices: np.ndarray, nthreads=1):
        v = np.ascontiguousarray(vertices, np.float32)
        self._ck(lib().so_draw_vertex_run_to_fragment(self.h, ctypes.byref(viewport), vs, ctypes.byref(uniforms), _fp(v),
                                                      v.shape[0], v.shape[1], nthreads))
        return self

    def set_vertices(self, verts: np.ndarray, space: int):
        v = np.ascontiguousarray(verts, np.float32)
        self._ck(lib().so_draw_set_vertices(self.h, _fp(v), v.shape[0], v.shape[1] - 4, space))
        return self

    def set_generated(self, which: int, verts: np.ndarray):
        v = np.ascontiguousarray(verts, np.float32)
        self._ck(lib().so_draw_set_generated(self.h, which, _fp(v), v.shape[0], v.shape[1] - 4))
        return self

    def geometry_run(self, gs, uniforms=None, nthreads=1):
        self._ck(lib().so_draw_geometry_run(self.h, gs, ctypes.byref(uniforms) if uniforms is not None else None, nthreads))
        return self

    def clip_primitives(self, nthreads=1, correct=False):
        return self.geometry_run(sr.GS_CLIP_SH if correct else sr.GS_CLIP, None, nthreads)

    def finish(self, viewport, nthreads=1):
        self._ck(lib().so_draw_finish(self.h, ctypes.byref(viewport), nthreads))
        return self

    def fragment_run(self, fb: OracleFramebuffer, fs, uniforms, stencil_test=0, stencil_op=0, texture=None, nthreads=1, sampler=None,
                     tile_slice=(0, 1), keep_winner=False):
        """tile_slice = (first, stride): only the tiles first, first+stride, ... of the reference's tile list (each still
        visits every primitive); the default is the whole frame."""
        tw, th = self.tile if self.tile else (max(fb.width, 1), max(fb.height, 1))
        st = SoRasterState(self.cull, self.blend, 1 if self.aa else 0, tw, th, stencil_test, stencil_op)
        if not keep_winner:
            fb.winner[:] = 0  # winner plane = primitives of THIS draw
        s = fb.struct()
        tex = None
        if texture is not None:
            tex, _keep = make_texture(texture, sampler)
        self._ck(lib().so_draw_fragment_run_tiles(self.h, ctypes.byref(s), ctypes.byref(st), fs, ctypes.byref(uniforms),
                                                  ctypes.byref(tex) if tex is not None else None, nthreads,
                                                  tile_slice[0], tile_slice[1]))
        return self

    def data(self, which: int) -> np.ndarray:
        n = lib().so_draw_count(self.h, which)
        S = 4 + lib().so_draw_nk(self.h)
        if n == 0:
            return np.zeros((0, S), np.float32)
        return np.ctypeslib.as_array(lib().so_draw_data(self.h, which), shape=(n, S)).copy()

    def bins(self, width, height, tw, th, cull=0):
        ntiles = ((width + tw - 1) // tw) * ((height + th - 1) // th)
        offsets = np.zeros(ntiles + 1, np.uint64)
        total = lib().so_draw_bins(self.h, width, height, tw, th, cull, offsets.ctypes.data_as(u64p), None)
        ids = np.zeros(max(total, 1), np.uint32)
        lib().so_draw_bins(self.h, width, height, tw, th, cull, offsets.ctypes.data_as(u64p), ids.ctypes.data_as(u32p))
        return offsets, ids[:total]


def tiles(width, height, tw=128, th=128) -> np.ndarray:
    n = lib().so_tiles(width, height, tw, th, None, 0)
    out = np.zeros((max(n, 1), 4), np.uint32)
    lib().so_tiles(width, height, tw, th, out.ctypes.data_as(u32p), n)
    return out[:n]
