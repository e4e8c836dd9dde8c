"""Scene construction for the benchmark configs and the parity tests (host side, numpy only).

Everything here produces *inputs* (meshes, uniform blobs, viewports) that are fed unchanged to both the
CUDA library and the test oracle.  Matrix helpers restate the nalgebra 0.12 constructors the reference
scenes call (examples/suzanne.rs:78-104, full_example/src/lib.rs:22-51), step by step in float32.
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass

import numpy as np

f32 = np.float32
SR_MAX_LIGHTS = 8


# --------------------------------------------------------------------------------------
# ctypes mirrors of include/softrender_b200_types.h
# --------------------------------------------------------------------------------------
class Light(ctypes.Structure):  # sr_light (full_example/src/light.rs:6-10)
    _fields_ = [("color", ctypes.c_float * 4), ("position", ctypes.c_float * 3), ("intensity", ctypes.c_float)]


class Uniforms(ctypes.Structure):  # sr_uniforms
    _fields_ = [
        ("camera", ctypes.c_float * 4),
        ("model", ctypes.c_float * 16),
        ("mit", ctypes.c_float * 16),
        ("view", ctypes.c_float * 16),
        ("projection", ctypes.c_float * 16),
        ("sz_light", ctypes.c_float * 4),
        ("sz_color", ctypes.c_float * 4),
        ("sz_intensity", ctypes.c_float),
        ("nlights", ctypes.c_uint32),
        ("reserved0", ctypes.c_uint32),
        ("reserved1", ctypes.c_uint32),
        ("lights", Light * SR_MAX_LIGHTS),
    ]

    def copy(self) -> "Uniforms":
        out = Uniforms()
        ctypes.memmove(ctypes.byref(out), ctypes.byref(self), ctypes.sizeof(Uniforms))
        return out


class Viewport(ctypes.Structure):  # sr_viewport (src/geometry/clipvertex.rs:40-48)
    _fields_ = [("x", ctypes.c_float), ("y", ctypes.c_float), ("width", ctypes.c_float),
                ("height", ctypes.c_float), ("near", ctypes.c_float), ("far", ctypes.c_float)]

    @staticmethod
    def new(width: int, height: int, near: float, far: float, x: int = 0, y: int = 0) -> "Viewport":
        """Viewport::new(dimensions, offset, near, far) (clipvertex.rs:50-59)."""
        return Viewport(float(x), float(y), float(width), float(height), float(near), float(far))


# --------------------------------------------------------------------------------------
# nalgebra 0.12 restatements (float32, same operation order)
# --------------------------------------------------------------------------------------
def _normalize3(v):
    v = np.asarray(v, dtype=f32)
    n = f32(np.sqrt(f32(f32(f32(v[0] * v[0]) + f32(v[1] * v[1])) + f32(v[2] * v[2]))))
    return (v / n).astype(f32)


def _cross(a, b):
    a = np.asarray(a, f32)
    b = np.asarray(b, f32)
    return np.array([f32(a[1] * b[2]) - f32(a[2] * b[1]),
                     f32(a[2] * b[0]) - f32(a[0] * b[2]),
                     f32(a[0] * b[1]) - f32(a[1] * b[0])], dtype=f32)


def mat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """nalgebra Matrix*Matrix: res(i,j) = sum_k a(i,k) b(k,j), zero-initialised, k ascending (row-major np arrays)."""
    out = np.zeros((4, 4), dtype=f32)
    for i in range(4):
        for j in range(4):
            acc = f32(0.0)
            for k in range(4):
                acc = f32(acc + f32(a[i, k] * b[k, j]))
            out[i, j] = acc
    return out


def rotation_from_scaled_axis(axisangle) -> np.ndarray:
    """UnitQuaternion::from_scaled_axis(..).to_rotation_matrix() (3x3)."""
    ax = np.asarray(axisangle, dtype=f32)
    angle = f32(np.sqrt(f32(f32(f32(ax[0] * ax[0]) + f32(ax[1] * ax[1])) + f32(ax[2] * ax[2]))))
    if angle == 0:
        return np.eye(3, dtype=f32)
    axis = (ax / angle).astype(f32)
    half = f32(angle / f32(2.0))
    s, w = f32(np.sin(half)), f32(np.cos(half))
    i, j, k = (axis * s).astype(f32)
    ww, ii, jj, kk = f32(w * w), f32(i * i), f32(j * j), f32(k * k)
    ij, wk, wj = f32(f32(i * j) * f32(2)), f32(f32(w * k) * f32(2)), f32(f32(w * j) * f32(2))
    ik, jk, wi = f32(f32(i * k) * f32(2)), f32(f32(j * k) * f32(2)), f32(f32(w * i) * f32(2))
    return np.array([[ww + ii - jj - kk, ij - wk, wj + ik],
                     [wk + ij, ww - ii + jj - kk, jk - wi],
                     [ik - wj, wi + jk, ww - ii - jj + kk]], dtype=f32)


def isometry(translation, axisangle):
    """Isometry3::new(translation, axisangle) -> (R 3x3, t 3)."""
    return rotation_from_scaled_axis(axisangle), np.asarray(translation, dtype=f32)


def isometry_to_homogeneous(R, t) -> np.ndarray:
    m = np.eye(4, dtype=f32)
    m[:3, :3] = R
    m[:3, 3] = t
    return m


def isometry_inverse(R, t):
    Ri = R.T.copy()
    ti = -(Ri @ t).astype(f32)
    return Ri, ti.astype(f32)


def look_at_rh(eye, target, up) -> np.ndarray:
    """Isometry3::look_at_rh(eye, target, up).to_homogeneous()."""
    eye = np.asarray(eye, f32)
    target = np.asarray(target, f32)
    z = _normalize3(eye - target)
    x = _normalize3(_cross(up, z))
    y = _normalize3(_cross(z, x))
    R = np.stack([x, y, z]).astype(f32)  # rows
    t = np.array([-(f32(f32(f32(R[r, 0] * eye[0]) + f32(R[r, 1] * eye[1])) + f32(R[r, 2] * eye[2]))) for r in range(3)],
                 dtype=f32)
    return isometry_to_homogeneous(R, t)


def perspective(aspect, fovy, znear, zfar) -> np.ndarray:
    """Perspective3::new(aspect, fovy, znear, zfar).to_homogeneous()."""
    aspect, fovy, znear, zfar = f32(aspect), f32(fovy), f32(znear), f32(zfar)
    m = np.zeros((4, 4), dtype=f32)
    m22 = f32(f32(1.0) / f32(np.tan(f32(fovy / f32(2.0)))))
    m[1, 1] = m22
    m[0, 0] = f32(m22 / aspect)
    m[2, 2] = f32(f32(zfar + znear) / f32(znear - zfar))
    m[2, 3] = f32(f32(f32(zfar * znear) * f32(2.0)) / f32(znear - zfar))
    m[3, 2] = f32(-1.0)
    return m


def _col_major(m: np.ndarray):
    return (ctypes.c_float * 16)(*[float(x) for x in np.asarray(m, f32).T.reshape(-1)])


def _vec(n, vals):
    return (ctypes.c_float * n)(*[float(f32(v)) for v in vals])


def model_matrix(rotation_y: float):
    """full_example/src/lib.rs:22-28: Isometry3::new(0, (0, rotation, 0)) -> (model, model inverse transpose)."""
    R, t = isometry((0, 0, 0), (0, rotation_y, 0))
    model = isometry_to_homogeneous(R, t)
    Ri, ti = isometry_inverse(R, t)
    mit = isometry_to_homogeneous(Ri, ti).T.copy()
    return model, mit


def suzanne_uniforms(width: int, height: int, rotation_y: float = 0.0) -> Uniforms:
    """GlobalUniforms + captured constants of examples/suzanne.rs:73-114."""
    u = Uniforms()
    model, mit = model_matrix(rotation_y)
    eye = (1.0, 0.0, 2.0)
    view = look_at_rh(eye, (0, 0, 0), (0, 1, 0))
    proj = perspective(f32(width) / f32(height), f32(np.deg2rad(f32(75.0))), 0.001, 1000.0)
    u.camera = _vec(4, (*eye, 1.0))
    u.model, u.mit, u.view, u.projection = _col_major(model), _col_major(mit), _col_major(view), _col_major(proj)
    u.sz_light = _vec(4, (5.0, 5.0, 5.0, 1.0))
    g = f32(2.2)
    u.sz_color = _vec(4, (np.power(f32(0.1), g), np.power(f32(0.5), g), np.power(f32(0.1), g), 1.0))
    u.sz_intensity = 4.0
    u.nlights = 0
    return u


def full_example_uniforms(aspect: float, camera_rotation: float, camera_distance: float, object_rotation: float,
                          fov: float, model_offset_x: float = 0.0) -> Uniforms:
    """generate_global_uniforms (full_example/src/lib.rs:30-71); `model_offset_x` translates the instance
    (config 2 composes three instances)."""
    u = Uniforms()
    R, _ = isometry((0, 0, 0), (0, object_rotation, 0))
    t = np.array([model_offset_x, 0, 0], dtype=f32)
    model = isometry_to_homogeneous(R, t)
    Ri, ti = isometry_inverse(R, t)
    mit = isometry_to_homogeneous(Ri, ti).T.copy()
    cr, cd = f32(camera_rotation), f32(camera_distance)
    eye = (f32(np.cos(cr)) * cd, cd, f32(np.sin(cr)) * cd)
    view = look_at_rh(eye, (0.0, 0.25, 0.0), (0, 1, 0))
    proj = perspective(aspect, fov, 0.1, 1000.0)
    u.camera = _vec(4, (*eye, 1.0))
    u.model, u.mit, u.view, u.projection = _col_major(model), _col_major(mit), _col_major(view), _col_major(proj)
    lights = [((1, 1, 1, 1), (-1, 1, -1), 9.0), ((0.6, 0.6, 1.0, 1.0), (1, 1, 1), 9.0),
              ((1.0, 0.3, 0.3, 1.0), (0, 3, -1), 25.0), ((0.7, 1.0, 0.7, 1.0), (-2, -1, 1), 25.0)]
    u.nlights = len(lights)
    for i, (c, p, inten) in enumerate(lights):
        u.lights[i].color = _vec(4, c)
        u.lights[i].position = _vec(3, p)
        u.lights[i].intensity = inten
    # the suzanne captures stay usable with this blob too
    u.sz_light = _vec(4, (5.0, 5.0, 5.0, 1.0))
    g = f32(2.2)
    u.sz_color = _vec(4, (np.power(f32(0.1), g), np.power(f32(0.5), g), np.power(f32(0.1), g), 1.0))
    u.sz_intensity = 4.0
    return u


def grid_uniforms(width: int, height: int) -> Uniforms:
    """Camera of configs 3/4: eye (0,0,2.2) -> origin, fovy 60 deg, near 0.1, far 100."""
    u = Uniforms()
    model, mit = model_matrix(0.0)
    eye = (0.0, 0.0, 2.2)
    view = look_at_rh(eye, (0, 0, 0), (0, 1, 0))
    proj = perspective(f32(width) / f32(height), f32(np.deg2rad(f32(60.0))), 0.1, 100.0)
    u.camera = _vec(4, (*eye, 1.0))
    u.model, u.mit, u.view, u.projection = _col_major(model), _col_major(mit), _col_major(view), _col_major(proj)
    u.sz_light = _vec(4, (5.0, 5.0, 5.0, 1.0))
    g = f32(2.2)
    u.sz_color = _vec(4, (np.power(f32(0.1), g), np.power(f32(0.5), g), np.power(f32(0.1), g), 1.0))
    u.sz_intensity = 4.0
    u.nlights = 0
    return u


# --------------------------------------------------------------------------------------
# meshes
# --------------------------------------------------------------------------------------
@dataclass
class MeshData:
    """Mesh<V> of src/mesh.rs:12-20: `vertices` is an AoS float32 array [nverts, vin_floats] with position.xyz
    first (SimpleVertex{position, data}), `indices` uint32 (the reference uses usize)."""
    vertices: np.ndarray
    indices: np.ndarray

    @property
    def ntris(self) -> int:
        return len(self.indices) // 3


def load_obj(path: str, with_uv: bool = False) -> MeshData:
    """OBJ loader with tobj 0.1.3 indexing: one vertex per distinct v/vt/vn triple in first-use order,
    polygons fan-triangulated (a, b, c), (a, c, d), ... (examples/suzanne.rs:24-47, full_example/src/mesh.rs:22-52)."""
    pos, nrm, tex = [], [], []
    verts, index_of, indices = [], {}, []
    with open(path) as fh:
        for line in fh:
            p = line.split()
            if not p:
                continue
            if p[0] == "v":
                pos.append([float(x) for x in p[1:4]])
            elif p[0] == "vn":
                nrm.append([float(x) for x in p[1:4]])
            elif p[0] == "vt":
                tex.append([float(x) for x in p[1:3]])
            elif p[0] == "f":
                face = []
                for tok in p[1:]:
                    parts = (tok.split("/") + ["", ""])[:3]
                    key = tuple(int(s) if s else 0 for s in parts)
                    if key not in index_of:
                        index_of[key] = len(verts)
                        v = list(pos[key[0] - 1]) + list(nrm[key[2] - 1] if key[2] else (0, 0, 0))
                        if with_uv:
                            v += list(tex[key[1] - 1] if key[1] else (0, 0))
                        verts.append(v)
                    face.append(index_of[key])
                for k in range(1, len(face) - 1):
                    indices += [face[0], face[k], face[k + 1]]
    return MeshData(np.asarray(verts, dtype=f32), np.asarray(indices, dtype=np.uint32))


def splitmix64_u01(seed: int, n: int) -> np.ndarray:
    """SplitMix64 stream -> u01 = (next() >> 40) * 2^-24 (common state)."""
    with np.errstate(over="ignore"):
        gamma = np.uint64(0x9E3779B97F4A7C15)
        z = np.uint64(seed) + gamma * np.arange(1, n + 1, dtype=np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(40)).astype(np.float64) * (2.0 ** -24)).astype(f32)


def make_grid(nx: int, ny: int, layers: int = 4, seed: int = 0x5EED0003, reverse: bool = False) -> MeshData:
    """Synthetic displaced-grid mesh of configs 3/4: `layers` regular grids of nx*ny
    cells, 2 triangles per cell, layers submitted front-to-back (or back-to-front when `reverse`)."""
    vx = 1.5 * (2.0 * np.arange(nx + 1, dtype=np.float64) / nx - 1.0)
    vy = 0.84 * (2.0 * np.arange(ny + 1, dtype=np.float64) / ny - 1.0)
    X, Y = np.meshgrid(vx, vy)  # [ny+1, nx+1]
    nvl = (nx + 1) * (ny + 1)
    rnd = splitmix64_u01(seed, layers * nvl).astype(np.float64).reshape(layers, ny + 1, nx + 1)
    verts = np.empty((layers, ny + 1, nx + 1, 6), dtype=f32)
    for l in range(layers):
        a, b = 7.0 * X + 0.9 * l, 5.0 * Y - 0.4 * l
        Z = -0.25 * l + 0.05 * np.sin(a) * np.cos(b) + 0.002 * (rnd[l] - 0.5)
        dzdx = 0.05 * 7.0 * np.cos(a) * np.cos(b)
        dzdy = -0.05 * 5.0 * np.sin(a) * np.sin(b)
        inv = 1.0 / np.sqrt(dzdx * dzdx + dzdy * dzdy + 1.0)
        verts[l, ..., 0], verts[l, ..., 1], verts[l, ..., 2] = X, Y, Z
        verts[l, ..., 3], verts[l, ..., 4], verts[l, ..., 5] = -dzdx * inv, -dzdy * inv, inv
    j, i = np.meshgrid(np.arange(ny, dtype=np.int64), np.arange(nx, dtype=np.int64), indexing="ij")
    v00 = j * (nx + 1) + i
    v10, v01, v11 = v00 + 1, v00 + (nx + 1), v00 + (nx + 1) + 1
    cell = np.stack([v00, v10, v11, v00, v11, v01], axis=-1).reshape(-1)  # (v00,v10,v11), (v00,v11,v01)
    order = range(layers - 1, -1, -1) if reverse else range(layers)
    idx = np.concatenate([cell + l * nvl for l in order]).astype(np.uint32)
    return MeshData(verts.reshape(-1, 6), idx)


def subdivide(mesh: MeshData, times: int = 1) -> MeshData:
    """Midpoint subdivision (1 triangle -> 4), per-face normals recomputed; the stand-in for the missing
    suzanne_highres.obj of config 2.  Vertices are de-indexed per face."""
    v, idx = mesh.vertices.astype(np.float64), mesh.indices.astype(np.int64)
    tris = v[idx].reshape(-1, 3, v.shape[1])
    for _ in range(times):
        a, b, c = tris[:, 0], tris[:, 1], tris[:, 2]
        ab, bc, ca = (a + b) / 2, (b + c) / 2, (c + a) / 2
        tris = np.concatenate([np.stack([a, ab, ca], 1), np.stack([ab, b, bc], 1),
                               np.stack([ca, bc, c], 1), np.stack([ab, bc, ca], 1)], 0)
    n = np.cross(tris[:, 1, :3] - tris[:, 0, :3], tris[:, 2, :3] - tris[:, 0, :3])
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    n = np.where(ln > 0, n / np.maximum(ln, 1e-30), np.array([0.0, 0.0, 1.0]))
    tris[:, :, 3:6] = n[:, None, :]
    verts = tris.reshape(-1, v.shape[1]).astype(f32)
    return MeshData(verts, np.arange(len(verts), dtype=np.uint32))


def checker_texture(size: int = 512, cells: int = 8) -> np.ndarray:
    """Procedural RGBA8 checker of config 2: colours (230,230,230)/(40,40,40)."""
    yy, xx = np.mgrid[0:size, 0:size]
    on = ((xx // (size // cells)) + (yy // (size // cells))) % 2 == 0
    img = np.empty((size, size, 4), dtype=np.uint8)
    img[..., :3] = np.where(on[..., None], 230, 40)
    img[..., 3] = 255
    return img


def deg(x: float) -> float:
    return float(f32(math.radians(x))) if False else float(np.deg2rad(f32(x)))
