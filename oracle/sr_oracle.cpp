// sr_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see sr_oracle.h).
//
// Literal f32 restatement of the draw hot path of novacrazy/rust-softrender.
// Every function cites the reference file:line it follows ("ref:" comments,
// paths relative to the reference repository root).  Build with
//   g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math
// so that no multiply-add is ever contracted: Rust never emits FMA for `a*b+c`.
//
// Third-party arithmetic that is not in the reference tree (nalgebra 0.12,
// num-traits 0.1, compiler-rt powi, libm) is restated from its published
// algorithm where noted.

#include "sr_oracle.h"

#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------

template <class F>
void run_threads(int nthreads, F &&fn) {
    // scoped_threadpool::Pool::scoped: every worker runs the same closure and
    // all are joined before the stage returns (ref: src/pipeline/mod.rs:115).
    if (nthreads <= 1) { fn(0); return; }
    std::vector<std::thread> pool;
    pool.reserve(nthreads);
    for (int t = 0; t < nthreads; ++t) pool.emplace_back([&fn, t] { fn(t); });
    for (auto &th : pool) th.join();
}

// nalgebra 0.12 Matrix * Matrix / Matrix * Vector: res(i,j) = sum_k a(i,k)*b(k,j),
// accumulator starts at zero, k ascending.  Column-major storage m[c*4+r].
inline void mat_vec(const float *m, const float *v, float *out) {
    for (int r = 0; r < 4; ++r) {
        float acc = 0.0f;
        for (int k = 0; k < 4; ++k) acc += m[k * 4 + r] * v[k];
        out[r] = acc;
    }
}
inline void mat_mat(const float *a, const float *b, float *out) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float acc = 0.0f;
            for (int k = 0; k < 4; ++k) acc += a[k * 4 + i] * b[j * 4 + k];
            out[j * 4 + i] = acc;
        }
}
// nalgebra dot / norm / normalize over all four components (index order, zero-initialised).
inline float dot4(const float *a, const float *b) {
    float acc = 0.0f;
    for (int i = 0; i < 4; ++i) acc += a[i] * b[i];
    return acc;
}
inline float norm4(const float *a) { return sqrtf(dot4(a, a)); }
inline void normalize4(const float *a, float *out) {
    float n = norm4(a);
    for (int i = 0; i < 4; ++i) out[i] = a[i] / n;
}
// compiler-rt __powisf2 (what f32::powi lowers to): square-and-multiply, LSB first.
inline float powi(float a, int b) {
    const bool recip = b < 0;
    float r = 1.0f;
    while (true) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0f / r : r;
}
// f32::hypot.  Evaluated in f64 (the squares are exact, one rounding in the sum, one in the
// sqrt, one to f32) so that the CUDA side can reproduce it bit-for-bit with IEEE double ops.
inline float hypot32(float x, float y) {
    double dx = x, dy = y;
    return (float)sqrt(dx * dx + dy * dy);
}

// Interpolate for scalars (ref: src/numeric/interpolate.rs:43-57)
inline float bary(float u, float ux, float v, float vx, float w, float wx) {
    return ux * u + vx * v + wx * w;  // (ux*u + vx*v) + wx*w
}
inline float lerp(float t, float x1, float x2) {
    return (1.0f - t) * x1 + t * x2;
}

struct Storage {  // ref: src/pipeline/storage.rs:8-12 (three flat Vecs of vertices)
    std::vector<float> points, lines, tris;
    void append(const Storage &o) {
        points.insert(points.end(), o.points.begin(), o.points.end());
        lines.insert(lines.end(), o.lines.begin(), o.lines.end());
        tris.insert(tris.end(), o.tris.begin(), o.tris.end());
    }
};

inline void push_vertex(std::vector<float> &dst, const float *v, uint32_t S) {
    dst.insert(dst.end(), v, v + S);
}

}  // namespace

struct so_draw {
    int primitive = SR_TRIANGLE;
    std::vector<uint32_t> indices;
    bool has_stencil_value = false;
    uint32_t stencil_value = 0;  // Option::unwrap_or_default (ref: src/pipeline/mod.rs:155)
    uint32_t nk = 0;
    bool have_indexed = false;   // indexed_vertices: Option<Vec<..>>
    std::vector<float> indexed;  // records of 4+nk floats
    Storage gen;                 // generated_primitives
    int space = 0;               // 0 = clip space, 1 = screen space (after finish)
};

namespace {

// ---------------------------------------------------------------------------
// a15: registered vertex shaders
// ---------------------------------------------------------------------------

uint32_t vs_nk(int vs, uint32_t vin_floats) {
    switch (vs) {
        case SR_VS_PASSTHROUGH: return vin_floats >= 4 ? vin_floats - 4 : 0;
        case SR_VS_SUZANNE: return 8;
        case SR_VS_FULL_EXAMPLE: return 10;
    }
    return 0;
}
uint32_t vs_vin(int vs) {
    switch (vs) {
        case SR_VS_SUZANNE: return 6;
        case SR_VS_FULL_EXAMPLE: return 8;
    }
    return 0;  // passthrough: any >= 4
}

void vertex_shader(int vs, const sr_uniforms *u, const float *in, uint32_t vin_floats, float *out) {
    switch (vs) {
        case SR_VS_PASSTHROUGH: {
            for (uint32_t i = 0; i < vin_floats; ++i) out[i] = in[i];
            break;
        }
        case SR_VS_SUZANNE: {
            // ref: examples/suzanne.rs:123-141
            float pos_h[4] = {in[0], in[1], in[2], 1.0f};  // Point3::to_homogeneous
            float nrm_h[4] = {in[3], in[4], in[5], 0.0f};  // Vector3::to_homogeneous
            float world[4], n[4], pv[16];
            mat_vec(u->model, pos_h, world);               // model * position
            mat_vec(u->mit, nrm_h, n);
            normalize4(n, out + 8);                        // (mit * normal).normalize()
            mat_mat(u->projection, u->view, pv);           // `projection * view * world` is left-associative
            mat_vec(pv, world, out);
            for (int i = 0; i < 4; ++i) out[4 + i] = world[i];
            break;
        }
        case SR_VS_FULL_EXAMPLE: {
            // ref: full_example/src/shaders.rs:8-31
            float pos_h[4] = {in[0], in[1], in[2], 1.0f};
            float nrm_h[4] = {in[3], in[4], in[5], 0.0f};
            float world[4], n[4], pv[16], mvp[16];
            mat_vec(u->model, pos_h, world);
            mat_vec(u->mit, nrm_h, n);
            normalize4(n, out + 8);
            mat_mat(u->projection, u->view, pv);
            mat_mat(pv, u->model, mvp);                    // projection * view * model
            mat_vec(mvp, pos_h, out);                      // mvp * position
            for (int i = 0; i < 4; ++i) out[4 + i] = world[i];
            out[12] = in[6];
            out[13] = in[7];
            break;
        }
    }
}

// ---------------------------------------------------------------------------
// a3: ClipVertex::normalize (ref: src/geometry/clipvertex.rs:89-127)
// ---------------------------------------------------------------------------
void normalize_vertex(float *v, const sr_viewport &vp) {
    const float x = v[0], y = v[1], z = v[2], w = v[3];
    const float left = vp.x, bottom = vp.y;
    const float right = left + vp.width;
    const float top = bottom + vp.height;
    // Matrix4::new takes rows; stored column-major here.
    float m[16] = {0};
    m[0 * 4 + 0] = (right - left) / 2.0f;
    m[3 * 4 + 0] = (right + left) / 2.0f;
    m[1 * 4 + 1] = (top - bottom) / -2.0f;
    m[3 * 4 + 1] = (top + bottom) / 2.0f;
    m[2 * 4 + 2] = (vp.far_ - vp.near_) / -2.0f;
    m[3 * 4 + 2] = (vp.far_ + vp.near_) / -2.0f;
    m[3 * 4 + 3] = 1.0f;
    float ndc[4] = {x / w, y / w, z / w, 1.0f};
    float screen[4];
    mat_vec(m, ndc, screen);
    screen[3] = 1.0f / w;
    for (int i = 0; i < 4; ++i) v[i] = screen[i];
}

// Mapper::map (ref: src/parallel.rs:57-83): chunks of 64*size_of::<U>() elements per fetch_add.
template <class F>
void mapper_map(uint64_t n, uint32_t out_floats, int nthreads, F &&fn) {
    const uint64_t chunk = 64ull * out_floats * 4ull;
    std::atomic<uint64_t> cursor{0};
    run_threads(nthreads, [&](int) {
        while (true) {
            uint64_t i = cursor.fetch_add(chunk, std::memory_order_relaxed);
            if (i >= n) break;
            uint64_t e = i + chunk < n ? i + chunk : n;
            for (; i < e; ++i) fn(i);
        }
    });
}

// ---------------------------------------------------------------------------
// a4: clipping planes (ref: src/geometry/clip.rs:33-63) and clip_primitives
//     (ref: src/pipeline/stages/geometry.rs:261-336)
// ---------------------------------------------------------------------------
inline bool has_inside(int plane, const float *v) {
    const float x = v[0], y = v[1], z = v[2], w = v[3];
    switch (plane) {
        case 0: return x >= -w;   // Left
        case 1: return x <= w;    // Right
        case 2: return y >= -w;   // Top
        case 3: return y <= w;    // Bottom
        case 4: return z >= 0.0f; // Near
        default: return z <= w;   // Far
    }
}
inline void intersect(int plane, const float *v1, const float *v2, uint32_t S, float *out) {
    const float x1 = v1[0], y1 = v1[1], z1 = v1[2], w1 = v1[3];
    const float x2 = v2[0], y2 = v2[1], z2 = v2[2], w2 = v2[3];
    float a, b;
    switch (plane) {
        case 0: a = w1 + x1; b = w2 + x2; break;
        case 1: a = w1 - x1; b = w2 - x2; break;
        case 2: a = w1 + y1; b = w2 + y2; break;
        case 3: a = w1 - y1; b = w2 - y2; break;
        case 4: a = z1; b = z2; break;
        default: a = w1 - z1; b = w2 - z2; break;
    }
    const float t = a / (a - b);
    for (uint32_t i = 0; i < S; ++i) out[i] = lerp(t, v1[i], v2[i]);  // position and every uniform
}

void clip_triangle(Storage &out, const float *a, const float *b, const float *c, uint32_t S) {
    std::vector<float> polygon;
    polygon.reserve(36 * S);
    std::vector<float> tmp(S);
    const float *edges[3][2] = {{a, b}, {b, c}, {c, a}};
    for (auto &e : edges) {
        const float *s = e[0], *p = e[1];
        for (int plane = 0; plane < 6; ++plane) {
            const bool s_in = has_inside(plane, s);
            const bool p_in = has_inside(plane, p);
            if (s_in != p_in) {
                intersect(plane, s, p, S, tmp.data());
                push_vertex(polygon, tmp.data(), S);
            }
            if (p_in) push_vertex(polygon, p, S);
        }
    }
    const size_t len = polygon.size() / S;
    if (len == 3) {
        out.tris.insert(out.tris.end(), polygon.begin(), polygon.end());
    } else if (len > 3) {
        const float *last = &polygon[(len - 1) * S];
        for (size_t i = 0; i < len - 2; ++i) {
            push_vertex(out.tris, last, S);
            push_vertex(out.tris, &polygon[i * S], S);
            push_vertex(out.tris, &polygon[(i + 1) * S], S);
        }
    }
}

// SR_GS_CLIP_SH: Sutherland-Hodgman against the six planes of clip.rs:37-42, one plane after the other; an edge runs
// from the previous vertex s to the current vertex p and crossings are intersect(plane, s, p) (clip.rs:47-63).  Not a
// restatement of reference behaviour (the reference's clipper is clip_triangle above) but of the opt-in correct
// clipper the library offers; the oracle carries the same definition so that it can be checked bit for bit.
void clip_triangle_sh(Storage &out, const float *a, const float *b, const float *c, uint32_t S) {
    std::vector<float> cur, nxt, tmp(S);
    push_vertex(cur, a, S); push_vertex(cur, b, S); push_vertex(cur, c, S);
    for (int plane = 0; plane < 6 && !cur.empty(); ++plane) {
        nxt.clear();
        const size_t n = cur.size() / S;
        for (size_t i = 0; i < n; ++i) {
            const float *s = &cur[((i + n - 1) % n) * S], *p = &cur[i * S];
            const bool s_in = has_inside(plane, s), p_in = has_inside(plane, p);
            if (p_in) {
                if (!s_in) { intersect(plane, s, p, S, tmp.data()); push_vertex(nxt, tmp.data(), S); }
                push_vertex(nxt, p, S);
            } else if (s_in) {
                intersect(plane, s, p, S, tmp.data());
                push_vertex(nxt, tmp.data(), S);
            }
        }
        cur.swap(nxt);
    }
    const size_t len = cur.size() / S;
    for (size_t i = 1; i + 1 < len; ++i) {  // fan around the first vertex
        push_vertex(out.tris, &cur[0], S);
        push_vertex(out.tris, &cur[i * S], S);
        push_vertex(out.tris, &cur[(i + 1) * S], S);
    }
}

void clip_line(Storage &out, const float *start_in, const float *end_in, uint32_t S) {
    std::vector<float> start(start_in, start_in + S), end(end_in, end_in + S), isect(S);
    int intersections = 0;
    for (int plane = 0; plane < 6; ++plane) {
        const bool s_in = has_inside(plane, start.data());
        const bool p_in = has_inside(plane, end.data());
        if (s_in != p_in) {
            intersect(plane, start.data(), end.data(), S, isect.data());
            if (s_in) end = isect;
            else if (p_in) start = isect;
            intersections += 1;
        } else if (!s_in) {
            return;
        }
        if (intersections > 2) break;
    }
    push_vertex(out.lines, start.data(), S);
    push_vertex(out.lines, end.data(), S);
}

void clip_point(Storage &out, const float *p, uint32_t S) {
    for (int plane = 0; plane < 6; ++plane)
        if (!has_inside(plane, p)) return;
    push_vertex(out.points, p, S);
}

// registered geometry shaders (a15)
constexpr float NORMAL_LENGTH = 0.05f;  // ref: full_example/src/shaders.rs:33

void gs_apply(int gs, Storage &out, int kind, const float *a, const float *b, const float *c, uint32_t S,
              const sr_uniforms *u) {
    // kind: 1 point (a), 2 line (a,b), 3 triangle (a,b,c)
    if (gs == SR_GS_CLIP || gs == SR_GS_CLIP_SH) {
        if (kind == 3) { if (gs == SR_GS_CLIP) clip_triangle(out, a, b, c, S); else clip_triangle_sh(out, a, b, c, S); }
        else if (kind == 2) clip_line(out, a, b, S);
        else clip_point(out, a, S);
        return;
    }
    if (kind != 3) {  // `_ => storage.re_emit(primitive)`
        if (kind == 1) push_vertex(out.points, a, S);
        else { push_vertex(out.lines, a, S); push_vertex(out.lines, b, S); }
        return;
    }
    float mv[16];
    mat_mat(u->projection, u->view, mv);  // let mv = projection * view;
    std::vector<float> rec(S);
    if (gs == SR_GS_FACE_NORMALS) {
        // ref: full_example/src/shaders.rs:63-89
        const float third = 1.0f / 3.0f;
        std::vector<float> center(S - 4);
        for (uint32_t i = 0; i < S - 4; ++i) center[i] = bary(third, a[4 + i], third, b[4 + i], third, c[4 + i]);
        float start[4], end[4], nn[4], tip[4];
        mat_vec(mv, center.data(), start);            // mv * center.position
        normalize4(center.data() + 4, nn);            // center.normal.normalize()
        for (int i = 0; i < 4; ++i) tip[i] = center[i] + nn[i] * NORMAL_LENGTH;
        mat_vec(mv, tip, end);
        for (int i = 0; i < 4; ++i) rec[i] = start[i];
        for (uint32_t i = 0; i < S - 4; ++i) rec[4 + i] = center[i];
        push_vertex(out.lines, rec.data(), S);
        for (int i = 0; i < 4; ++i) rec[i] = end[i];
        push_vertex(out.lines, rec.data(), S);
    } else {  // SR_GS_VERTEX_NORMALS, ref: full_example/src/shaders.rs:35-61
        const float *vs[3] = {a, b, c};
        for (const float *v : vs) {
            float start[4], end[4], tip[4];
            mat_vec(mv, v + 4, start);                // mv * position
            for (int i = 0; i < 4; ++i) tip[i] = v[4 + i] + v[8 + i] * NORMAL_LENGTH;
            mat_vec(mv, tip, end);
            for (uint32_t i = 4; i < S; ++i) rec[i] = v[i];
            for (int i = 0; i < 4; ++i) rec[i] = start[i];
            push_vertex(out.lines, rec.data(), S);
            for (int i = 0; i < 4; ++i) rec[i] = end[i];
            push_vertex(out.lines, rec.data(), S);
        }
    }
}

// ---------------------------------------------------------------------------
// a14: stencil (ref: src/stencil.rs:112-123,147-158) for the unsigned stencil types u8 / u16 / u32 (the Stencil trait,
// src/stencil.rs:9-60: wrapping_add / wrapping_sub / saturating_* / not of the primitive type); values are held in 32 bits,
// smax = the type's MAX
// ---------------------------------------------------------------------------
inline bool stencil_test(uint32_t test, uint32_t value, uint32_t mask) {
    switch (test) {
        case SR_STENCIL_ALWAYS: return true;
        case SR_STENCIL_NEVER: return false;
        case SR_STENCIL_LESS_THAN: return mask < value;
        case SR_STENCIL_LESS_THAN_EQ: return mask <= value;
        case SR_STENCIL_GREATER_THAN: return mask > value;
        case SR_STENCIL_GREATER_THAN_EQ: return mask >= value;
        case SR_STENCIL_EQUAL: return mask == value;
        case SR_STENCIL_NOT_EQUAL: return mask != value;
    }
    return false;
}
inline uint32_t stencil_op(uint32_t op, uint32_t value, uint32_t mask, uint32_t smax = 0xFFu) {
    switch (op) {
        case SR_STENCIL_KEEP: return value;
        case SR_STENCIL_INVERT: return ~value & smax;                                     // Stencil::not
        case SR_STENCIL_ZERO: return 0;
        case SR_STENCIL_REPLACE: return mask;
        case SR_STENCIL_INCREMENT_WRAP: return (value + 1u) & smax;                       // wrapping_add(one)
        case SR_STENCIL_DECREMENT_WRAP: return (value - 1u) & smax;                       // wrapping_sub(one)
        case SR_STENCIL_INCREMENT_SAT: return value == smax ? smax : value + 1u;          // saturating_add(one)
        case SR_STENCIL_DECREMENT_SAT: return value == 0 ? 0u : value - 1u;               // saturating_sub(one)
    }
    return value;
}
inline uint32_t stencil_bytes_of(const so_framebuffer *fb) { return fb->stencil_bytes ? fb->stencil_bytes : 1u; }
inline uint32_t stencil_load(const so_framebuffer *fb, uint64_t i) {
    switch (stencil_bytes_of(fb)) {
        case 2: return reinterpret_cast<const uint16_t *>(fb->stencil)[i];
        case 4: return reinterpret_cast<const uint32_t *>(fb->stencil)[i];
    }
    return fb->stencil[i];
}
inline void stencil_store(so_framebuffer *fb, uint64_t i, uint32_t v) {
    switch (stencil_bytes_of(fb)) {
        case 2: reinterpret_cast<uint16_t *>(fb->stencil)[i] = (uint16_t)v; return;
        case 4: reinterpret_cast<uint32_t *>(fb->stencil)[i] = v; return;
    }
    fb->stencil[i] = (uint8_t)v;
}

// ---------------------------------------------------------------------------
// a13: blend (ref: src/color/blend.rs:28-31; full_example/src/color.rs:5-17)
// ---------------------------------------------------------------------------
inline void blend(uint32_t mode, const float *a /*src*/, const float *b /*dst*/, float *out) {
    if (mode == SR_BLEND_ALPHA_OVER) {
        auto over = [](float x, float y, float a_, float b_) {
            float a1 = 1.0f - a_;
            return (x * a_ + y * b_ * a1) / (a_ + b_ * a1);
        };
        float r[4] = {over(a[0], b[0], a[3], b[3]), over(a[1], b[1], a[3], b[3]), over(a[2], b[2], a[3], b[3]),
                      a[3] + b[3] * (1.0f - a[3])};
        for (int i = 0; i < 4; ++i) out[i] = r[i];
    } else if (mode == SR_BLEND_ADDITIVE) {  // a user blend, GenericBlend::new(|a, b| a + b) (ref: src/color/blend.rs:57-76)
        for (int i = 0; i < 4; ++i) out[i] = a[i] + b[i];
    } else {
        for (int i = 0; i < 4; ++i) out[i] = a[i];
    }
}

// ---------------------------------------------------------------------------
// a15: registered fragment shaders.  Returns false for Fragment::Discard.
// `sv` = interpolated ScreenVertex record (position[4] + K).
// ---------------------------------------------------------------------------
inline float saturate(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }
inline float fresnel_schlick(float cos_theta, float ior) {
    float f0 = powi((1.0f - ior) / (1.0f + ior), 2);
    return f0 + (1.0f - f0) * powi(1.0f - cos_theta, 5);
}

// Rust's `f as u32` (saturating; NaN -> 0), which the sampler of full_example/src/texture.rs:52-62 relies on
inline uint32_t as_u32(float f) {
    if (!(f > 0.0f)) return 0u;
    if (f >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)f;
}

// texture(t, coord, filter, edge): src/texture.rs:14-45 names the call, full_example/src/texture.rs:25-84 is the only
// sampler the reference ships (TextureRead::sample is unimplemented!(), src/texture.rs:49-52), restated here for every
// Filter / Edge and both texel formats.  Definitions where the reference has none: Edge::Border(C) returns C for a
// coordinate outside [0,1]^2 (or NaN) and samples as Clamp inside (GL_CLAMP_TO_BORDER, src/texture.rs:43-44); a
// framebuffer colour plane (TextureBufferRef, texturebuffer.rs:12-58) yields its f32 colour as stored -- no /255 and no
// decode_gamma, which full_example applies to 8-bit sRGB images only.  Deviation: the reference reads texel x+1 / y+1
// unclamped (the image crate would panic at u==1 or v==1); texel indices it would read out of bounds are clamped to
// the last row / column.
void texture_sample(const so_texture *t, float u, float v, float *out) {
    if (t->edge == SR_EDGE_WRAP) {
        u = u - truncf(u);  // f32::fract
        v = v - truncf(v);
    } else {
        if (t->edge == SR_EDGE_BORDER && !(u >= 0.0f && u <= 1.0f && v >= 0.0f && v <= 1.0f)) {
            for (int ch = 0; ch < 4; ++ch) out[ch] = t->border[ch];
            return;
        }
        u = fmaxf(fminf(u, 1.0f), 0.0f);
        v = fmaxf(fminf(v, 1.0f), 0.0f);
    }
    const uint32_t lastx = t->width - 1, lasty = t->height - 1;
    const bool image = t->rgba != nullptr;
    auto texel = [&](uint32_t px, uint32_t py, int ch) {
        px = px < lastx ? px : lastx;
        py = py < lasty ? py : lasty;
        const size_t i = (size_t)py * t->width + px;
        return image ? (float)t->rgba[i * 4 + ch] / 255.0f : t->texels_f32[i * t->stride + ch];
    };
    float val[4];
    if (t->filter == SR_FILTER_NEAREST) {
        const uint32_t x = as_u32(roundf(u * (float)lastx)), y = as_u32(roundf(v * (float)lasty));
        for (int ch = 0; ch < 4; ++ch) val[ch] = texel(x, y, ch);
    } else {
        const float uu = (u * (float)lastx) + 0.5f, vv = (v * (float)lasty) + 0.5f;
        const uint32_t x = as_u32(floorf(uu)), y = as_u32(floorf(vv));
        const float u_ratio = uu - (float)x, v_ratio = vv - (float)y;
        const float u_opp = 1.0f - u_ratio, v_opp = 1.0f - v_ratio;
        const uint32_t x1 = x == 0xFFFFFFFFu ? x : x + 1, y1 = y == 0xFFFFFFFFu ? y : y + 1;
        for (int ch = 0; ch < 4; ++ch) {
            float xy = texel(x, y, ch), x1y = texel(x1, y, ch), xy1 = texel(x, y1, ch), x1y1 = texel(x1, y1, ch);
            val[ch] = (xy * u_opp + x1y * u_ratio) * v_opp + (xy1 * u_opp + x1y1 * u_ratio) * v_ratio;
        }
    }
    for (int ch = 0; ch < 4; ++ch) out[ch] = (image && ch < 3) ? powf(val[ch], 2.2f) : val[ch];
}

// `out` = the colour the shader returns; a shader returning a tuple of two colours (SR_FS_SUZANNE_GBUFFER, for a texture buffer
// with two colour planes: texturebuffer.rs:129-147) writes the second one to out[4..8)
bool fragment_shader(int fs, const float *sv, const sr_uniforms *u, const so_texture *tex, float *out) {
    const float *K = sv + 4;
    switch (fs) {
        case SR_FS_SUZANNE_GBUFFER:
            for (int i = 0; i < 4; ++i) out[4 + i] = K[4 + i];  // .1 = the interpolated normal
            return fragment_shader(SR_FS_SUZANNE, sv, u, tex, out);  // .0 = the suzanne colour
        case SR_FS_FLAT:
            for (int i = 0; i < 4; ++i) out[i] = K[i];
            return true;
        case SR_FS_GREEN:
            out[0] = 0.0f; out[1] = 1.0f; out[2] = 0.0f; out[3] = 1.0f;
            return true;
        case SR_FS_TEXTURE_UNLIT:
            if (tex && (tex->rgba || tex->texels_f32)) texture_sample(tex, K[0], K[1], out);
            else out[0] = out[1] = out[2] = out[3] = 0.0f;
            return true;
        case SR_FS_DISCARD_CHECKER: {
            int xi = (int)floorf(sv[0]), yi = (int)floorf(sv[1]);
            if ((xi + yi) & 1) return false;
            for (int i = 0; i < 4; ++i) out[i] = K[i];
            return true;
        }
        case SR_FS_SUZANNE: {
            // ref: examples/suzanne.rs:147-183
            const float *position = K, *normal = K + 4;
            float d[4], view_dir[4], light_dir[4], h[4], halfway[4];
            for (int i = 0; i < 4; ++i) d[i] = u->camera[i] - position[i];
            normalize4(d, view_dir);
            for (int i = 0; i < 4; ++i) d[i] = u->sz_light[i] - position[i];
            normalize4(d, light_dir);
            for (int i = 0; i < 4; ++i) h[i] = light_dir[i] + view_dir[i];
            normalize4(h, halfway);
            float NdotL = fmaxf(fminf(dot4(light_dir, normal), 1.0f), 0.0f);
            float NdotH = fmaxf(fminf(dot4(normal, halfway), 1.0f), 0.0f);
            float VdotH = fmaxf(fminf(dot4(view_dir, halfway), 1.0f), 0.0f);
            float f = fresnel_schlick(VdotH, 1.45f);
            float diffuse = NdotL * (1.0f - f);
            float specular = f * powf(NdotH, 32.0f * 2.0f);
            const float inv_gamma = 1.0f / 2.2f;
            for (int i = 0; i < 3; ++i)
                out[i] = powf(u->sz_intensity * (specular + (diffuse * u->sz_color[i])), inv_gamma);
            out[3] = 1.0f;
            return true;
        }
        case SR_FS_FULL_EXAMPLE:
        case SR_FS_FULL_EXAMPLE_TEXTURED: {
            // ref: full_example/src/shaders.rs:108-162
            const float *position = K, *normal = K + 4;
            float d[4], view_dir[4];
            for (int i = 0; i < 4; ++i) d[i] = u->camera[i] - position[i];
            normalize4(d, view_dir);
            const float m = powf(0.25f, 2.2f);  // decode_gamma(material colour)
            float material[3] = {m, m, m};
            if (fs == SR_FS_FULL_EXAMPLE_TEXTURED && tex && (tex->rgba || tex->texels_f32)) {
                float t[4];
                texture_sample(tex, K[8], K[9], t);
                for (int i = 0; i < 3; ++i) material[i] = material[i] * t[i];
            }
            const float albedo = 0.7f;
            float color[3] = {0.0f, 0.0f, 0.0f};
            for (uint32_t l = 0; l < u->nlights && l < SR_MAX_LIGHTS; ++l) {
                const sr_light &light = u->lights[l];
                float lp[4] = {light.position[0], light.position[1], light.position[2], 1.0f};
                float ld[4], light_dir[4], h[4], halfway[4];
                for (int i = 0; i < 4; ++i) ld[i] = lp[i] - position[i];
                float light_distance = norm4(ld);
                normalize4(ld, light_dir);
                for (int i = 0; i < 4; ++i) h[i] = light_dir[i] + view_dir[i];
                normalize4(h, halfway);
                float intensity = light.intensity / powi(light_distance, 2);
                float NdotL = saturate(dot4(light_dir, normal));
                float NdotH = saturate(dot4(normal, halfway));
                float VdotH = saturate(dot4(view_dir, halfway));
                float f = fresnel_schlick(VdotH, 1.45f);
                float diffuse = (1.0f - f) * NdotL;
                float specular = f * powi(NdotH, 32 * 2);
                for (int i = 0; i < 3; ++i)
                    color[i] += intensity * light.color[i] * (specular + (diffuse * albedo * material[i]));
            }
            const float inv_gamma = 1.0f / 2.2f;
            for (int i = 0; i < 3; ++i) {
                // aces_filmic_tonemap_component, full_example/src/color.rs:20-28
                float x = color[i];
                float tm = (x * (2.51f * x + 0.03f)) / (x * (2.43f * x + 0.59f) + 0.14f);
                out[i] = powf(tm, inv_gamma);
            }
            out[3] = 1.0f;
            return true;
        }
    }
    return false;
}

// ---------------------------------------------------------------------------
// RasterArguments (ref: src/pipeline/stages/rasterization/mod.rs:14-23)
// ---------------------------------------------------------------------------
struct RasterArgs {
    uint32_t width, height;
    uint32_t tx0, ty0, tx1, ty1;  // tile, inclusive
    float bx0, by0, bx1, by1;     // bounds = tile cast to float
    uint32_t stencil_value;
    uint32_t stencil_test, stencil_op;
    bool aa_lines;
    uint32_t cull;
    uint32_t blend;
    int fs;
    const sr_uniforms *uniforms;
    const so_texture *tex;
    uint32_t S;
};

// Shared tail of all three rasterisers after the stencil step: z<0, depth test, shade, blend.
// RGBAu8Color targets (ref: src/color/predefined.rs:26).  The registered shaders compute f32 colours; their u8 form is
// `(c * 255.0) as u8` per channel -- Rust's saturating, truncating `as` (NaN -> 0), the conversion of the reference's own
// presentation loop (realtime_example/src/main.rs:100-116).
inline uint8_t as_u8(float c) {
    const float v = c * 255.0f;
    if (!(v > 0.0f)) return 0;  // negative, zero, NaN
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}
// AlphaMultiply for u8 (ref: src/color/helper.rs:36-42): (channel as f32 * (alpha as f32 / 255.0)) as u8
inline uint8_t mul_alpha_u8(uint8_t channel, uint8_t alpha) {
    const float v = (float)channel * ((float)alpha / 255.0f);
    if (!(v > 0.0f)) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}

inline void shade_and_write(const RasterArgs &A, so_framebuffer *fb, uint64_t index, const float *sv, float alpha,
                            bool use_alpha, uint32_t prim_id, double alpha_f64 = 1.0) {
    const float z = sv[2];
    if (z < 0.0f) {
        const float d = z;  // Depth::from_scalar
        const float dt = fb->depth[index];
        if (d >= dt) {
            float c[8];
            if (fb->color1) {  // tuple colour, Blend = (): set_pixel_unchecked stores each colour into its plane (texturebuffer.rs:141-147)
                if (!fragment_shader(A.fs, sv, A.uniforms, A.tex, c)) return;
                if (use_alpha) c[3] = c[3] * alpha;
                for (int i = 0; i < 4; ++i) { fb->color[index * 4 + i] = c[i]; fb->color1[index * 4 + i] = c[4 + i]; }
                fb->depth[index] = d;
                if (fb->winner) fb->winner[index] = prim_id + 1;
                return;
            }
            if (fb->color_u8) {  // Blend = () on a u8 colour: blend(c.mul_alpha(..), p) = the source colour
                if (!fragment_shader(A.fs, sv, A.uniforms, A.tex, c)) return;
                uint8_t q[4] = {as_u8(c[0]), as_u8(c[1]), as_u8(c[2]), as_u8(c[3])};
                // lines: c.mul_alpha(ColorAlpha::from_scalar(alpha)) (line.rs:100); from_scalar is NumCast f64 -> u8, i.e. the
                // coverage truncates to 0 or 1 (src/color/mod.rs:26-33) -- kept as the reference has it
                if (use_alpha) q[3] = mul_alpha_u8(q[3], (uint8_t)alpha_f64);
                for (int i = 0; i < 4; ++i) fb->color_u8[index * 4 + i] = q[i];
                fb->depth[index] = d;
                if (fb->winner) fb->winner[index] = prim_id + 1;
                return;
            }
            if (fragment_shader(A.fs, sv, A.uniforms, A.tex, c)) {
                if (use_alpha) c[3] = c[3] * alpha;  // Color::mul_alpha scales the alpha channel only (predefined.rs:82-86)
                float outc[4];
                blend(A.blend, c, &fb->color[index * 4], outc);
                for (int i = 0; i < 4; ++i) fb->color[index * 4 + i] = outc[i];
                fb->depth[index] = d;
                if (fb->winner) fb->winner[index] = prim_id + 1;
            }
        }
    }
}

inline bool stencil_step(const RasterArgs &A, so_framebuffer *fb, uint64_t index) {
    if (!fb->stencil) return true;  // stencil type (): test Always, op Keep (ref: src/stencil.rs:65-86,169-175)
    const uint32_t bytes = stencil_bytes_of(fb), smax = bytes == 1 ? 0xFFu : bytes == 2 ? 0xFFFFu : 0xFFFFFFFFu;
    const uint32_t sval = stencil_load(fb, index), mesh = A.stencil_value & smax;  // the mesh's value is of the buffer's type
    if (!stencil_test(A.stencil_test, sval, mesh)) return false;
    stencil_store(fb, index, stencil_op(A.stencil_op, sval, mesh, smax));
    return true;
}

// a6: rasterize_triangle (ref: src/pipeline/stages/rasterization/triangle.rs:23-155)
void rasterize_triangle(const RasterArgs &A, so_framebuffer *fb, const float *a, const float *b, const float *c,
                        uint32_t prim_id, float *scratch) {
    const float x1 = a[0], y1 = a[1], x2 = b[0], y2 = b[1], x3 = c[0], y3 = c[1];
    // Deviation: NaN coordinates make the reference panic in cast(NaN).unwrap(); here the primitive is skipped.
    if (std::isnan(x1) || std::isnan(y1) || std::isnan(x2) || std::isnan(y2) || std::isnan(x3) || std::isnan(y3)) return;

    if (A.cull != SR_CULL_NONE) {
        const float area = x1 * y2 + x2 * y3 + x3 * y1 - x2 * y1 - x3 * y2 - x1 * y3;
        const uint32_t winding = std::signbit(area) ? SR_CLOCKWISE : SR_COUNTER_CLOCKWISE;
        if (winding == A.cull) return;
    }
    const float det = (y2 - y3) * (x1 - x3) + (x3 - x2) * (y1 - y3);

    auto clamp_as_int = [](float value, uint32_t lo, uint32_t hi) -> uint32_t {
        if (value < (float)lo) return lo;
        if (value > (float)hi) return hi;
        return (uint32_t)value;  // cast: truncation toward zero
    };
    const uint32_t minx = clamp_as_int(fminf(fminf(x1, x2), x3), A.tx0, A.tx1);
    const uint32_t miny = clamp_as_int(fminf(fminf(y1, y2), y3), A.ty0, A.ty1);
    const uint32_t maxx = clamp_as_int(fmaxf(fmaxf(x1, x2), x3), A.tx0, A.tx1);
    const uint32_t maxy = clamp_as_int(fmaxf(fmaxf(y1, y2), y3), A.ty0, A.ty1);

    const uint32_t S = A.S;
    for (uint32_t py = miny; py <= maxy; ++py) {
        for (uint32_t px = minx; px <= maxx; ++px) {
            const uint64_t index = (uint64_t)px + (uint64_t)py * A.width;
            if (!stencil_step(A, fb, index)) continue;
            const float x = (float)px + 0.5f, y = (float)py + 0.5f;
            const float u = ((y2 - y3) * (x - x3) + (x3 - x2) * (y - y3)) / det;
            const float v = ((y3 - y1) * (x - x3) + (x1 - x3) * (y - y3)) / det;
            const float w = 1.0f - u - v;
            if (!(u < 0.0f || v < 0.0f || w < 0.0f)) {
                for (int i = 0; i < 4; ++i) scratch[i] = bary(u, a[i], v, b[i], w, c[i]);
                const float z = scratch[2];
                if (z < 0.0f && z >= fb->depth[index]) {
                    for (uint32_t i = 4; i < S; ++i) scratch[i] = bary(u, a[i], v, b[i], w, c[i]);
                }
                shade_and_write(A, fb, index, scratch, 1.0f, false, prim_id);
            }
        }
    }
}

// liang_barsky_iterative (ref: src/geometry/line.rs:6-54)
bool liang_barsky(float x1, float y1, float x2, float y2, float xmin, float ymin, float xmax, float ymax, float *o) {
    float t0 = 0.0f, t1 = 1.0f;
    const float dx = x2 - x1, dy = y2 - y1;
    for (int edge = 0; edge < 4; ++edge) {
        float p, q;
        switch (edge) {
            case 0: p = -dx; q = x1 - xmin; break;
            case 1: p = dx; q = xmax - x1; break;
            case 2: p = -dy; q = y1 - ymin; break;
            default: p = dy; q = ymax - y1; break;
        }
        if (p == 0.0f && q < 0.0f) return false;
        const float r = q / p;
        if (p < 0.0f) {
            if (r > t1) return false;
            else if (r > t0) t0 = r;
        } else if (p > 0.0f) {
            if (r < t0) return false;
            else if (r < t1) t1 = r;
        }
    }
    o[0] = x1 + t0 * dx; o[1] = y1 + t0 * dy; o[2] = x1 + t1 * dx; o[3] = y1 + t1 * dy;
    return true;
}

// draw_line_bresenham (ref: src/pipeline/stages/rasterization/line.rs:125-151)
template <class P>
void draw_line_bresenham(int64_t x0, int64_t y0, int64_t x1, int64_t y1, P &&plot) {
    const int64_t dx = llabs(x1 - x0);
    const int64_t dy = -llabs(y1 - y0);
    const int64_t sx = x0 < x1 ? 1 : -1;
    const int64_t sy = y0 < y1 ? 1 : -1;
    int64_t err = dx + dy;
    while (true) {
        plot(x0, y0, 1.0);
        if (x0 == x1 && y0 == y1) break;
        const int64_t e2 = 2 * err;
        if (e2 >= dy) { err += dy; x0 += sx; }
        if (e2 <= dx) { err += dx; y0 += sy; }
    }
}

// draw_line_xiaolin_wu (ref: line.rs:159-239), f64 throughout
inline double fract64(double x) { return x - trunc(x); }
template <class P>
void draw_line_xiaolin_wu(double x0, double y0, double x1, double y1, P &&plot) {
    auto plot_float = [&](double x, double y, double opacity) { plot((int64_t)x, (int64_t)y, opacity); };
    const bool steep = fabs(y1 - y0) > fabs(x1 - x0);
    if (steep) { std::swap(x0, y0); std::swap(x1, y1); }
    if (x0 > x1) { std::swap(x0, x1); std::swap(y0, y1); }
    const double dx = x1 - x0, dy = y1 - y0;
    const double gradient = dx < 0.0001 ? 1.0 : dy / dx;
    double xend = round(x0);
    double yend = y0 + gradient * (xend - x0);
    double xgap = 1.0 - fract64(x0 + 0.5);
    const double xpxl1 = xend;
    const double ypxl1 = trunc(yend);
    if (steep) {
        plot_float(ypxl1, xpxl1, (1.0 - fract64(yend)) * xgap);
        plot_float(ypxl1 + 1.0, xpxl1, fract64(yend) * xgap);
    } else {
        plot_float(xpxl1, ypxl1, (1.0 - fract64(yend)) * xgap);
        plot_float(xpxl1, ypxl1 + 1.0, fract64(yend) * xgap);
    }
    double intery = yend + gradient;
    xend = round(x1);
    yend = y1 + gradient * (xend - x1);
    xgap = fract64(x1 + 0.5);
    const double xpxl2 = xend;
    const double ypxl2 = trunc(yend);
    if (steep) {
        plot_float(ypxl2, xpxl2, (1.0 - fract64(yend)) * xgap);
        plot_float(ypxl2 + 1.0, xpxl2, fract64(yend) * xgap);
    } else {
        plot_float(xpxl2, ypxl2, (1.0 - fract64(yend)) * xgap);
        plot_float(xpxl2, ypxl2 + 1.0, fract64(yend) * xgap);
    }
    double x = xpxl1 + 1.0;
    while (x <= (xpxl2 - 1.0)) {
        const double y = trunc(intery);
        if (steep) {
            plot_float(y, x, 1.0 - fract64(intery));
            plot_float(y + 1.0, x, fract64(intery));
        } else {
            plot_float(x, y, 1.0 - fract64(intery));
            plot_float(x, y + 1.0, fract64(intery));
        }
        intery += gradient;
        x += 1.0;
    }
}

// a9: rasterize_line (ref: src/pipeline/stages/rasterization/line.rs:22-119)
void rasterize_line(const RasterArgs &A, so_framebuffer *fb, const float *start, const float *end, uint32_t prim_id,
                    float *scratch) {
    float cl[4];
    if (std::isnan(start[0]) || std::isnan(start[1]) || std::isnan(end[0]) || std::isnan(end[1])) return;  // see triangle
    if (!liang_barsky(start[0], start[1], end[0], end[1], A.bx0, A.by0, A.bx1, A.by1, cl)) return;
    const float x1 = cl[0], y1 = cl[1], x2 = cl[2], y2 = cl[3];
    // Deviation: non-finite clipped end-points make the reference panic in cast(..).unwrap(); skipped here.
    if (!std::isfinite(x1) || !std::isfinite(y1) || !std::isfinite(x2) || !std::isfinite(y2)) return;
    const float d = hypot32(x1 - x2, y1 - y2);
    const uint32_t S = A.S;
    auto plot = [&](int64_t x, int64_t y, double alpha) {
        if (x >= 0 && y >= 0) {
            // Deviation: the reference indexes the framebuffer unchecked; coordinates past the frame
            // (possible only through Wu's y+1 / x+1 neighbour on the last row/column) are skipped here.
            if ((uint64_t)x >= A.width || (uint64_t)y >= A.height) return;
            const uint64_t index = (uint64_t)x + (uint64_t)y * A.width;
            if (!stencil_step(A, fb, index)) return;
            const float xf = (float)x + 0.5f, yf = (float)y + 0.5f;
            const float t = hypot32(x1 - xf, y1 - yf) / d;
            for (uint32_t i = 0; i < S; ++i) scratch[i] = lerp(t, start[i], end[i]);
            shade_and_write(A, fb, index, scratch, (float)alpha, true, prim_id, alpha);
        }
    };
    if (A.aa_lines) draw_line_xiaolin_wu((double)x1, (double)y1, (double)x2, (double)y2, plot);
    else draw_line_bresenham((int64_t)x1, (int64_t)y1, (int64_t)x2, (int64_t)y2, plot);
}

// a10: rasterize_point (ref: src/pipeline/stages/rasterization/point.rs:21-86)
void rasterize_point(const RasterArgs &A, so_framebuffer *fb, const float *p, uint32_t prim_id) {
    const float x = p[0], y = p[1];
    if (A.bx0 <= x && x < A.bx1 && A.by0 <= y && y < A.by1) {
        const uint64_t index = (uint64_t)(uint32_t)x + (uint64_t)(uint32_t)y * A.width;
        if (!stencil_step(A, fb, index)) return;
        shade_and_write(A, fb, index, p, 1.0f, false, prim_id);
    }
}

struct Tile { uint32_t x0, y0, x1, y1; };

// a5: tile list (ref: src/pipeline/stages/fragment.rs:188-216)
std::vector<Tile> make_tiles(uint32_t width, uint32_t height, uint32_t tw, uint32_t th) {
    std::vector<Tile> tiles;
    if (width == 0 || height == 0 || tw == 0 || th == 0) return tiles;
    const uint64_t xmax = width - 1, ymax = height - 1;
    uint64_t y = 0;
    while (y < ymax) {
        uint64_t x = 0;
        const uint64_t next_y = std::min<uint64_t>(y + th, ymax);
        while (x < xmax) {
            const uint64_t next_x = std::min<uint64_t>(x + tw, xmax);
            tiles.push_back({(uint32_t)x, (uint32_t)y, (uint32_t)next_x, (uint32_t)next_y});
            x = next_x;
        }
        y = next_y;
    }
    return tiles;
}

}  // namespace

// ===========================================================================
// C API
// ===========================================================================
extern "C" {

so_draw *so_draw_create(int primitive, const uint32_t *indices, uint64_t nindices, int has_sv, uint32_t sv) {
    if (primitive < SR_POINT || primitive > SR_TRIANGLE) return nullptr;
    if (nindices % (uint64_t)primitive != 0) return nullptr;  // assert_eq!(mesh.indices.len() % T::num_vertices(), 0)
    so_draw *d = new so_draw();
    d->primitive = primitive;
    d->indices.assign(indices, indices + nindices);
    d->has_stencil_value = has_sv != 0;
    d->stencil_value = has_sv ? sv : 0;
    return d;
}
void so_draw_destroy(so_draw *d) { delete d; }

static int vertex_stage(so_draw *d, const sr_viewport *vp, int vs, const sr_uniforms *u, const float *vin,
                        uint64_t nverts, uint32_t vin_floats, int nthreads) {
    if (!d || !u || (!vin && nverts)) return SR_ERR_INVALID_ARGUMENT;
    if (vs < SR_VS_PASSTHROUGH || vs > SR_VS_FULL_EXAMPLE) return SR_ERR_INVALID_ARGUMENT;
    if (vs == SR_VS_PASSTHROUGH ? vin_floats < 4 : vin_floats != vs_vin(vs)) return SR_ERR_INVALID_ARGUMENT;
    for (uint32_t idx : d->indices)
        if (idx >= nverts) return SR_ERR_INVALID_ARGUMENT;
    d->nk = vs_nk(vs, vin_floats);
    const uint32_t S = 4 + d->nk;
    d->indexed.assign((size_t)nverts * S, 0.0f);
    d->have_indexed = true;
    d->gen = Storage();
    mapper_map(nverts, S, nthreads, [&](uint64_t i) {
        float *out = &d->indexed[i * S];
        vertex_shader(vs, u, vin + i * vin_floats, vin_floats, out);
        if (vp) normalize_vertex(out, *vp);
    });
    d->space = vp ? 1 : 0;
    return SR_OK;
}

int so_draw_vertex_run(so_draw *d, int vs, const sr_uniforms *u, const float *vin, uint64_t nverts,
                       uint32_t vin_floats, int nthreads) {
    return vertex_stage(d, nullptr, vs, u, vin, nverts, vin_floats, nthreads);
}
int so_draw_vertex_run_to_fragment(so_draw *d, const sr_viewport *vp, int vs, const sr_uniforms *u, const float *vin,
                                   uint64_t nverts, uint32_t vin_floats, int nthreads) {
    if (!vp) return SR_ERR_INVALID_ARGUMENT;
    return vertex_stage(d, vp, vs, u, vin, nverts, vin_floats, nthreads);
}

int so_draw_set_vertices(so_draw *d, const float *verts, uint64_t nverts, uint32_t nk, int space) {
    if (!d) return SR_ERR_INVALID_ARGUMENT;
    for (uint32_t idx : d->indices)
        if (idx >= nverts) return SR_ERR_INVALID_ARGUMENT;
    d->nk = nk;
    d->indexed.assign(verts, verts + (size_t)nverts * (4 + nk));
    d->have_indexed = true;
    d->space = space;
    return SR_OK;
}
int so_draw_set_generated(so_draw *d, int which, const float *verts, uint64_t nverts, uint32_t nk) {
    if (!d || which < 1 || which > 3 || nverts % (uint64_t)which != 0) return SR_ERR_INVALID_ARGUMENT;
    if (d->have_indexed && d->nk != nk) return SR_ERR_INVALID_ARGUMENT;
    d->nk = nk;
    std::vector<float> &dst = which == 1 ? d->gen.points : which == 2 ? d->gen.lines : d->gen.tris;
    dst.assign(verts, verts + (size_t)nverts * (4 + nk));
    return SR_OK;
}

// GeometryShader::run (ref: src/pipeline/stages/geometry.rs:132-258).  nthreads<=1 gives the canonical
// order (generated points, lines, tris, then the indexed mesh); with threads, per-thread storages are
// concatenated in thread order (the reference merges in mutex-acquisition order).
int so_draw_geometry_run(so_draw *d, int gs, const sr_uniforms *u, int nthreads) {
    if (!d || d->space != 0) return SR_ERR_INVALID_STATE;
    if (gs < SR_GS_CLIP || gs > SR_GS_CLIP_SH) return SR_ERR_INVALID_ARGUMENT;
    if (gs != SR_GS_CLIP && gs != SR_GS_CLIP_SH && (!u || d->nk < 8)) return SR_ERR_INVALID_ARGUMENT;
    const uint32_t S = 4 + d->nk;
    const int nt = nthreads < 1 ? 1 : nthreads;
    std::vector<Storage> locals(nt);
    std::atomic<uint64_t> pc{0}, lc{0}, tc{0}, ic{0};
    const uint64_t npoints = d->gen.points.size() / S, nlines = d->gen.lines.size() / S, ntris = d->gen.tris.size() / S;
    const uint64_t n = (uint64_t)d->primitive;
    run_threads(nt, [&](int t) {
        Storage &st = locals[t];
        uint64_t i;
        while ((i = pc.fetch_add(1)) < npoints) gs_apply(gs, st, 1, &d->gen.points[i * S], nullptr, nullptr, S, u);
        while ((i = lc.fetch_add(2)) < nlines)
            gs_apply(gs, st, 2, &d->gen.lines[i * S], &d->gen.lines[(i + 1) * S], nullptr, S, u);
        while ((i = tc.fetch_add(3)) < ntris)
            gs_apply(gs, st, 3, &d->gen.tris[i * S], &d->gen.tris[(i + 1) * S], &d->gen.tris[(i + 2) * S], S, u);
        if (d->have_indexed) {
            while ((i = ic.fetch_add(n)) < d->indices.size()) {
                const float *a = &d->indexed[(size_t)d->indices[i] * S];
                const float *b = n > 1 ? &d->indexed[(size_t)d->indices[i + 1] * S] : nullptr;
                const float *c = n > 2 ? &d->indexed[(size_t)d->indices[i + 2] * S] : nullptr;
                gs_apply(gs, st, (int)n, a, b, c, S, u);
            }
        }
    });
    Storage merged;
    for (auto &st : locals) merged.append(st);
    d->gen = std::move(merged);
    d->have_indexed = false;  // indexed_vertices: None
    d->indexed.clear();
    return SR_OK;
}

// GeometryShader::finish (ref: geometry.rs:60-129)
int so_draw_finish(so_draw *d, const sr_viewport *vp, int nthreads) {
    if (!d || !vp || d->space != 0) return SR_ERR_INVALID_STATE;
    const uint32_t S = 4 + d->nk;
    auto norm_all = [&](std::vector<float> &v) {
        mapper_map(v.size() / S, S, nthreads, [&](uint64_t i) { normalize_vertex(&v[i * S], *vp); });
    };
    norm_all(d->gen.points);
    norm_all(d->gen.lines);
    norm_all(d->gen.tris);
    if (d->have_indexed) norm_all(d->indexed);
    d->space = 1;
    return SR_OK;
}

// FragmentShader::run (ref: src/pipeline/stages/fragment.rs:168-319)
int so_draw_fragment_run(so_draw *d, so_framebuffer *fb, const so_raster_state *st, int fs, const sr_uniforms *u,
                         const so_texture *tex, int nthreads) {
    return so_draw_fragment_run_tiles(d, fb, st, fs, u, tex, nthreads, 0, 1);
}

// The same driver restricted to the tiles first, first+stride, first+2*stride, ... of the reference's tile list
// (fragment.rs:188-216): every selected tile still visits every primitive (fragment.rs:268-311).  bench.py's CPU arm
// times a whole frame as `stride` such slices, one per step; (0, 1) is the whole frame.
int so_draw_fragment_run_tiles(so_draw *d, so_framebuffer *fb, const so_raster_state *st, int fs, const sr_uniforms *u,
                               const so_texture *tex, int nthreads, uint64_t tile_first, uint64_t tile_stride) {
    if (!d || !fb || !st || !u || tile_stride == 0) return SR_ERR_INVALID_ARGUMENT;
    if (d->space != 1) return SR_ERR_INVALID_STATE;
    if (fs < SR_FS_FLAT || fs > SR_FS_SUZANNE_GBUFFER) return SR_ERR_INVALID_ARGUMENT;
    if ((fs == SR_FS_SUZANNE_GBUFFER) != (fb->color1 != nullptr)) return SR_ERR_INVALID_STATE;  // a type error in the reference
    if (fb->color1 && (st->blend != SR_BLEND_REPLACE || fb->color_u8)) return SR_ERR_INVALID_ARGUMENT;
    if (fs == SR_FS_SUZANNE_GBUFFER && d->nk < 8) return SR_ERR_INVALID_ARGUMENT;
    if (fs == SR_FS_TEXTURE_UNLIT && d->nk < 2) return SR_ERR_INVALID_ARGUMENT;
    const uint32_t S = 4 + d->nk;
    if ((fs == SR_FS_FLAT || fs == SR_FS_DISCARD_CHECKER) && d->nk < 4) return SR_ERR_INVALID_ARGUMENT;
    if ((fs == SR_FS_SUZANNE || fs == SR_FS_FULL_EXAMPLE) && d->nk < 8) return SR_ERR_INVALID_ARGUMENT;
    if (fs == SR_FS_FULL_EXAMPLE_TEXTURED && d->nk < 10) return SR_ERR_INVALID_ARGUMENT;

    const std::vector<Tile> tiles = make_tiles(fb->width, fb->height, st->tile_width, st->tile_height);
    std::atomic<uint64_t> cursor{0};
    const uint64_t nidx = d->indices.size();
    const uint64_t ntri_idx = (d->primitive == SR_TRIANGLE && d->have_indexed) ? nidx / 3 : 0;
    const uint64_t ngen_tris = d->gen.tris.size() / S / 3;
    const uint64_t nline_idx = (d->primitive == SR_LINE && d->have_indexed) ? nidx / 2 : 0;
    const uint64_t ngen_lines = d->gen.lines.size() / S / 2;
    const uint64_t npoint_idx = (d->primitive == SR_POINT && d->have_indexed) ? nidx : 0;
    const uint64_t ngen_points = d->gen.points.size() / S;

    run_threads(nthreads < 1 ? 1 : nthreads, [&](int) {
        std::vector<float> scratch(S);
        while (true) {
            const uint64_t ti = tile_first + cursor.fetch_add(1, std::memory_order_relaxed) * tile_stride;
            if (ti >= tiles.size()) break;
            const Tile &tile = tiles[ti];
            RasterArgs A;
            A.width = fb->width; A.height = fb->height;
            A.tx0 = tile.x0; A.ty0 = tile.y0; A.tx1 = tile.x1; A.ty1 = tile.y1;
            A.bx0 = (float)tile.x0; A.by0 = (float)tile.y0; A.bx1 = (float)tile.x1; A.by1 = (float)tile.y1;
            A.stencil_value = d->stencil_value;
            A.stencil_test = st->stencil_test; A.stencil_op = st->stencil_op;
            A.aa_lines = st->antialiased_lines != 0;
            A.cull = st->cull_faces; A.blend = st->blend; A.fs = fs; A.uniforms = u; A.tex = tex; A.S = S;
            uint32_t prim = 0;
            for (uint64_t t = 0; t < ntri_idx; ++t, ++prim)
                rasterize_triangle(A, fb, &d->indexed[(size_t)d->indices[3 * t] * S],
                                   &d->indexed[(size_t)d->indices[3 * t + 1] * S],
                                   &d->indexed[(size_t)d->indices[3 * t + 2] * S], prim, scratch.data());
            for (uint64_t t = 0; t < ngen_tris; ++t, ++prim)
                rasterize_triangle(A, fb, &d->gen.tris[(3 * t) * S], &d->gen.tris[(3 * t + 1) * S],
                                   &d->gen.tris[(3 * t + 2) * S], prim, scratch.data());
            for (uint64_t l = 0; l < nline_idx; ++l, ++prim)
                rasterize_line(A, fb, &d->indexed[(size_t)d->indices[2 * l] * S],
                               &d->indexed[(size_t)d->indices[2 * l + 1] * S], prim, scratch.data());
            for (uint64_t l = 0; l < ngen_lines; ++l, ++prim)
                rasterize_line(A, fb, &d->gen.lines[(2 * l) * S], &d->gen.lines[(2 * l + 1) * S], prim, scratch.data());
            for (uint64_t p = 0; p < npoint_idx; ++p, ++prim)
                rasterize_point(A, fb, &d->indexed[(size_t)d->indices[p] * S], prim);
            for (uint64_t p = 0; p < ngen_points; ++p, ++prim)
                rasterize_point(A, fb, &d->gen.points[p * S], prim);
        }
    });
    return SR_OK;
}

void so_texture_sample(const so_texture *t, float u, float v, float out[4]) { texture_sample(t, u, v, out); }

uint64_t so_draw_count(const so_draw *d, int which) {
    if (!d) return 0;
    const uint32_t S = 4 + d->nk;
    switch (which) {
        case 0: return d->have_indexed ? d->indexed.size() / S : 0;
        case 1: return d->gen.points.size() / S;
        case 2: return d->gen.lines.size() / S;
        case 3: return d->gen.tris.size() / S;
    }
    return 0;
}
const float *so_draw_data(const so_draw *d, int which) {
    if (!d) return nullptr;
    switch (which) {
        case 0: return d->indexed.data();
        case 1: return d->gen.points.data();
        case 2: return d->gen.lines.data();
        case 3: return d->gen.tris.data();
    }
    return nullptr;
}
uint32_t so_draw_nk(const so_draw *d) { return d ? d->nk : 0; }

uint64_t so_tiles(uint32_t width, uint32_t height, uint32_t tw, uint32_t th, uint32_t *out, uint64_t cap) {
    const std::vector<Tile> tiles = make_tiles(width, height, tw, th);
    for (uint64_t i = 0; i < tiles.size() && i < cap; ++i) {
        out[i * 4 + 0] = tiles[i].x0; out[i * 4 + 1] = tiles[i].y0;
        out[i * 4 + 2] = tiles[i].x1; out[i * 4 + 3] = tiles[i].y1;
    }
    return tiles.size();
}

// a7 (derived): frame-clamped integer bbox of triangle.rs:66-78 with tile = whole frame, intersected with a
// disjoint grid of tw x th tiles.  Culled and NaN triangles are in no bin.
uint64_t so_draw_bins(const so_draw *d, uint32_t width, uint32_t height, uint32_t tw, uint32_t th, uint32_t cull,
                      uint64_t *offsets, uint32_t *ids) {
    if (!d || d->space != 1 || !width || !height || !tw || !th) return 0;
    const uint32_t S = 4 + d->nk;
    const uint32_t ntx = (width + tw - 1) / tw, nty = (height + th - 1) / th;
    const uint64_t ntiles = (uint64_t)ntx * nty;
    const uint64_t ntri_idx = (d->primitive == SR_TRIANGLE && d->have_indexed) ? d->indices.size() / 3 : 0;
    const uint64_t ngen = d->gen.tris.size() / S / 3;
    std::vector<std::vector<uint32_t>> bins(ntiles);
    auto clamp_as_int = [](float value, uint32_t lo, uint32_t hi) -> uint32_t {
        if (value < (float)lo) return lo;
        if (value > (float)hi) return hi;
        return (uint32_t)value;
    };
    for (uint64_t t = 0; t < ntri_idx + ngen; ++t) {
        const float *a, *b, *c;
        if (t < ntri_idx) {
            a = &d->indexed[(size_t)d->indices[3 * t] * S];
            b = &d->indexed[(size_t)d->indices[3 * t + 1] * S];
            c = &d->indexed[(size_t)d->indices[3 * t + 2] * S];
        } else {
            const uint64_t g = t - ntri_idx;
            a = &d->gen.tris[(3 * g) * S]; b = &d->gen.tris[(3 * g + 1) * S]; c = &d->gen.tris[(3 * g + 2) * S];
        }
        const float x1 = a[0], y1 = a[1], x2 = b[0], y2 = b[1], x3 = c[0], y3 = c[1];
        if (std::isnan(x1) || std::isnan(y1) || std::isnan(x2) || std::isnan(y2) || std::isnan(x3) || std::isnan(y3)) continue;
        if (cull != SR_CULL_NONE) {
            const float area = x1 * y2 + x2 * y3 + x3 * y1 - x2 * y1 - x3 * y2 - x1 * y3;
            const uint32_t winding = std::signbit(area) ? SR_CLOCKWISE : SR_COUNTER_CLOCKWISE;
            if (winding == cull) continue;
        }
        const uint32_t minx = clamp_as_int(fminf(fminf(x1, x2), x3), 0, width - 1);
        const uint32_t miny = clamp_as_int(fminf(fminf(y1, y2), y3), 0, height - 1);
        const uint32_t maxx = clamp_as_int(fmaxf(fmaxf(x1, x2), x3), 0, width - 1);
        const uint32_t maxy = clamp_as_int(fmaxf(fmaxf(y1, y2), y3), 0, height - 1);
        if (maxx < minx || maxy < miny) continue;  // inverted box: the reference's loops run zero times
        for (uint32_t ty = miny / th; ty <= maxy / th; ++ty)
            for (uint32_t tx = minx / tw; tx <= maxx / tw; ++tx) bins[(uint64_t)ty * ntx + tx].push_back((uint32_t)t);
    }
    uint64_t total = 0;
    for (uint64_t i = 0; i < ntiles; ++i) {
        if (offsets) offsets[i] = total;
        if (ids) memcpy(ids + total, bins[i].data(), bins[i].size() * sizeof(uint32_t));
        total += bins[i].size();
    }
    if (offsets) offsets[ntiles] = total;
    return total;
}

uint64_t so_coordinate_index(uint32_t x, uint32_t y, uint32_t width) {
    return (uint64_t)x + (uint64_t)y * (uint64_t)width;  // ref: src/geometry/coordinate.rs:47-51
}
int so_stencil_test(uint32_t test, uint8_t value, uint8_t mask) { return stencil_test(test, value, mask) ? 1 : 0; }
uint8_t so_stencil_op(uint32_t op, uint8_t value, uint8_t mask) { return (uint8_t)stencil_op(op, value, mask, 0xFFu); }
uint32_t so_stencil_op_wide(uint32_t op, uint32_t value, uint32_t mask, uint32_t bits) {
    return stencil_op(op, value, mask, bits == 8 ? 0xFFu : bits == 16 ? 0xFFFFu : 0xFFFFFFFFu);
}

float so_depth_far(void) {
    // Depth::far() = Bounded::min_value() = f32::MIN (ref: src/framebuffer/attachments/depth.rs:31)
    uint32_t bits = 0xFF7FFFFFu;
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

void so_framebuffer_clear(so_framebuffer *fb, const float color[4]) {
    // ref: src/framebuffer/renderbuffer/mod.rs:126-133
    const uint64_t n = (uint64_t)fb->width * fb->height;
    const float far_ = so_depth_far();
    for (uint64_t i = 0; i < n; ++i) {
        if (fb->color_u8) for (int c = 0; c < 4; ++c) fb->color_u8[i * 4 + c] = as_u8(color[c]);
        else for (int c = 0; c < 4; ++c) fb->color[i * 4 + c] = color[c];
        fb->depth[i] = far_;
        if (fb->stencil) stencil_store(fb, i, 0);
        if (fb->winner) fb->winner[i] = 0;
    }
}

void so_framebuffer_clear2(so_framebuffer *fb, const float color[4], const float color1[4]) {
    // ref: src/framebuffer/texturebuffer.rs:181-197 -- the tuple is destructured by name, each plane overwritten with its colour
    so_framebuffer_clear(fb, color);
    const uint64_t n = (uint64_t)fb->width * fb->height;
    if (fb->color1) for (uint64_t i = 0; i < n; ++i) for (int c = 0; c < 4; ++c) fb->color1[i * 4 + c] = color1[c];
}

}  // extern "C"
