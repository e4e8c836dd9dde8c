// sr_common.cuh -- shared device/host definitions for libsoftrender_b200 (sm_100a).
//
// Exactness contract: the translation unit is compiled with -fmad=false and the default
// -prec-div=true -prec-sqrt=true -ftz=false, so every `a*b+c` below is an IEEE multiply followed
// by an IEEE add and `/`, sqrtf are correctly rounded -- the same f32 expression trees the Rust
// reference evaluates (it never contracts to FMA).  Coverage, depth and every interpolated
// attribute are therefore bit-identical to the reference arithmetic; only powf differs (libm).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <stdarg.h>

#include "../../include/softrender_b200.h"

#ifndef SR_TILE_W
#define SR_TILE_W 64
#endif
#ifndef SR_TILE_H
#define SR_TILE_H 32
#endif
#define SR_TILE_PIXELS (SR_TILE_W * SR_TILE_H)
#define SR_GROUP 32          // primitives per bin entry (one warp's worth of consecutive primitives)
#define SR_MAX_NK 16         // interpolated floats per vertex (4 float4 planes)
#define SR_RECT_INVALID 0x00000001u  // packed tile rect with tx0 > tx1

// ---- order-preserving map f32 -> u32 (all non-NaN values) -----------------------------------
__host__ __device__ __forceinline__ uint32_t sr_depth_key(float z) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(z);
#else
    uint32_t u; memcpy(&u, &z, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float sr_key_depth(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// ---- nalgebra 0.12 style products: zero-initialised accumulator, k ascending, column-major ----
__host__ __device__ __forceinline__ void sr_mat_vec(const float *m, const float *v, float *out) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) acc = acc + m[k * 4 + r] * v[k];
        out[r] = acc;
    }
}
__host__ __device__ __forceinline__ void sr_mat_mat(const float *a, const float *b, float *out) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float acc = 0.0f;
            for (int k = 0; k < 4; ++k) acc = acc + a[k * 4 + i] * b[j * 4 + k];
            out[j * 4 + i] = acc;
        }
}
__device__ __forceinline__ float sr_dot4(const float *a, const float *b) {
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc = acc + a[i] * b[i];
    return acc;
}
// Correctly rounded n/det without the generic division sequence.  r = RN(1/det); after one
// Newton correction q is within one ulp of n/det, the residual n - q*det is then exact in an FMA, and
// q' = RN(q + rem*r) is the correctly rounded quotient (Markstein's theorem).  Valid for |det| in
// [2^-40, 2^40] and |n| in [2^-60, 2^60] (no over/underflow anywhere); checked against __fdiv_rn on the GPU
// by tests/test_gpu_parity.py::test_exact_division_shortcut.  A zero numerator is outside that range (the sign of the zero
// quotient would not follow IEEE for n = -0): every coverage call site tests |n| >= 2^-60 first and divides generically
// otherwise; the texel/255 call sites only ever pass n >= +0 over a positive divisor, where +0 comes out.
__device__ __forceinline__ float sr_div_exact(float n, float det, float rdet) {
    float q = n * rdet;
    float rem = fmaf(-q, det, n);
    q = fmaf(rem, rdet, q);
    rem = fmaf(-q, det, n);
    return fmaf(rem, rdet, q);
}
__device__ __forceinline__ float sr_norm4(const float *a) { return sqrtf(sr_dot4(a, a)); }
__device__ __forceinline__ void sr_normalize4(const float *a, float *out) {
    const float n = sr_norm4(a);
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = a[i] / n;
}
// compiler-rt __powisf2 for the constant exponents the shaders use
__device__ __forceinline__ float sr_powi2(float a) { return a * a; }
__device__ __forceinline__ float sr_powi5(float a) { float a2 = a * a; float a4 = a2 * a2; return a * a4; }
__device__ __forceinline__ float sr_powi64(float a) {
#pragma unroll
    for (int i = 0; i < 6; ++i) a = a * a;
    return a;
}
// f32::hypot evaluated in f64 exactly like the oracle
__device__ __forceinline__ float sr_hypot32(float x, float y) {
    const double dx = x, dy = y;
    return (float)sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}
// Interpolate (src/numeric/interpolate.rs:43-57)
__device__ __forceinline__ float sr_bary(float u, float ux, float v, float vx, float w, float wx) {
    return (ux * u + vx * v) + wx * w;
}
// contracted form for values that only feed lit shading (colour parity 1/255), never position or depth
__device__ __forceinline__ float sr_bary_fast(float u, float ux, float v, float vx, float w, float wx) { return fmaf(wx, w, fmaf(vx, v, ux * u)); }
__device__ __forceinline__ float sr_lerp(float t, float x1, float x2) { return (1.0f - t) * x1 + t * x2; }

// ---- vertex storage in HBM ----------------------------------------------------------------------
// One float4 position per vertex (its own array: setup and coverage read nothing else) plus one attribute RECORD per
// vertex: `np` = ceil(nk/4) consecutive float4.  With np = 2 (the shipped shaders: world position + normal) a record is
// exactly one 32-byte sector and moves with one 256-bit load / store.
struct SrVertexSet {
    const float4 *pos;
    const float4 *attr;
    uint64_t np;
};
__host__ __device__ __forceinline__ uint64_t sr_attr_at(uint64_t np, uint64_t vertex, uint32_t plane) { return vertex * np + plane; }
// the first two float4 of a vertex's record (np == 2: the whole record) with one 256-bit load
__device__ __forceinline__ void sr_ldg_record2(const float4 *rec, float4 &a, float4 &b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(rec));
}
__device__ __forceinline__ void sr_stg_record2(float4 *rec, const float4 &a, const float4 &b) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(rec), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x),
                 "f"(b.y), "f"(b.z), "f"(b.w)
                 : "memory");
}
// The primitives a fragment stage consumes, in the reference's canonical order
// (src/pipeline/stages/fragment.rs:268-311): first the indexed mesh primitives, then the generated ones.
struct SrPrimSource {
    const uint32_t *indices;  // segment 0: n0 primitives, NV indices each, into vs0
    SrVertexSet vs0;
    uint32_t n0;
    SrVertexSet vs1;          // segment 1: n1 primitives, vertices stored consecutively (NV per primitive)
    uint32_t n1;
    const uint32_t *seq1;     // optional: literal sequence number of each generated primitive
    const uint32_t *n1_dev;   // optional: the number of generated primitives really present, in device memory.  n1 is then an upper
                              // bound known to the host (clip_primitives without a host synchronisation: the output stream is sized
                              // for the worst case and its unused tail holds NaN positions, which every consumer skips); kernels
                              // that loop over the primitives read the true count so that they do not walk the tail
    uint32_t nplanes;
    uint32_t nk;
};
// number of primitives to walk (see n1_dev)
__device__ __forceinline__ uint32_t sr_prim_count(const SrPrimSource &s) {
    return s.n0 + (s.n1_dev != nullptr ? min(s.n1, __ldg(s.n1_dev)) : s.n1);
}

template <int NV>
__device__ __forceinline__ void sr_prim_vertices(const SrPrimSource &s, uint32_t t, const SrVertexSet *&vs, uint32_t *vi) {
    if (t < s.n0) {
        vs = &s.vs0;
#pragma unroll
        for (int k = 0; k < NV; ++k) vi[k] = __ldg(s.indices + (uint64_t)t * NV + k);
    } else {
        vs = &s.vs1;
        const uint32_t g = t - s.n0;
#pragma unroll
        for (int k = 0; k < NV; ++k) vi[k] = g * NV + k;
    }
}
// canonical primitive number reported in the winner plane
__device__ __forceinline__ uint32_t sr_prim_canonical(const SrPrimSource &s, uint32_t t, uint32_t base) {
    if (t < s.n0 || s.seq1 == nullptr) return base + t;
    return base + s.n0 + s.seq1[t - s.n0];
}

// ---- framebuffer view ---------------------------------------------------------------------------
struct SrFbView {
    float *aos;        // width*height*5 floats {r,g,b,a,depth}; may be a peer (NVLink) address
    uint32_t soa;      // texture-buffer storage (declare_texture_buffer!, src/framebuffer/texturebuffer.rs:72-110): the number of colour planes
                       // (0: not a texture buffer).  Plane k holds width*height float4 at `aos` + k*4*width*height floats (each re-usable as
                       // a texture without copying), the depths follow the last plane
    float clear1[4];   // clear colour of the second colour plane
    uint32_t u8color;  // colour attachment RGBAu8Color (src/color/predefined.rs:26): the AoS pixel is {rgba8, f32 depth} = 8 bytes instead of 20
    uint8_t *stencil;  // or null; elements of stencil_bytes (1, 2 or 4) bytes
    uint32_t stencil_bytes;
    uint32_t *winner;  // or null
    uint32_t width, height;
    uint32_t ntx, nty;       // tile grid
    uint32_t pending_clear;  // contents are the lazily-recorded clear colour, not yet in HBM
    float clear[4];
};

// ---- u8 colour targets ---------------------------------------------------------------------------------------
// The registered fragment shaders compute f32 colours; on an RGBAu8Color target a shader's result is `(c * 255.0) as u8` per
// channel (the conversion the reference's own presentation loop applies, realtime_example/src/main.rs:100-116; Rust's `as`
// truncates toward zero and saturates, NaN -> 0: exactly cvt.rzi.u32.f32 + min).
__device__ __forceinline__ uint32_t sr_as_u8(float c) { return min(__float2uint_rz(c * 255.0f), 255u); }
// Color::mul_alpha of a u8 colour (src/color/predefined.rs:82-86 -> AlphaMultiply for u8, src/color/helper.rs:36-42):
// (channel as f32 * (alpha as f32 / 255.0)) as u8, applied to the alpha channel only
__device__ __forceinline__ uint32_t sr_mul_alpha_u8(uint32_t channel, uint32_t alpha) {
    return min(__float2uint_rz((float)channel * ((float)alpha / 255.0f)), 255u);
}
// shader colour -> the four u8 channels, held as exact small floats (what the tile kernels keep in registers / shared memory);
// line_alpha: the fragment of a line, whose coverage `alpha_u8` = ColorAlpha::from_scalar(alpha) (NumCast f64 -> u8: truncation,
// so 0 or 1 -- a quirk of the reference that is kept: src/color/mod.rs:26-33, rasterization/line.rs:100) multiplies the alpha channel
__device__ __forceinline__ void sr_quantise_u8(float *c, bool line_alpha, uint32_t alpha_u8) {
    uint32_t q[4] = {sr_as_u8(c[0]), sr_as_u8(c[1]), sr_as_u8(c[2]), sr_as_u8(c[3])};
    if (line_alpha) q[3] = sr_mul_alpha_u8(q[3], alpha_u8);
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = (float)q[i];
}
__device__ __forceinline__ uint32_t sr_pack_u8(const float *q) {  // q: channel values 0..255 as floats
    return (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
}
__device__ __forceinline__ void sr_unpack_u8(uint32_t v, float *q) {
    q[0] = (float)(v & 255u); q[1] = (float)((v >> 8) & 255u); q[2] = (float)((v >> 16) & 255u); q[3] = (float)(v >> 24);
}
// one pixel of either AoS layout; colours of a u8 target travel as channel values 0..255 in floats
__device__ __forceinline__ float *sr_fb_depth_plane(const SrFbView &fb) { return fb.aos + 4ull * fb.soa * fb.width * fb.height; }  // (soa)
// a pixel of a two-plane texture buffer: o = {colour .0 (4), depth, colour .1 (4)}
__device__ __forceinline__ void sr_fb_store_pixel2(const SrFbView &fb, uint64_t index, const float *o) {
    reinterpret_cast<float4 *>(fb.aos)[index] = make_float4(o[0], o[1], o[2], o[3]);
    reinterpret_cast<float4 *>(fb.aos + 4ull * fb.width * fb.height)[index] = make_float4(o[5], o[6], o[7], o[8]);
    sr_fb_depth_plane(fb)[index] = o[4];
}
__device__ __forceinline__ void sr_fb_store_pixel(const SrFbView &fb, uint64_t index, const float *o /* r,g,b,a,depth */) {
    if (fb.soa) {
        reinterpret_cast<float4 *>(fb.aos)[index] = make_float4(o[0], o[1], o[2], o[3]);
        sr_fb_depth_plane(fb)[index] = o[4];
    } else if (fb.u8color) {
        *reinterpret_cast<uint2 *>(reinterpret_cast<unsigned char *>(fb.aos) + index * 8) = make_uint2(sr_pack_u8(o), __float_as_uint(o[4]));
    } else {
        float *dst = fb.aos + index * 5;
        dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2]; dst[3] = o[3]; dst[4] = o[4];
    }
}
__device__ __forceinline__ void sr_fb_load_pixel(const SrFbView &fb, uint64_t index, float *o) {
    if (fb.soa) {
        const float4 v = reinterpret_cast<const float4 *>(fb.aos)[index];
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
        o[4] = sr_fb_depth_plane(fb)[index];
    } else if (fb.u8color) {
        const uint2 v = *reinterpret_cast<const uint2 *>(reinterpret_cast<const unsigned char *>(fb.aos) + index * 8);
        sr_unpack_u8(v.x, o);
        o[4] = __uint_as_float(v.y);
    } else {
        const float *src = fb.aos + index * 5;
        o[0] = src[0]; o[1] = src[1]; o[2] = src[2]; o[3] = src[3]; o[4] = src[4];
    }
}
__device__ __forceinline__ float sr_fb_load_depth(const SrFbView &fb, uint64_t index) {
    if (fb.soa) return sr_fb_depth_plane(fb)[index];
    return fb.u8color ? reinterpret_cast<const float *>(reinterpret_cast<const unsigned char *>(fb.aos) + index * 8)[1] : fb.aos[index * 5 + 4];
}

// packed tile rectangle of a primitive: tx0 | ty0<<8 | tx1<<16 | ty1<<24
__device__ __forceinline__ uint32_t sr_pack_rect(uint32_t tx0, uint32_t ty0, uint32_t tx1, uint32_t ty1) {
    return tx0 | (ty0 << 8) | (tx1 << 16) | (ty1 << 24);
}
__device__ __forceinline__ bool sr_rect_hits(uint32_t r, uint32_t tx, uint32_t ty) {
    const uint32_t tx0 = r & 255u, ty0 = (r >> 8) & 255u, tx1 = (r >> 16) & 255u, ty1 = r >> 24;
    return tx0 <= tx && tx <= tx1 && ty0 <= ty && ty <= ty1;
}

// clamp_as_int! of rasterization/triangle.rs:66-72
__device__ __forceinline__ uint32_t sr_clamp_as_int(float value, uint32_t lo, uint32_t hi) {
    if (value < (float)lo) return lo;
    if (value > (float)hi) return hi;
    return __float2uint_rz(value);
}
