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
    uint32_t nplanes;
    uint32_t nk;
};

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
    uint8_t *stencil;  // or null; elements of stencil_bytes (1, 2 or 4) bytes
    uint32_t stencil_bytes;
    uint32_t *winner;  // or null
    uint32_t width, height;
    uint32_t ntx, nty;       // tile grid
    uint32_t pending_clear;  // contents are the lazily-recorded clear colour, not yet in HBM
    float clear[4];
};

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
