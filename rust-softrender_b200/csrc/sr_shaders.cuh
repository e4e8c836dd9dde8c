// sr_shaders.cuh -- the registered set of device-function shaders that replaces the reference's
// Rust shader closures.  Each mirrors one closure the reference ships.
#pragma once

#include "sr_common.cuh"

// ---- vertex stage constants: uniforms + products that are identical for every vertex -------------
// `projection * view * world` is left-associative in Rust, so the reference recomputes the 4x4
// product per vertex (examples/suzanne.rs:134, full_example/src/shaders.rs:20); it is the same
// value every time, so it is computed once on the host with the same f32 operation order.
struct SrVsConst {
    sr_uniforms u;
    float pv[16];   // projection * view
    float pvm[16];  // (projection * view) * model
    float vpm[16];  // viewport matrix of ClipVertex::normalize (clipvertex.rs:107-112)
    int normalize;  // run_to_fragment: apply ClipVertex::normalize after the shader
};

#define SR_TEX_RGBA8 0u  // image texture: u8 texels / 255, then decode_gamma (full_example/src/texture.rs:33-40,84)
#define SR_TEX_F32 1u    // a framebuffer's colour bound in place (texturebuffer.rs:12-58): f32 RGBA every tex_stride floats, linear
struct SrFsConst {
    sr_uniforms u;
    const uint8_t *tex;  // RGBA8 texels, or the f32 AoS pixels of the source framebuffer
    uint32_t tex_w, tex_h;
    uint32_t tex_kind, tex_stride;    // SR_TEX_*; floats per texel of an SR_TEX_F32 source
    uint32_t tex_filter, tex_edge;    // sr_texture_filter, sr_texture_edge (src/texture.rs:21-45)
    float tex_border[4];              // Edge::Border(C)
};

template <int VS> struct SrVsInfo;  // VIN = input floats (0 = runtime), NK = interpolated floats (0 = runtime)
template <> struct SrVsInfo<SR_VS_PASSTHROUGH> { static constexpr int VIN = 0, NK = 0; };
template <> struct SrVsInfo<SR_VS_SUZANNE> { static constexpr int VIN = 6, NK = 8; };
template <> struct SrVsInfo<SR_VS_FULL_EXAMPLE> { static constexpr int VIN = 8, NK = 10; };

// ClipVertex::normalize (src/geometry/clipvertex.rs:89-127)
__device__ __forceinline__ void sr_normalize_vertex(const float *vpm, float *p) {
    const float w = p[3];
    float ndc[4] = {p[0] / w, p[1] / w, p[2] / w, 1.0f};
    float screen[4];
    sr_mat_vec(vpm, ndc, screen);
    p[0] = screen[0];
    p[1] = screen[1];
    p[2] = screen[2];
    p[3] = 1.0f / w;
}

// examples/suzanne.rs:123-141 and full_example/src/shaders.rs:8-31; `in` = {pos3, normal3[, uv2]},
// `out` = {clip4, world4, normal4[, uv2]}
template <int VS>
__device__ __forceinline__ void sr_vertex_shader(const SrVsConst &c, const float *in, float *out) {
    float pos_h[4] = {in[0], in[1], in[2], 1.0f};  // Point3::to_homogeneous
    float nrm_h[4] = {in[3], in[4], in[5], 0.0f};  // Vector3::to_homogeneous
    float n[4];
    sr_mat_vec(c.u.model, pos_h, out + 4);          // world_position = model * position
    sr_mat_vec(c.u.mit, nrm_h, n);
    sr_normalize4(n, out + 8);                      // (mit * normal).normalize()
    if (VS == SR_VS_SUZANNE) {
        sr_mat_vec(c.pv, out + 4, out);             // (projection * view) * world_position
    } else {
        sr_mat_vec(c.pvm, pos_h, out);              // mvp * position
        out[12] = in[6];
        out[13] = in[7];
    }
}

// ---- fragment shaders ---------------------------------------------------------------------------------
// NK = interpolated floats the shader reads; DISCARDS = may return Fragment::Discard; LIT = the colour goes through
// transcendental shading arithmetic (parity bar 1/255), so the K interpolation feeding it may use FMA contraction; the
// pass-through test shaders need K bit-exact.
template <int FS> struct SrFsInfo;
template <> struct SrFsInfo<SR_FS_FLAT> { static constexpr int NK = 4; static constexpr bool DISCARDS = false, LIT = false; };
template <> struct SrFsInfo<SR_FS_SUZANNE> { static constexpr int NK = 8; static constexpr bool DISCARDS = false, LIT = true; };
template <> struct SrFsInfo<SR_FS_FULL_EXAMPLE> { static constexpr int NK = 8; static constexpr bool DISCARDS = false, LIT = true; };
template <> struct SrFsInfo<SR_FS_FULL_EXAMPLE_TEXTURED> { static constexpr int NK = 10; static constexpr bool DISCARDS = false, LIT = true; };
template <> struct SrFsInfo<SR_FS_GREEN> { static constexpr int NK = 0; static constexpr bool DISCARDS = false, LIT = false; };
template <> struct SrFsInfo<SR_FS_DISCARD_CHECKER> { static constexpr int NK = 4; static constexpr bool DISCARDS = true, LIT = false; };
template <> struct SrFsInfo<SR_FS_TEXTURE_UNLIT> { static constexpr int NK = 2; static constexpr bool DISCARDS = false, LIT = false; };
// (LIT = false: the second output IS the interpolated normal, so the attributes are interpolated exactly, without FMA contraction)
template <> struct SrFsInfo<SR_FS_SUZANNE_GBUFFER> { static constexpr int NK = 8; static constexpr bool DISCARDS = false, LIT = false; };
// colour outputs of a fragment shader: 2 = it returns a tuple of two colours (out[0..4) and out[5..9); out[4] carries the depth)
template <int FS> struct SrFsOutputs { static constexpr int N = 1; };
template <> struct SrFsOutputs<SR_FS_SUZANNE_GBUFFER> { static constexpr int N = 2; };

// Fragment-shader arithmetic is NOT on the bit-exact path (colour parity is 1/255 per channel, and the reference's
// powf is libm's): normalisation uses rsqrtf and powers use exp2(y*log2(x)) on the SFU unless SR_FS_EXACT is set.
#ifdef SR_FS_EXACT
__device__ __forceinline__ float sr_fs_pow(float x, float y) { return powf(x, y); }
__device__ __forceinline__ float sr_fs_dot4(const float *a, const float *b) { return sr_dot4(a, b); }
__device__ __forceinline__ void sr_fs_normalize4(const float *a, float *out) { sr_normalize4(a, out); }
__device__ __forceinline__ float sr_fs_div(float a, float b) { return a / b; }
#else
__device__ __forceinline__ float sr_fs_pow(float x, float y) { return __powf(x, y); }
__device__ __forceinline__ float sr_fs_dot4(const float *a, const float *b) { return fmaf(a[3], b[3], fmaf(a[2], b[2], fmaf(a[1], b[1], a[0] * b[0]))); }
__device__ __forceinline__ void sr_fs_normalize4(const float *a, float *out) {
    const float inv = rsqrtf(sr_fs_dot4(a, a));
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = a[i] * inv;
}
__device__ __forceinline__ float sr_fs_div(float a, float b) { return __fdividef(a, b); }
#endif
__device__ __forceinline__ float sr_saturate(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }
__device__ __forceinline__ float sr_fresnel_schlick(float cos_theta, float ior) {
    const float f0 = sr_powi2((1.0f - ior) / (1.0f + ior));
    return f0 + (1.0f - f0) * sr_powi5(1.0f - cos_theta);
}

// full_example/src/texture.rs:47-84 (Bilinear, Clamp); the x+1 / y+1 neighbour is clamped to the
// last texel where the reference would index out of bounds (documented deviation).
__device__ __forceinline__ void sr_texture_bilinear_clamp(const SrFsConst &c, float u, float v, float *out) {
    u = fmaxf(fminf(u, 1.0f), 0.0f);
    v = fmaxf(fminf(v, 1.0f), 0.0f);
    const float uu = (u * (float)(c.tex_w - 1)) + 0.5f;
    const float vv = (v * (float)(c.tex_h - 1)) + 0.5f;
    const uint32_t x = (uint32_t)floorf(uu), y = (uint32_t)floorf(vv);
    const float u_ratio = uu - (float)x, v_ratio = vv - (float)y;
    const float u_opp = 1.0f - u_ratio, v_opp = 1.0f - v_ratio;
    const uint32_t x0 = x < c.tex_w ? x : c.tex_w - 1, y0 = y < c.tex_h ? y : c.tex_h - 1;
    const uint32_t x1 = x + 1 < c.tex_w ? x + 1 : c.tex_w - 1, y1 = y + 1 < c.tex_h ? y + 1 : c.tex_h - 1;
    const uchar4 t00 = __ldg((const uchar4 *)c.tex + (size_t)y0 * c.tex_w + x0);
    const uchar4 t10 = __ldg((const uchar4 *)c.tex + (size_t)y0 * c.tex_w + x1);
    const uchar4 t01 = __ldg((const uchar4 *)c.tex + (size_t)y1 * c.tex_w + x0);
    const uchar4 t11 = __ldg((const uchar4 *)c.tex + (size_t)y1 * c.tex_w + x1);
    // texel / 255 (texture.rs:66-69): correctly rounded quotients through the shared-divisor shortcut (sr_div_exact with
    // r = RN(1/255); 0 / 255 comes out as +0) instead of sixteen IEEE division sequences -- same bits
    const float r255 = 0x1.010102p-8f;  // RN(1/255)
    auto unorm = [&](const uchar4 &t, float *a) {
        a[0] = sr_div_exact((float)t.x, 255.0f, r255); a[1] = sr_div_exact((float)t.y, 255.0f, r255);
        a[2] = sr_div_exact((float)t.z, 255.0f, r255); a[3] = sr_div_exact((float)t.w, 255.0f, r255);
    };
    float a00[4], a10[4], a01[4], a11[4];
    unorm(t00, a00); unorm(t10, a10); unorm(t01, a01); unorm(t11, a11);
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        const float val = (a00[ch] * u_opp + a10[ch] * u_ratio) * v_opp + (a01[ch] * u_opp + a11[ch] * u_ratio) * v_ratio;
        out[ch] = ch < 3 ? sr_fs_pow(val, 2.2f) : val;  // decode_gamma (full_example/src/color.rs:48-55)
    }
}

// texture(t, coord, filter, edge) (src/texture.rs:14-18) for any Filter / Edge / texel format: the arithmetic of
// full_example/src/texture.rs:25-84 with Rust's saturating `as u32` (negative and NaN -> 0; CUDA's float->uint conversion
// saturates the same way) and every texel index clamped to the last row / column where the reference would index out
// of bounds.
__device__ __forceinline__ float4 sr_texture_sample_body(const SrFsConst &c, float u, float v) {
    if (c.tex_edge == SR_EDGE_WRAP) {
        u = u - truncf(u);  // f32::fract
        v = v - truncf(v);
    } else {
        if (c.tex_edge == SR_EDGE_BORDER && !(u >= 0.0f && u <= 1.0f && v >= 0.0f && v <= 1.0f)) {
            return make_float4(c.tex_border[0], c.tex_border[1], c.tex_border[2], c.tex_border[3]);
        }
        u = fmaxf(fminf(u, 1.0f), 0.0f);
        v = fmaxf(fminf(v, 1.0f), 0.0f);
    }
    const uint32_t lastx = c.tex_w - 1, lasty = c.tex_h - 1;
    const bool f32 = c.tex_kind == SR_TEX_F32;
    const float r255 = 0x1.010102p-8f;  // RN(1/255) (what __frcp_rn(255.0f) returns; as a constant it costs nothing on the f32 path)
    const uint32_t texel_bytes = f32 ? c.tex_stride * 4u : 4u;
    auto texel = [&](uint32_t x, uint32_t y, float *a) {
        x = min(x, lastx); y = min(y, lasty);
        const uint8_t *at = c.tex + (size_t)y * (c.tex_w * texel_bytes) + x * texel_bytes;  // (a row is below 4 GB)
        if (f32) {
            if (c.tex_stride == 4u) {  // the colour plane of a texture buffer (what the reference samples): one 16-byte load per texel
                const float4 t = __ldg(reinterpret_cast<const float4 *>(at));
                a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w;
            } else {
                const float *t = reinterpret_cast<const float *>(at);
                a[0] = __ldg(t); a[1] = __ldg(t + 1); a[2] = __ldg(t + 2); a[3] = __ldg(t + 3);
            }
        } else {
            const uchar4 t = __ldg(reinterpret_cast<const uchar4 *>(at));
            a[0] = sr_div_exact((float)t.x, 255.0f, r255); a[1] = sr_div_exact((float)t.y, 255.0f, r255);
            a[2] = sr_div_exact((float)t.z, 255.0f, r255); a[3] = sr_div_exact((float)t.w, 255.0f, r255);
        }
    };
    float val[4];
    if (c.tex_filter == SR_FILTER_NEAREST) {
        texel((uint32_t)roundf(u * (float)lastx), (uint32_t)roundf(v * (float)lasty), val);
    } else {
        const float uu = (u * (float)lastx) + 0.5f, vv = (v * (float)lasty) + 0.5f;
        const uint32_t x = (uint32_t)floorf(uu), y = (uint32_t)floorf(vv);
        const float u_ratio = uu - (float)x, v_ratio = vv - (float)y;
        const float u_opp = 1.0f - u_ratio, v_opp = 1.0f - v_ratio;
        const uint32_t x1 = x == 0xFFFFFFFFu ? x : x + 1, y1 = y == 0xFFFFFFFFu ? y : y + 1;
        float a00[4], a10[4], a01[4], a11[4];
        texel(x, y, a00); texel(x1, y, a10); texel(x, y1, a01); texel(x1, y1, a11);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
            val[ch] = (a00[ch] * u_opp + a10[ch] * u_ratio) * v_opp + (a01[ch] * u_opp + a11[ch] * u_ratio) * v_ratio;
    }
    if (!f32) {  // decode_gamma (full_example/src/color.rs:48-55): image textures only
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) val[ch] = sr_fs_pow(val[ch], 2.2f);
    }
    return make_float4(val[0], val[1], val[2], val[3]);
}
// Out of line for the lit shaders: the shipped scene's Bilinear + Clamp + RGBA8 case keeps its own straight-line body above and
// the kernels their register budget; the unlit second-pass shader (nothing else to keep live) inlines the body.
__device__ __noinline__ float4 sr_texture_sample_general(const SrFsConst &c, float u, float v) { return sr_texture_sample_body(c, u, v); }
__device__ __forceinline__ void sr_texture_sample(const SrFsConst &c, float u, float v, float *out) {
    if (c.tex_kind == SR_TEX_RGBA8 && c.tex_filter == SR_FILTER_BILINEAR && c.tex_edge == SR_EDGE_CLAMP) sr_texture_bilinear_clamp(c, u, v, out);
    else {
        const float4 t = sr_texture_sample_general(c, u, v);
        out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
    }
}

// `sv` = interpolated ScreenVertex: position[4] then K.  Returns false for Fragment::Discard.
template <int FS>
__device__ __forceinline__ bool sr_fragment_shader(const SrFsConst &c, const float *sv, float *out) {
    const float *K = sv + 4;
    if (FS == SR_FS_FLAT) {
#pragma unroll
        for (int i = 0; i < 4; ++i) out[i] = K[i];
        return true;
    } else if (FS == SR_FS_GREEN) {  // full_example/src/shaders.rs:102
        out[0] = 0.0f; out[1] = 1.0f; out[2] = 0.0f; out[3] = 1.0f;
        return true;
    } else if (FS == SR_FS_TEXTURE_UNLIT) {  // texture(t, uv, filter, edge), src/texture.rs:14-18
        if (c.tex != nullptr) {
            const float4 t = sr_texture_sample_body(c, K[0], K[1]);
            out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
        } else { out[0] = 0.0f; out[1] = 0.0f; out[2] = 0.0f; out[3] = 0.0f; }
        return true;
    } else if (FS == SR_FS_DISCARD_CHECKER) {
        const int xi = (int)floorf(sv[0]), yi = (int)floorf(sv[1]);
        if ((xi + yi) & 1) return false;
#pragma unroll
        for (int i = 0; i < 4; ++i) out[i] = K[i];
        return true;
    } else if (FS == SR_FS_SUZANNE || FS == SR_FS_SUZANNE_GBUFFER) {
        // examples/suzanne.rs:147-183 (the G-buffer variant returns the interpolated normal as its second colour)
        if (FS == SR_FS_SUZANNE_GBUFFER) {
#pragma unroll
            for (int i = 0; i < 4; ++i) out[5 + i] = K[4 + i];
        }
        const float *position = K, *normal = K + 4;
        float d[4], view_dir[4], light_dir[4], h[4], halfway[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = c.u.camera[i] - position[i];
        sr_fs_normalize4(d, view_dir);
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = c.u.sz_light[i] - position[i];
        sr_fs_normalize4(d, light_dir);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = light_dir[i] + view_dir[i];
        sr_fs_normalize4(h, halfway);
        const float NdotL = fmaxf(fminf(sr_fs_dot4(light_dir, normal), 1.0f), 0.0f);
        const float NdotH = fmaxf(fminf(sr_fs_dot4(normal, halfway), 1.0f), 0.0f);
        const float VdotH = fmaxf(fminf(sr_fs_dot4(view_dir, halfway), 1.0f), 0.0f);
        const float f = sr_fresnel_schlick(VdotH, 1.45f);
        const float diffuse = NdotL * (1.0f - f);
        const float specular = f * sr_fs_pow(NdotH, 32.0f * 2.0f);
        const float inv_gamma = 1.0f / 2.2f;
#pragma unroll
        for (int i = 0; i < 3; ++i) out[i] = sr_fs_pow(c.u.sz_intensity * (specular + (diffuse * c.u.sz_color[i])), inv_gamma);
        out[3] = 1.0f;
        return true;
    } else {
        // full_example/src/shaders.rs:108-162
        const float *position = K, *normal = K + 4;
        float d[4], view_dir[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = c.u.camera[i] - position[i];
        sr_fs_normalize4(d, view_dir);
        const float m = sr_fs_pow(0.25f, 2.2f);
        float material[3] = {m, m, m};
        if (FS == SR_FS_FULL_EXAMPLE_TEXTURED) {
            if (c.tex != nullptr) {
                float t[4];
                sr_texture_sample(c, K[8], K[9], t);
#pragma unroll
                for (int i = 0; i < 3; ++i) material[i] = material[i] * t[i];
            }
        }
        const float albedo = 0.7f;
        float color[3] = {0.0f, 0.0f, 0.0f};
        const uint32_t nl = c.u.nlights < SR_MAX_LIGHTS ? c.u.nlights : SR_MAX_LIGHTS;
        for (uint32_t l = 0; l < nl; ++l) {
            const sr_light &light = c.u.lights[l];
            const float lp[4] = {light.position[0], light.position[1], light.position[2], 1.0f};
            float ld[4], light_dir[4], h[4], halfway[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) ld[i] = lp[i] - position[i];
            const float light_distance = sr_norm4(ld);
            sr_fs_normalize4(ld, light_dir);
#pragma unroll
            for (int i = 0; i < 4; ++i) h[i] = light_dir[i] + view_dir[i];
            sr_fs_normalize4(h, halfway);
            const float intensity = sr_fs_div(light.intensity, sr_powi2(light_distance));
            const float NdotL = sr_saturate(sr_fs_dot4(light_dir, normal));
            const float NdotH = sr_saturate(sr_fs_dot4(normal, halfway));
            const float VdotH = sr_saturate(sr_fs_dot4(view_dir, halfway));
            const float f = sr_fresnel_schlick(VdotH, 1.45f);
            const float diffuse = (1.0f - f) * NdotL;
            const float specular = f * sr_powi64(NdotH);
#pragma unroll
            for (int i = 0; i < 3; ++i)
                color[i] = color[i] + intensity * light.color[i] * (specular + (diffuse * albedo * material[i]));
        }
        const float inv_gamma = 1.0f / 2.2f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float x = color[i];  // aces_filmic_tonemap_component (full_example/src/color.rs:20-28)
            const float tm = sr_fs_div(x * (2.51f * x + 0.03f), x * (2.43f * x + 0.59f) + 0.14f);
            out[i] = sr_fs_pow(tm, inv_gamma);
        }
        out[3] = 1.0f;
        return true;
    }
}

// Blend (src/color/blend.rs:28-31 for `()`; full_example/src/color.rs:5-17 for alpha-over)
__device__ __forceinline__ void sr_blend(uint32_t mode, const float *a, const float *b, float *out) {
    if (mode == SR_BLEND_ALPHA_OVER) {
        const float a1 = 1.0f - a[3];
        float r[4];
#pragma unroll
        for (int i = 0; i < 3; ++i) r[i] = (a[i] * a[3] + b[i] * b[3] * a1) / (a[3] + b[3] * a1);
        r[3] = a[3] + b[3] * (1.0f - a[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) out[i] = r[i];
    } else if (mode == SR_BLEND_ADDITIVE) {  // GenericBlend::new(|a, b| a + b), component-wise Vector4 addition
#pragma unroll
        for (int i = 0; i < 4; ++i) out[i] = a[i] + b[i];
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) out[i] = a[i];
    }
}

// StencilTest::test / StencilOp::op (src/stencil.rs:112-123,147-158) on u8
// value / mask are u8, u16 or u32 stencil values held in 32 bits; smax = the type's MAX (wrapping and saturation bounds)
__device__ __forceinline__ bool sr_stencil_test_fn(uint32_t test, uint32_t value, uint32_t mask) {
    switch (test) {
        case SR_STENCIL_ALWAYS: return true;
        case SR_STENCIL_NEVER: return false;
        case SR_STENCIL_LESS_THAN: return mask < value;
        case SR_STENCIL_LESS_THAN_EQ: return mask <= value;
        case SR_STENCIL_GREATER_THAN: return mask > value;
        case SR_STENCIL_GREATER_THAN_EQ: return mask >= value;
        case SR_STENCIL_EQUAL: return mask == value;
        default: return mask != value;
    }
}
__device__ __forceinline__ uint32_t sr_stencil_op_fn(uint32_t op, uint32_t value, uint32_t mask, uint32_t smax) {
    switch (op) {
        case SR_STENCIL_KEEP: return value;
        case SR_STENCIL_INVERT: return ~value & smax;
        case SR_STENCIL_ZERO: return 0;
        case SR_STENCIL_REPLACE: return mask;
        case SR_STENCIL_INCREMENT_WRAP: return (value + 1u) & smax;
        case SR_STENCIL_DECREMENT_WRAP: return (value - 1u) & smax;
        case SR_STENCIL_INCREMENT_SAT: return value == smax ? smax : value + 1u;
        default: return value == 0 ? 0u : value - 1u;
    }
}
__device__ __forceinline__ uint32_t sr_stencil_load(const uint8_t *base, uint32_t bytes, uint64_t i) {
    return bytes == 1 ? (uint32_t)base[i] : bytes == 2 ? (uint32_t)reinterpret_cast<const uint16_t *>(base)[i] : reinterpret_cast<const uint32_t *>(base)[i];
}
__device__ __forceinline__ void sr_stencil_store(uint8_t *base, uint32_t bytes, uint64_t i, uint32_t v) {
    if (bytes == 1) base[i] = (uint8_t)v;
    else if (bytes == 2) reinterpret_cast<uint16_t *>(base)[i] = (uint16_t)v;
    else reinterpret_cast<uint32_t *>(base)[i] = v;
}
