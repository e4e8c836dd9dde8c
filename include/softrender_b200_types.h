/*
 * softrender_b200_types.h -- plain-old-data types shared by the C ABI
 * (softrender_b200.h), the C++ builder mirror (softrender_b200.hpp) and the
 * test oracle (oracle/sr_oracle.h).
 *
 * Everything here is a C restatement of a Rust type or enum of the reference
 * (novacrazy/rust-softrender); each item cites the reference file:line it
 * stands for.  All arithmetic on the draw path is f32.
 */
#ifndef SOFTRENDER_B200_TYPES_H
#define SOFTRENDER_B200_TYPES_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (the reference panics; the C ABI returns codes) ------- */
enum sr_status {
    SR_OK = 0,
    SR_ERR_INVALID_ARGUMENT = 1, /* contract violation: reference asserts (src/pipeline/mod.rs:129-130,148) */
    SR_ERR_INVALID_STATE = 2,    /* stage called out of order (the Rust types make this unrepresentable) */
    SR_ERR_CUDA = 3,             /* CUDA runtime failure; text in sr_last_error() */
    SR_ERR_OUT_OF_MEMORY = 4,
    SR_ERR_INVALID_PIXEL_COORDINATE = 5, /* RenderError::InvalidPixelCoordinate, src/error.rs:9 */
    SR_ERR_UNSUPPORTED = 6
};

/* ---- primitives: zero-sized marker types Point/Line/Triangle
 *      (src/primitive.rs:76,102,133) ------------------------------------- */
enum sr_primitive {
    SR_POINT = 1,    /* Primitive::num_vertices() == 1 */
    SR_LINE = 2,     /* 2 */
    SR_TRIANGLE = 3  /* 3 */
};

/* ---- FaceWinding (src/geometry/winding.rs:11) + None -------------------- */
enum sr_winding {
    SR_CULL_NONE = 0,
    SR_CLOCKWISE = 1,
    SR_COUNTER_CLOCKWISE = 2
};

/* ---- StencilTest / StencilOp (src/stencil.rs:90-124,129-159) ------------ */
enum sr_stencil_test {
    SR_STENCIL_ALWAYS = 0,
    SR_STENCIL_NEVER = 1,
    SR_STENCIL_LESS_THAN = 2,        /* mask <  value */
    SR_STENCIL_GREATER_THAN = 3,     /* mask >  value */
    SR_STENCIL_LESS_THAN_EQ = 4,     /* mask <= value */
    SR_STENCIL_GREATER_THAN_EQ = 5,  /* mask >= value */
    SR_STENCIL_EQUAL = 6,
    SR_STENCIL_NOT_EQUAL = 7
};

enum sr_stencil_op {
    SR_STENCIL_KEEP = 0,
    SR_STENCIL_INVERT = 1,
    SR_STENCIL_ZERO = 2,
    SR_STENCIL_REPLACE = 3,
    SR_STENCIL_INCREMENT_WRAP = 4,
    SR_STENCIL_DECREMENT_WRAP = 5,
    SR_STENCIL_INCREMENT_SAT = 6,
    SR_STENCIL_DECREMENT_SAT = 7
};

/* ---- framebuffer formats: RenderBuffer<ColorDepth[Stencil]Attachments<RGBAf32Color,f32[,u8]>>
 *      (src/framebuffer/renderbuffer/mod.rs:16-30, attachments/predefined.rs:11-26) */
enum sr_fb_format {
    SR_FB_RGBAF32_DF32 = 0,     /* stencil type (): 20 B/pixel AoS {r,g,b,a,depth} */
    SR_FB_RGBAF32_DF32_S8 = 1,  /* stencil type u8: colour+depth AoS as above, stencil in its own u8 plane */
    SR_FB_RGBAF32_DF32_S16 = 2, /* stencil type u16 (the Stencil trait covers every integer width, src/stencil.rs:9-60) */
    SR_FB_RGBAF32_DF32_S32 = 3, /* stencil type u32 */
    SR_FB_RGBAU8_DF32 = 4,      /* colour RGBAu8Color (src/color/predefined.rs:26): 8 B/pixel AoS {r,g,b,a as u8, f32 depth}.  A registered
                                 * shader's f32 colour c is stored as `(c * 255.0) as u8` per channel; Blend = () only; lines scale the
                                 * alpha channel with the integer rule of src/color/helper.rs:36-42 */
    SR_FB_RGBAU8_DF32_S8 = 5,   /* the same with a u8 stencil plane */
    SR_FB_TEXTURE_RGBAF32_DF32 = 6,    /* RGBAf32TextureBuffer (src/framebuffer/texturebuffer.rs:200-210, declare_texture_buffer! :72-198): the colour
                                        * attachment is its own plane of width*height Vector4<f32> ("re-used as textures without copying", :63-66) and
                                        * the depths another: 16 + 4 B/pixel, structure of arrays */
    SR_FB_TEXTURE_RGBAF32_DF32_S8 = 7, /* the same with a u8 stencil plane */
    SR_FB_TEXTURE_2xRGBAF32_DF32 = 8   /* a texture buffer declared with TWO colour planes (declare_texture_buffer! { pub a: RGBAf32Color, pub b:
                                        * RGBAf32Color }): PixelBuffer::Color is the tuple (texturebuffer.rs:129-133), written by fragment shaders
                                        * with two outputs (SR_FS_SUZANNE_GBUFFER); Blend = (), no stencil; 16 + 16 + 4 B/pixel in three planes */
};

/* ---- Viewport (src/geometry/clipvertex.rs:40-48) ------------------------ */
typedef struct sr_viewport {
    float x, y, width, height, near_, far_;
} sr_viewport;

/* ---- registered shaders --------------------------------------------------
 * The reference takes Rust closures; here a closed set of device functions
 * mirrors the closures the reference ships.
 * Vin = floats per input vertex (position.xyz first), K = interpolated floats.
 */
enum sr_vertex_shader {
    /* test shader: Vin = {x,y,z, w, k0..k(n-1)}; clip = (x,y,z,w); K = k  (any nk = Vin-4) */
    SR_VS_PASSTHROUGH = 0,
    /* examples/suzanne.rs:123-141: Vin = pos3+normal3; K = {world_pos4, normal4} */
    SR_VS_SUZANNE = 1,
    /* full_example/src/shaders.rs:8-31: Vin = pos3+normal3+uv2; K = {world_pos4, normal4, uv2} */
    SR_VS_FULL_EXAMPLE = 2
};

enum sr_fragment_shader {
    SR_FS_FLAT = 0,          /* test shader: colour = K[0..4) */
    SR_FS_SUZANNE = 1,       /* examples/suzanne.rs:147-183 */
    SR_FS_FULL_EXAMPLE = 2,  /* full_example/src/shaders.rs:108-162 (4-light Blinn-Phong + ACES + gamma) */
    SR_FS_FULL_EXAMPLE_TEXTURED = 3, /* same, material colour * bilinear/clamp texture sample (full_example/src/texture.rs:47-84) */
    SR_FS_GREEN = 4,         /* full_example/src/shaders.rs:102 */
    SR_FS_DISCARD_CHECKER = 5, /* test shader: Fragment::Discard on odd (floor(x)+floor(y)), else colour = K[0..4) (fragment.rs:61-66) */
    SR_FS_TEXTURE_UNLIT = 6  /* second-pass shader of a render-to-texture chain: colour = texture(bound texture, uv = K[0..2), filter, edge)
                              * -- the call of src/texture.rs:14-18 and nothing else */,
    SR_FS_SUZANNE_GBUFFER = 7 /* a fragment shader with TWO colour outputs, for a texture buffer with two colour planes (the tuple colour of
                               * declare_texture_buffer!, src/framebuffer/texturebuffer.rs:129-147): .0 = the suzanne Blinn-Phong colour
                               * (examples/suzanne.rs:147-183), .1 = the interpolated world-space normal K[4..8) -- a G-buffer pass */
};

/* texture sampling state (src/texture.rs:21-45); the arithmetic is the sampler the reference ships,
 * full_example/src/texture.rs:25-84 (TextureRead::sample itself is unimplemented!() there, src/texture.rs:49-52) */
enum sr_texture_filter {
    SR_FILTER_NEAREST = 0,   /* Filter::Nearest: texel (round(u (w-1)), round(v (h-1))), full_example/src/texture.rs:52-57 */
    SR_FILTER_BILINEAR = 1   /* Filter::Bilinear, full_example/src/texture.rs:58-82 (what the shipped scene uses; the pipeline default) */
};
enum sr_texture_edge {
    SR_EDGE_CLAMP = 0,       /* Edge::Clamp  = GL_CLAMP_TO_EDGE: (u.min(1).max(0), v.min(1).max(0)), full_example/src/texture.rs:28 */
    SR_EDGE_WRAP = 1,        /* Edge::Wrap   = GL_REPEAT: (u.fract(), v.fract()), full_example/src/texture.rs:29 (fract of a negative is negative) */
    SR_EDGE_BORDER = 2       /* Edge::Border(C) = GL_CLAMP_TO_BORDER (src/texture.rs:43-44): a coordinate outside [0,1]^2 (or NaN)
                              * returns the border colour unchanged, inside it samples as Clamp */
};

enum sr_geometry_shader {
    SR_GS_CLIP = 0,            /* GeometryShader::clip_primitives, src/pipeline/stages/geometry.rs:261-336 */
    SR_GS_FACE_NORMALS = 1,    /* full_example/src/shaders.rs:63-89 */
    SR_GS_VERTEX_NORMALS = 2,  /* full_example/src/shaders.rs:35-61 */
    SR_GS_CLIP_SH = 3          /* opt-in: a CORRECT clipper (Sutherland-Hodgman against the same six planes, src/geometry/clip.rs:37-42)
                                * for triangles -- the fix the reference asks for (src/lib.rs "Glaring Problems: Clipping");
                                * lines and points as SR_GS_CLIP */
};

enum sr_blend {
    SR_BLEND_REPLACE = 0,     /* Blend for (): src/color/blend.rs:28-31 */
    SR_BLEND_ALPHA_OVER = 1,  /* full_example/src/color.rs:5-17 */
    SR_BLEND_ADDITIVE = 2     /* a user blend function, GenericBlend::new(|a, b| a + b) (src/color/blend.rs:57-76): the worked example of
                               * how a third blend is registered -- INTEGRATION.md "Adding a blend function" */
};

/* ---- global uniforms ------------------------------------------------------
 * One POD superset of the GlobalUniforms structs of examples/suzanne.rs:50-59
 * and full_example/src/uniforms.rs:7-17 plus the values the example closures
 * capture.  Matrices are COLUMN-MAJOR (nalgebra storage): m[c*4 + r].
 */
#define SR_MAX_LIGHTS 8

typedef struct sr_light {      /* full_example/src/light.rs:6-10 */
    float color[4];
    float position[3];
    float intensity;
} sr_light;

typedef struct sr_uniforms {
    float camera[4];
    float model[16];
    float mit[16];             /* model inverse transpose */
    float view[16];
    float projection[16];
    /* captured by the suzanne fragment closure (examples/suzanne.rs:110-114) */
    float sz_light[4];         /* Point3(5,5,5).to_homogeneous() */
    float sz_color[4];         /* (0.1^2.2, 0.5^2.2, 0.1^2.2, 1) */
    float sz_intensity;        /* 4.0 */
    uint32_t nlights;          /* full_example: lights.len() (<= SR_MAX_LIGHTS) */
    uint32_t reserved0, reserved1;
    sr_light lights[SR_MAX_LIGHTS];
} sr_uniforms;

/* ---- per-stage device timings of the last draw (CUDA events, ms) -------- */
typedef struct sr_stage_times {
    float vertex_ms;
    float geometry_ms;
    float bin_ms;     /* per-tile lists of the ordered path (triangles with blend/stencil/discard, lines, points) */
    float vis_init_ms;/* opaque path: visibility-buffer initialisation (k_vis_init) */
    float micro_ms;   /* opaque path: per-triangle setup / small-triangle rasterisation (k_micro) */
    float raster_ms;  /* tile kernels: large triangles, resolve (shading) and the single write-back */
    float total_ms;
} sr_stage_times;

/* one entry of the shader registry (sr_registry_entry) */
typedef struct sr_shader_info {
    uint32_t id;
    uint32_t vin_floats;   /* vertex shaders: floats per input vertex (0 = any: pass-through) */
    uint32_t nk;           /* vertex shaders: interpolated floats produced; fragment shaders: interpolated floats read */
    uint32_t discards;     /* fragment shaders: may return Fragment::Discard (such draws take the ordered path) */
    uint32_t needs_texture;
    char name[40];
    char reference[72];    /* the closure of the reference this device function mirrors (file:line) */
} sr_shader_info;
enum sr_registry_kind { SR_REGISTRY_VERTEX = 0, SR_REGISTRY_GEOMETRY = 1, SR_REGISTRY_FRAGMENT = 2, SR_REGISTRY_BLEND = 3 };

#ifdef __cplusplus
}
#endif
#endif /* SOFTRENDER_B200_TYPES_H */
