/*
 * sr_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the draw hot path of novacrazy/rust-softrender, used as
 * the parity checker for the CUDA implementation and as the "restated
 * reference CPU path" timed by bench.py (cpu_baseline / --impl reference).
 * Nothing in the product library (rust-softrender_b200/csrc) links, loads or
 * calls this code; only tests/, __graft_entry__.smoke() and bench.py's CPU
 * legs may.
 *
 * PARITY PINNING: the reference cannot be built here (no Rust toolchain; its
 * nalgebra 0.12 / alga 0.5 / num-traits 0.1 / tobj 0.1.3 dependencies are not
 * vendored) and its own test-suite pins only the pixel index layout
 * (src/geometry/coordinate.rs:71-85).  The oracle is therefore pinned by
 *   (1) that index test, restated in tests/test_oracle_kat.py;
 *   (2) known-answer vectors derived line-by-line from the cited reference
 *       source;
 *   (3) the reference's golden image examples/suzanne.png (loose visual
 *       golden: silhouette bbox, lit-coverage IoU, median colour error).
 * For coverage/depth/clipping/blending/stencil beyond (1)-(3): PARITY UNPINNED
 * by reference-run outputs.
 */
#ifndef SR_ORACLE_H
#define SR_ORACLE_H

#include "../include/softrender_b200_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Framebuffer planes owned by the caller (numpy arrays in the tests).
 * Pixel index = x + y*width (src/geometry/coordinate.rs:47-51). */
typedef struct so_framebuffer {
    uint32_t width, height;
    float *color;      /* width*height*4, RGBAf32 */
    float *depth;      /* width*height */
    uint8_t *stencil;  /* width*height or NULL (stencil type `()`) */
    uint32_t *winner;  /* optional: 1 + canonical index of the last primitive that wrote the pixel in this draw */
    uint32_t stencil_bytes; /* element size of `stencil`: 1 (u8; 0 means 1), 2 (u16) or 4 (u32) -- src/stencil.rs:9-60 */
    uint8_t *color_u8;      /* non-NULL: the colour attachment is RGBAu8Color (src/color/predefined.rs:26), width*height*4 bytes;
                             * `color` is then unused.  Blend = () only. */
    float *color1;          /* non-NULL: a texture buffer declared with TWO colour planes (declare_texture_buffer!,
                             * src/framebuffer/texturebuffer.rs:72-147): PixelBuffer::Color is the tuple (color, color1), the fragment
                             * shader returns both (SR_FS_SUZANNE_GBUFFER) and set_pixel_unchecked stores each into its own plane
                             * (:141-147).  Blend = () only. */
} so_framebuffer;

typedef struct so_texture {
    uint32_t width, height;
    const uint8_t *rgba; /* width*height*4 image texels (full_example/src/texture.rs:6), or NULL */
    /* render-to-texture source (src/framebuffer/texturebuffer.rs:12-58): f32 RGBA every `stride` floats, used when rgba is NULL */
    const float *texels_f32;
    uint32_t stride;
    uint32_t filter;     /* sr_texture_filter (src/texture.rs:21-25) */
    uint32_t edge;       /* sr_texture_edge (src/texture.rs:33-41) */
    float border[4];     /* Edge::Border(C) */
} so_texture;

/* FragmentShader builder state (src/pipeline/stages/fragment.rs:45-56) +
 * the pipeline's stencil config (src/pipeline/mod.rs:38-40). */
typedef struct so_raster_state {
    uint32_t cull_faces;        /* sr_winding */
    uint32_t blend;             /* sr_blend */
    uint32_t antialiased_lines; /* bool */
    uint32_t tile_width, tile_height; /* DEFAULT_TILE_SIZE 128x128 (fragment.rs:29); >= frame size gives the canonical one-tile semantics */
    uint32_t stencil_test;      /* sr_stencil_test */
    uint32_t stencil_op;        /* sr_stencil_op */
} so_raster_state;

typedef struct so_draw so_draw;

/* Pipeline::render_mesh (src/pipeline/mod.rs:146): returns NULL when
 * nindices % num_vertices(primitive) != 0 (the reference asserts). */
so_draw *so_draw_create(int primitive, const uint32_t *indices, uint64_t nindices,
                        int has_stencil_value, uint32_t stencil_value);
void so_draw_destroy(so_draw *);

/* VertexShader::run (src/pipeline/stages/vertex.rs:87) */
int so_draw_vertex_run(so_draw *, int vs, const sr_uniforms *, const float *vin,
                       uint64_t nverts, uint32_t vin_floats, int nthreads);
/* VertexShader::run_to_fragment (vertex.rs:123) */
int so_draw_vertex_run_to_fragment(so_draw *, const sr_viewport *, int vs, const sr_uniforms *,
                                   const float *vin, uint64_t nverts, uint32_t vin_floats, int nthreads);
/* test injection: indexed vertices already in clip (space=0) or screen (space=1) space; record = 4+nk floats */
int so_draw_set_vertices(so_draw *, const float *verts, uint64_t nverts, uint32_t nk, int space);
/* test injection: generated primitives (which: 1 points, 2 lines, 3 tris) */
int so_draw_set_generated(so_draw *, int which, const float *verts, uint64_t nverts, uint32_t nk);
/* GeometryShader::run / clip_primitives (geometry.rs:132,261) */
int so_draw_geometry_run(so_draw *, int gs, const sr_uniforms *, int nthreads);
/* GeometryShader::finish (geometry.rs:60) */
int so_draw_finish(so_draw *, const sr_viewport *, int nthreads);
/* FragmentShader::run (fragment.rs:168) */
int so_draw_fragment_run(so_draw *, so_framebuffer *, const so_raster_state *, int fs,
                         const sr_uniforms *, const so_texture *, int nthreads);

/* FragmentShader::run over the tiles first, first+stride, ... of its tile list only (each still visits every primitive);
 * (0, 1) = the whole frame.  bench.py's CPU arm times one frame as `stride` slices. */
int so_draw_fragment_run_tiles(so_draw *, so_framebuffer *, const so_raster_state *, int fs,
                               const sr_uniforms *, const so_texture *, int nthreads,
                               uint64_t tile_first, uint64_t tile_stride);

/* texture(t, coord, filter, edge) (src/texture.rs:14-18) with the arithmetic of full_example/src/texture.rs:25-84 */
void so_texture_sample(const so_texture *, float u, float v, float out[4]);

/* introspection: which = 0 indexed vertices, 1 points, 2 lines, 3 tris */
uint64_t so_draw_count(const so_draw *, int which);
const float *so_draw_data(const so_draw *, int which);
uint32_t so_draw_nk(const so_draw *);

/* tile list of FragmentShader::run (fragment.rs:188-216); out = 4 u32 per tile (x0,y0,x1,y1 inclusive);
 * returns the tile count (writes at most cap tiles). */
uint64_t so_tiles(uint32_t width, uint32_t height, uint32_t tw, uint32_t th, uint32_t *out, uint64_t cap);

/* derived tile bins for the triangles of a finished draw on a disjoint grid of
 * tw x th pixel tiles: CSR offsets[ntiles+1], ids = canonical triangle index, ascending per tile.
 * Call with ids==NULL to obtain the total. Returns total entries. */
uint64_t so_draw_bins(const so_draw *, uint32_t width, uint32_t height, uint32_t tw, uint32_t th,
                      uint32_t cull_faces, uint64_t *offsets, uint32_t *ids);

/* Coordinate::into_index (src/geometry/coordinate.rs:47-51) */
uint64_t so_coordinate_index(uint32_t x, uint32_t y, uint32_t width);

/* StencilTest::test / StencilOp::op on u8 (src/stencil.rs:112-123,147-158) */
int so_stencil_test(uint32_t test, uint8_t value, uint8_t mask);
uint8_t so_stencil_op(uint32_t op, uint8_t value, uint8_t mask);
/* the same on the u8 / u16 / u32 stencil types (bits = 8, 16, 32) */
uint32_t so_stencil_op_wide(uint32_t op, uint32_t value, uint32_t mask, uint32_t bits);

/* f32::MIN bit pattern used by Depth::far (src/framebuffer/attachments/depth.rs:31) */
float so_depth_far(void);

/* RenderBuffer::clear (src/framebuffer/renderbuffer/mod.rs:126-133) */
void so_framebuffer_clear(so_framebuffer *, const float color[4]);
/* Framebuffer::clear of a texture buffer with two colour planes: the tuple of colours, one per plane (texturebuffer.rs:181-197) */
void so_framebuffer_clear2(so_framebuffer *, const float color[4], const float color1[4]);

#ifdef __cplusplus
}
#endif
#endif
