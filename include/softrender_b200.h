/*
 * softrender_b200.h -- C ABI of libsoftrender_b200.so, the B200-native (sm_100a CUDA)
 * implementation of the draw hot path of novacrazy/rust-softrender.
 *
 * The reference has no FFI/plugin interface: its "operator API" is the typed Rust
 * builder (Pipeline -> VertexShader -> GeometryShader -> FragmentShader).  Each entry
 * point below is what a Rust `-sys` crate would bind so that the body of the cited
 * builder method becomes one call into this library (INTEGRATION.md shows the binding).
 * Paths are relative to the reference repository root.
 *
 * Conventions
 *  - every function returns an sr_status (0 = SR_OK); sr_last_error() gives the text of the
 *    last failure on the calling thread.  Nothing aborts or throws across the boundary
 *    (the reference panics on contract violations).
 *  - handles are opaque, freed by the matching *_destroy.  Buffers passed in are copied
 *    before the call returns; downloads write into caller memory.
 *  - a context is externally synchronised (one caller thread at a time), matching the
 *    reference's `&mut Pipeline`.
 *  - draw calls are enqueued on the context's CUDA stream; the framebuffer contents are
 *    defined for every later call on the same context (downloads synchronise), which is
 *    the observable behaviour of the reference's `pool.scoped` joins.
 *  - there is NO CPU fallback: every entry point needs a CUDA device.
 *  - lifetimes: a handle may be destroyed as soon as the calls that take it have returned, EXCEPT that a pipeline keeps
 *    plain references to its framebuffer and bound texture / texture-source framebuffer, and a draw to its pipeline:
 *    destroy draws before their pipeline, a pipeline before its framebuffer and textures (the reference's borrows,
 *    `&'a mut P` in src/pipeline/stages/, enforce the same order at compile time).  Device buffers themselves are
 *    reference counted: a mesh, texture or framebuffer may be destroyed while enqueued work still reads it, and a
 *    context stays valid until its last child is gone.  All objects of one draw must belong to one context;
 *    handles of different contexts are rejected with SR_ERR_INVALID_STATE (sr_framebuffer_alias / _ipc_open give another
 *    context a handle on the same pixels; detach a shard group from its contexts before destroying it).
 */
#ifndef SOFTRENDER_B200_H
#define SOFTRENDER_B200_H

#include "softrender_b200_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sr_context sr_context;
typedef struct sr_framebuffer sr_framebuffer;
typedef struct sr_mesh sr_mesh;
typedef struct sr_texture sr_texture;
typedef struct sr_pipeline sr_pipeline;
typedef struct sr_draw sr_draw;

/* ---- library ----------------------------------------------------------------------- */
const char *sr_last_error(void);
int sr_version(void);
/* internal GPU tile (pixels).  Results do not depend on it. */
int sr_tile_size(uint32_t *width, uint32_t *height);

/* ---- shader registry --------------------------------------------------------------------
 * Closures cannot cross a C ABI: the reference's vertex / geometry / fragment closures and Blend impls are a registered set
 * of device functions.  Enumeration: call with index = 0, 1, ... until SR_ERR_INVALID_ARGUMENT; `info->id` is the value
 * to pass to sr_vertex_run / sr_geometry_run / sr_fragment_run / sr_fragment_set_blend.  Needs no device. */
int sr_registry_entry(uint32_t kind /* sr_registry_kind */, uint32_t index, sr_shader_info *info);

/* ---- context: owns the device, stream and scratch memory.  Replaces the thread pool
 *      Pipeline::new creates (src/pipeline/mod.rs:110-117). ---------------------------- */
int sr_context_create(int device_ordinal, sr_context **out);
int sr_context_destroy(sr_context *);
int sr_context_synchronize(sr_context *);
/* the context's cudaStream_t (for CUDA-event timing on the launching stream) */
void *sr_context_stream(sr_context *);
/* sort-first tile sharding: this context rasterises only the GPU tiles with
 * tile_index % world == rank (geometry stages still run for the whole mesh). */
int sr_context_set_tile_shard(sr_context *, uint32_t rank, uint32_t world);

/* ---- tile-sharded frames whose per-triangle front end is sharded too -------------------------------------
 * The reference parallelises FragmentShader::run by tile (src/pipeline/stages/fragment.rs:240-253: a pool of
 * threads pulls tiles from an atomic cursor, every tile visits every primitive).  Across GPUs the same split is
 * "rank r owns the tiles with index % world == r" (sr_context_set_tile_shard); an sr_shard group additionally
 * splits the per-triangle work: rank r rasterises the triangles [r*T/world, (r+1)*T/world) of an opaque draw into
 * a full-frame key buffer of its own, and the owner of a tile pulls the other ranks' keys of that tile over
 * NVLink inside its tile kernel (TMA bulk loads from peer-mapped memory), max-merges them and resolves.  The
 * merged key -- max over all fragments of (depth, submission index) -- is exactly what one GPU reduces, so the
 * frame is bit-identical to the single-GPU frame.  Ranks synchronise through progress words in each other's
 * exchange block (no host round trip, no NCCL call per frame).
 *
 *   sr_shard_create   allocates this rank's exchange block (`lanes` key buffers + progress words) for frames
 *                     of width x height; the context must already carry its (rank, world) tile shard.
 *   sr_shard_export   64-byte CUDA IPC handle of the block (one process per GPU: all-gather these).
 *   sr_shard_connect  opens the peers' handles (array of world x 64 bytes, entry `rank` ignored).
 *   sr_shard_connect_local  the same for ranks that live in ONE process (tests; peers[] = world handles).
 *   sr_context_attach_shard  eligible draws of this context (opaque triangles onto a freshly cleared frame,
 *                     >= 65536 triangles) take the range-sharded path on `lane`; every rank must issue the same
 *                     sequence of such draws per lane.  Contexts of one rank that work on different lanes keep
 *                     independent frames in flight.  shard = NULL detaches.
 *   sr_shard_status   0, or 1 + the peer a wait gave up on (a rank died or never issued its draw).
 * Everything not eligible falls back to plain sort-first tile sharding, which needs no exchange. */
typedef struct sr_shard sr_shard;
int sr_shard_create(sr_context *, uint32_t width, uint32_t height, uint32_t lanes, sr_shard **out);
int sr_shard_export(sr_shard *, void *handle64);
int sr_shard_connect(sr_shard *, const void *handles64, uint32_t count);
int sr_shard_connect_local(sr_shard *, sr_shard *const *peers, uint32_t count);
int sr_context_attach_shard(sr_context *, sr_shard *, uint32_t lane);
int sr_shard_status(sr_shard *, uint32_t *status);
int sr_shard_destroy(sr_shard *);
/* a second handle on the pixels of `src` for another context of the SAME process (a rank of a single-process
 * shard group draws into rank 0's framebuffer through it); colour + depth only, does not own the memory */
int sr_framebuffer_alias(sr_context *, sr_framebuffer *src, sr_framebuffer **out);
/* tuning of the opaque triangle path: triangles whose frame-clamped bounding box holds at most
 * `area` pixels are rasterised per-triangle into the visibility buffer, the rest through per-tile lists.
 * area = 0 sends every triangle through the tile lists; area = SR_MICRO_AREA_AUTO (the default) lets the library pick
 * the split per draw from the triangle count.  Draws onto existing (not freshly cleared) contents
 * use the visibility buffer only from `min_triangles` on.  `precheck` bit 0: read a pixel's key before the atomic
 * (measured slower); bit 1: switch OFF the early depth rejection of whole small triangles.  Results never depend on
 * these values. */
#define SR_MICRO_AREA_AUTO 0xFFFFFFFFu
int sr_context_set_micro(sr_context *, uint32_t area, uint32_t min_triangles, uint32_t precheck);
/* record the per-stage CUDA events read by sr_context_stage_times / sr_context_stage_timestamps (off by default:
 * eight timed events per draw are a measurable share of a small frame) */
int sr_context_set_stage_timing(sr_context *, int enable);
/* timeline introspection: time in ms from `base_event` (a cudaEvent_t recorded by the caller, timing enabled) to the
 * internal stage events of the most recent draw: [0] vertex begin, [1] vertex end, [2] geometry end, [3] fragment begin,
 * [4] ordered bins end, [7] visibility init end, [5] raster front end (k_micro) end, [6] fragment end; -1 = not recorded */
int sr_context_stage_timestamps(sr_context *, void *base_event, float ms[8]);
/* Cross-context ordering for frames in flight (one context = one CUDA stream; the reference has one frame in flight
 * per pipeline, the stage structs in src/pipeline/stages/ hold `&mut P`).  Work enqueued on `waiter` after this call starts only when
 * `other` has reached `point`: 0 = everything enqueued on it so far; 1 = the raster front end (per-triangle setup and
 * small-triangle rasterisation) of its most recent opaque draw -- the software-pipeline schedule "frame n+1's vertex
 * stage and front end overlap frame n's tile resolve". */
int sr_context_wait_for(sr_context *waiter, sr_context *other, uint32_t point);
/* capacity (u32 entries) of the arena that holds the opaque path's per-tile triangle lists.  A draw is enqueued
 * against the current capacity without a host synchronisation; if its lists do not fit, the tile pass skips itself on
 * the device and is enqueued again with a larger arena at the next call that touches the context.
 * Setting a tiny capacity exercises that path in tests. */
int sr_context_set_list_capacity(sr_context *, uint32_t entries);
int sr_context_list_capacity(sr_context *, uint32_t *entries);
/* the same for the ordered path's per-tile group lists: capacities of the point, line and triangle arenas
 * (sr_context_set_list_capacity sets all of them) */
int sr_context_ordered_list_capacity(sr_context *, uint32_t entries[3]);
/* number of kernels launched by this context since creation (bench.py's gpu_launches) */
int sr_context_launch_count(sr_context *, uint64_t *out);

/* ---- framebuffer: RenderBuffer<ColorDepth[Stencil]Attachments<RGBAf32Color,f32[,u8]>>
 *      (src/framebuffer/renderbuffer/mod.rs:34-133) ------------------------------------ */
/* RenderBuffer::with_dimensions (renderbuffer/mod.rs:58-63): colour = Color::empty() (zeros),
 * depth = Depth::far() = f32::MIN, stencil = 0. */
int sr_framebuffer_create(sr_context *, uint32_t width, uint32_t height, uint32_t format, sr_framebuffer **out);
int sr_framebuffer_destroy(sr_framebuffer *);
/* RenderBuffer::clear (renderbuffer/mod.rs:126-133) */
int sr_framebuffer_clear(sr_framebuffer *, const float color[4]);
/* HasDimensions::dimensions (src/geometry/dimension.rs:29) */
int sr_framebuffer_dimensions(const sr_framebuffer *, uint32_t *width, uint32_t *height);
/* read-back of the RenderBuffer's Vec<{color,depth}>: row-major, index = x + y*width
 * (src/geometry/coordinate.rs:47-51), 20 B/pixel {r,g,b,a,depth} -- 8 B/pixel {r,g,b,a as u8, f32 depth} for an RGBAu8Color
 * target.  nbytes must be width*height*20 (resp. *8). */
int sr_framebuffer_download(sr_framebuffer *, void *dst, size_t nbytes);
/* presentation read-back: the loop of realtime_example/src/main.rs:100-116, `(c.r * 255.0) as u8` per channel
 * (truncating, saturating, NaN -> 0), converted on the device so that 4 instead of 20 bytes per pixel cross PCIe.
 * order 0: bytes r,g,b,a (image crate Rgba<u8>, src/image/color.rs:60-90); order 1: a,b,g,r (the example's SDL
 * RGBA8888 streaming texture).  nbytes must be width*height*4. */
int sr_framebuffer_download_rgba8(sr_framebuffer *, uint8_t *dst, size_t nbytes, uint32_t order);
/* plane views (texturebuffer.rs:72-198 layout): any pointer may be NULL.  `color` holds width*height colours of the format's
 * colour type (4 x f32, or 4 x u8 for an RGBAu8Color target), `stencil` width*height elements of its stencil type (u8, u16 or
 * u32: src/stencil.rs:9-60). */
int sr_framebuffer_download_planes(sr_framebuffer *, void *color, float *depth, void *stencil);
int sr_framebuffer_upload_planes(sr_framebuffer *, const void *color, const float *depth, const void *stencil);
/* checked accessors: PixelRead::pixel_ref / FramebufferAccessor::{get_depth, get_stencil} (src/pixels/mod.rs:56-63,
 * src/framebuffer/accessor.rs:28-38) and PixelWrite::pixel_mut / FramebufferAccessorMut::{set_depth, set_stencil}
 * (src/pixels/mod.rs:77-98, src/framebuffer/accessor.rs:52-70); out-of-range -> SR_ERR_INVALID_PIXEL_COORDINATE.
 * The stencil value travels as u32 whatever the attachment's width (set: it must fit); the channels of an RGBAu8Color target
 * travel as their values 0..255 in floats (set: they must be such values); any pointer may be NULL. */
int sr_framebuffer_get_pixel(sr_framebuffer *, uint32_t x, uint32_t y, float rgba[4], float *depth, uint32_t *stencil);
int sr_framebuffer_set_pixel(sr_framebuffer *, uint32_t x, uint32_t y, const float rgba[4], const float *depth, const uint32_t *stencil);
/* Texture buffers declared with more than one colour plane (declare_texture_buffer! takes one OR MORE named colours,
 * src/framebuffer/texturebuffer.rs:72-110; format SR_FB_TEXTURE_2xRGBAF32_DF32 has two).  The planes are addressed by index in
 * declaration order -- the C-ABI form of the named accessors the macro generates (:110-117).
 * _clear_attachment: Framebuffer::clear takes the tuple of all colours (:181-197); this records the colour of ONE plane and, like
 *   sr_framebuffer_clear, a clear of the whole framebuffer (the other planes keep the colour last recorded for them, Color::empty()
 *   initially).  sr_framebuffer_clear(fb, c) is _clear_attachment(fb, 0, c).
 * _download_attachment: width*height colours (4 x f32) of plane `index`.
 * A draw into such a framebuffer needs a fragment shader that returns the tuple (SR_FS_SUZANNE_GBUFFER) and vice versa
 * (a type error in the reference, SR_ERR_INVALID_STATE here); Blend = (), stencil (), triangles. */
int sr_framebuffer_clear_attachment(sr_framebuffer *, uint32_t index, const float color[4]);
int sr_framebuffer_download_attachment(sr_framebuffer *, uint32_t index, float *color);
/* parity introspection: per pixel, 1 + canonical index of the last primitive of the most recent
 * draw that wrote it (0 = untouched by that draw).  Enable before drawing. */
int sr_framebuffer_enable_winner(sr_framebuffer *, int enable);
int sr_framebuffer_download_winner(sr_framebuffer *, uint32_t *dst);
/* device address of the AoS pixel store (for zero-copy consumers / collectives) */
void *sr_framebuffer_device_ptr(sr_framebuffer *);

/* ---- multi-GPU composite over NVLink peer memory -------------------------------------
 * One process per GPU.  Rank 0 exports its framebuffer; the other ranks open it and make it
 * their write-back target, so the tile rasteriser stores finished tiles straight into rank
 * 0's HBM (no staging buffer, no separate collective). */
int sr_framebuffer_ipc_export(sr_framebuffer *, void *handle64 /* 64 bytes */);
int sr_framebuffer_ipc_open(sr_context *, const void *handle64, uint32_t width, uint32_t height, uint32_t format,
                            sr_framebuffer **out);

/* ---- mesh: Arc<Mesh<V>> (src/mesh.rs:12-20,61-63) ----------------------------------------
 * vertices: AoS, `vin_floats` f32 per vertex, position.xyz first (SimpleVertex{position,data});
 * indices: u32 (index_bytes 4) or usize/u64 (index_bytes 8).  Converted to SoA planes in HBM. */
int sr_mesh_create(sr_context *, const float *vertices, uint64_t nverts, uint32_t vin_floats,
                   const void *indices, uint64_t nindices, uint32_t index_bytes, sr_mesh **out);
int sr_mesh_destroy(sr_mesh *);

/* ---- texture: RGBA8 image sampled by the textured fragment shader
 *      (full_example/src/texture.rs:25-84) ----------------------------------------------- */
int sr_texture_create(sr_context *, const uint8_t *rgba, uint32_t width, uint32_t height, sr_texture **out);
int sr_texture_destroy(sr_texture *);

/* ---- pipeline: Pipeline<U, F, S> (src/pipeline/mod.rs:60-140) ------------------------------ */
/* Pipeline::from_framebuffer (mod.rs:120); errors if width or height is 0 (asserts mod.rs:129-130) */
int sr_pipeline_create(sr_context *, sr_framebuffer *, const sr_uniforms *, sr_pipeline **out);
int sr_pipeline_destroy(sr_pipeline *);
/* PipelineObject::uniforms_mut (mod.rs:45) */
int sr_pipeline_set_uniforms(sr_pipeline *, const sr_uniforms *);
/* Pipeline::with_framebuffer (mod.rs:126-140): also RESETS the stencil config to default */
int sr_pipeline_set_framebuffer(sr_pipeline *, sr_framebuffer *);
/* PipelineObject::stencil_config_mut with GenericStencilConfig{op,test} (src/stencil.rs:177-199) */
int sr_pipeline_set_stencil_config(sr_pipeline *, uint32_t test, uint32_t op);
int sr_pipeline_bind_texture(sr_pipeline *, sr_texture *);
/* Render-to-texture: the colour attachment of `src` becomes the pipeline's texture IN PLACE -- the zero-cost
 * TextureBufferRef of src/framebuffer/texturebuffer.rs:12-58 ("re-used as textures without copying", :63-66).
 * Texels are the f32 RGBA colours as rendered (no /255, no gamma decode: a render target holds linear colour).
 * NULL unbinds; binding replaces an image texture bound with sr_pipeline_bind_texture and vice versa.  A draw that
 * renders into `src` itself, or whose pipeline lives in another context than `src`, fails with SR_ERR_INVALID_STATE;
 * a recorded clear of `src` is materialised before the sampling draw.  Like an image texture, `src` must stay alive while it
 * is bound (the pipeline borrows it, as `TextureBufferRef<'a>` borrows its parent); unbind with NULL before destroying it. */
int sr_pipeline_bind_framebuffer_texture(sr_pipeline *, sr_framebuffer *src);
/* the same for colour plane `index` of a texture buffer (`buffer.normals()` of a buffer declared with `pub normals: ...`,
 * texturebuffer.rs:110-117); index 0 of a one-plane texture buffer is sr_pipeline_bind_framebuffer_texture */
int sr_pipeline_bind_framebuffer_attachment(sr_pipeline *, sr_framebuffer *src, uint32_t index);
/* Filter and Edge of texture(t, coord, filter, edge) (src/texture.rs:14-45) for every sampling shader of the pipeline.
 * border_rgba: 4 floats, read only for SR_EDGE_BORDER (NULL = transparent black).  Default: NEAREST, CLAMP -- `impl Default for Filter` /
 * `for Edge` (src/texture.rs:27-31,43-45); the config-2 scene sets BILINEAR explicitly. */
int sr_pipeline_set_sampler(sr_pipeline *, uint32_t filter, uint32_t edge, const float *border_rgba);

/* ---- draw: the VertexShader -> GeometryShader -> FragmentShader chain ---------------------- */
/* Pipeline::render_mesh (mod.rs:146-157); errors if nindices % num_vertices(primitive) != 0 (assert mod.rs:148).
 * has_stencil_value=0 is `None` (value defaults to 0). */
int sr_render_mesh(sr_pipeline *, sr_mesh *, uint32_t primitive, int has_stencil_value, uint32_t stencil_value,
                   sr_draw **out);
/* VertexShader::run (src/pipeline/stages/vertex.rs:87-120) */
int sr_vertex_run(sr_draw *, uint32_t vertex_shader);
/* VertexShader::run_to_fragment (vertex.rs:123-160) */
int sr_vertex_run_to_fragment(sr_draw *, const sr_viewport *, uint32_t vertex_shader);
/* GeometryShader::run (src/pipeline/stages/geometry.rs:132-258) with a registered shader */
int sr_geometry_run(sr_draw *, uint32_t geometry_shader);
/* GeometryShader::clip_primitives (geometry.rs:261-336): the reference's clipper, restated literally.
 * sr_geometry_run(d, SR_GS_CLIP_SH) is the opt-in correct one (Sutherland-Hodgman, same planes and intersect()). */
int sr_geometry_clip_primitives(sr_draw *);
/* GeometryShader::finish (geometry.rs:60-129) */
int sr_geometry_finish(sr_draw *, const sr_viewport *);
/* {Vertex,Geometry,Fragment}Shader::duplicate (vertex.rs:44, geometry.rs:43, fragment.rs:122) */
int sr_draw_duplicate(sr_draw *, sr_draw **out);
/* FragmentShader::cull_faces / set_cull_faces (src/pipeline/stages/fragment.rs:82-90) */
int sr_fragment_set_cull_faces(sr_draw *, uint32_t winding);
/* FragmentShader::antialiased_lines (fragment.rs:96-104) */
int sr_fragment_set_antialiased_lines(sr_draw *, int enable);
/* FragmentShader::tile_size (fragment.rs:107-116).  Accepted for API parity; the GPU tiles pixels
 * disjointly, which equals the reference with one frame-sized tile (DESIGN.md "canonical semantics"). */
int sr_fragment_set_tile_size(sr_draw *, uint32_t width, uint32_t height);
/* FragmentShader::with_blend / with_default_blend (fragment.rs:140-160) with a registered blend */
int sr_fragment_set_blend(sr_draw *, uint32_t blend);
/* FragmentShader::run (fragment.rs:168-319) */
int sr_fragment_run(sr_draw *, uint32_t fragment_shader);
int sr_draw_destroy(sr_draw *);

/* ---- parity-test injection and introspection ------------------------------------------------ */
/* start a draw from already-shaded vertices (records of 4+nk floats) in clip (space=0) or screen
 * (space=1) space; the state VertexShader::run / GeometryShader::finish would have produced. */
int sr_draw_from_vertices(sr_pipeline *, uint32_t primitive, const float *verts, uint64_t nverts, uint32_t nk,
                          int space, const uint32_t *indices, uint64_t nindices, int has_stencil_value,
                          uint32_t stencil_value, sr_draw **out);
/* replace one generated-primitive stream (which: 1 points, 2 lines, 3 triangles) */
int sr_draw_set_generated(sr_draw *, int which, const float *verts, uint64_t nverts, uint32_t nk);
/* which: 0 indexed vertices, 1 points, 2 lines, 3 triangles (vertex counts; records of 4+nk floats) */
int sr_draw_count(sr_draw *, int which, uint64_t *nverts, uint32_t *nk);
int sr_draw_download(sr_draw *, int which, float *dst, uint64_t capacity_floats);
/* for generated triangles: position of each kept triangle in the reference's literal output
 * sequence (the clipper's zero-area triangles may be dropped) */
int sr_draw_download_sequence(sr_draw *, uint32_t *dst, uint64_t capacity);
/* per-GPU-tile triangle lists of a finished (screen-space) draw: CSR offsets[ntiles+1] + ids
 * (canonical triangle index, ascending per tile).  ids may be NULL to query *total. */
int sr_draw_bins(sr_draw *, uint64_t *offsets, uint32_t *ids, uint64_t ids_capacity, uint64_t *total);
/* the same for the lists the opaque fast path actually built for the latest opaque draw of the context (k_bin_small, or
 * k_micro + k_large_fill: only the triangles whose frame-clamped bounding box exceeds *micro_area pixels go through per-tile
 * lists there, the rest are rasterised per triangle; *micro_area = 0: every triangle).  ids ascending per tile;
 * offsets holds ntiles + 1 entries (ntiles from the framebuffer's size and sr_tile_size). */
int sr_context_last_opaque_lists(sr_context *, uint64_t *offsets, uint32_t *ids, uint64_t ids_capacity, uint64_t *total,
                                 uint32_t *micro_area);
/* self-test: compares the rasteriser's exact-division shortcut with IEEE division on `count` random operand
 * pairs over its whole validity range; *mismatches must come back 0 */
int sr_selftest_division(sr_context *, uint64_t seed, uint64_t count, uint64_t *mismatches);
/* device time of the stages of the most recent fragment_run of this context */
int sr_context_stage_times(sr_context *, sr_stage_times *out);

#ifdef __cplusplus
}
#endif
#endif /* SOFTRENDER_B200_H */
