// sr_api.cu -- host side of libsoftrender_b200.so: handle objects, stage orchestration, the C ABI
// declared in include/softrender_b200.h.  No CPU fallback exists: every entry point drives CUDA.
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <algorithm>

#include <nvtx3/nvToolsExt.h>

#include "sr_common.cuh"
#include "sr_shaders.cuh"
#include "sr_stages.cuh"
#include "sr_raster.cuh"

// ---------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int sr_fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}
#define SR_CUDA(expr)                                                                                       \
    do {                                                                                                    \
        cudaError_t e__ = (expr);                                                                           \
        if (e__ != cudaSuccess) return sr_fail(SR_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)
#define SR_TRY(expr)               \
    do {                           \
        int rc__ = (expr);         \
        if (rc__ != SR_OK) return rc__; \
    } while (0)

// ---------------------------------------------------------------------------------------------------------
// objects
// ---------------------------------------------------------------------------------------------------------
struct DevBuf {
    sr_context *ctx = nullptr;
    void *ptr = nullptr;
    size_t bytes = 0;
    // an index buffer remembers the vertex range its triangles [t0, t1) reference (range-sharded frames; computed once)
    uint32_t vr_t0 = 0, vr_t1 = 0, vr_lo = 0, vr_hi = 0;
    // ... and, for the chunk-culled front end: the vertex range of every 1024-triangle chunk, the object-space box of every 256-vertex
    // block of the mesh it indexes (static per mesh, computed once), and whether the mesh is coherent enough for that mode
    std::shared_ptr<DevBuf> chunk_vr, blk_aabb;
    const void *aabb_planes = nullptr;
    ~DevBuf();
    template <class T> T *as() const { return reinterpret_cast<T *>(ptr); }
};
using Buf = std::shared_ptr<DevBuf>;

struct sr_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::multimap<size_t, void *> free_list;  // stream-ordered reuse of scratch allocations
    int sm_count = 148;
    uint32_t shard_rank = 0, shard_world = 1;
    uint64_t launches = 0;
    cudaEvent_t ev[8] = {};  // vertex begin/end, geometry end, fragment begin, bins end, micro end, raster end
    bool ev_valid[8] = {};
    bool stage_timing = false;  // sr_context_set_stage_timing
    uint32_t micro_area = SR_MICRO_AREA_DEFAULT;  // bbox pixels up to which k_micro rasterises a triangle itself (0: off)
    bool micro_auto = true;                       // choose the split per draw from the triangle count (sr_micro_area_for)
    uint32_t micro_min_tris = 65536;              // draws onto existing contents use the visibility buffer from this size on
    uint32_t micro_precheck = 0;  // read the key before the atomic: measured slower (load latency inside the per-lane loop)
    uint32_t *pinned = nullptr;                   // pinned host words for device->host counters
    // per-tile lists of the opaque path live in one arena that only grows; a draw is launched optimistically against
    // the current capacity and validated later (settle) so that no host synchronisation sits inside a frame
    Buf list_arena;
    uint32_t list_cap = 0;
    struct PendingOpaque *pending = nullptr;
    struct PendingOrdered *pending_ord = nullptr;
    Buf ord_arena[3];                             // grow-only group-list arenas of the ordered path (points, lines, triangles)
    uint32_t ord_cap[3] = {0, 0, 0};
    cudaEvent_t ev_front = nullptr;               // recorded after the raster front end (k_micro) of the latest opaque draw
    bool ev_front_valid = false;
    Buf zero_off;                                 // all-zero CSR offsets for empty primitive kinds
    uint32_t zero_off_tiles = 0;
    Buf last_off, last_list;                      // per-tile triangle lists of the latest opaque draw (sr_context_last_opaque_lists)
    uint32_t last_ntiles = 0, last_micro_area = 0;
    struct sr_shard *shard = nullptr;             // range-sharded front end (sr_context_attach_shard)
    uint32_t shard_lane = 0;
    cudaStream_t aux = nullptr;                   // second stream: rank 0's clear pre-fill of foreign tiles runs beside its k_micro
    cudaEvent_t ev_aux[2] = {};
    cudaEvent_t ev_dbg[6] = {};                   // SR_SHARD_DEBUG: finer timestamps inside a range-sharded frame (diagnosis only)
    bool ev_dbg_valid = false;
    sr_stage_times times = {};
    int alloc(size_t bytes, Buf *out);
    // Lifetime: every device buffer keeps its context alive (refs), so children destroyed after sr_context_destroy -- a
    // garbage collector tearing objects down in arbitrary order -- still find a valid context; `closed` contexts give
    // memory straight back to the driver.
    int refs = 1;
    bool closed = false;
    void release(void *p, size_t bytes) {
        if (closed) cudaFree(p);
        else free_list.emplace(bytes, p);
    }
};
static void ctx_unref(sr_context *c) {
    if (--c->refs == 0) delete c;
}

DevBuf::~DevBuf() {
    if (ptr && ctx) {
        ctx->release(ptr, bytes);
        ctx_unref(ctx);
    }
}

int sr_context::alloc(size_t bytes, Buf *out) {
    if (bytes == 0) bytes = 256;
    bytes = (bytes + 511) & ~(size_t)511;
    void *p = nullptr;
    auto it = free_list.lower_bound(bytes);
    if (it != free_list.end() && it->first <= bytes + bytes / 4 + 4096) {
        p = it->second;
        bytes = it->first;
        free_list.erase(it);
    } else {
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            // drop the cache and retry once
            for (auto &kv : free_list) cudaFree(kv.second);
            free_list.clear();
            e = cudaMalloc(&p, bytes);
            if (e != cudaSuccess) return sr_fail(SR_ERR_OUT_OF_MEMORY, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        }
    }
    auto b = std::make_shared<DevBuf>();
    b->ctx = this;
    ++refs;
    b->ptr = p;
    b->bytes = bytes;
    *out = b;
    return SR_OK;
}

#define SR_LAUNCH(ctx, kernel, grid, block, smem, ...)                                                    \
    do {                                                                                                  \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                                  \
        (ctx)->launches++;                                                                                \
        cudaError_t e__ = cudaGetLastError();                                                             \
        if (e__ != cudaSuccess) return sr_fail(SR_ERR_CUDA, "launch %s failed: %s", #kernel, cudaGetErrorString(e__)); \
    } while (0)

static inline uint32_t ceil_div(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

// NVTX range around the host side of a stage (tracing): visible on the timeline of any NVTX-aware tool,
// free when none is attached (NVTX3 is header-only and resolves its injection library lazily)
struct SrRange {
    explicit SrRange(const char *name) { nvtxRangePushA(name); }
    ~SrRange() { nvtxRangePop(); }
    SrRange(const SrRange &) = delete;
    SrRange &operator=(const SrRange &) = delete;
};

struct sr_framebuffer {
    sr_context *ctx;
    uint32_t width, height, format;
    uint32_t ntx, nty;
    Buf aos_buf, stencil_buf, winner_buf;
    uint32_t stencil_bytes = 0;  // element size of the stencil attachment: 1, 2 or 4 (0: stencil type `()`)
    bool u8color = false;        // RGBAu8Color target: 8-byte AoS pixels {rgba8, f32 depth}
    uint32_t soa = 0;            // texture-buffer storage: number of colour planes (float4 per pixel each), followed by the depth plane
    float clear1[4] = {0, 0, 0, 0};  // clear colour of the second colour plane (Color::empty() until set)
    size_t px_bytes() const { return u8color ? 8 : soa == 2 ? 36 : 20; }
    Buf vis_buf;                // visibility buffer of the opaque path (allocated on first use)
    bool vis_clean = false;     // every key of the tiles of shard (vis_rank, vis_world) is "far": the resolve hands it back that way
    uint32_t vis_rank = 0, vis_world = 0;
    float *aos = nullptr;       // device (possibly peer) pointer
    bool is_peer = false;       // opened through cudaIpcOpenMemHandle
    bool pending_clear = false;
    float clear[4] = {0, 0, 0, 0};
    bool winner_enabled = false;
    SrFbView view() const {
        SrFbView v;
        v.aos = aos;
        v.stencil = stencil_buf ? stencil_buf->as<uint8_t>() : nullptr;
        v.stencil_bytes = stencil_bytes;
        v.u8color = u8color ? 1u : 0u;
        v.soa = soa;
        for (int i = 0; i < 4; ++i) v.clear1[i] = clear1[i];
        v.winner = (winner_enabled && winner_buf) ? winner_buf->as<uint32_t>() : nullptr;
        v.width = width; v.height = height; v.ntx = ntx; v.nty = nty;
        v.pending_clear = pending_clear ? 1u : 0u;
        for (int i = 0; i < 4; ++i) v.clear[i] = clear[i];
        return v;
    }
};

// Exchange block of one rank of a shard group (include/softrender_b200.h): per lane one full-frame key buffer and two
// arrays of progress words (ready: "my keys of frame n are complete", done: "I have finished reading everybody's keys of
// frame n"), then one error word.  One cudaMalloc so that one IPC handle exports it.
struct sr_shard {
    sr_context *ctx = nullptr;
    uint32_t rank = 0, world = 1, width = 0, height = 0, ntx = 0, nty = 0, lanes = 1;
    unsigned char *block = nullptr;
    size_t vis_bytes = 0, lane_stride = 0, block_bytes = 0;
    unsigned char *peer[SR_SHARD_MAX_WORLD] = {};  // peers' blocks mapped into this process (null: me / not connected)
    bool peer_ipc[SR_SHARD_MAX_WORLD] = {};
    bool connected = false;
    uint32_t frame[8] = {};  // frames issued per lane
    int last_by_rows[8] = {-1, -1, -1, -1, -1, -1, -1, -1};  // ownership function of the lane's previous frame (a change clears every key)
    SrTileOwners owners;     // tile -> rank of the range-sharded frames (identical on every rank)
    uint32_t owned_cap = 0;  // upper bound of the tiles this rank owns per period (diagnostics)
    uint32_t merge_ctas = 0; // resident merge CTAs per SM (0: the kernel's natural 4); fewer leave room for the next frame's front end
    bool fused_merge = false; // merge the peers' keys inside the resolve (PHASE 2 pulls) instead of the streaming k_shard_merge
    unsigned long long *vis(unsigned char *base, uint32_t lane) const { return reinterpret_cast<unsigned long long *>(base + lane * lane_stride); }
    uint32_t *ready(unsigned char *base, uint32_t lane) const { return reinterpret_cast<uint32_t *>(base + lane * lane_stride + vis_bytes); }
    uint32_t *done(unsigned char *base, uint32_t lane) const { return ready(base, lane) + SR_SHARD_MAX_WORLD; }
    uint32_t *touched(unsigned char *base, uint32_t lane) const { return reinterpret_cast<uint32_t *>(base + lane * lane_stride + vis_bytes + 256); }
    uint32_t *error() const { return reinterpret_cast<uint32_t *>(block + lanes * lane_stride); }
};

struct sr_mesh {
    sr_context *ctx;
    Buf planes, indices;
    uint64_t nverts = 0, pstride = 0, nindices = 0;
    uint32_t vin = 0;
};

struct sr_texture {
    sr_context *ctx;
    Buf rgba;
    uint32_t width, height;
};

struct sr_pipeline {
    sr_context *ctx;
    sr_framebuffer *fb;
    sr_uniforms uniforms;
    uint32_t stencil_test = SR_STENCIL_ALWAYS, stencil_op = SR_STENCIL_KEEP;
    sr_texture *texture = nullptr;
    sr_framebuffer *fb_texture = nullptr;  // render-to-texture source, sampled in place (texturebuffer.rs:12-58)
    uint32_t fb_texture_plane = 0;         // which colour plane of a texture buffer
    uint32_t tex_filter = SR_FILTER_NEAREST, tex_edge = SR_EDGE_CLAMP;  // impl Default for Filter / Edge (src/texture.rs:27-31,43-45)
    float tex_border[4] = {0, 0, 0, 0};
};

struct VertexStream {  // one flat Vec of vertices in HBM: position array + attribute records (np float4 per vertex)
    Buf pos, attr;
    uint64_t n = 0, stride = 0, np = 1;  // stride = n padded to 4 (allocation size in vertices)
    SrVertexSet set() const {
        SrVertexSet s;
        s.pos = pos ? pos->as<float4>() : nullptr;
        s.attr = attr ? attr->as<float4>() : nullptr;
        s.np = np;
        return s;
    }
};

enum DrawStage { STAGE_VERTEX = 0, STAGE_GEOMETRY = 1, STAGE_FRAGMENT = 2 };

struct sr_draw {
    sr_pipeline *pipeline;
    uint32_t primitive;
    bool has_stencil_value = false;
    uint32_t stencil_value = 0;
    // the mesh (Arc<Mesh<V>>): input planes for the vertex stage, indices for every stage
    Buf mesh_planes, indices;
    uint64_t mesh_nverts = 0, mesh_pstride = 0, nindices = 0;
    uint32_t vin = 0;
    DrawStage stage = STAGE_VERTEX;
    uint32_t nk = 0;
    bool have_indexed = false;  // indexed_vertices: Option<Vec<..>>
    VertexStream indexed;
    VertexStream gen[3];        // generated points, lines, tris
    Buf tri_seq;                // literal sequence number of each generated triangle (null = identity)
    uint32_t tri_literal_total = 0;
    // clip_primitives without a host synchronisation (small meshes): gen[2] is sized for the clipper's worst case, the unused
    // tail of its positions is NaN, and the true numbers {kept, literal} of output triangles live here, on the device, until
    // somebody needs them on the host (resolve_tri_count)
    Buf tri_count_dev;
    // FragmentShader builder state (fragment.rs:45-56)
    uint32_t cull = SR_CULL_NONE, blend = SR_BLEND_REPLACE, aa_lines = 0;
    uint32_t tile_w = 128, tile_h = 128;  // DEFAULT_TILE_SIZE (fragment.rs:29); accepted, not used
    // run_to_fragment on a context with a shard group attached: the vertex stage is recorded, not run -- a range-sharded
    // fragment stage shades only the vertices this rank needs; every other consumer shades the whole mesh first
    bool vertex_lazy = false;
    uint32_t lazy_vs = 0;
    SrVsConst lazy_vc;
};

static uint32_t nplanes_of(uint32_t nk) { return (nk + 3) / 4; }
// the one host synchronisation clip_primitives skipped, for the callers that need the counts on the host (introspection, a second
// geometry pass, draws that also carry lines or points -- their canonical numbers follow the literal triangle count)
static int resolve_tri_count(sr_draw *d);

// ---------------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------------
static void record(sr_context *c, int i) {
    if (!c->stage_timing) return;  // eight timed events per draw are a measurable share of a small frame
    if (!c->ev[i]) cudaEventCreate(&c->ev[i]);
    cudaEventRecord(c->ev[i], c->stream);
    c->ev_valid[i] = true;
}

static int alloc_stream(sr_context *c, uint64_t n, uint32_t nk, VertexStream *out) {
    out->n = n;
    out->stride = (n + 3) & ~(uint64_t)3;
    if (out->stride == 0) out->stride = 4;
    out->np = std::max(1u, nplanes_of(nk));
    SR_TRY(c->alloc(out->stride * sizeof(float4), &out->pos));
    SR_TRY(c->alloc(out->stride * sizeof(float4) * out->np, &out->attr));
    return SR_OK;
}

static void fill_vs_const(const sr_pipeline *p, const sr_viewport *vp, SrVsConst *c) {
    memset(c, 0, sizeof(*c));
    c->u = p->uniforms;
    sr_mat_mat(p->uniforms.projection, p->uniforms.view, c->pv);  // projection * view
    sr_mat_mat(c->pv, p->uniforms.model, c->pvm);                 // (projection * view) * model
    c->normalize = vp ? 1 : 0;
    if (vp) {
        // viewport matrix of ClipVertex::normalize (clipvertex.rs:104-112), column-major
        const float left = vp->x, bottom = vp->y;
        const float right = left + vp->width, top = bottom + vp->height;
        c->vpm[0 * 4 + 0] = (right - left) / 2.0f;
        c->vpm[3 * 4 + 0] = (right + left) / 2.0f;
        c->vpm[1 * 4 + 1] = (top - bottom) / -2.0f;
        c->vpm[3 * 4 + 1] = (top + bottom) / 2.0f;
        c->vpm[2 * 4 + 2] = (vp->far_ - vp->near_) / -2.0f;
        c->vpm[3 * 4 + 2] = (vp->far_ + vp->near_) / -2.0f;
        c->vpm[3 * 4 + 3] = 1.0f;
    }
}

// enqueue an exclusive scan; the grand total lands in pinned host word `slot` once the stream gets there
static int exclusive_scan_async(sr_context *c, const uint32_t *in, uint64_t n, uint32_t *out, int slot) {
    const uint32_t nblocks = std::max(1u, ceil_div(n, SR_SCAN_BLOCK));
    Buf sums, total;
    SR_TRY(c->alloc((size_t)nblocks * 4, &sums));
    SR_TRY(c->alloc(4, &total));
    SR_LAUNCH(c, k_scan_reduce, nblocks, SR_SCAN_THREADS, 0, in, n, sums->as<uint32_t>());
    SR_LAUNCH(c, k_scan_sums, 1, SR_SCAN_THREADS, 0, sums->as<uint32_t>(), nblocks, total->as<uint32_t>());
    SR_LAUNCH(c, k_scan_apply, nblocks, SR_SCAN_THREADS, 0, in, n, sums->as<uint32_t>(), out);
    if (!c->pinned) SR_CUDA(cudaHostAlloc((void **)&c->pinned, 64, cudaHostAllocDefault));
    SR_CUDA(cudaMemcpyAsync(&c->pinned[slot], total->ptr, 4, cudaMemcpyDeviceToHost, c->stream));
    return SR_OK;
}
static int exclusive_scan(sr_context *c, const uint32_t *in, uint64_t n, uint32_t *out, uint32_t *total_host) {
    SR_TRY(exclusive_scan_async(c, in, n, out, 3));
    SR_CUDA(cudaStreamSynchronize(c->stream));
    *total_host = c->pinned[3];
    return SR_OK;
}

static int settle(sr_context *c);
static int materialize_clear(sr_framebuffer *fb) {
    SR_TRY(settle(fb->ctx));  // every read-back / host-visible access of a framebuffer passes through here
    if (!fb->pending_clear) return SR_OK;
    const uint64_t n = (uint64_t)fb->width * fb->height;
    SrFbView v = fb->view();
    if (fb->winner_buf) v.winner = fb->winner_buf->as<uint32_t>();
    SR_LAUNCH(fb->ctx, k_fb_fill, ceil_div(n, 256), 256, 0, v);
    fb->pending_clear = false;
    return SR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// bins
// ---------------------------------------------------------------------------------------------------------
struct Bins {
    Buf rects, off, list, count;
    uint32_t total = 0;
    uint32_t capacity = 0xFFFFFFFFu;
    SrBinParams params;  // of the fill pass (kept for a replay)
    uint32_t grid = 0;
    int nv = 0;          // 0: empty kind
    bool small = false;  // built by one k_bin_small_groups launch (replayed as a whole)
};

static int zero_offsets(sr_context *c, uint32_t ntiles, Bins *b) {
    // one persistent all-zero offsets array serves every empty primitive kind (nothing ever writes to it)
    if (!c->zero_off || c->zero_off_tiles < ntiles) {
        SR_TRY(c->alloc((size_t)(ntiles + 1) * 4, &c->zero_off));
        SR_CUDA(cudaMemsetAsync(c->zero_off->ptr, 0, (size_t)(ntiles + 1) * 4, c->stream));
        c->zero_off_tiles = ntiles;
    }
    b->off = c->zero_off;
    b->rects = c->zero_off;
    b->list = c->zero_off;
    b->total = 0;
    b->nv = 0;
    return SR_OK;
}

// Per-tile group lists of one primitive kind.  exact = true sizes the list with a host synchronisation (introspection:
// sr_draw_bins).  exact = false is the draw path: the list lives in a grow-only arena of the context, the fill pass and
// the tile kernel are enqueued against its current capacity and skip themselves on the device if the scanned total does
// not fit; the total travels to pinned memory (`total_slot`) and settle() re-runs the skipped passes with a larger arena.
template <int NV>
static int launch_bin_small_groups(sr_context *c, Bins *b) {
    static bool configured[16] = {};
    if (!configured[c->device & 15]) {
        SR_CUDA(cudaFuncSetAttribute(k_bin_small_groups<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, SR_BIN_SMALL_MAX_TILES * 4));
        configured[c->device & 15] = true;
    }
    SR_LAUNCH(c, k_bin_small_groups<NV>, 1, SR_BIN_SMALL_G_THREADS, (size_t)b->params.ntx * b->params.nty * 4, b->params, b->off->as<uint32_t>());
    return SR_OK;
}
template <int NV>
static int build_bins(sr_context *c, const sr_framebuffer *fb, const SrPrimSource &src, uint32_t nprims, uint32_t cull, Bins *b,
                      bool exact = true, uint32_t *total_slot = nullptr) {
    const uint32_t ntiles = fb->ntx * fb->nty;
    if (nprims == 0) return zero_offsets(c, ntiles, b);
    SR_TRY(c->alloc(((size_t)nprims + 64) * 4, &b->rects));
    SR_TRY(c->alloc((size_t)ntiles * 4, &b->count));
    SR_TRY(c->alloc((size_t)(ntiles + 1) * 4, &b->off));
    SrBinParams &p = b->params;
    memset(&p, 0, sizeof(p));
    p.src = src;
    p.nprims = nprims;
    p.cull = cull;
    p.width = fb->width; p.height = fb->height; p.ntx = fb->ntx; p.nty = fb->nty;
    p.shard_rank = c->shard_rank; p.shard_world = c->shard_world;
    p.rects = b->rects->as<uint32_t>();
    p.tile_count = b->count->as<uint32_t>();
    b->grid = ceil_div(nprims, 256);
    b->nv = NV;
    if (!exact && nprims <= (src.n1_dev ? SR_BIN_SMALL_MAX_TRIS_DEV : SR_BIN_SMALL_MAX_TRIS) && ntiles <= SR_BIN_SMALL_MAX_TILES) {
        // small draw: rectangles, counts, scan and fill in one single-CTA launch (no memset, no separate scan / fill)
        Buf &arena = c->ord_arena[NV - 1];
        if (!arena) {
            c->ord_cap[NV - 1] = 1u << 18;
            SR_TRY(c->alloc((size_t)c->ord_cap[NV - 1] * 4, &arena));
        }
        b->list = arena;
        b->capacity = c->ord_cap[NV - 1];
        b->small = true;
        p.tile_off = b->off->as<uint32_t>();
        p.list = b->list->as<uint32_t>();
        p.capacity = b->capacity;
        SR_TRY(launch_bin_small_groups<NV>(c, b));
        SR_CUDA(cudaMemcpyAsync(total_slot, b->off->as<uint32_t>() + ntiles, 4, cudaMemcpyDeviceToHost, c->stream));
        return SR_OK;
    }
    SR_CUDA(cudaMemsetAsync(b->count->ptr, 0, (size_t)ntiles * 4, c->stream));
    SR_LAUNCH(c, k_bin_setup<NV>, b->grid, 256, 0, p);
    SR_LAUNCH(c, k_tile_offsets, 1, SR_OFFSETS_THREADS, 0, b->count->as<uint32_t>(), ntiles, b->off->as<uint32_t>(), b->count->as<uint32_t>());
    p.tile_off = b->off->as<uint32_t>();
    if (exact) {
        SR_CUDA(cudaMemcpyAsync(&b->total, b->off->as<uint32_t>() + ntiles, 4, cudaMemcpyDeviceToHost, c->stream));
        SR_CUDA(cudaStreamSynchronize(c->stream));
        SR_TRY(c->alloc((size_t)std::max(b->total, 1u) * 4, &b->list));
        b->capacity = std::max(b->total, 1u);
    } else {
        Buf &arena = c->ord_arena[NV - 1];
        if (!arena) {
            c->ord_cap[NV - 1] = 1u << 18;
            SR_TRY(c->alloc((size_t)c->ord_cap[NV - 1] * 4, &arena));
        }
        b->list = arena;
        b->capacity = c->ord_cap[NV - 1];
        SR_CUDA(cudaMemcpyAsync(total_slot, b->off->as<uint32_t>() + ntiles, 4, cudaMemcpyDeviceToHost, c->stream));
    }
    p.list = b->list->as<uint32_t>();
    p.capacity = b->capacity;
    SR_LAUNCH(c, k_bin_fill, b->grid, 256, 0, p);
    return SR_OK;
}

static SrPrimSource prim_source(const sr_draw *d, uint32_t kind /*1 point,2 line,3 tri*/) {
    SrPrimSource s;
    memset(&s, 0, sizeof(s));
    s.nk = d->nk;
    s.nplanes = nplanes_of(d->nk);
    if (d->have_indexed && d->primitive == kind) {
        s.indices = d->indices->as<uint32_t>();
        s.vs0 = d->indexed.set();
        s.n0 = (uint32_t)(d->nindices / kind);
    }
    const VertexStream &g = d->gen[kind - 1];
    s.vs1 = g.set();
    s.n1 = (uint32_t)(g.n / kind);
    s.seq1 = (kind == 3 && d->tri_seq) ? d->tri_seq->as<uint32_t>() : nullptr;
    s.n1_dev = (kind == 3 && d->tri_count_dev) ? d->tri_count_dev->as<uint32_t>() : nullptr;
    return s;
}

// ---------------------------------------------------------------------------------------------------------
// kernel dispatch by registered shader id
// ---------------------------------------------------------------------------------------------------------
template <int FS>
static int launch_tiles(sr_context *c, uint32_t ntiles_owned, const SrTileParams &p) {
    static bool configured[16] = {};  // per device: opt in to the large dynamic shared memory once
    if (!configured[c->device & 15]) {
        SR_CUDA(cudaFuncSetAttribute(k_tile_ordered<FS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SR_ORD_SMEM_BYTES + 3 * SR_TILE_PIXELS)));
        configured[c->device & 15] = true;
    }
    // the tile's stencil values end the dynamic shared memory: 1 byte per pixel is in SR_ORD_SMEM_BYTES, wider types add the rest
    const size_t smem = SR_ORD_SMEM_BYTES + (p.fb.stencil && p.fb.stencil_bytes > 1 ? (size_t)(p.fb.stencil_bytes - 1) * SR_TILE_PIXELS : 0);
    SR_LAUNCH(c, k_tile_ordered<FS>, ntiles_owned, SR_RASTER_THREADS, smem, p);
    return SR_OK;
}
static int launch_tiles_fs(sr_context *c, uint32_t fs, uint32_t ntiles_owned, const SrTileParams &p) {
    switch (fs) {
        case SR_FS_FLAT: return launch_tiles<SR_FS_FLAT>(c, ntiles_owned, p);
        case SR_FS_SUZANNE: return launch_tiles<SR_FS_SUZANNE>(c, ntiles_owned, p);
        case SR_FS_FULL_EXAMPLE: return launch_tiles<SR_FS_FULL_EXAMPLE>(c, ntiles_owned, p);
        case SR_FS_FULL_EXAMPLE_TEXTURED: return launch_tiles<SR_FS_FULL_EXAMPLE_TEXTURED>(c, ntiles_owned, p);
        case SR_FS_GREEN: return launch_tiles<SR_FS_GREEN>(c, ntiles_owned, p);
        case SR_FS_DISCARD_CHECKER: return launch_tiles<SR_FS_DISCARD_CHECKER>(c, ntiles_owned, p);
        case SR_FS_TEXTURE_UNLIT: return launch_tiles<SR_FS_TEXTURE_UNLIT>(c, ntiles_owned, p);
    }
    return sr_fail(SR_ERR_INVALID_ARGUMENT, "unknown fragment shader %u", fs);
}
template <int FS, bool EXTRA>
static int launch_opaque(sr_context *c, uint32_t ntiles_owned, const SrOpaqueParams &p) {
    static bool configured[16] = {};
    if (!configured[c->device & 15]) {
        SR_CUDA(cudaFuncSetAttribute(k_tile_opaque<FS, EXTRA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SR_OPQ_SMEM_BYTES));
        configured[c->device & 15] = true;
    }
    SR_LAUNCH(c, (k_tile_opaque<FS, EXTRA>), ntiles_owned, SR_OPQ_THREADS, SR_OPQ_SMEM_BYTES, p);
    return SR_OK;
}
// range-sharded frames: PHASE 1 (list sweep into the rank's own keys; no shading, one instantiation) and PHASE 2 (merge + resolve)
static int launch_opaque_sweep(sr_context *c, uint32_t ntiles, const SrOpaqueParams &p) {
    static bool configured[16] = {};
    if (!configured[c->device & 15]) {
        SR_CUDA(cudaFuncSetAttribute(k_tile_opaque<SR_FS_FLAT, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SR_OPQ_SMEM_BYTES));
        configured[c->device & 15] = true;
    }
    SR_LAUNCH(c, (k_tile_opaque<SR_FS_FLAT, false, 1>), ntiles, SR_OPQ_THREADS, SR_OPQ_SMEM_BYTES, p);
    return SR_OK;
}
// `ctas_per_sm` > 0 pads the dynamic shared memory so that at most that many merge CTAs are resident per SM: on the ranks whose
// write-back crosses NVLink the resolve waits on the link, and leaving registers free lets the next frame's front end (another lane)
// run beside it
template <int FS>
static int launch_opaque_merge(sr_context *c, uint32_t ntiles, const SrOpaqueParams &p, uint32_t ctas_per_sm) {
    static bool configured[16] = {};
    if (!configured[c->device & 15]) {
        SR_CUDA(cudaFuncSetAttribute(k_tile_opaque<FS, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured[c->device & 15] = true;
    }
    size_t smem = p.npeers > 0 ? SR_OPQ_MERGE_SMEM_BYTES : SR_OPQ_SMEM_BYTES;  // the peer-key staging buffer only when peers are pulled here
    if (ctas_per_sm > 0) smem = std::max<size_t>(smem, std::min<size_t>(200 * 1024, (size_t)(227 * 1024) / (ctas_per_sm + 1) + 1024));
    SR_LAUNCH(c, (k_tile_opaque<FS, false, 2>), ntiles, SR_OPQ_THREADS, smem, p);
    return SR_OK;
}
static int launch_opaque_merge_fs(sr_context *c, uint32_t fs, uint32_t ntiles, const SrOpaqueParams &p, uint32_t ctas_per_sm) {
    switch (fs) {
        case SR_FS_FLAT: return launch_opaque_merge<SR_FS_FLAT>(c, ntiles, p, ctas_per_sm);
        case SR_FS_SUZANNE: return launch_opaque_merge<SR_FS_SUZANNE>(c, ntiles, p, ctas_per_sm);
        case SR_FS_FULL_EXAMPLE: return launch_opaque_merge<SR_FS_FULL_EXAMPLE>(c, ntiles, p, ctas_per_sm);
        case SR_FS_FULL_EXAMPLE_TEXTURED: return launch_opaque_merge<SR_FS_FULL_EXAMPLE_TEXTURED>(c, ntiles, p, ctas_per_sm);
        case SR_FS_GREEN: return launch_opaque_merge<SR_FS_GREEN>(c, ntiles, p, ctas_per_sm);
        case SR_FS_TEXTURE_UNLIT: return launch_opaque_merge<SR_FS_TEXTURE_UNLIT>(c, ntiles, p, ctas_per_sm);
        case SR_FS_SUZANNE_GBUFFER: return launch_opaque_merge<SR_FS_SUZANNE_GBUFFER>(c, ntiles, p, ctas_per_sm);
    }
    return sr_fail(SR_ERR_INVALID_ARGUMENT, "fragment shader %u cannot run on the opaque path", fs);
}
template <int FS>
static int launch_few(sr_context *c, uint32_t ntiles, const SrFewParams &p) {
    SR_LAUNCH(c, k_tile_few<FS>, ntiles, SR_FEW_THREADS, 0, p);
    return SR_OK;
}
static int launch_few_fs(sr_context *c, uint32_t fs, uint32_t ntiles, const SrFewParams &p) {
    switch (fs) {
        case SR_FS_FLAT: return launch_few<SR_FS_FLAT>(c, ntiles, p);
        case SR_FS_SUZANNE: return launch_few<SR_FS_SUZANNE>(c, ntiles, p);
        case SR_FS_FULL_EXAMPLE: return launch_few<SR_FS_FULL_EXAMPLE>(c, ntiles, p);
        case SR_FS_FULL_EXAMPLE_TEXTURED: return launch_few<SR_FS_FULL_EXAMPLE_TEXTURED>(c, ntiles, p);
        case SR_FS_GREEN: return launch_few<SR_FS_GREEN>(c, ntiles, p);
        case SR_FS_TEXTURE_UNLIT: return launch_few<SR_FS_TEXTURE_UNLIT>(c, ntiles, p);
    }
    return sr_fail(SR_ERR_INVALID_ARGUMENT, "fragment shader %u cannot run on the opaque path", fs);
}
static int launch_opaque_fs(sr_context *c, uint32_t fs, uint32_t ntiles_owned, const SrOpaqueParams &p) {
    const bool extra = p.nlines + p.npoints > 0;
#define SR_OPQ_CASE(F) case F: return extra ? launch_opaque<F, true>(c, ntiles_owned, p) : launch_opaque<F, false>(c, ntiles_owned, p)
    switch (fs) {
        SR_OPQ_CASE(SR_FS_FLAT);
        SR_OPQ_CASE(SR_FS_SUZANNE);
        SR_OPQ_CASE(SR_FS_FULL_EXAMPLE);
        SR_OPQ_CASE(SR_FS_FULL_EXAMPLE_TEXTURED);
        SR_OPQ_CASE(SR_FS_GREEN);
        SR_OPQ_CASE(SR_FS_TEXTURE_UNLIT);
        case SR_FS_SUZANNE_GBUFFER:  // (two colour outputs: triangles only, the sharded-resolve and plain instantiations)
            if (extra) return sr_fail(SR_ERR_UNSUPPORTED, "two-output fragment shaders draw triangles only");
            return launch_opaque<SR_FS_SUZANNE_GBUFFER, false>(c, ntiles_owned, p);
    }
#undef SR_OPQ_CASE
    return sr_fail(SR_ERR_INVALID_ARGUMENT, "fragment shader %u cannot run on the opaque path", fs);
}

// A draw whose per-tile lists were sized optimistically.  The device kernels skip themselves when the lists do not fit
// (k_large_fill / k_tile_opaque compare the scanned total with the capacity they were given); settle() reads the totals
// once the stream has passed them and, in that rare case, grows the arena and enqueues the skipped pass again.
struct PendingOpaque {
    cudaEvent_t counted = nullptr;  // recorded after the totals were copied to pinned memory
    uint32_t capacity = 0;
    uint32_t fs = 0, owned = 0, ntiles = 0;
    SrOpaqueParams op;
    std::vector<Buf> keep;          // every device buffer the pass reads (the draw may be destroyed meanwhile)
    Buf count, off, lcount, lids, lrects;
    sr_framebuffer *fb = nullptr;
    bool small = false;             // front end = one k_bin_small launch (re-run as a whole on overflow)
    bool ranged = false;            // range-sharded frame: cannot be replayed (the peers have moved on) -- overflow is an error
    SrMicroParams mp;
};
static int launch_opaque_pass(sr_context *c, PendingOpaque *q);
static int launch_bin_small(sr_context *c, PendingOpaque *q);
static int opaque_triangles_ranged(sr_context *c, sr_framebuffer *fb, const SrTileParams &tp, uint32_t cull, uint32_t fs, uint32_t owned,
                                   const std::vector<Buf> &keep, sr_draw *d);
static int materialize_vertices(sr_draw *d);
// The same for the ordered tile pass: its three group lists (points, lines, triangles) live in grow-only arenas.
struct PendingOrdered {
    cudaEvent_t counted = nullptr;
    SrTileParams tp;
    uint32_t fs = 0, owned = 0;
    Bins bins[3];  // points, lines, triangles
    std::vector<Buf> keep;
};
static int launch_tiles_fs(sr_context *c, uint32_t fs, uint32_t ntiles_owned, const SrTileParams &p);
static int settle_ordered(sr_context *c) {
    PendingOrdered *q = c->pending_ord;
    if (!q) return SR_OK;
    c->pending_ord = nullptr;
    std::unique_ptr<PendingOrdered> guard(q);
    SR_CUDA(cudaSetDevice(c->device));
    SR_CUDA(cudaEventSynchronize(q->counted));
    cudaEventDestroy(q->counted);
    q->counted = nullptr;
    bool replay = false;
    for (int k = 0; k < 3; ++k) {
        Bins &b = q->bins[k];
        if (b.nv == 0) continue;
        const uint32_t total = c->pinned[8 + k];
        if (total <= b.capacity) continue;
        // this kind's fill pass (and therefore the tile kernel) skipped itself: larger arena, fill again
        replay = true;
        const uint32_t cap = std::max<uint32_t>(total + total / 2, 1u << 18);
        SR_TRY(c->alloc((size_t)cap * 4, &c->ord_arena[k]));
        c->ord_cap[k] = cap;
        b.list = c->ord_arena[k];
        b.capacity = cap;
        b.params.list = b.list->as<uint32_t>();
        b.params.capacity = cap;
        if (b.small) {
            if (b.nv == 3) SR_TRY(launch_bin_small_groups<3>(c, &b));
            else if (b.nv == 2) SR_TRY(launch_bin_small_groups<2>(c, &b));
            else SR_TRY(launch_bin_small_groups<1>(c, &b));
        } else {
            SR_LAUNCH(c, k_bin_fill, b.grid, 256, 0, b.params);
        }
    }
    if (!replay) return SR_OK;
    q->tp.point_list = q->bins[0].list->as<uint32_t>(); q->tp.point_cap = q->bins[0].capacity;
    q->tp.line_list = q->bins[1].list->as<uint32_t>(); q->tp.line_cap = q->bins[1].capacity;
    q->tp.tri_list = q->bins[2].list->as<uint32_t>(); q->tp.tri_cap = q->bins[2].capacity;
    return launch_tiles_fs(c, q->fs, q->owned, q->tp);
}
static int settle(sr_context *c) {
    if (c && c->pending_ord && !c->pending) return settle_ordered(c);
    PendingOpaque *q = c ? c->pending : nullptr;
    if (!q) return SR_OK;
    c->pending = nullptr;
    std::unique_ptr<PendingOpaque> guard(q);
    SR_CUDA(cudaSetDevice(c->device));
    SR_CUDA(cudaEventSynchronize(q->counted));
    cudaEventDestroy(q->counted);
    q->counted = nullptr;
    const uint32_t total = c->pinned[0];
    if (total <= q->capacity) return settle_ordered(c);
    // the pass skipped itself: grow the arena and run it again (framebuffer and visibility buffer are untouched)
    Buf bigger;
    const uint32_t cap = std::max<uint32_t>(total + total / 2, 1u << 20);
    SR_TRY(c->alloc((size_t)cap * 4, &bigger));
    c->list_arena = bigger;
    c->list_cap = cap;
    if (q->ranged)
        return sr_fail(SR_ERR_INVALID_STATE, "range-sharded draw: the per-tile lists of its large triangles (%u entries) exceeded the arena (%u); "
                       "the frame is incomplete and cannot be replayed (the peers have moved on) -- the arena has been grown, draw the frame again",
                       total, q->capacity);
    q->capacity = cap;
    q->op.list = bigger->as<uint32_t>();
    q->op.list_capacity = cap;
    q->keep.push_back(bigger);
    SR_TRY(launch_opaque_pass(c, q));
    return settle_ordered(c);
}

// Where the split between the per-triangle front end and the per-tile lists lies.  One thread walking a whole bounding box
// is the cheapest way through a triangle as long as there are enough triangles to fill the machine (148 SMs x 2048
// threads); below that the serial walks are the critical path and the tile kernel's cooperative sweep wins.  Measured
// on 4-layer grids at 3840x2160 (profiles/scripts/area_sweep.py): 52k triangles of 280 px^2: 0.30 ms at area 16, 0.28 ms at 1024;
// 100k x 146 px^2: 0.36 -> 0.25; 207k x 70 px^2: 0.46 -> 0.23; 1M x 15 px^2: 0.67 -> 0.23; 31k x 464 px^2: 0.27 -> 0.33.
static uint32_t sr_micro_area_for(uint32_t ntris) { return ntris >= 49152u ? 1024u : SR_MICRO_AREA_DEFAULT; }

// The opaque triangle path (sr_raster.cuh): visibility-buffer init, k_micro (per-triangle setup + direct
// rasterisation of small triangles + compaction/counting of the large ones), per-tile lists of the large
// triangles, then the tile kernel (large triangles + resolve + single write-back).
static bool ranged_eligible(const sr_context *c, const sr_framebuffer *fb, const SrTileParams &tp, bool extra) {
    const uint32_t ntiles = fb->ntx * fb->nty;
    return c->shard && c->shard->connected && !extra && fb->pending_clear && tp.ntris >= 65536u && tp.tris.n1 == 0 && (c->micro_auto || c->micro_area > 0) &&
           c->shard->world == c->shard_world && c->shard->rank == c->shard_rank && c->shard->ntx == fb->ntx && c->shard->nty == fb->nty &&
           ntiles >= c->shard_world;
}
static int opaque_triangles(sr_context *c, sr_framebuffer *fb, const SrTileParams &tp, uint32_t cull, uint32_t fs, uint32_t owned,
                            const std::vector<Buf> &keep, bool extra, sr_draw *d = nullptr) {
    const uint32_t ntiles = fb->ntx * fb->nty;
    // a recorded clear also resets the stencil attachment (renderbuffer/mod.rs:126-133: stencil = Default::default()); the opaque path
    // never touches the stencil plane (stencil Always / Keep), so the reset is made here -- found by compute-sanitizer's memcheck run, whose
    // allocations are not zero-filled
    if (fb->pending_clear && fb->stencil_buf)
        SR_CUDA(cudaMemsetAsync(fb->stencil_buf->ptr, 0, (size_t)fb->width * fb->height * fb->stencil_bytes, c->stream));
    if (ranged_eligible(c, fb, tp, extra)) return opaque_triangles_ranged(c, fb, tp, cull, fs, owned, keep, d);
    if (d) SR_TRY(materialize_vertices(d));
    // a handful of triangles (a full-screen pass, UI rectangles) onto a plain RGBAf32 RenderBuffer: one launch, no lists (k_tile_few)
    static const bool no_few = getenv("SR_NO_FEW") != nullptr;  // A/B switch
    if (!no_few && !extra && c->micro_auto && tp.ntris > 0 && tp.ntris <= SR_FEW_MAX && !tp.tris.n1_dev && !fb->u8color && !fb->soa &&
        !(fb->winner_enabled && fb->winner_buf) && c->shard_world == 1 && !fb->is_peer && fs != SR_FS_SUZANNE_GBUFFER) {
        record(c, 7);
        record(c, 5);
        SrFewParams fp;
        memset(&fp, 0, sizeof(fp));
        fp.tris = tp.tris;
        fp.ntris = tp.ntris;
        fp.cull = cull;
        fp.fb = fb->view();
        fp.fs = tp.fs;
        SR_TRY(launch_few_fs(c, fs, ntiles, fp));
        fb->pending_clear = false;
        return SR_OK;
    }
    // small draws: one single-CTA launch builds the per-tile lists (k_bin_small); no visibility buffer
    if (!extra && c->micro_auto && tp.ntris > 0 && tp.ntris <= (tp.tris.n1_dev ? SR_BIN_SMALL_MAX_TRIS_DEV : SR_BIN_SMALL_MAX_TRIS) &&
        ntiles <= SR_BIN_SMALL_MAX_TILES) {
        record(c, 7);
        auto q = std::make_unique<PendingOpaque>();
        SR_TRY(c->alloc((size_t)(ntiles + 1) * 4, &q->off));
        if (!c->list_arena) {
            c->list_cap = 1u << 20;
            SR_TRY(c->alloc((size_t)c->list_cap * 4, &c->list_arena));
        }
        q->small = true;
        q->capacity = c->list_cap;
        q->fs = fs; q->owned = owned; q->ntiles = ntiles;
        memset(&q->mp, 0, sizeof(q->mp));
        q->mp.src = tp.tris;
        q->mp.ntris = tp.ntris;
        q->mp.cull = cull;
        q->mp.width = fb->width; q->mp.height = fb->height; q->mp.ntx = fb->ntx; q->mp.nty = fb->nty;
        q->mp.shard_rank = c->shard_rank; q->mp.shard_world = c->shard_world;
        q->keep = keep;
        q->keep.push_back(c->list_arena);
        q->fb = fb;
        memset(&q->op, 0, sizeof(q->op));
        q->op.tris = tp.tris;
        q->op.ntris = tp.ntris;
        q->op.tile_off = q->off->as<uint32_t>();
        q->op.list = c->list_arena->as<uint32_t>();
        q->op.list_capacity = c->list_cap;
        q->op.ntiles = ntiles;
        q->op.fb = fb->view();
        q->op.shard_rank = c->shard_rank; q->op.shard_world = c->shard_world;
        q->op.fs = tp.fs;
        SR_TRY(launch_bin_small(c, q.get()));
        c->last_off = q->off; c->last_list = c->list_arena; c->last_ntiles = ntiles; c->last_micro_area = 0;
        record(c, 5);
        if (!c->ev_front) SR_CUDA(cudaEventCreateWithFlags(&c->ev_front, cudaEventDisableTiming));
        SR_CUDA(cudaEventRecord(c->ev_front, c->stream));
        c->ev_front_valid = true;
        if (!c->pinned) SR_CUDA(cudaHostAlloc((void **)&c->pinned, 64, cudaHostAllocDefault));
        SR_CUDA(cudaMemcpyAsync(&c->pinned[0], q->off->as<uint32_t>() + ntiles, 4, cudaMemcpyDeviceToHost, c->stream));
        SR_CUDA(cudaEventCreateWithFlags(&q->counted, cudaEventDisableTiming));
        SR_CUDA(cudaEventRecord(q->counted, c->stream));
        SR_TRY(launch_opaque_fs(c, fs, owned, q->op));
        fb->pending_clear = false;
        c->pending = q.release();
        return SR_OK;
    }
    const uint32_t micro_area = c->micro_auto ? sr_micro_area_for(tp.ntris) : c->micro_area;
    // `extra`: the draw's non-antialiased lines and points go through the visibility buffer as well (k_lines_vis, k_points_vis)
    const bool use_micro = extra || (micro_area > 0 && tp.ntris > 0 && (fb->pending_clear || tp.ntris >= c->micro_min_tris));
    if (use_micro) {
        if (!fb->vis_buf) {
            SR_TRY(c->alloc((size_t)ntiles * SR_TILE_PIXELS * 8, &fb->vis_buf));
            fb->vis_clean = false;
        }
        const bool clean = fb->vis_clean && fb->vis_rank == c->shard_rank && fb->vis_world == c->shard_world;
        if (!(clean && fb->pending_clear))
            SR_LAUNCH(c, k_vis_init, owned, 256, 0, fb->vis_buf->as<unsigned long long>(), fb->view(), c->shard_rank, c->shard_world);
        fb->vis_clean = true;  // the tile kernel below resets the keys it consumes
        fb->vis_rank = c->shard_rank;
        fb->vis_world = c->shard_world;
    }
    record(c, 7);
    Buf count, off, lcount, lids, lrects;
    SR_TRY(c->alloc((size_t)(ntiles + 1) * 4, &count));
    SR_TRY(c->alloc((size_t)(ntiles + 1) * 4, &off));
    SR_TRY(c->alloc(4, &lcount));
    SR_TRY(c->alloc((size_t)std::max(tp.ntris, 1u) * 4, &lids));
    SR_TRY(c->alloc((size_t)std::max(tp.ntris, 1u) * 4, &lrects));
    SR_CUDA(cudaMemsetAsync(count->ptr, 0, (size_t)(ntiles + 1) * 4, c->stream));
    SR_CUDA(cudaMemsetAsync(lcount->ptr, 0, 4, c->stream));
    auto q = std::make_unique<PendingOpaque>();
    if (tp.ntris) {
        SrMicroParams mp;
        memset(&mp, 0, sizeof(mp));
        mp.src = tp.tris;
        mp.ntris = tp.ntris;
        mp.tri_begin = 0; mp.tri_end = tp.ntris;
        mp.cull = cull;
        mp.width = fb->width; mp.height = fb->height; mp.ntx = fb->ntx; mp.nty = fb->nty;
        mp.shard_rank = c->shard_rank; mp.shard_world = c->shard_world;
        mp.micro_area = use_micro ? micro_area : 0u;  // (0 with `extra` and the split switched off: every triangle through the lists)
        mp.vis = use_micro ? fb->vis_buf->as<unsigned long long>() : nullptr;
        mp.large_count = lcount->as<uint32_t>();
        mp.large_ids = lids->as<uint32_t>();
        mp.large_rects = lrects->as<uint32_t>();
        mp.tile_count = count->as<uint32_t>();
        // micro_precheck: bit 0 = per-fragment key pre-check before the atomic, bit 1 = early depth rejection OFF
        const uint32_t grid = ceil_div(tp.ntris, SR_MICRO_THREADS);
        switch (c->micro_precheck & 3u) {
            case 0: SR_LAUNCH(c, (k_micro<false, true>), grid, SR_MICRO_THREADS, 0, mp); break;
            case 1: SR_LAUNCH(c, (k_micro<true, true>), grid, SR_MICRO_THREADS, 0, mp); break;
            case 2: SR_LAUNCH(c, (k_micro<false, false>), grid, SR_MICRO_THREADS, 0, mp); break;
            default: SR_LAUNCH(c, (k_micro<true, false>), grid, SR_MICRO_THREADS, 0, mp); break;
        }
    }
    if (extra) {
        SrExtraParams ep;
        memset(&ep, 0, sizeof(ep));
        ep.lines = tp.lines; ep.points = tp.points;
        ep.nlines = tp.nlines; ep.npoints = tp.npoints; ep.ntris = tp.ntris;
        ep.width = fb->width; ep.height = fb->height; ep.ntx = fb->ntx;
        ep.shard_rank = c->shard_rank; ep.shard_world = c->shard_world;
        ep.vis = fb->vis_buf->as<unsigned long long>();
        if (tp.nlines) SR_LAUNCH(c, k_lines_vis, ceil_div(tp.nlines, 128), 128, 0, ep);
        if (tp.npoints) SR_LAUNCH(c, k_points_vis, ceil_div(tp.npoints, 128), 128, 0, ep);
    }
    record(c, 5);
    if (!c->ev_front) SR_CUDA(cudaEventCreateWithFlags(&c->ev_front, cudaEventDisableTiming));
    SR_CUDA(cudaEventRecord(c->ev_front, c->stream));
    c->ev_front_valid = true;
    SR_LAUNCH(c, k_tile_offsets, 1, SR_OFFSETS_THREADS, 0, count->as<uint32_t>(), ntiles, off->as<uint32_t>(), count->as<uint32_t>());
    if (!c->pinned) SR_CUDA(cudaHostAlloc((void **)&c->pinned, 64, cudaHostAllocDefault));
    SR_CUDA(cudaMemcpyAsync(&c->pinned[0], off->as<uint32_t>() + ntiles, 4, cudaMemcpyDeviceToHost, c->stream));
    SR_CUDA(cudaEventCreateWithFlags(&q->counted, cudaEventDisableTiming));
    SR_CUDA(cudaEventRecord(q->counted, c->stream));
    if (!c->list_arena) {
        c->list_cap = 1u << 20;
        SR_TRY(c->alloc((size_t)c->list_cap * 4, &c->list_arena));
    }
    q->capacity = c->list_cap;
    q->fs = fs; q->owned = owned; q->ntiles = ntiles;
    q->count = count; q->off = off; q->lcount = lcount; q->lids = lids; q->lrects = lrects;
    q->keep = keep;
    q->keep.push_back(c->list_arena);
    if (use_micro) q->keep.push_back(fb->vis_buf);
    q->fb = fb;
    memset(&q->op, 0, sizeof(q->op));
    q->op.tris = tp.tris;
    q->op.ntris = tp.ntris;
    q->op.vis = use_micro ? fb->vis_buf->as<unsigned long long>() : nullptr;
    q->op.reset_vis = use_micro ? 1u : 0u;
    q->op.tile_off = off->as<uint32_t>();
    q->op.list = c->list_arena->as<uint32_t>();
    q->op.list_capacity = c->list_cap;
    q->op.ntiles = ntiles;
    q->op.fb = fb->view();
    q->op.shard_rank = c->shard_rank; q->op.shard_world = c->shard_world;
    q->op.fs = tp.fs;
    if (extra) {
        q->op.lines = tp.lines; q->op.points = tp.points;
        q->op.nlines = tp.nlines; q->op.npoints = tp.npoints;
        q->op.line_base = tp.line_base; q->op.point_base = tp.point_base;
    }
    SR_TRY(launch_opaque_pass(c, q.get()));
    c->last_off = off; c->last_list = c->list_arena; c->last_ntiles = ntiles; c->last_micro_area = use_micro ? micro_area : 0u;
    fb->pending_clear = false;
    c->pending = q.release();
    return SR_OK;
}
static int launch_vertex(sr_context *c, sr_draw *d, uint32_t vs, const SrVsConst &vc, const SrVertexSpan &span) {
    if (span.end <= span.begin) return SR_OK;
    SrMeshView mv;
    mv.planes = d->mesh_planes->as<float>();
    mv.pstride = d->mesh_pstride;
    mv.nverts = d->mesh_nverts;
    mv.vin = d->vin;
    float4 *pos = d->indexed.pos->as<float4>(), *attr = d->indexed.attr->as<float4>();
    if (span.mask != nullptr) {  // one warp per 1024 vertices
        const uint32_t grid = ceil_div(ceil_div(span.end - span.begin, 1024), 8);
        if (vs == SR_VS_SUZANNE) SR_LAUNCH(c, k_vertex_marked<SR_VS_SUZANNE>, grid, 256, 0, vc, mv, pos, attr, d->indexed.np, span);
        else if (vs == SR_VS_FULL_EXAMPLE) SR_LAUNCH(c, k_vertex_marked<SR_VS_FULL_EXAMPLE>, grid, 256, 0, vc, mv, pos, attr, d->indexed.np, span);
        else return sr_fail(SR_ERR_INVALID_ARGUMENT, "vertex shader %u", vs);
        return SR_OK;
    }
    const uint32_t grid = ceil_div(span.end - span.begin, 256);
    if (vs == SR_VS_SUZANNE) SR_LAUNCH(c, k_vertex<SR_VS_SUZANNE>, grid, 256, 0, vc, mv, pos, attr, d->indexed.np, span);
    else if (vs == SR_VS_FULL_EXAMPLE) SR_LAUNCH(c, k_vertex<SR_VS_FULL_EXAMPLE>, grid, 256, 0, vc, mv, pos, attr, d->indexed.np, span);
    else return sr_fail(SR_ERR_INVALID_ARGUMENT, "vertex shader %u", vs);
    return SR_OK;
}
// a recorded (lazy) vertex stage becomes vertices: the whole mesh, for every consumer but the range-sharded fragment stage
static int resolve_tri_count(sr_draw *d) {
    if (!d->tri_count_dev) return SR_OK;
    sr_context *c = d->pipeline->ctx;
    SR_CUDA(cudaSetDevice(c->device));
    uint32_t host[2] = {0, 0};
    SR_CUDA(cudaMemcpyAsync(host, d->tri_count_dev->ptr, 8, cudaMemcpyDeviceToHost, c->stream));
    SR_CUDA(cudaStreamSynchronize(c->stream));
    d->gen[2].n = (uint64_t)host[0] * 3;  // (the streams stay allocated for the worst case; the NaN tail is simply no longer looked at)
    d->tri_literal_total = host[1];
    d->tri_count_dev.reset();
    return SR_OK;
}
static int materialize_vertices(sr_draw *d) {
    if (!d->vertex_lazy) return SR_OK;
    d->vertex_lazy = false;
    const SrVertexSpan whole = {0, d->mesh_nverts, nullptr, 0, 0, nullptr};
    return launch_vertex(d->pipeline->ctx, d, d->lazy_vs, d->lazy_vc, whole);
}
// The opaque triangle path of a range-sharded frame (include/softrender_b200.h).  Per frame n of a lane:
//   a. (lazy vertex stage) shade the vertex range MY triangles reference
//   b. wait until every peer has finished reading my keys of frame n-1            (done words)
//   c. hand the keys of the tiles I do not own back "far" (my own tiles were reset by my resolve); rank 0 also pre-fills the
//      framebuffer pixels of the tiles it does not own with the clear, so that their owners only send what was drawn on
//   d. k_micro over MY triangle range, all tiles, global triangle ids; large triangles -> lists of all tiles -> PHASE 1 sweep
//   e. publish "frame n ready" to every peer                                      (ready words)
//   f. wait until every peer's keys of frame n are ready
//   g. merge the peers' keys of my tiles over NVLink (k_shard_merge, or inside the resolve: sr_shard::fused_merge) and mark the
//      vertices the winners reference; (lazy vertex stage) shade exactly those
//   h. resolve my tiles + write-back (to rank 0's framebuffer)
//   i. publish "frame n done"
// All of it is enqueued on the context's stream; nothing synchronises with the host (except, once per mesh, the reduction that
// finds the vertex range of the rank's triangles).
static int opaque_triangles_ranged(sr_context *c, sr_framebuffer *fb, const SrTileParams &tp, uint32_t cull, uint32_t fs, uint32_t owned,
                                   const std::vector<Buf> &keep, sr_draw *d) {
    SrRange nvtx("softrender: range-sharded opaque pass");
    sr_shard *sh = c->shard;
    const uint32_t lane = c->shard_lane, ntiles = fb->ntx * fb->nty;
    const uint32_t t0 = (uint32_t)((uint64_t)tp.ntris * sh->rank / sh->world), t1 = (uint32_t)((uint64_t)tp.ntris * (sh->rank + 1) / sh->world);
    const bool lazy = d && d->vertex_lazy;
    uint32_t vlo = 0, vhi = 0, nchunk = 0;
    bool chunked = false;
    Buf mark, chunk_list, blocks;
    if (lazy) {
        DevBuf *ib = d->indices.get();
        if (!(ib->vr_t1 > ib->vr_t0 && ib->vr_t0 == t0 && ib->vr_t1 == t1)) {  // once per mesh and (rank, world)
            Buf mm;
            SR_TRY(c->alloc(8, &mm));
            const uint32_t init[2] = {0xFFFFFFFFu, 0u};
            SR_CUDA(cudaMemcpyAsync(mm->ptr, init, 8, cudaMemcpyHostToDevice, c->stream));
            if (t1 > t0) SR_LAUNCH(c, k_index_minmax, 148 * 8, 256, 0, d->indices->as<uint32_t>() + (uint64_t)t0 * 3, (uint64_t)(t1 - t0) * 3, mm->as<uint32_t>());
            uint32_t out[2] = {0, 0};
            SR_CUDA(cudaMemcpyAsync(out, mm->ptr, 8, cudaMemcpyDeviceToHost, c->stream));
            SR_CUDA(cudaStreamSynchronize(c->stream));
            ib->vr_t0 = t0; ib->vr_t1 = t1;
            ib->vr_lo = t1 > t0 ? out[0] : 0; ib->vr_hi = t1 > t0 ? std::min<uint64_t>(out[1], d->mesh_nverts) : 0;
        }
        vlo = ib->vr_lo; vhi = ib->vr_hi;
        record(c, 0);
        // ---- chunk-culled front end: which 1024-triangle chunks can reach MY tile rows (row y belongs to rank y % world)? ----
        // Opt-in (SR_SHARD_CHUNKS=1): measured on configs 3 / 4 it LOSES to plain triangle ranges (N = 2, config 4: 2.63 against 1.94 ms) --
        // those index buffers run in long strips along wavy grid rows, so a chunk reaches 3-4 tile rows, nearly every rank keeps nearly
        // every chunk and the vertex blocks of a kept chunk cover whole grid rows.  It needs index buffers whose chunks are compact on
        // screen; kept because it is exact and tested (tests/test_gpu_range_shard.py), not because it is the default.
        const bool no_chunks = getenv("SR_SHARD_CHUNKS") == nullptr;
        nchunk = ceil_div(tp.ntris, SR_CHUNK_TRIS);
        const uint32_t nblk = ceil_div(d->mesh_nverts, SR_VERTEX_BLOCK);
        if (!no_chunks && fb->nty >= sh->world) {
            if (!ib->chunk_vr || !ib->blk_aabb || ib->aabb_planes != d->mesh_planes->ptr) {  // static per mesh
                SR_TRY(c->alloc((size_t)nchunk * 8, &ib->chunk_vr));
                SR_TRY(c->alloc((size_t)nblk * 24, &ib->blk_aabb));
                SR_LAUNCH(c, k_chunk_vrange, nchunk, 256, 0, d->indices->as<uint32_t>(), tp.ntris, (uint32_t)SR_CHUNK_TRIS, ib->chunk_vr->as<uint2>());
                SR_LAUNCH(c, k_block_aabb, nblk, SR_VERTEX_BLOCK, 0, d->mesh_planes->as<float>(), d->mesh_pstride, d->mesh_nverts, ib->blk_aabb->as<float>());
                ib->aabb_planes = d->mesh_planes->ptr;
            }
            Buf rows, flag, pos;
            SR_TRY(c->alloc((size_t)nblk * 4, &rows));
            SR_TRY(c->alloc((size_t)nchunk * 4, &flag));
            SR_TRY(c->alloc((size_t)nchunk * 4, &pos));
            SR_TRY(c->alloc((size_t)nchunk * 4 + 4, &chunk_list));
            SR_TRY(c->alloc((size_t)nblk + 16, &blocks));
            SR_CUDA(cudaMemsetAsync(blocks->ptr, 0, (size_t)nblk + 16, c->stream));
            SR_LAUNCH(c, k_block_rows, ceil_div(nblk, 256), 256, 0, ib->blk_aabb->as<float>(), nblk, d->lazy_vc, fb->height, fb->nty, (uint32_t)SR_TILE_H,
                      rows->as<uint32_t>());
            SR_LAUNCH(c, k_chunk_select, ceil_div(nchunk, 256), 256, 0, ib->chunk_vr->as<uint2>(), rows->as<uint32_t>(), nchunk, sh->rank, sh->world,
                      flag->as<uint32_t>(), blocks->as<uint8_t>());
            SR_TRY(exclusive_scan_async(c, flag->as<uint32_t>(), nchunk, pos->as<uint32_t>(), 4));
            SR_LAUNCH(c, k_chunk_compact, ceil_div(nchunk, 256), 256, 0, flag->as<uint32_t>(), pos->as<uint32_t>(), nchunk, chunk_list->as<uint32_t>(),
                      chunk_list->as<uint32_t>() + nchunk);
            chunked = true;
        }
        if (chunked) {
            const SrVertexSpan mine = {0, d->mesh_nverts, nullptr, 0, 0, blocks->as<uint8_t>()};
            SR_TRY(launch_vertex(c, d, d->lazy_vs, d->lazy_vc, mine));                                                   // a (chunks)
            vlo = vhi = 0;
        } else {
            blocks.reset();
            const SrVertexSpan own = {vlo, vhi, nullptr, 0, 0, nullptr};
            SR_TRY(launch_vertex(c, d, d->lazy_vs, d->lazy_vc, own));                                                    // a (range)
        }
        record(c, 1);
        record(c, 2);
        record(c, 3);
        record(c, 4);
        const size_t mark_bytes = ((size_t)d->mesh_nverts + 127) / 128 * 16 + 16;  // one bit per vertex, whole 16-byte groups
        SR_TRY(c->alloc(mark_bytes, &mark));
        SR_CUDA(cudaMemsetAsync(mark->ptr, 0, mark_bytes, c->stream));
        // (the draw stays "lazy": only part of its vertices is shaded, any other consumer shades the whole mesh first)
    }
    const uint32_t n = ++sh->frame[lane];
    static const bool dbg = getenv("SR_SHARD_DEBUG") != nullptr;
    auto stamp = [&](int i) {
        if (!dbg) return;
        if (!c->ev_dbg[i]) cudaEventCreate(&c->ev_dbg[i]);
        cudaEventRecord(c->ev_dbg[i], c->stream);
    };
    if (dbg && c->ev_dbg_valid) {  // the previous frame of this context, printed once it has completed
        cudaEventSynchronize(c->ev_dbg[5]);
        float t[5] = {};
        for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&t[i], c->ev_dbg[i], c->ev_dbg[i + 1]);
        fprintf(stderr, "[shard rank %u lane %u frame %u] wait_ready %.3f  merge+mark %.3f  vertex(winners) %.3f  resolve %.3f  signal %.3f ms\n",
                sh->rank, lane, n - 1, t[0], t[1], t[2], t[3], t[4]);
    }
    unsigned long long *vis = sh->vis(sh->block, lane);
    // a wait gives up after 10 s (a peer died); SR_SHARD_TIMEOUT_MS shortens it for tools that serialise kernels (compute-sanitizer),
    // under which the ranks of a single-process group can never overlap and every wait runs into its timeout
    static const unsigned long long timeout_ns = [] {
        const char *e = getenv("SR_SHARD_TIMEOUT_MS");
        return (e && atoll(e) > 0 ? (unsigned long long)atoll(e) : 10000ull) * 1000ull * 1000ull;
    }();
    SrShardPeers ready_peers, done_peers;
    memset(&ready_peers, 0, sizeof(ready_peers));
    memset(&done_peers, 0, sizeof(done_peers));
    for (uint32_t p = 0; p < sh->world; ++p)
        if (p != sh->rank) {
            ready_peers.word[p] = sh->ready(sh->peer[p], lane);
            done_peers.word[p] = sh->done(sh->peer[p], lane);
        }
    SrTileOwners owners = sh->owners;
    owners.by_rows = chunked ? 1u : 0u;  // the resolve follows the front end: a rank resolves the rows it rasterised (hardly any key crosses NVLink)
    owners.ntx = fb->ntx; owners.world = sh->world;
    const bool owners_changed = sh->last_by_rows[lane] != (int)owners.by_rows;  // (tiles reset under the other function may be dirty)
    sh->last_by_rows[lane] = (int)owners.by_rows;
    if (n > 1) SR_LAUNCH(c, k_shard_wait, 1, 32, 0, sh->done(sh->block, lane), sh->world, sh->rank, n - 1, timeout_ns, sh->error());   // b
    SR_LAUNCH(c, k_vis_clear_foreign, ntiles, 256, 0, vis, fb->view(), owners, sh->rank, (n == 1 || owners_changed) ? 1u : 0u, 0u);     // c
    if (sh->rank == 0) {  // the framebuffer pre-fill is pure HBM writes, k_micro is latency bound: they share the GPU well
        if (!c->aux) {
            SR_CUDA(cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking));
            SR_CUDA(cudaEventCreateWithFlags(&c->ev_aux[0], cudaEventDisableTiming));
            SR_CUDA(cudaEventCreateWithFlags(&c->ev_aux[1], cudaEventDisableTiming));
        }
        SR_CUDA(cudaEventRecord(c->ev_aux[0], c->stream));
        SR_CUDA(cudaStreamWaitEvent(c->aux, c->ev_aux[0], 0));
        k_fb_fill_foreign<<<ntiles, 256, 0, c->aux>>>(fb->view(), owners, sh->rank);
        c->launches++;
        SR_CUDA(cudaGetLastError());
        SR_CUDA(cudaEventRecord(c->ev_aux[1], c->aux));
    }
    record(c, 7);
    Buf count, off, lcount, lids, lrects;
    SR_TRY(c->alloc((size_t)(ntiles + 1) * 4, &count));
    SR_TRY(c->alloc((size_t)(ntiles + 1) * 4, &off));
    SR_TRY(c->alloc(4, &lcount));
    const uint32_t lcap = chunked ? tp.ntris : std::max(t1 - t0, 1u);  // (chunks: the rank's triangle count is only known on the device)
    SR_TRY(c->alloc((size_t)lcap * 4, &lids));
    SR_TRY(c->alloc((size_t)lcap * 4, &lrects));
    SR_CUDA(cudaMemsetAsync(count->ptr, 0, (size_t)(ntiles + 1) * 4, c->stream));
    SR_CUDA(cudaMemsetAsync(lcount->ptr, 0, 4, c->stream));
    auto q = std::make_unique<PendingOpaque>();
    SrMicroParams mp;
    memset(&mp, 0, sizeof(mp));
    mp.src = tp.tris;
    mp.ntris = tp.ntris;
    mp.tri_begin = t0; mp.tri_end = t1;
    mp.cull = cull;
    mp.width = fb->width; mp.height = fb->height; mp.ntx = fb->ntx; mp.nty = fb->nty;
    mp.shard_rank = 0; mp.shard_world = 1;  // every tile: ownership only matters from the merge on
    mp.micro_area = c->micro_auto ? sr_micro_area_for(tp.ntris) : c->micro_area;
    mp.vis = vis;
    mp.large_count = lcount->as<uint32_t>();
    mp.large_ids = lids->as<uint32_t>();
    mp.large_rects = lrects->as<uint32_t>();
    mp.tile_count = count->as<uint32_t>();
    if (chunked) {                                                                                                       // d (chunks)
        mp.tri_begin = 0; mp.tri_end = tp.ntris;
        const uint32_t grid = nchunk * (SR_CHUNK_TRIS / SR_MICRO_THREADS);
        const uint32_t *list = chunk_list->as<uint32_t>();
        if (c->micro_precheck & 2u) SR_LAUNCH(c, (k_micro_chunks<false, false>), grid, SR_MICRO_THREADS, 0, mp, list, list + nchunk);
        else SR_LAUNCH(c, (k_micro_chunks<false, true>), grid, SR_MICRO_THREADS, 0, mp, list, list + nchunk);
    } else if (t1 > t0) {                                                                                                // d (range)
        const uint32_t grid = ceil_div(t1 - t0, SR_MICRO_THREADS);
        if (c->micro_precheck & 2u) SR_LAUNCH(c, (k_micro<false, false>), grid, SR_MICRO_THREADS, 0, mp);
        else SR_LAUNCH(c, (k_micro<false, true>), grid, SR_MICRO_THREADS, 0, mp);
    }
    record(c, 5);
    SR_LAUNCH(c, k_tile_offsets, 1, SR_OFFSETS_THREADS, 0, count->as<uint32_t>(), ntiles, off->as<uint32_t>(), count->as<uint32_t>());
    SR_CUDA(cudaMemcpyAsync(&c->pinned[0], off->as<uint32_t>() + ntiles, 4, cudaMemcpyDeviceToHost, c->stream));
    SR_CUDA(cudaEventCreateWithFlags(&q->counted, cudaEventDisableTiming));
    SR_CUDA(cudaEventRecord(q->counted, c->stream));
    if (!c->list_arena) {
        c->list_cap = 1u << 20;
        SR_TRY(c->alloc((size_t)c->list_cap * 4, &c->list_arena));
    }
    q->capacity = c->list_cap;
    q->ranged = true;
    q->fs = fs; q->owned = owned; q->ntiles = ntiles;
    q->count = count; q->off = off; q->lcount = lcount; q->lids = lids; q->lrects = lrects;
    q->keep = keep;
    q->keep.push_back(c->list_arena);
    if (mark) q->keep.push_back(mark);
    if (chunk_list) q->keep.push_back(chunk_list);
    if (blocks) q->keep.push_back(blocks);
    q->fb = fb;
    SrOpaqueParams &op = q->op;
    memset(&op, 0, sizeof(op));
    op.tris = tp.tris;
    op.ntris = tp.ntris;
    op.vis = vis;
    op.tile_off = off->as<uint32_t>();
    op.list = c->list_arena->as<uint32_t>();
    op.list_capacity = c->list_cap;
    op.ntiles = ntiles;
    op.fb = fb->view();
    op.fs = tp.fs;
    op.shard_rank = 0; op.shard_world = 1;  // (d, continued) the large triangles of my range, every tile
    op.reset_vis = 0;
    SR_LAUNCH(c, k_large_fill, std::min<uint32_t>(ceil_div(lcap, 8), 148u * 4u), 256, 0, lcount->as<uint32_t>(), lids->as<uint32_t>(),
              lrects->as<uint32_t>(), fb->ntx, ntiles, 0u, 1u, off->as<uint32_t>(), count->as<uint32_t>(), c->list_arena->as<uint32_t>(), c->list_cap);
    SR_TRY(launch_opaque_sweep(c, ntiles, op));
    SR_LAUNCH(c, k_vis_rows_touched, ntiles, 256, 0, vis, fb->ntx, owners, sh->rank, sh->touched(sh->block, lane));
    if (sh->rank == 0) SR_CUDA(cudaStreamWaitEvent(c->stream, c->ev_aux[1], 0));  // the pre-fill is in place before "ready" goes out
    SR_LAUNCH(c, k_shard_signal, 1, 32, 0, ready_peers, sh->rank, n);                                                    // e
    stamp(0);
    SR_LAUNCH(c, k_shard_wait, 1, 32, 0, sh->ready(sh->block, lane), sh->world, sh->rank, n, timeout_ns, sh->error());   // f
    stamp(1);
    op.shard_rank = sh->rank; op.shard_world = sh->world;
    op.reset_vis = 1;
    op.owners = owners;
    op.elide_clear = sh->rank != 0 ? 1u : 0u;
    const bool fused = sh->fused_merge && !lazy;  // (the winners' vertices must be known before the resolve when they are shaded on demand)
    if (fused) {
        for (uint32_t p = 0; p < sh->world; ++p)
            if (p != sh->rank) op.peer_vis[op.npeers++] = sh->vis(sh->peer[p], lane);
    } else {                                                                                                             // g
        SrMergeParams mg;
        memset(&mg, 0, sizeof(mg));
        mg.vis = vis;
        for (uint32_t p = 0; p < sh->world; ++p)
            if (p != sh->rank) {
                mg.peer_touched[mg.npeers] = sh->touched(sh->peer[p], lane);
                mg.peer_vis[mg.npeers++] = sh->vis(sh->peer[p], lane);
            }
        mg.ntx = fb->ntx; mg.rank = sh->rank;
        mg.owners = owners;
        mg.shaded_blocks = chunked ? blocks->as<uint8_t>() : nullptr;
        mg.indices = tp.tris.indices;
        mg.ntris = tp.tris.n0;
        mg.mark = lazy ? mark->as<uint32_t>() : nullptr;
        mg.skip_lo = vlo; mg.skip_hi = vhi;
        SR_LAUNCH(c, k_shard_merge, ntiles, 256, 0, mg);
        stamp(2);
        if (lazy) {
            const SrVertexSpan winners = {0, d->mesh_nverts, mark->as<uint32_t>(), vlo, vhi, chunked ? blocks->as<uint8_t>() : nullptr};
            SR_TRY(launch_vertex(c, d, d->lazy_vs, d->lazy_vc, winners));
        }
    }
    if (fused) stamp(2);
    stamp(3);
    SR_TRY(launch_opaque_merge_fs(c, fs, ntiles, op, sh->merge_ctas));                                                   // h
    stamp(4);
    SR_LAUNCH(c, k_shard_signal, 1, 32, 0, done_peers, sh->rank, n);                                                     // i
    stamp(5);
    c->ev_dbg_valid = dbg;
    fb->pending_clear = false;
    c->pending = q.release();
    return SR_OK;
}

// per-tile lists of the large triangles + the tile kernel (both skip themselves if the lists do not fit the arena)
static int launch_bin_small(sr_context *c, PendingOpaque *q) {
    static bool configured[16] = {};
    if (!configured[c->device & 15]) {
        SR_CUDA(cudaFuncSetAttribute(k_bin_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (SR_BIN_SMALL_MAX_TILES + SR_BIN_SMALL_MAX_TRIS_DEV) * 4));
        configured[c->device & 15] = true;
    }
    // per-tile counters + one tile rectangle per triangle (whole rounds of SR_BIN_SMALL_THREADS)
    const size_t smem = ((size_t)q->ntiles + (size_t)ceil_div(q->mp.ntris, SR_BIN_SMALL_THREADS) * SR_BIN_SMALL_THREADS) * 4;
    SR_LAUNCH(c, k_bin_small, 1, SR_BIN_SMALL_THREADS, smem, q->mp, q->off->as<uint32_t>(), const_cast<uint32_t *>(q->op.list),
              q->capacity);
    return SR_OK;
}
static int launch_opaque_pass(sr_context *c, PendingOpaque *q) {
    if (q->small) {
        SR_TRY(launch_bin_small(c, q));
        return launch_opaque_fs(c, q->fs, q->owned, q->op);
    }
    if (q->op.ntris)
        SR_LAUNCH(c, k_large_fill, std::min<uint32_t>(ceil_div(q->op.ntris, 8), 148u * 4u), 256, 0, q->lcount->as<uint32_t>(),
                  q->lids->as<uint32_t>(), q->lrects->as<uint32_t>(), q->op.fb.ntx, q->ntiles, q->op.shard_rank, q->op.shard_world,
                  q->off->as<uint32_t>(), q->count->as<uint32_t>(), const_cast<uint32_t *>(q->op.list), q->capacity);
    return launch_opaque_fs(c, q->fs, q->owned, q->op);
}
static int fs_nk(uint32_t fs) {
    switch (fs) {
        case SR_FS_FLAT: return SrFsInfo<SR_FS_FLAT>::NK;
        case SR_FS_SUZANNE: return SrFsInfo<SR_FS_SUZANNE>::NK;
        case SR_FS_FULL_EXAMPLE: return SrFsInfo<SR_FS_FULL_EXAMPLE>::NK;
        case SR_FS_FULL_EXAMPLE_TEXTURED: return SrFsInfo<SR_FS_FULL_EXAMPLE_TEXTURED>::NK;
        case SR_FS_GREEN: return SrFsInfo<SR_FS_GREEN>::NK;
        case SR_FS_DISCARD_CHECKER: return SrFsInfo<SR_FS_DISCARD_CHECKER>::NK;
        case SR_FS_TEXTURE_UNLIT: return SrFsInfo<SR_FS_TEXTURE_UNLIT>::NK;
        case SR_FS_SUZANNE_GBUFFER: return SrFsInfo<SR_FS_SUZANNE_GBUFFER>::NK;
    }
    return -1;
}

// CUDA loads kernels lazily, and the first launch of a not-yet-loaded kernel can wait for the device to go idle.  A rank whose
// wait kernel is spinning for a peer must never be held up like that (in a single process it would be a deadlock until the
// wait's timeout), so every kernel a range-sharded frame can launch is loaded before the first such frame.
template <class K>
static void preload(K kernel) {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, kernel);
}
static void preload_ranged_kernels() {
    preload(k_shard_wait);
    preload(k_shard_signal);
    preload(k_vis_clear_foreign);
    preload(k_shard_merge);
    preload(k_micro_chunks<false, true>);
    preload(k_micro_chunks<false, false>);
    preload(k_chunk_vrange);
    preload(k_block_aabb);
    preload(k_block_rows);
    preload(k_chunk_select);
    preload(k_chunk_compact);
    preload(k_vis_rows_touched);
    preload(k_fb_fill_foreign);
    preload(k_vertex_marked<SR_VS_SUZANNE>);
    preload(k_vertex_marked<SR_VS_FULL_EXAMPLE>);
    preload(k_index_minmax);
    preload(k_micro<false, true>);
    preload(k_micro<false, false>);
    preload(k_tile_offsets);
    preload(k_large_fill);
    preload(k_tile_opaque<SR_FS_FLAT, false, 1>);
    preload(k_tile_opaque<SR_FS_FLAT, false, 2>);
    preload(k_tile_opaque<SR_FS_SUZANNE, false, 2>);
    preload(k_tile_opaque<SR_FS_FULL_EXAMPLE, false, 2>);
    preload(k_tile_opaque<SR_FS_FULL_EXAMPLE_TEXTURED, false, 2>);
    preload(k_tile_opaque<SR_FS_GREEN, false, 2>);
    preload(k_tile_opaque<SR_FS_TEXTURE_UNLIT, false, 2>);
    preload(k_tile_opaque<SR_FS_SUZANNE_GBUFFER, false, 2>);
    preload(k_vertex<SR_VS_SUZANNE>);
    preload(k_vertex<SR_VS_FULL_EXAMPLE>);
    preload(k_vertex_passthrough);
    preload(k_normalize);
    preload(k_aos_to_planes);
    preload(k_records_to_planes);
    preload(k_fb_fill);
    preload(k_scan_reduce);
    preload(k_scan_sums);
    preload(k_scan_apply);
    cudaGetLastError();
}

// =========================================================================================================
// C ABI
// =========================================================================================================
extern "C" {

const char *sr_last_error(void) { return g_last_error.c_str(); }
int sr_version(void) { return 100; }
int sr_tile_size(uint32_t *w, uint32_t *h) {
    if (w) *w = SR_TILE_W;
    if (h) *h = SR_TILE_H;
    return SR_OK;
}

// ---- shader registry -------------------------------------------------------------------------------------
int sr_registry_entry(uint32_t kind, uint32_t index, sr_shader_info *info) {
    if (!info) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    struct Row { uint32_t id, vin, nk, discards, tex; const char *name, *ref; };
    static const Row vs[] = {
        {SR_VS_PASSTHROUGH, 0, 0, 0, 0, "passthrough", "test shader: Vin = {x,y,z,w,k...} copied"},
        {SR_VS_SUZANNE, SrVsInfo<SR_VS_SUZANNE>::VIN, SrVsInfo<SR_VS_SUZANNE>::NK, 0, 0, "suzanne", "examples/suzanne.rs:123-141"},
        {SR_VS_FULL_EXAMPLE, SrVsInfo<SR_VS_FULL_EXAMPLE>::VIN, SrVsInfo<SR_VS_FULL_EXAMPLE>::NK, 0, 0, "full_example", "full_example/src/shaders.rs:8-31"},
    };
    static const Row gs[] = {
        {SR_GS_CLIP, 0, 0, 0, 0, "clip_primitives", "src/pipeline/stages/geometry.rs:261-336"},
        {SR_GS_FACE_NORMALS, 0, 8, 0, 0, "face_normals", "full_example/src/shaders.rs:63-89"},
        {SR_GS_VERTEX_NORMALS, 0, 8, 0, 0, "vertex_normals", "full_example/src/shaders.rs:35-61"},
        {SR_GS_CLIP_SH, 0, 0, 0, 0, "clip_sutherland_hodgman", "opt-in correct clipper, planes of src/geometry/clip.rs:33-63"},
    };
    static const Row fsr[] = {
        {SR_FS_FLAT, 0, SrFsInfo<SR_FS_FLAT>::NK, 0, 0, "flat", "test shader: colour = K[0..4)"},
        {SR_FS_SUZANNE, 0, SrFsInfo<SR_FS_SUZANNE>::NK, 0, 0, "suzanne_blinn_phong", "examples/suzanne.rs:147-183"},
        {SR_FS_FULL_EXAMPLE, 0, SrFsInfo<SR_FS_FULL_EXAMPLE>::NK, 0, 0, "full_example_4light", "full_example/src/shaders.rs:108-162"},
        {SR_FS_FULL_EXAMPLE_TEXTURED, 0, SrFsInfo<SR_FS_FULL_EXAMPLE_TEXTURED>::NK, 0, 1, "full_example_4light_textured",
         "full_example/src/shaders.rs:108-162 + texture.rs:47-84"},
        {SR_FS_GREEN, 0, SrFsInfo<SR_FS_GREEN>::NK, 0, 0, "green", "full_example/src/shaders.rs:102"},
        {SR_FS_DISCARD_CHECKER, 0, SrFsInfo<SR_FS_DISCARD_CHECKER>::NK, 1, 0, "discard_checker", "test shader: Fragment::Discard (fragment.rs:61-66)"},
        {SR_FS_TEXTURE_UNLIT, 0, SrFsInfo<SR_FS_TEXTURE_UNLIT>::NK, 0, 1, "texture_unlit", "texture(t, uv, filter, edge): src/texture.rs:14-18 + full_example/src/texture.rs:25-84"},
        {SR_FS_SUZANNE_GBUFFER, 0, SrFsInfo<SR_FS_SUZANNE_GBUFFER>::NK, 0, 0, "suzanne_gbuffer", "two outputs (colour, normal) for a two-plane texture buffer: texturebuffer.rs:129-147"},
    };
    static const Row bl[] = {
        {SR_BLEND_REPLACE, 0, 0, 0, 0, "replace", "Blend for (): src/color/blend.rs:28-31"},
        {SR_BLEND_ALPHA_OVER, 0, 0, 0, 0, "alpha_over", "full_example/src/color.rs:5-17"},
        {SR_BLEND_ADDITIVE, 0, 0, 0, 0, "additive", "GenericBlend::new(|a, b| a + b): src/color/blend.rs:57-76"},
    };
    const Row *rows = nullptr;
    uint32_t n = 0;
    switch (kind) {
        case SR_REGISTRY_VERTEX: rows = vs; n = sizeof(vs) / sizeof(vs[0]); break;
        case SR_REGISTRY_GEOMETRY: rows = gs; n = sizeof(gs) / sizeof(gs[0]); break;
        case SR_REGISTRY_FRAGMENT: rows = fsr; n = sizeof(fsr) / sizeof(fsr[0]); break;
        case SR_REGISTRY_BLEND: rows = bl; n = sizeof(bl) / sizeof(bl[0]); break;
        default: return sr_fail(SR_ERR_INVALID_ARGUMENT, "registry kind %u", kind);
    }
    if (index >= n) return sr_fail(SR_ERR_INVALID_ARGUMENT, "registry index %u of %u", index, n);
    memset(info, 0, sizeof(*info));
    info->id = rows[index].id; info->vin_floats = rows[index].vin; info->nk = rows[index].nk;
    info->discards = rows[index].discards; info->needs_texture = rows[index].tex;
    snprintf(info->name, sizeof(info->name), "%s", rows[index].name);
    snprintf(info->reference, sizeof(info->reference), "%s", rows[index].ref);
    return SR_OK;
}

// ---- context ---------------------------------------------------------------------------------------------
int sr_context_create(int device, sr_context **out) {
    if (!out) return sr_fail(SR_ERR_INVALID_ARGUMENT, "out is null");
    int count = 0;
    SR_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return sr_fail(SR_ERR_INVALID_ARGUMENT, "device %d of %d", device, count);
    SR_CUDA(cudaSetDevice(device));
    auto *c = new sr_context();
    c->device = device;
    if (cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || c->sm_count <= 0) c->sm_count = 148;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete c;
        return sr_fail(SR_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    // pinned words for device->host counters: allocated here, not lazily -- a page-locked allocation may synchronise with
    // the device, which must not happen while a peer's wait kernel is spinning (range-sharded frames)
    e = cudaHostAlloc((void **)&c->pinned, 64, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaStreamDestroy(c->stream);
        delete c;
        return sr_fail(SR_ERR_CUDA, "cudaHostAlloc: %s", cudaGetErrorString(e));
    }
    *out = c;
    return SR_OK;
}
int sr_context_destroy(sr_context *c) {
    settle(c);
    if (!c) return SR_OK;
    if (c->closed) return sr_fail(SR_ERR_INVALID_STATE, "context already destroyed");
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    // the context's own buffers first (they hold references to it), then the cache, stream and events; the object itself
    // goes when the last buffer of a still-living child (framebuffer, mesh, draw ...) has been released
    c->list_arena.reset();
    c->last_off.reset();
    c->last_list.reset();
    for (auto &a : c->ord_arena) a.reset();
    c->zero_off.reset();
    c->closed = true;
    for (auto &kv : c->free_list) cudaFree(kv.second);
    c->free_list.clear();
    for (auto &e : c->ev)
        if (e) { cudaEventDestroy(e); e = nullptr; }
    if (c->ev_front) { cudaEventDestroy(c->ev_front); c->ev_front = nullptr; }
    cudaStreamDestroy(c->stream);
    c->stream = nullptr;
    if (c->aux) { cudaStreamDestroy(c->aux); c->aux = nullptr; }
    for (auto &e : c->ev_aux)
        if (e) { cudaEventDestroy(e); e = nullptr; }
    if (c->pinned) { cudaFreeHost(c->pinned); c->pinned = nullptr; }
    ctx_unref(c);
    return SR_OK;
}
int sr_context_synchronize(sr_context *c) {
    SR_TRY(settle(c));
    if (!c) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null context");
    SR_CUDA(cudaStreamSynchronize(c->stream));
    return SR_OK;
}
void *sr_context_stream(sr_context *c) { return c ? (void *)c->stream : nullptr; }
int sr_context_set_tile_shard(sr_context *c, uint32_t rank, uint32_t world) {
    if (!c || world == 0 || rank >= world) return sr_fail(SR_ERR_INVALID_ARGUMENT, "bad shard %u/%u", rank, world);
    c->shard_rank = rank;
    c->shard_world = world;
    return SR_OK;
}
int sr_context_set_micro(sr_context *c, uint32_t area, uint32_t min_triangles, uint32_t precheck) {
    if (!c) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null context");
    if (area != SR_MICRO_AREA_AUTO && area > SR_MICRO_AREA_MAX) return sr_fail(SR_ERR_INVALID_ARGUMENT, "micro area %u > %u", area, SR_MICRO_AREA_MAX);
    c->micro_auto = area == SR_MICRO_AREA_AUTO;
    if (!c->micro_auto) c->micro_area = area;
    c->micro_min_tris = min_triangles;
    c->micro_precheck = precheck & 3u;
    return SR_OK;
}
int sr_context_set_stage_timing(sr_context *c, int enable) {
    if (!c) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null context");
    c->stage_timing = enable != 0;
    if (!c->stage_timing)
        for (bool &v : c->ev_valid) v = false;
    return SR_OK;
}
int sr_context_stage_timestamps(sr_context *c, void *base_event, float ms[8]) {
    if (!c || !base_event || !ms) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    SR_TRY(settle(c));
    SR_CUDA(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 8; ++i) {
        ms[i] = -1.0f;
        if (c->ev_valid[i]) cudaEventElapsedTime(&ms[i], (cudaEvent_t)base_event, c->ev[i]);
    }
    return SR_OK;
}
int sr_context_wait_for(sr_context *waiter, sr_context *other, uint32_t point) {
    if (!waiter || !other || point > 1) return sr_fail(SR_ERR_INVALID_ARGUMENT, "bad wait");
    if (waiter == other) return SR_OK;  // one stream: already ordered
    if (point == 1) {
        if (other->ev_front_valid) {
            SR_CUDA(cudaSetDevice(waiter->device));
            SR_CUDA(cudaStreamWaitEvent(waiter->stream, other->ev_front, 0));
        }
        return SR_OK;
    }
    // "everything enqueued so far" includes a pass that skipped itself on the device because its tile lists overflowed: it is
    // re-enqueued here (settle), before the event, so that the waiter can never start ahead of the replay
    SR_TRY(settle(other));
    cudaEvent_t ev;
    SR_CUDA(cudaSetDevice(other->device));
    SR_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    SR_CUDA(cudaEventRecord(ev, other->stream));
    SR_CUDA(cudaSetDevice(waiter->device));
    SR_CUDA(cudaStreamWaitEvent(waiter->stream, ev, 0));
    SR_CUDA(cudaEventDestroy(ev));  // released once the wait has been satisfied
    return SR_OK;
}
int sr_context_set_list_capacity(sr_context *c, uint32_t entries) {
    if (!c || entries == 0) return sr_fail(SR_ERR_INVALID_ARGUMENT, "bad capacity");
    SR_TRY(settle(c));
    SR_CUDA(cudaSetDevice(c->device));
    Buf arena;
    SR_TRY(c->alloc((size_t)entries * 4, &arena));
    c->list_arena = arena;
    c->list_cap = entries;
    for (int k = 0; k < 3; ++k) {  // the ordered path's group-list arenas follow
        SR_TRY(c->alloc((size_t)entries * 4, &c->ord_arena[k]));
        c->ord_cap[k] = entries;
    }
    return SR_OK;
}
int sr_context_list_capacity(sr_context *c, uint32_t *entries) {
    if (!c || !entries) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    SR_TRY(settle(c));
    *entries = c->list_cap;
    return SR_OK;
}
int sr_context_ordered_list_capacity(sr_context *c, uint32_t entries[3]) {
    if (!c || !entries) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    SR_TRY(settle(c));
    for (int k = 0; k < 3; ++k) entries[k] = c->ord_cap[k];
    return SR_OK;
}
int sr_context_launch_count(sr_context *c, uint64_t *out) {
    if (!c || !out) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    *out = c->launches;
    return SR_OK;
}
int sr_context_stage_times(sr_context *c, sr_stage_times *out) {
    SR_TRY(settle(c));
    if (!c || !out) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    SR_CUDA(cudaStreamSynchronize(c->stream));
    sr_stage_times t = {};
    auto span = [&](int a, int b, float *dst) {
        if (c->ev_valid[a] && c->ev_valid[b]) cudaEventElapsedTime(dst, c->ev[a], c->ev[b]);
    };
    span(0, 1, &t.vertex_ms);
    span(1, 2, &t.geometry_ms);
    span(3, 4, &t.bin_ms);
    span(4, 7, &t.vis_init_ms);
    span(7, 5, &t.micro_ms);
    span(5, 6, &t.raster_ms);
    t.total_ms = t.vertex_ms + t.geometry_ms + t.bin_ms + t.vis_init_ms + t.micro_ms + t.raster_ms;
    *out = t;
    return SR_OK;
}

// ---- framebuffer -----------------------------------------------------------------------------------------
int sr_framebuffer_create(sr_context *c, uint32_t width, uint32_t height, uint32_t format, sr_framebuffer **out) {
    if (!c || !out) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (format > SR_FB_TEXTURE_2xRGBAF32_DF32) return sr_fail(SR_ERR_INVALID_ARGUMENT, "unknown format %u", format);
    if (width > 256u * SR_TILE_W || height > 256u * SR_TILE_H || width > 65535u || height > 65535u)
        return sr_fail(SR_ERR_UNSUPPORTED, "framebuffer %ux%u exceeds %ux%u", width, height, 256u * SR_TILE_W, 256u * SR_TILE_H);
    SR_CUDA(cudaSetDevice(c->device));
    auto fb = std::make_unique<sr_framebuffer>();
    fb->ctx = c;
    fb->width = width; fb->height = height; fb->format = format;
    fb->ntx = ceil_div(width, SR_TILE_W); fb->nty = ceil_div(height, SR_TILE_H);
    const uint64_t n = (uint64_t)width * height;
    fb->u8color = format == SR_FB_RGBAU8_DF32 || format == SR_FB_RGBAU8_DF32_S8;
    fb->soa = (format == SR_FB_TEXTURE_RGBAF32_DF32 || format == SR_FB_TEXTURE_RGBAF32_DF32_S8) ? 1u : format == SR_FB_TEXTURE_2xRGBAF32_DF32 ? 2u : 0u;
    SR_TRY(c->alloc(std::max<uint64_t>(n, 1) * fb->px_bytes(), &fb->aos_buf));
    fb->aos = fb->aos_buf->as<float>();
    fb->stencil_bytes = (format == SR_FB_RGBAF32_DF32_S8 || format == SR_FB_RGBAU8_DF32_S8 || format == SR_FB_TEXTURE_RGBAF32_DF32_S8) ? 1u
                        : format == SR_FB_RGBAF32_DF32_S16 ? 2u
                        : format == SR_FB_RGBAF32_DF32_S32 ? 4u : 0u;
    if (fb->stencil_bytes) SR_TRY(c->alloc(std::max<uint64_t>(n, 1) * fb->stencil_bytes, &fb->stencil_buf));
    // RenderBuffer::with_dimensions: Color::empty() (zeros), Depth::far(), stencil default -- recorded lazily
    fb->pending_clear = true;
    *out = fb.release();
    return SR_OK;
}
int sr_framebuffer_destroy(sr_framebuffer *fb) {
    if (!fb) return SR_OK;
    sr_context *c = fb->ctx;
    const bool peer = fb->is_peer;
    if (!c->closed) settle(c);
    if (peer && fb->aos) cudaIpcCloseMemHandle(fb->aos);
    delete fb;
    if (peer) ctx_unref(c);
    return SR_OK;
}
int sr_framebuffer_clear(sr_framebuffer *fb, const float color[4]) {
    if (!fb || !color) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    for (int i = 0; i < 4; ++i) fb->clear[i] = color[i];
    fb->pending_clear = true;  // produced on chip by the next draw (or materialised before a read-back)
    return SR_OK;
}
int sr_framebuffer_dimensions(const sr_framebuffer *fb, uint32_t *w, uint32_t *h) {
    if (!fb) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (w) *w = fb->width;
    if (h) *h = fb->height;
    return SR_OK;
}
int sr_framebuffer_download(sr_framebuffer *fb, void *dst, size_t nbytes) {
    SrRange nvtx("softrender: framebuffer download");
    if (!fb || !dst) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    const size_t need = (size_t)fb->width * fb->height * (fb->soa ? 20 : fb->px_bytes());
    if (nbytes != need) return sr_fail(SR_ERR_INVALID_ARGUMENT, "download size %zu, expected %zu", nbytes, need);
    SR_CUDA(cudaSetDevice(fb->ctx->device));
    SR_TRY(materialize_clear(fb));
    if (fb->soa == 2) return sr_fail(SR_ERR_UNSUPPORTED, "a texture buffer with two colour planes has no 20-byte pixel view: use sr_framebuffer_download_planes / _download_attachment");
    if (fb->soa) {  // the reference's PixelRead view of a texture buffer: gathered into the same 20-byte records
        Buf tmp;
        SR_TRY(fb->ctx->alloc(need, &tmp));
        SR_LAUNCH(fb->ctx, k_soa_to_aos, ceil_div((uint64_t)fb->width * fb->height, 256), 256, 0, fb->view(), tmp->as<float>());
        SR_CUDA(cudaMemcpyAsync(dst, tmp->ptr, need, cudaMemcpyDeviceToHost, fb->ctx->stream));
        SR_CUDA(cudaStreamSynchronize(fb->ctx->stream));
        return SR_OK;
    }
    SR_CUDA(cudaMemcpyAsync(dst, fb->aos, need, cudaMemcpyDeviceToHost, fb->ctx->stream));
    SR_CUDA(cudaStreamSynchronize(fb->ctx->stream));
    return SR_OK;
}
int sr_framebuffer_download_rgba8(sr_framebuffer *fb, uint8_t *dst, size_t nbytes, uint32_t order) {
    if (!fb || !dst) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    const uint64_t n = (uint64_t)fb->width * fb->height;
    if (nbytes != n * 4) return sr_fail(SR_ERR_INVALID_ARGUMENT, "download size %zu, expected %zu", nbytes, (size_t)(n * 4));
    if (order > 1) return sr_fail(SR_ERR_INVALID_ARGUMENT, "byte order %u", order);
    sr_context *c = fb->ctx;
    SR_CUDA(cudaSetDevice(c->device));
    SR_TRY(materialize_clear(fb));
    Buf packed;
    SR_TRY(c->alloc(n * 4, &packed));
    if (fb->soa) SR_LAUNCH(c, k_soa_to_rgba8, ceil_div(n, 256), 256, 0, fb->view(), order, packed->as<uint32_t>());
    else if (fb->u8color) SR_LAUNCH(c, k_fb8_to_rgba8, ceil_div(n, 256), 256, 0, reinterpret_cast<const uint2 *>(fb->aos), n, order, packed->as<uint32_t>());
    else SR_LAUNCH(c, k_fb_to_rgba8, ceil_div(ceil_div(n, 4), 256), 256, 0, fb->aos, n, order, packed->as<uint32_t>());
    SR_CUDA(cudaMemcpyAsync(dst, packed->ptr, n * 4, cudaMemcpyDeviceToHost, c->stream));
    SR_CUDA(cudaStreamSynchronize(c->stream));
    return SR_OK;
}
int sr_framebuffer_download_planes(sr_framebuffer *fb, void *color, float *depth, void *stencil) {
    if (!fb) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    sr_context *c = fb->ctx;
    SR_CUDA(cudaSetDevice(c->device));
    SR_TRY(materialize_clear(fb));
    const uint64_t n = (uint64_t)fb->width * fb->height;
    if (fb->soa) {  // the planes ARE the storage: straight copies
        if (color) SR_CUDA(cudaMemcpyAsync(color, fb->aos, n * 16, cudaMemcpyDeviceToHost, c->stream));
        if (depth) SR_CUDA(cudaMemcpyAsync(depth, fb->aos + 4 * n * fb->soa, n * 4, cudaMemcpyDeviceToHost, c->stream));
        if (stencil) {
            if (!fb->stencil_buf) return sr_fail(SR_ERR_INVALID_ARGUMENT, "framebuffer has no stencil attachment");
            SR_CUDA(cudaMemcpyAsync(stencil, fb->stencil_buf->ptr, n * fb->stencil_bytes, cudaMemcpyDeviceToHost, c->stream));
        }
        SR_CUDA(cudaStreamSynchronize(c->stream));
        return SR_OK;
    }
    Buf dc, dd;
    const size_t cbytes = fb->u8color ? 4 : 16;  // colour plane element: Vector4<u8> or Vector4<f32>
    if (color) SR_TRY(c->alloc(n * cbytes, &dc));
    if (depth) SR_TRY(c->alloc(n * 4, &dd));
    if ((color || depth) && fb->u8color)
        SR_LAUNCH(c, k_fb8_split, ceil_div(n, 256), 256, 0, reinterpret_cast<const uint2 *>(fb->aos), n, color ? dc->as<uint32_t>() : nullptr,
                  depth ? dd->as<float>() : nullptr);
    else if (color || depth)
        SR_LAUNCH(c, k_fb_split, ceil_div(n, 256), 256, 0, fb->aos, n, color ? dc->as<float>() : nullptr, depth ? dd->as<float>() : nullptr);
    if (color) SR_CUDA(cudaMemcpyAsync(color, dc->ptr, n * cbytes, cudaMemcpyDeviceToHost, c->stream));
    if (depth) SR_CUDA(cudaMemcpyAsync(depth, dd->ptr, n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (stencil) {
        if (!fb->stencil_buf) return sr_fail(SR_ERR_INVALID_ARGUMENT, "framebuffer has no stencil attachment");
        SR_CUDA(cudaMemcpyAsync(stencil, fb->stencil_buf->ptr, n * fb->stencil_bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    SR_CUDA(cudaStreamSynchronize(c->stream));
    return SR_OK;
}
int sr_framebuffer_upload_planes(sr_framebuffer *fb, const void *color, const float *depth, const void *stencil) {
    if (!fb) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (stencil && !fb->stencil_buf) return sr_fail(SR_ERR_INVALID_ARGUMENT, "framebuffer has no stencil attachment");  // before anything is enqueued
    sr_context *c = fb->ctx;
    SR_CUDA(cudaSetDevice(c->device));
    SR_TRY(materialize_clear(fb));
    const uint64_t n = (uint64_t)fb->width * fb->height;
    if (fb->soa) {
        if (color) SR_CUDA(cudaMemcpyAsync(fb->aos, color, n * 16, cudaMemcpyHostToDevice, c->stream));
        if (depth) SR_CUDA(cudaMemcpyAsync(fb->aos + 4 * n * fb->soa, depth, n * 4, cudaMemcpyHostToDevice, c->stream));
        if (stencil) SR_CUDA(cudaMemcpyAsync(fb->stencil_buf->ptr, stencil, n * fb->stencil_bytes, cudaMemcpyHostToDevice, c->stream));
        SR_CUDA(cudaStreamSynchronize(c->stream));
        return SR_OK;
    }
    Buf dc, dd;
    const size_t cbytes = fb->u8color ? 4 : 16;
    if (color) {
        SR_TRY(c->alloc(n * cbytes, &dc));
        SR_CUDA(cudaMemcpyAsync(dc->ptr, color, n * cbytes, cudaMemcpyHostToDevice, c->stream));
    }
    if (depth) {
        SR_TRY(c->alloc(n * 4, &dd));
        SR_CUDA(cudaMemcpyAsync(dd->ptr, depth, n * 4, cudaMemcpyHostToDevice, c->stream));
    }
    if ((color || depth) && fb->u8color)
        SR_LAUNCH(c, k_fb8_merge, ceil_div(n, 256), 256, 0, reinterpret_cast<uint2 *>(fb->aos), n, color ? dc->as<uint32_t>() : nullptr,
                  depth ? dd->as<float>() : nullptr);
    else if (color || depth)
        SR_LAUNCH(c, k_fb_merge, ceil_div(n, 256), 256, 0, fb->aos, n, color ? dc->as<float>() : nullptr, depth ? dd->as<float>() : nullptr);
    if (stencil) {
        if (!fb->stencil_buf) return sr_fail(SR_ERR_INVALID_ARGUMENT, "framebuffer has no stencil attachment");
        SR_CUDA(cudaMemcpyAsync(fb->stencil_buf->ptr, stencil, n * fb->stencil_bytes, cudaMemcpyHostToDevice, c->stream));
    }
    SR_CUDA(cudaStreamSynchronize(c->stream));
    return SR_OK;
}
int sr_framebuffer_get_pixel(sr_framebuffer *fb, uint32_t x, uint32_t y, float rgba[4], float *depth, uint32_t *stencil) {
    if (!fb) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (x >= fb->width || y >= fb->height)
        return sr_fail(SR_ERR_INVALID_PIXEL_COORDINATE, "pixel (%u,%u) outside %ux%u", x, y, fb->width, fb->height);
    sr_context *c = fb->ctx;
    SR_CUDA(cudaSetDevice(c->device));
    SR_TRY(materialize_clear(fb));
    float px[5] = {0, 0, 0, 0, 0};
    uint32_t px8[2] = {0, 0};
    const uint64_t idx = (uint64_t)x + (uint64_t)y * fb->width;
    const uint64_t npix = (uint64_t)fb->width * fb->height;
    if (fb->soa) {
        SR_CUDA(cudaMemcpyAsync(px, fb->aos + idx * 4, 16, cudaMemcpyDeviceToHost, c->stream));
        SR_CUDA(cudaMemcpyAsync(px + 4, fb->aos + 4 * npix * fb->soa + idx, 4, cudaMemcpyDeviceToHost, c->stream));
    } else if (fb->u8color) SR_CUDA(cudaMemcpyAsync(px8, reinterpret_cast<unsigned char *>(fb->aos) + idx * 8, 8, cudaMemcpyDeviceToHost, c->stream));
    else SR_CUDA(cudaMemcpyAsync(px, fb->aos + idx * 5, 20, cudaMemcpyDeviceToHost, c->stream));
    uint32_t s = 0;  // (little-endian: the low bytes of `s` receive a u8 / u16 element)
    if (stencil && fb->stencil_buf)
        SR_CUDA(cudaMemcpyAsync(&s, fb->stencil_buf->as<uint8_t>() + idx * fb->stencil_bytes, fb->stencil_bytes, cudaMemcpyDeviceToHost, c->stream));
    SR_CUDA(cudaStreamSynchronize(c->stream));
    if (fb->u8color) {  // channel values 0..255, as floats
        for (int i = 0; i < 4; ++i) px[i] = (float)((px8[0] >> (8 * i)) & 255u);
        memcpy(&px[4], &px8[1], 4);
    }
    if (rgba) memcpy(rgba, px, 16);
    if (depth) *depth = px[4];
    if (stencil) *stencil = s;
    return SR_OK;
}
// PixelWrite::set_pixel + FramebufferAccessorMut::{set_depth, set_stencil} (src/pixels/mod.rs:77-98, src/framebuffer/accessor.rs:18-60)
int sr_framebuffer_set_pixel(sr_framebuffer *fb, uint32_t x, uint32_t y, const float rgba[4], const float *depth, const uint32_t *stencil) {
    if (!fb) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (x >= fb->width || y >= fb->height)
        return sr_fail(SR_ERR_INVALID_PIXEL_COORDINATE, "pixel (%u,%u) outside %ux%u", x, y, fb->width, fb->height);
    if (stencil && !fb->stencil_buf) return sr_fail(SR_ERR_INVALID_ARGUMENT, "framebuffer has no stencil attachment");
    sr_context *c = fb->ctx;
    SR_CUDA(cudaSetDevice(c->device));
    SR_TRY(materialize_clear(fb));
    const uint64_t idx = (uint64_t)x + (uint64_t)y * fb->width;
    uint32_t packed = 0;
    if (fb->soa) {
        const uint64_t npix = (uint64_t)fb->width * fb->height;
        if (rgba) SR_CUDA(cudaMemcpyAsync(fb->aos + idx * 4, rgba, 16, cudaMemcpyHostToDevice, c->stream));
        if (depth) SR_CUDA(cudaMemcpyAsync(fb->aos + 4 * npix * fb->soa + idx, depth, 4, cudaMemcpyHostToDevice, c->stream));
    } else if (fb->u8color) {
        unsigned char *px = reinterpret_cast<unsigned char *>(fb->aos) + idx * 8;
        if (rgba) {
            for (int i = 0; i < 4; ++i) {
                if (!(rgba[i] >= 0.0f && rgba[i] <= 255.0f) || rgba[i] != (float)(uint32_t)rgba[i])
                    return sr_fail(SR_ERR_INVALID_ARGUMENT, "channel %d = %g is not a u8 value (RGBAu8Color target: pass 0..255)", i, (double)rgba[i]);
                packed |= (uint32_t)rgba[i] << (8 * i);
            }
            SR_CUDA(cudaMemcpyAsync(px, &packed, 4, cudaMemcpyHostToDevice, c->stream));
        }
        if (depth) SR_CUDA(cudaMemcpyAsync(px + 4, depth, 4, cudaMemcpyHostToDevice, c->stream));
    } else {
        if (rgba) SR_CUDA(cudaMemcpyAsync(fb->aos + idx * 5, rgba, 16, cudaMemcpyHostToDevice, c->stream));
        if (depth) SR_CUDA(cudaMemcpyAsync(fb->aos + idx * 5 + 4, depth, 4, cudaMemcpyHostToDevice, c->stream));
    }
    if (stencil) {
        const uint32_t smax = fb->stencil_bytes == 1 ? 0xFFu : fb->stencil_bytes == 2 ? 0xFFFFu : 0xFFFFFFFFu;
        if (*stencil > smax) return sr_fail(SR_ERR_INVALID_ARGUMENT, "stencil value %u does not fit the attachment's %u-bit type", *stencil, fb->stencil_bytes * 8);
        SR_CUDA(cudaMemcpyAsync(fb->stencil_buf->as<uint8_t>() + idx * fb->stencil_bytes, stencil, fb->stencil_bytes, cudaMemcpyHostToDevice, c->stream));
    }
    SR_CUDA(cudaStreamSynchronize(c->stream));  // the caller's buffers are copied before the call returns
    return SR_OK;
}
// the colour planes of a texture buffer by index (the named attachments of declare_texture_buffer!, texturebuffer.rs:110-117)
int sr_framebuffer_clear_attachment(sr_framebuffer *fb, uint32_t index, const float color[4]) {
    if (!fb || !color) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (index >= std::max(fb->soa, 1u)) return sr_fail(SR_ERR_INVALID_ARGUMENT, "colour attachment %u of %u", index, std::max(fb->soa, 1u));
    for (int i = 0; i < 4; ++i) (index ? fb->clear1 : fb->clear)[i] = color[i];
    fb->pending_clear = true;  // Framebuffer::clear takes the tuple of all colours (texturebuffer.rs:181-197): the other planes are cleared to their recorded colours
    return SR_OK;
}
int sr_framebuffer_download_attachment(sr_framebuffer *fb, uint32_t index, float *color) {
    if (!fb || !color) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (!fb->soa || index >= fb->soa) return sr_fail(SR_ERR_INVALID_ARGUMENT, "colour plane %u of a framebuffer with %u planes", index, fb->soa);
    sr_context *c = fb->ctx;
    SR_CUDA(cudaSetDevice(c->device));
    SR_TRY(materialize_clear(fb));
    const uint64_t n = (uint64_t)fb->width * fb->height;
    SR_CUDA(cudaMemcpyAsync(color, fb->aos + 4 * n * index, n * 16, cudaMemcpyDeviceToHost, c->stream));
    SR_CUDA(cudaStreamSynchronize(c->stream));
    return SR_OK;
}
int sr_framebuffer_enable_winner(sr_framebuffer *fb, int enable) {
    if (!fb) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (enable && !fb->winner_buf) {
        const uint64_t n = (uint64_t)fb->width * fb->height;
        SR_TRY(fb->ctx->alloc(std::max<uint64_t>(n, 1) * 4, &fb->winner_buf));
        SR_CUDA(cudaMemsetAsync(fb->winner_buf->ptr, 0, std::max<uint64_t>(n, 1) * 4, fb->ctx->stream));
    }
    fb->winner_enabled = enable != 0;
    return SR_OK;
}
int sr_framebuffer_download_winner(sr_framebuffer *fb, uint32_t *dst) {
    if (!fb || !dst) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (!fb->winner_buf) return sr_fail(SR_ERR_INVALID_STATE, "winner plane not enabled");
    SR_TRY(materialize_clear(fb));
    SR_CUDA(cudaMemcpyAsync(dst, fb->winner_buf->ptr, (size_t)fb->width * fb->height * 4, cudaMemcpyDeviceToHost, fb->ctx->stream));
    SR_CUDA(cudaStreamSynchronize(fb->ctx->stream));
    return SR_OK;
}
void *sr_framebuffer_device_ptr(sr_framebuffer *fb) { return fb ? (void *)fb->aos : nullptr; }

int sr_framebuffer_ipc_export(sr_framebuffer *fb, void *handle64) {
    if (!fb || !handle64) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    SR_CUDA(cudaSetDevice(fb->ctx->device));
    SR_TRY(materialize_clear(fb));
    SR_CUDA(cudaStreamSynchronize(fb->ctx->stream));
    cudaIpcMemHandle_t h;
    SR_CUDA(cudaIpcGetMemHandle(&h, fb->aos));
    memcpy(handle64, &h, 64);
    return SR_OK;
}
int sr_framebuffer_ipc_open(sr_context *c, const void *handle64, uint32_t width, uint32_t height, uint32_t format,
                            sr_framebuffer **out) {
    if (!c || !handle64 || !out) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (format != SR_FB_RGBAF32_DF32 && format != SR_FB_RGBAU8_DF32 && format != SR_FB_TEXTURE_RGBAF32_DF32)
        return sr_fail(SR_ERR_UNSUPPORTED, "peer framebuffers carry colour+depth only");
    SR_CUDA(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *p = nullptr;
    SR_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    auto *fb = new sr_framebuffer();
    fb->ctx = c;
    fb->width = width; fb->height = height; fb->format = format;
    fb->ntx = ceil_div(width, SR_TILE_W); fb->nty = ceil_div(height, SR_TILE_H);
    fb->aos = reinterpret_cast<float *>(p);
    fb->u8color = format == SR_FB_RGBAU8_DF32;
    fb->soa = format == SR_FB_TEXTURE_RGBAF32_DF32 ? 1u : 0u;
    fb->is_peer = true;
    fb->pending_clear = false;
    ++c->refs;  // owns no buffer of the context, so it holds the reference itself (released by sr_framebuffer_destroy)
    *out = fb;
    return SR_OK;
}

int sr_framebuffer_alias(sr_context *c, sr_framebuffer *src, sr_framebuffer **out) {
    if (!c || !src || !out) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (src->format != SR_FB_RGBAF32_DF32 && src->format != SR_FB_RGBAU8_DF32 && src->format != SR_FB_TEXTURE_RGBAF32_DF32)
        return sr_fail(SR_ERR_UNSUPPORTED, "aliased framebuffers carry colour+depth only");
    SR_TRY(materialize_clear(src));
    SR_CUDA(cudaSetDevice(src->ctx->device));
    SR_CUDA(cudaStreamSynchronize(src->ctx->stream));
    if (c->device != src->ctx->device) {
        SR_CUDA(cudaSetDevice(c->device));
        cudaError_t e = cudaDeviceEnablePeerAccess(src->ctx->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return sr_fail(SR_ERR_CUDA, "peer access %d -> %d: %s", c->device, src->ctx->device, cudaGetErrorString(e));
        cudaGetLastError();
    }
    auto *fb = new sr_framebuffer();
    fb->ctx = c;
    fb->width = src->width; fb->height = src->height; fb->format = src->format;
    fb->ntx = src->ntx; fb->nty = src->nty;
    fb->aos = src->aos;
    fb->u8color = src->u8color;
    fb->soa = src->soa;
    fb->aos_buf = src->aos_buf;  // shares ownership: the pixels outlive either handle
    fb->pending_clear = false;
    *out = fb;
    return SR_OK;
}

// ---- shard groups (range-sharded front end) --------------------------------------------------------------
// ownership pattern: rank 0 holds `k0` of every k0 + k*(world-1) consecutive tiles, every other rank k, spread evenly
static int shard_set_shares(sr_shard *sh, uint32_t k0, uint32_t k) {
    const uint32_t period = k0 + k * (sh->world - 1);
    if (k == 0 || period == 0 || period > SR_OWNER_PERIOD_MAX) return sr_fail(SR_ERR_INVALID_ARGUMENT, "tile shares %u:%u need a period of 1..%d", k0, k, SR_OWNER_PERIOD_MAX);
    uint32_t want[SR_SHARD_MAX_WORLD], got[SR_SHARD_MAX_WORLD] = {};
    for (uint32_t r = 0; r < sh->world; ++r) want[r] = r == 0 ? k0 : k;
    for (uint32_t i = 0; i < period; ++i) {  // slot i goes to the rank that is furthest behind its share
        uint32_t best = 0;
        double worst = -1e30;
        for (uint32_t r = 0; r < sh->world; ++r) {
            if (got[r] >= want[r]) continue;
            const double deficit = (double)want[r] * (i + 1) / period - got[r];
            if (deficit > worst + 1e-12) { worst = deficit; best = r; }
        }
        sh->owners.owner[i] = (uint8_t)best;
        ++got[best];
    }
    sh->owners.period = period;
    return SR_OK;
}
int sr_shard_create(sr_context *c, uint32_t width, uint32_t height, uint32_t lanes, sr_shard **out) {
    if (!c || !out || lanes == 0 || lanes > 8) return sr_fail(SR_ERR_INVALID_ARGUMENT, "bad shard group");
    if (c->shard_world < 2 || c->shard_world > SR_SHARD_MAX_WORLD)
        return sr_fail(SR_ERR_INVALID_STATE, "set the context's tile shard (2..%d ranks) before creating its shard group", SR_SHARD_MAX_WORLD);
    SR_CUDA(cudaSetDevice(c->device));
    preload_ranged_kernels();
    auto sh = std::make_unique<sr_shard>();
    sh->ctx = c;
    sh->rank = c->shard_rank; sh->world = c->shard_world;
    sh->width = width; sh->height = height; sh->lanes = lanes;
    sh->ntx = ceil_div(width, SR_TILE_W); sh->nty = ceil_div(height, SR_TILE_H);
    sh->vis_bytes = (size_t)sh->ntx * sh->nty * SR_TILE_PIXELS * 8;
    sh->lane_stride = sh->vis_bytes + 256 + (((size_t)sh->ntx * sh->nty * 4 + 255) & ~(size_t)255);  // keys, progress words, row masks
    sh->block_bytes = sh->lane_stride * lanes + 256;
    shard_set_shares(sh.get(), 1, 1);
    if (const char *env = getenv("SR_SHARD_SHARES")) {  // tuning experiments: "rank0_slots,other_slots[,merge_ctas]"
        unsigned a = 1, b = 1, m = 0;
        if (sscanf(env, "%u,%u,%u", &a, &b, &m) >= 2) {
            if (shard_set_shares(sh.get(), a, b) != SR_OK) return SR_ERR_INVALID_ARGUMENT;
            sh->merge_ctas = m & 15u;
            sh->fused_merge = (m & 16u) != 0;
        }
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, sh->block_bytes);
    if (e != cudaSuccess) return sr_fail(SR_ERR_OUT_OF_MEMORY, "cudaMalloc(%zu) for the exchange block failed: %s", sh->block_bytes, cudaGetErrorString(e));
    sh->block = reinterpret_cast<unsigned char *>(p);
    for (uint32_t l = 0; l < lanes; ++l) SR_CUDA(cudaMemsetAsync(sh->block + l * sh->lane_stride + sh->vis_bytes, 0, 256, c->stream));
    SR_CUDA(cudaMemsetAsync(sh->error(), 0, 256, c->stream));
    SR_CUDA(cudaStreamSynchronize(c->stream));
    ++c->refs;
    *out = sh.release();
    return SR_OK;
}
int sr_shard_export(sr_shard *sh, void *handle64) {
    if (!sh || !handle64) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    SR_CUDA(cudaSetDevice(sh->ctx->device));
    cudaIpcMemHandle_t h;
    SR_CUDA(cudaIpcGetMemHandle(&h, sh->block));
    memcpy(handle64, &h, 64);
    return SR_OK;
}
int sr_shard_connect(sr_shard *sh, const void *handles64, uint32_t count) {
    if (!sh || !handles64) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (count != sh->world) return sr_fail(SR_ERR_INVALID_ARGUMENT, "%u handles for a group of %u ranks", count, sh->world);
    if (sh->connected) return sr_fail(SR_ERR_INVALID_STATE, "shard group already connected");
    SR_CUDA(cudaSetDevice(sh->ctx->device));
    for (uint32_t p = 0; p < sh->world; ++p) {
        if (p == sh->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, reinterpret_cast<const unsigned char *>(handles64) + (size_t)p * 64, 64);
        void *ptr = nullptr;
        SR_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        sh->peer[p] = reinterpret_cast<unsigned char *>(ptr);
        sh->peer_ipc[p] = true;
    }
    sh->connected = true;
    return SR_OK;
}
int sr_shard_connect_local(sr_shard *sh, sr_shard *const *peers, uint32_t count) {
    if (!sh || !peers) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (count != sh->world) return sr_fail(SR_ERR_INVALID_ARGUMENT, "%u peers for a group of %u ranks", count, sh->world);
    if (sh->connected) return sr_fail(SR_ERR_INVALID_STATE, "shard group already connected");
    SR_CUDA(cudaSetDevice(sh->ctx->device));
    for (uint32_t p = 0; p < sh->world; ++p) {
        if (p == sh->rank) continue;
        const sr_shard *o = peers[p];
        if (!o || o->rank != p || o->world != sh->world || o->block_bytes != sh->block_bytes || o->lanes != sh->lanes)
            return sr_fail(SR_ERR_INVALID_ARGUMENT, "peer %u does not belong to this group", p);
        if (o->ctx->device != sh->ctx->device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(o->ctx->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return sr_fail(SR_ERR_CUDA, "peer access %d -> %d: %s", sh->ctx->device, o->ctx->device, cudaGetErrorString(e));
            cudaGetLastError();
        }
        sh->peer[p] = o->block;
    }
    sh->connected = true;
    return SR_OK;
}
int sr_context_attach_shard(sr_context *c, sr_shard *sh, uint32_t lane) {
    if (!c) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    SR_TRY(settle(c));
    if (!sh) { c->shard = nullptr; return SR_OK; }
    if (lane >= sh->lanes) return sr_fail(SR_ERR_INVALID_ARGUMENT, "lane %u of %u", lane, sh->lanes);
    if (c->device != sh->ctx->device) return sr_fail(SR_ERR_INVALID_STATE, "the shard group lives on device %d", sh->ctx->device);
    if (c->shard_rank != sh->rank || c->shard_world != sh->world) return sr_fail(SR_ERR_INVALID_STATE, "the context's tile shard is %u/%u, the group's %u/%u", c->shard_rank, c->shard_world, sh->rank, sh->world);
    c->shard = sh;
    c->shard_lane = lane;
    return SR_OK;
}
int sr_shard_status(sr_shard *sh, uint32_t *status) {
    if (!sh || !status) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    SR_CUDA(cudaSetDevice(sh->ctx->device));
    SR_CUDA(cudaMemcpy(status, sh->error(), 4, cudaMemcpyDeviceToHost));
    return SR_OK;
}
int sr_shard_destroy(sr_shard *sh) {
    if (!sh) return SR_OK;
    cudaSetDevice(sh->ctx->device);
    cudaDeviceSynchronize();
    for (uint32_t p = 0; p < SR_SHARD_MAX_WORLD; ++p)
        if (sh->peer[p] && sh->peer_ipc[p]) cudaIpcCloseMemHandle(sh->peer[p]);
    cudaFree(sh->block);
    ctx_unref(sh->ctx);
    delete sh;
    return SR_OK;
}

// ---- mesh ------------------------------------------------------------------------------------------------
int sr_mesh_create(sr_context *c, const float *vertices, uint64_t nverts, uint32_t vin_floats, const void *indices,
                   uint64_t nindices, uint32_t index_bytes, sr_mesh **out) {
    if (!c || !out || (!vertices && nverts) || (!indices && nindices)) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (vin_floats < 3 || vin_floats > 4 + SR_MAX_NK) return sr_fail(SR_ERR_INVALID_ARGUMENT, "vin_floats %u", vin_floats);
    if (index_bytes != 4 && index_bytes != 8) return sr_fail(SR_ERR_INVALID_ARGUMENT, "index_bytes %u", index_bytes);
    if (nverts > 0xFFFFFFF0ull || nindices > 0xFFFFFFF0ull * 3) return sr_fail(SR_ERR_UNSUPPORTED, "mesh too large for 32-bit ids");
    SR_CUDA(cudaSetDevice(c->device));
    auto m = std::make_unique<sr_mesh>();
    m->ctx = c;
    m->nverts = nverts;
    m->vin = vin_floats;
    m->nindices = nindices;
    m->pstride = std::max<uint64_t>((nverts + 3) & ~(uint64_t)3, 4);
    SR_TRY(c->alloc(m->pstride * vin_floats * 4, &m->planes));
    SR_TRY(c->alloc((nindices + 128) * 4, &m->indices));  // padded: the tile kernel bulk-copies whole groups (96 indices)
    if (nverts) {
        Buf tmp;
        SR_TRY(c->alloc(nverts * vin_floats * 4, &tmp));
        SR_CUDA(cudaMemsetAsync(m->planes->ptr, 0, m->pstride * vin_floats * 4, c->stream));
        SR_CUDA(cudaMemcpyAsync(tmp->ptr, vertices, nverts * vin_floats * 4, cudaMemcpyHostToDevice, c->stream));
        SR_LAUNCH(c, k_aos_to_planes, ceil_div(nverts * vin_floats, 256), 256, 0, tmp->as<float>(), nverts, vin_floats,
                  m->planes->as<float>(), m->pstride);
    }
    if (nindices) {
        if (index_bytes == 4) {
            SR_CUDA(cudaMemcpyAsync(m->indices->ptr, indices, nindices * 4, cudaMemcpyHostToDevice, c->stream));
        } else {
            Buf tmp;
            SR_TRY(c->alloc(nindices * 8, &tmp));
            SR_CUDA(cudaMemcpyAsync(tmp->ptr, indices, nindices * 8, cudaMemcpyHostToDevice, c->stream));
            SR_LAUNCH(c, k_narrow_indices, ceil_div(nindices, 256), 256, 0, tmp->as<uint64_t>(), nindices, m->indices->as<uint32_t>());
        }
    }
    // range-check the indices on the device (the reference would panic on an out-of-range index)
    uint32_t maxidx = 0;
    if (!c->pinned) SR_CUDA(cudaHostAlloc((void **)&c->pinned, 64, cudaHostAllocDefault));
    if (nindices) {
        Buf mx;
        SR_TRY(c->alloc(4, &mx));
        SR_CUDA(cudaMemsetAsync(mx->ptr, 0, 4, c->stream));
        SR_LAUNCH(c, k_index_max, std::min<uint32_t>(ceil_div(nindices, 256), 1184u), 256, 0, m->indices->as<uint32_t>(), nindices, mx->as<uint32_t>());
        // pinned destination: a pageable one would make this call block inside the driver until the upload has finished,
        // stalling every other thread's CUDA calls (frames in flight on other contexts)
        SR_CUDA(cudaMemcpyAsync(&c->pinned[2], mx->ptr, 4, cudaMemcpyDeviceToHost, c->stream));
    }
    // "buffers passed in are copied before return"
    SR_CUDA(cudaStreamSynchronize(c->stream));
    if (nindices) maxidx = c->pinned[2];
    if (nindices && maxidx >= nverts)
        return sr_fail(SR_ERR_INVALID_ARGUMENT, "index %u out of range (%llu vertices)", maxidx, (unsigned long long)nverts);
    *out = m.release();
    return SR_OK;
}
int sr_mesh_destroy(sr_mesh *m) {
    delete m;
    return SR_OK;
}

// ---- texture ---------------------------------------------------------------------------------------------
int sr_texture_create(sr_context *c, const uint8_t *rgba, uint32_t width, uint32_t height, sr_texture **out) {
    if (!c || !rgba || !out || !width || !height) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null or empty texture");
    SR_CUDA(cudaSetDevice(c->device));
    auto t = std::make_unique<sr_texture>();
    t->ctx = c;
    t->width = width; t->height = height;
    SR_TRY(c->alloc((size_t)width * height * 4, &t->rgba));
    SR_CUDA(cudaMemcpyAsync(t->rgba->ptr, rgba, (size_t)width * height * 4, cudaMemcpyHostToDevice, c->stream));
    SR_CUDA(cudaStreamSynchronize(c->stream));
    *out = t.release();
    return SR_OK;
}
int sr_texture_destroy(sr_texture *t) {
    delete t;
    return SR_OK;
}

// ---- pipeline --------------------------------------------------------------------------------------------
int sr_pipeline_create(sr_context *c, sr_framebuffer *fb, const sr_uniforms *u, sr_pipeline **out) {
    if (!c || !fb || !u || !out) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (fb->width == 0) return sr_fail(SR_ERR_INVALID_ARGUMENT, "Framebuffer must have a non-zero width");
    if (fb->height == 0) return sr_fail(SR_ERR_INVALID_ARGUMENT, "Framebuffer must have a non-zero height");
    if (fb->ctx != c) return sr_fail(SR_ERR_INVALID_STATE, "framebuffer belongs to another context (use sr_framebuffer_alias / sr_framebuffer_ipc_open)");
    auto *p = new sr_pipeline();
    p->ctx = c;
    p->fb = fb;
    p->uniforms = *u;
    *out = p;
    return SR_OK;
}
int sr_pipeline_destroy(sr_pipeline *p) {
    delete p;
    return SR_OK;
}
int sr_pipeline_set_uniforms(sr_pipeline *p, const sr_uniforms *u) {
    if (!p || !u) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    p->uniforms = *u;
    return SR_OK;
}
int sr_pipeline_set_framebuffer(sr_pipeline *p, sr_framebuffer *fb) {
    if (!p || !fb) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (fb->width == 0) return sr_fail(SR_ERR_INVALID_ARGUMENT, "Framebuffer must have a non-zero width");
    if (fb->height == 0) return sr_fail(SR_ERR_INVALID_ARGUMENT, "Framebuffer must have a non-zero height");
    if (fb->ctx != p->ctx) return sr_fail(SR_ERR_INVALID_STATE, "framebuffer belongs to another context (use sr_framebuffer_alias / sr_framebuffer_ipc_open)");
    p->fb = fb;
    p->stencil_test = SR_STENCIL_ALWAYS;  // with_framebuffer: stencil_config: Default::default() (mod.rs:137)
    p->stencil_op = SR_STENCIL_KEEP;
    return SR_OK;
}
int sr_pipeline_set_stencil_config(sr_pipeline *p, uint32_t test, uint32_t op) {
    if (!p || test > SR_STENCIL_NOT_EQUAL || op > SR_STENCIL_DECREMENT_SAT) return sr_fail(SR_ERR_INVALID_ARGUMENT, "bad stencil config");
    p->stencil_test = test;
    p->stencil_op = op;
    return SR_OK;
}
int sr_pipeline_bind_texture(sr_pipeline *p, sr_texture *t) {
    if (!p) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (t && t->ctx != p->ctx) return sr_fail(SR_ERR_INVALID_STATE, "texture belongs to another context");
    p->texture = t;
    p->fb_texture = nullptr;
    return SR_OK;
}
int sr_pipeline_bind_framebuffer_texture(sr_pipeline *p, sr_framebuffer *src) {
    if (!p) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (src && (src->width == 0 || src->height == 0)) return sr_fail(SR_ERR_INVALID_ARGUMENT, "texture source framebuffer is empty");
    p->fb_texture = src;
    p->fb_texture_plane = 0;
    p->texture = nullptr;
    return SR_OK;
}
// the same for colour plane `index` of a texture buffer (the named TextureBufferRef accessors, texturebuffer.rs:110-117)
int sr_pipeline_bind_framebuffer_attachment(sr_pipeline *p, sr_framebuffer *src, uint32_t index) {
    if (!p || !src) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (!src->soa || index >= src->soa) return sr_fail(SR_ERR_INVALID_ARGUMENT, "colour plane %u of a framebuffer with %u planes", index, src->soa);
    SR_TRY(sr_pipeline_bind_framebuffer_texture(p, src));
    p->fb_texture_plane = index;
    return SR_OK;
}
int sr_pipeline_set_sampler(sr_pipeline *p, uint32_t filter, uint32_t edge, const float *border_rgba) {
    if (!p || filter > SR_FILTER_BILINEAR || edge > SR_EDGE_BORDER) return sr_fail(SR_ERR_INVALID_ARGUMENT, "bad sampler state");
    p->tex_filter = filter;
    p->tex_edge = edge;
    for (int i = 0; i < 4; ++i) p->tex_border[i] = border_rgba ? border_rgba[i] : 0.0f;
    return SR_OK;
}

// ---- draw ------------------------------------------------------------------------------------------------
int sr_render_mesh(sr_pipeline *p, sr_mesh *m, uint32_t primitive, int has_sv, uint32_t sv, sr_draw **out) {
    if (!p || !m || !out) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (primitive < SR_POINT || primitive > SR_TRIANGLE) return sr_fail(SR_ERR_INVALID_ARGUMENT, "primitive %u", primitive);
    if (m->ctx != p->ctx) return sr_fail(SR_ERR_INVALID_STATE, "mesh and pipeline belong to different contexts (one stream orders their work)");
    if (m->nindices % primitive != 0)
        return sr_fail(SR_ERR_INVALID_ARGUMENT, "assertion failed: mesh.indices.len() %% T::num_vertices() == 0 (%llu %% %u)",
                       (unsigned long long)m->nindices, primitive);
    auto *d = new sr_draw();
    d->pipeline = p;
    d->primitive = primitive;
    d->has_stencil_value = has_sv != 0;
    d->stencil_value = has_sv ? sv : 0;  // stencil.unwrap_or_default()
    d->mesh_planes = m->planes;
    d->indices = m->indices;
    d->mesh_nverts = m->nverts;
    d->mesh_pstride = m->pstride;
    d->nindices = m->nindices;
    d->vin = m->vin;
    *out = d;
    return SR_OK;
}

static int vertex_stage(sr_draw *d, const sr_viewport *vp, uint32_t vs) {
    SrRange nvtx("softrender: vertex stage");
    if (!d) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (d->stage != STAGE_VERTEX || !d->mesh_planes) return sr_fail(SR_ERR_INVALID_STATE, "vertex stage already consumed");
    sr_pipeline *p = d->pipeline;
    sr_context *c = p->ctx;
    SR_CUDA(cudaSetDevice(c->device));
    uint32_t nk;
    switch (vs) {
        case SR_VS_PASSTHROUGH:
            if (d->vin < 4) return sr_fail(SR_ERR_INVALID_ARGUMENT, "passthrough needs >= 4 input floats");
            nk = d->vin - 4;
            break;
        case SR_VS_SUZANNE:
            if (d->vin != 6) return sr_fail(SR_ERR_INVALID_ARGUMENT, "suzanne vertex shader needs pos3+normal3");
            nk = 8;
            break;
        case SR_VS_FULL_EXAMPLE:
            if (d->vin != 8) return sr_fail(SR_ERR_INVALID_ARGUMENT, "full_example vertex shader needs pos3+normal3+uv2");
            nk = 10;
            break;
        default: return sr_fail(SR_ERR_INVALID_ARGUMENT, "unknown vertex shader %u", vs);
    }
    d->nk = nk;
    SR_TRY(alloc_stream(c, d->mesh_nverts, nk, &d->indexed));
    SrVsConst vc;
    fill_vs_const(p, vp, &vc);
    SrMeshView mv;
    mv.planes = d->mesh_planes->as<float>();
    mv.pstride = d->mesh_pstride;
    mv.nverts = d->mesh_nverts;
    mv.vin = d->vin;
    record(c, 0);
    static const bool no_lazy = getenv("SR_SHARD_NO_LAZY_VERTEX") != nullptr;  // A/B switch (tuning): shade the whole mesh on every rank
    if (!no_lazy && vp && c->shard && c->shard->connected && vs != SR_VS_PASSTHROUGH && d->primitive == SR_TRIANGLE && d->nindices / 3 >= 65536u) {
        d->vertex_lazy = true;  // decided by the fragment stage (opaque_triangles_ranged / materialize_vertices)
        d->lazy_vs = vs;
        d->lazy_vc = vc;
    } else if (d->mesh_nverts) {
        if (vs == SR_VS_PASSTHROUGH) {
            SR_LAUNCH(c, k_vertex_passthrough, ceil_div(d->mesh_nverts, 128), 128, 0, vc, mv, d->indexed.pos->as<float4>(),
                      d->indexed.attr->as<float4>(), d->indexed.np);
        } else {
            const SrVertexSpan whole = {0, d->mesh_nverts, nullptr, 0, 0, nullptr};
            SR_TRY(launch_vertex(c, d, vs, vc, whole));
        }
    }
    record(c, 1);
    record(c, 2);
    d->have_indexed = true;
    d->stage = vp ? STAGE_FRAGMENT : STAGE_GEOMETRY;
    return SR_OK;
}
int sr_vertex_run(sr_draw *d, uint32_t vs) { return vertex_stage(d, nullptr, vs); }
int sr_vertex_run_to_fragment(sr_draw *d, const sr_viewport *vp, uint32_t vs) {
    if (!vp) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null viewport");
    return vertex_stage(d, vp, vs);
}

}  // extern "C"
// count -> scan -> emit for the 0/1-output clippers
template <int NV>
static int clip_small(sr_context *c, const SrGeoIn &in, uint32_t nk, VertexStream *out) {
    const uint32_t n = in.ngen + in.nidx;
    if (n == 0) return alloc_stream(c, 0, nk, out);
    Buf cnt, off;
    SR_TRY(c->alloc((size_t)n * 4, &cnt));
    SR_TRY(c->alloc((size_t)n * 4, &off));
    SrGeoOut none = {};
    const uint32_t grid = ceil_div(n, 128);
    if (NV == 3) SR_LAUNCH(c, k_clip_tri_sh<0>, grid, 128, 0, in, cnt->as<uint32_t>(), nullptr, none);  // SR_GS_CLIP_SH
    else if (NV == 2) SR_LAUNCH(c, k_clip_line<0>, grid, 128, 0, in, cnt->as<uint32_t>(), nullptr, none);
    else SR_LAUNCH(c, k_clip_point<0>, grid, 128, 0, in, cnt->as<uint32_t>(), nullptr, none);
    uint32_t total = 0;
    SR_TRY(exclusive_scan(c, cnt->as<uint32_t>(), n, off->as<uint32_t>(), &total));
    SR_TRY(alloc_stream(c, (uint64_t)total * NV, nk, out));
    SrGeoOut o = {out->pos->as<float4>(), out->attr->as<float4>(), out->np};
    if (NV == 3) SR_LAUNCH(c, k_clip_tri_sh<1>, grid, 128, 0, in, nullptr, off->as<uint32_t>(), o);
    else if (NV == 2) SR_LAUNCH(c, k_clip_line<1>, grid, 128, 0, in, nullptr, off->as<uint32_t>(), o);
    else SR_LAUNCH(c, k_clip_point<1>, grid, 128, 0, in, nullptr, off->as<uint32_t>(), o);
    return SR_OK;
}
extern "C" {

int sr_geometry_run(sr_draw *d, uint32_t gs) {
    SrRange nvtx("softrender: geometry stage");
    if (!d) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (d->stage != STAGE_GEOMETRY) return sr_fail(SR_ERR_INVALID_STATE, "geometry stage needs clip-space vertices");
    if (gs > SR_GS_CLIP_SH) return sr_fail(SR_ERR_INVALID_ARGUMENT, "unknown geometry shader %u", gs);
    if (gs != SR_GS_CLIP && gs != SR_GS_CLIP_SH && d->nk < 8) return sr_fail(SR_ERR_INVALID_ARGUMENT, "normal visualisation needs K = {position4, normal4, ..}");
    sr_pipeline *p = d->pipeline;
    sr_context *c = p->ctx;
    SR_CUDA(cudaSetDevice(c->device));
    SR_TRY(resolve_tri_count(d));  // (a second geometry pass over the output of an unsynchronised clip)
    const uint32_t np = nplanes_of(d->nk);
    auto geo_in = [&](uint32_t kind, bool with_gen, bool with_idx) {
        SrGeoIn in;
        memset(&in, 0, sizeof(in));
        in.nplanes = np;
        if (with_gen) {
            in.gen = d->gen[kind - 1].set();
            in.ngen = (uint32_t)(d->gen[kind - 1].n / kind);
        }
        if (with_idx && d->have_indexed && d->primitive == kind) {
            in.idx = d->indexed.set();
            in.indices = d->indices->as<uint32_t>();
            in.nidx = (uint32_t)(d->nindices / kind);
        }
        return in;
    };
    VertexStream npoints, nlines, ntris;
    Buf nseq, count_dev;
    uint32_t literal_total = 0;
    if (gs == SR_GS_CLIP || gs == SR_GS_CLIP_SH) {
        SR_TRY(clip_small<1>(c, geo_in(1, true, true), d->nk, &npoints));
        SR_TRY(clip_small<2>(c, geo_in(2, true, true), d->nk, &nlines));
        const SrGeoIn tin = geo_in(3, true, true);
        const uint32_t n = tin.ngen + tin.nidx;
        if (gs == SR_GS_CLIP_SH) {
            // the opt-in correct clipper: every output triangle is kept and numbered in output order (no literal sequence)
            SR_TRY(clip_small<3>(c, tin, d->nk, &ntris));
        } else if (n == 0) {
            SR_TRY(alloc_stream(c, 0, d->nk, &ntris));
        } else {
            // zero-area outputs of the literal clipper are dropped unless a stencil op could observe them
            const bool stencil_active = p->fb->stencil_buf && p->stencil_op != SR_STENCIL_KEEP;
            const uint32_t drop = stencil_active ? 0u : 1u;
            // Small meshes (a model of a thousand triangles: configs 1 and 5) neither synchronise with the host nor take more than one
            // launch -- the round trip was 18 of a Suzanne frame's 110 us, the chain count -> scan -> emit three of its launches.  The
            // literal clipper emits at most 34 triangles per input triangle (a polygon of at most 36 entries, geometry.rs:265-298), so
            // the output stream is sized for that, its positions are pre-filled with NaN (every consumer skips a NaN primitive) and the
            // true counts stay on the device (sr_draw::tri_count_dev): the kernels that walk the stream read them there, the host only
            // ever uses the bound.  Draws that also carry points or lines keep the synchronisation (their canonical numbers follow the
            // literal triangle count).
            static const bool force_sync = getenv("SR_CLIP_SYNC") != nullptr;  // A/B switch
            if (!force_sync && n <= SR_CLIP_SMALL_MAX && n * 34u <= SR_BIN_SMALL_MAX_TRIS_DEV && npoints.n == 0 && nlines.n == 0 && c->shard_world == 1) {
                const uint32_t bound = n * 34u;
                SR_TRY(c->alloc(8, &count_dev));
                SR_TRY(alloc_stream(c, (uint64_t)bound * 3, d->nk, &ntris));
                SR_CUDA(cudaMemsetAsync(ntris.pos->ptr, 0xFF, (size_t)bound * 3 * sizeof(float4), c->stream));  // NaN positions
                SR_TRY(c->alloc((size_t)bound * 4, &nseq));
                SrGeoOut o = {ntris.pos->as<float4>(), ntris.attr->as<float4>(), ntris.np};
                SR_LAUNCH(c, k_clip_tri_small, 1, SR_CLIP_SMALL_MAX, 0, tin, drop, o, nseq->as<uint32_t>(), count_dev->as<uint32_t>());
                goto clipped;
            }
            Buf kept, lit, kept_off, lit_off;
            SR_TRY(c->alloc((size_t)n * 4, &kept));
            SR_TRY(c->alloc((size_t)n * 4, &lit));
            SR_TRY(c->alloc((size_t)n * 4, &kept_off));
            SR_TRY(c->alloc((size_t)n * 4, &lit_off));
            const uint32_t grid = ceil_div(n, 128);
            SR_LAUNCH(c, k_clip_tri_count, grid, 128, 0, tin, drop, kept->as<uint32_t>(), lit->as<uint32_t>());
            // both scans are enqueued before the one synchronisation that sizes the output stream
            if (n <= SR_SCAN_BLOCK) {
                Buf totals;
                SR_TRY(c->alloc(8, &totals));
                SR_LAUNCH(c, k_scan_pair_small, 1, SR_SCAN_THREADS, 0, kept->as<uint32_t>(), lit->as<uint32_t>(), n, kept_off->as<uint32_t>(),
                          lit_off->as<uint32_t>(), totals->as<uint32_t>());
                if (!c->pinned) SR_CUDA(cudaHostAlloc((void **)&c->pinned, 64, cudaHostAllocDefault));
                SR_CUDA(cudaMemcpyAsync(&c->pinned[3], totals->ptr, 8, cudaMemcpyDeviceToHost, c->stream));
            } else {
                SR_TRY(exclusive_scan_async(c, kept->as<uint32_t>(), n, kept_off->as<uint32_t>(), 3));
                SR_TRY(exclusive_scan_async(c, lit->as<uint32_t>(), n, lit_off->as<uint32_t>(), 4));
            }
            SR_CUDA(cudaStreamSynchronize(c->stream));
            const uint32_t kept_total = c->pinned[3];
            literal_total = c->pinned[4];
            SR_TRY(alloc_stream(c, (uint64_t)kept_total * 3, d->nk, &ntris));
            SR_TRY(c->alloc((size_t)std::max(kept_total, 1u) * 4, &nseq));
            SrGeoOut o = {ntris.pos->as<float4>(), ntris.attr->as<float4>(), ntris.np};
            SR_LAUNCH(c, k_clip_tri_emit, grid, 128, 0, tin, drop, kept_off->as<uint32_t>(), lit_off->as<uint32_t>(), o, nseq->as<uint32_t>());
        }
    } else {
        SrVsConst vc;
        fill_vs_const(p, nullptr, &vc);
        // points: re-emitted
        {
            const SrGeoIn in = geo_in(1, true, true);
            const uint32_t n = in.ngen + in.nidx;
            SR_TRY(alloc_stream(c, n, d->nk, &npoints));
            SrGeoOut o = {npoints.pos->as<float4>(), npoints.attr->as<float4>(), npoints.np};
            if (n) SR_LAUNCH(c, k_geo_reemit<1>, ceil_div(n, 128), 128, 0, in, o, (uint64_t)0);
        }
        // lines: [generated lines re-emitted | normals of generated triangles | indexed lines re-emitted or normals of indexed triangles]
        {
            const SrGeoIn gl = geo_in(2, true, false), gt = geo_in(3, true, false);
            const SrGeoIn il = geo_in(2, false, true), it = geo_in(3, false, true);
            const uint32_t per_tri = gs == SR_GS_FACE_NORMALS ? 1u : 3u;
            const uint64_t total_lines = (uint64_t)gl.ngen + (uint64_t)gt.ngen * per_tri + il.nidx + (uint64_t)it.nidx * per_tri;
            SR_TRY(alloc_stream(c, total_lines * 2, d->nk, &nlines));
            SrGeoOut o = {nlines.pos->as<float4>(), nlines.attr->as<float4>(), nlines.np};
            uint64_t base = 0;
            if (gl.ngen) SR_LAUNCH(c, k_geo_reemit<2>, ceil_div(gl.ngen, 128), 128, 0, gl, o, base);
            base += (uint64_t)gl.ngen * 2;
            auto normals = [&](const SrGeoIn &in, uint64_t at) -> int {
                const uint32_t n = in.ngen + in.nidx;
                if (!n) return SR_OK;
                if (gs == SR_GS_FACE_NORMALS) SR_LAUNCH(c, k_geo_normals<SR_GS_FACE_NORMALS>, ceil_div(n, 128), 128, 0, in, vc, o, at);
                else SR_LAUNCH(c, k_geo_normals<SR_GS_VERTEX_NORMALS>, ceil_div(n, 128), 128, 0, in, vc, o, at);
                return SR_OK;
            };
            SR_TRY(normals(gt, base));
            base += (uint64_t)gt.ngen * per_tri * 2;
            if (il.nidx) SR_LAUNCH(c, k_geo_reemit<2>, ceil_div(il.nidx, 128), 128, 0, il, o, base);
            base += (uint64_t)il.nidx * 2;
            SR_TRY(normals(it, base));
        }
        SR_TRY(alloc_stream(c, 0, d->nk, &ntris));
    }
clipped:
    d->gen[0] = npoints;
    d->gen[1] = nlines;
    d->gen[2] = ntris;
    d->tri_seq = nseq;
    d->tri_count_dev = count_dev;
    d->tri_literal_total = literal_total;
    d->have_indexed = false;  // indexed_vertices: None (geometry.rs:250-257)
    d->indexed = VertexStream();
    record(c, 2);
    return SR_OK;
}
int sr_geometry_clip_primitives(sr_draw *d) { return sr_geometry_run(d, SR_GS_CLIP); }

int sr_geometry_finish(sr_draw *d, const sr_viewport *vp) {
    SrRange nvtx("softrender: finish");
    if (!d || !vp) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (d->stage != STAGE_GEOMETRY) return sr_fail(SR_ERR_INVALID_STATE, "finish needs clip-space vertices");
    sr_pipeline *p = d->pipeline;
    sr_context *c = p->ctx;
    SR_CUDA(cudaSetDevice(c->device));
    SrVsConst vc;
    fill_vs_const(p, vp, &vc);
    auto norm = [&](VertexStream &s) -> int {
        if (!s.pos || s.n == 0) return SR_OK;
        if (s.pos.use_count() > 1) {  // shared with a duplicate(): copy on write (geometry.rs:43 deep-copies)
            Buf np_;
            SR_TRY(c->alloc(s.stride * sizeof(float4), &np_));
            SR_CUDA(cudaMemcpyAsync(np_->ptr, s.pos->ptr, s.stride * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
            s.pos = np_;
        }
        SR_LAUNCH(c, k_normalize, ceil_div(s.n, 256), 256, 0, s.pos->as<float4>(), s.n, vc);
        return SR_OK;
    };
    SR_TRY(norm(d->gen[0]));
    SR_TRY(norm(d->gen[1]));
    SR_TRY(norm(d->gen[2]));
    if (d->have_indexed) SR_TRY(norm(d->indexed));
    record(c, 2);
    d->stage = STAGE_FRAGMENT;
    return SR_OK;
}

int sr_draw_duplicate(sr_draw *d, sr_draw **out) {
    if (!d || !out) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    *out = new sr_draw(*d);  // device buffers are shared and immutable (finish copies on write)
    return SR_OK;
}
int sr_fragment_set_cull_faces(sr_draw *d, uint32_t winding) {
    if (!d || winding > SR_COUNTER_CLOCKWISE) return sr_fail(SR_ERR_INVALID_ARGUMENT, "bad winding");
    d->cull = winding;
    return SR_OK;
}
int sr_fragment_set_antialiased_lines(sr_draw *d, int enable) {
    if (!d) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    d->aa_lines = enable ? 1 : 0;
    return SR_OK;
}
int sr_fragment_set_tile_size(sr_draw *d, uint32_t w, uint32_t h) {
    if (!d || !w || !h) return sr_fail(SR_ERR_INVALID_ARGUMENT, "bad tile size");
    d->tile_w = w;
    d->tile_h = h;
    return SR_OK;
}
int sr_fragment_set_blend(sr_draw *d, uint32_t blend) {
    if (!d || blend > SR_BLEND_ADDITIVE) return sr_fail(SR_ERR_INVALID_ARGUMENT, "bad blend");
    d->blend = blend;
    return SR_OK;
}

int sr_fragment_run(sr_draw *d, uint32_t fs) {
    SrRange nvtx("softrender: fragment stage");
    if (!d) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (d->stage != STAGE_FRAGMENT) return sr_fail(SR_ERR_INVALID_STATE, "fragment stage needs screen-space vertices (finish / run_to_fragment)");
    const int need = fs_nk(fs);
    if (need < 0) return sr_fail(SR_ERR_INVALID_ARGUMENT, "unknown fragment shader %u", fs);
    if ((int)d->nk < need) return sr_fail(SR_ERR_INVALID_ARGUMENT, "fragment shader %u reads %d interpolated floats, draw carries %u", fs, need, d->nk);
    sr_pipeline *p = d->pipeline;
    sr_context *c = p->ctx;
    sr_framebuffer *fb = p->fb;
    SR_CUDA(cudaSetDevice(c->device));
    SR_TRY(settle(c));
    if (fb->width < 2 || fb->height < 2) return SR_OK;  // fragment.rs:188-216: a 1-pixel-wide frame has no tiles, nothing is drawn
    {   // a recorded vertex stage stays recorded only for the range-sharded opaque triangle pass
        const bool st_active = fb->stencil_buf && !(p->stencil_test == SR_STENCIL_ALWAYS && p->stencil_op == SR_STENCIL_KEEP);
        if (d->vertex_lazy && !(d->blend == SR_BLEND_REPLACE && !st_active && fs != SR_FS_DISCARD_CHECKER)) SR_TRY(materialize_vertices(d));
    }
    {   // the colour type of the target and the shader's return type must agree (a type error in the reference)
        const bool two_out = fs == SR_FS_SUZANNE_GBUFFER;
        if (two_out != (fb->soa == 2))
            return sr_fail(SR_ERR_INVALID_STATE, two_out ? "a fragment shader with two colour outputs needs a texture buffer with two colour planes"
                                                         : "a texture buffer with two colour planes needs a fragment shader with two colour outputs");
        if (two_out) {
            const bool st_active = fb->stencil_buf && !(p->stencil_test == SR_STENCIL_ALWAYS && p->stencil_op == SR_STENCIL_KEEP);
            if (d->blend != SR_BLEND_REPLACE || st_active || d->primitive != SR_TRIANGLE || d->gen[0].n || d->gen[1].n)
                return sr_fail(SR_ERR_UNSUPPORTED, "two-plane texture buffers: triangles with Blend = () and stencil () only");
        }
    }
    if (fb->u8color && d->blend != SR_BLEND_REPLACE)
        return sr_fail(SR_ERR_UNSUPPORTED, "RGBAu8Color targets take Blend = () only: the registered blend functions are defined on f32 colours "
                       "(full_example/src/color.rs:5-17); see INTEGRATION.md for registering a u8 blend");
    const bool samples = fs == SR_FS_FULL_EXAMPLE_TEXTURED || fs == SR_FS_TEXTURE_UNLIT;
    if (samples && p->fb_texture && p->fb_texture->u8color)
        return sr_fail(SR_ERR_UNSUPPORTED, "an RGBAu8Color target cannot be bound as a texture source (sampler semantics are defined for f32 targets and RGBA8 images)");
    if (samples && !p->texture && !p->fb_texture) return sr_fail(SR_ERR_INVALID_STATE, "textured shader without a bound texture");
    if (samples && p->fb_texture) {
        sr_framebuffer *src = p->fb_texture;
        if (src == fb || src->aos == fb->aos) return sr_fail(SR_ERR_INVALID_STATE, "a framebuffer cannot be sampled by a draw that renders into it");
        if (src->ctx != c) return sr_fail(SR_ERR_INVALID_STATE, "texture source framebuffer belongs to another context");
        SR_TRY(materialize_clear(src));  // a recorded clear becomes pixels before they are sampled
    }

    // an unsynchronised clip left the triangle counts on the device: fine for a draw of triangles on one GPU, everything else
    // gets them to the host first (canonical numbers of lines / points follow the literal triangle count; tile-sharded contexts)
    if (d->tri_count_dev && (d->gen[0].n || d->gen[1].n || (d->have_indexed && d->primitive != SR_TRIANGLE) || c->shard_world > 1 || c->shard))
        SR_TRY(resolve_tri_count(d));
    SrTileParams tp;
    memset(&tp, 0, sizeof(tp));
    tp.tris = prim_source(d, 3);
    tp.lines = prim_source(d, 2);
    tp.points = prim_source(d, 1);
    tp.ntris = tp.tris.n0 + tp.tris.n1;
    tp.nlines = tp.lines.n0 + tp.lines.n1;
    tp.npoints = tp.points.n0 + tp.points.n1;
    const uint32_t tri_canonical = tp.tris.n0 + (tp.tris.seq1 ? d->tri_literal_total : tp.tris.n1);
    tp.line_base = tri_canonical;
    tp.point_base = tri_canonical + tp.nlines;
    tp.shard_rank = c->shard_rank; tp.shard_world = c->shard_world;
    tp.blend = d->blend;
    tp.stencil_test = p->stencil_test; tp.stencil_op = p->stencil_op; tp.stencil_value = d->stencil_value;
    tp.aa_lines = d->aa_lines;
    tp.fs.u = p->uniforms;
    tp.fs.tex_filter = p->tex_filter; tp.fs.tex_edge = p->tex_edge;
    for (int i = 0; i < 4; ++i) tp.fs.tex_border[i] = p->tex_border[i];
    if (p->fb_texture && samples) {
        tp.fs.tex = reinterpret_cast<const uint8_t *>(p->fb_texture->aos + 4ull * p->fb_texture->width * p->fb_texture->height * p->fb_texture_plane);
        tp.fs.tex_w = p->fb_texture->width;
        tp.fs.tex_h = p->fb_texture->height;
        tp.fs.tex_kind = SR_TEX_F32;
        tp.fs.tex_stride = p->fb_texture->soa ? 4 : 5;  // texture buffer: the colour plane itself (TextureBufferRef, texturebuffer.rs:12-47); RenderBuffer: the 20-byte AoS pixel
    } else if (p->texture) {
        tp.fs.tex = p->texture->rgba->as<uint8_t>();
        tp.fs.tex_w = p->texture->width;
        tp.fs.tex_h = p->texture->height;
        tp.fs.tex_kind = SR_TEX_RGBA8;
    }

    if (fb->winner_enabled && fb->winner_buf)  // winner plane reports the primitives of THIS draw
        SR_CUDA(cudaMemsetAsync(fb->winner_buf->ptr, 0, (size_t)fb->width * fb->height * 4, c->stream));
    record(c, 3);
    const uint32_t ntiles = fb->ntx * fb->nty;
    const uint32_t owned = ntiles > c->shard_rank ? (ntiles - c->shard_rank + c->shard_world - 1) / c->shard_world : 0;
    const bool stencil_active = fb->stencil_buf && !(p->stencil_test == SR_STENCIL_ALWAYS && p->stencil_op == SR_STENCIL_KEEP);
    const bool opaque_ok = d->blend == SR_BLEND_REPLACE && !stencil_active && fs != SR_FS_DISCARD_CHECKER;
    // non-antialiased lines and points of an opaque draw are order independent too: they join the triangles in the
    // visibility buffer (k_lines_vis / k_points_vis) instead of the ordered tile pass
    const bool extra_vis = opaque_ok && !d->aa_lines && tp.nlines + tp.npoints > 0 && (uint64_t)tp.ntris + tp.nlines + tp.npoints < 0xFFFFFFFFull;
    const bool ordered_pass = !opaque_ok || (!extra_vis && tp.nlines + tp.npoints > 0);
    auto q = std::make_unique<PendingOrdered>();
    Bins &bp = q->bins[0], &bl = q->bins[1], &bt = q->bins[2];
    if (!c->pinned) SR_CUDA(cudaHostAlloc((void **)&c->pinned, 64, cudaHostAllocDefault));
    if (opaque_ok) SR_TRY(zero_offsets(c, ntiles, &bt));  // triangles go through opaque_triangles below
    else SR_TRY(build_bins<3>(c, fb, tp.tris, tp.ntris, d->cull, &bt, false, &c->pinned[10]));
    SR_TRY(build_bins<2>(c, fb, tp.lines, extra_vis ? 0u : tp.nlines, SR_CULL_NONE, &bl, false, &c->pinned[9]));
    SR_TRY(build_bins<1>(c, fb, tp.points, extra_vis ? 0u : tp.npoints, SR_CULL_NONE, &bp, false, &c->pinned[8]));
    record(c, 4);
    record(c, 7);
    record(c, 5);
    tp.tri_rects = bt.rects->as<uint32_t>(); tp.tri_off = bt.off->as<uint32_t>(); tp.tri_list = bt.list->as<uint32_t>(); tp.tri_cap = bt.capacity;
    tp.line_rects = bl.rects->as<uint32_t>(); tp.line_off = bl.off->as<uint32_t>(); tp.line_list = bl.list->as<uint32_t>(); tp.line_cap = bl.capacity;
    tp.point_rects = bp.rects->as<uint32_t>(); tp.point_off = bp.off->as<uint32_t>(); tp.point_list = bp.list->as<uint32_t>(); tp.point_cap = bp.capacity;

    if (owned) {
        std::vector<Buf> keep = {d->indices, d->indexed.pos, d->indexed.attr, d->gen[0].pos, d->gen[0].attr, d->gen[1].pos, d->gen[1].attr,
                                 d->gen[2].pos, d->gen[2].attr, d->tri_seq, d->tri_count_dev};
        if (p->texture) keep.push_back(p->texture->rgba);
        if (p->fb_texture && samples && p->fb_texture->aos_buf) keep.push_back(p->fb_texture->aos_buf);
        SrTileParams ordered = tp;
        if (opaque_ok) {
            // triangles through the order-independent resolve; lines/points (always after all triangles,
            // fragment.rs:268-311) through the ordered kernel
            if (tp.ntris || fb->pending_clear || extra_vis) {
                SR_TRY(opaque_triangles(c, fb, tp, d->cull, fs, owned, keep, extra_vis, d));
                if (ordered_pass) SR_TRY(settle(c));  // a replayed triangle pass must not land after the lines
            }
            ordered.ntris = 0;
        }
        if (ordered_pass) {
            ordered.fb = fb->view();
            // the scanned totals are on their way to pinned memory (build_bins): validated at the next call on this context
            SR_CUDA(cudaEventCreateWithFlags(&q->counted, cudaEventDisableTiming));
            SR_CUDA(cudaEventRecord(q->counted, c->stream));
            SR_TRY(launch_tiles_fs(c, fs, owned, ordered));
            fb->pending_clear = false;
            q->tp = ordered;
            q->fs = fs; q->owned = owned;
            q->keep = keep;
            c->pending_ord = q.release();
        }
    }
    record(c, 6);
    return SR_OK;
}

int sr_draw_destroy(sr_draw *d) {
    delete d;
    return SR_OK;
}

// ---- injection / introspection ---------------------------------------------------------------------------
static int upload_records(sr_context *c, const float *verts, uint64_t n, uint32_t nk, VertexStream *out) {
    SR_TRY(alloc_stream(c, n, nk, out));
    if (n == 0) return SR_OK;
    Buf tmp;
    SR_TRY(c->alloc(n * (4 + nk) * 4, &tmp));
    SR_CUDA(cudaMemcpyAsync(tmp->ptr, verts, n * (4 + nk) * 4, cudaMemcpyHostToDevice, c->stream));
    SR_LAUNCH(c, k_records_to_planes, ceil_div(n, 256), 256, 0, tmp->as<float>(), n, nk, out->pos->as<float4>(), out->attr->as<float4>(), out->np);
    SR_CUDA(cudaStreamSynchronize(c->stream));
    return SR_OK;
}

int sr_draw_from_vertices(sr_pipeline *p, uint32_t primitive, const float *verts, uint64_t nverts, uint32_t nk, int space,
                          const uint32_t *indices, uint64_t nindices, int has_sv, uint32_t sv, sr_draw **out) {
    if (!p || !out || (!verts && nverts) || (!indices && nindices)) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (primitive < SR_POINT || primitive > SR_TRIANGLE) return sr_fail(SR_ERR_INVALID_ARGUMENT, "primitive %u", primitive);
    if (nindices % primitive != 0) return sr_fail(SR_ERR_INVALID_ARGUMENT, "indices.len() %% num_vertices != 0");
    if (nk > SR_MAX_NK) return sr_fail(SR_ERR_UNSUPPORTED, "nk %u > %u", nk, SR_MAX_NK);
    for (uint64_t i = 0; i < nindices; ++i)
        if (indices[i] >= nverts) return sr_fail(SR_ERR_INVALID_ARGUMENT, "index out of range");
    sr_context *c = p->ctx;
    SR_CUDA(cudaSetDevice(c->device));
    auto d = std::make_unique<sr_draw>();
    d->pipeline = p;
    d->primitive = primitive;
    d->has_stencil_value = has_sv != 0;
    d->stencil_value = has_sv ? sv : 0;
    d->nk = nk;
    d->nindices = nindices;
    SR_TRY(c->alloc((nindices + 128) * 4, &d->indices));
    if (nindices) SR_CUDA(cudaMemcpyAsync(d->indices->ptr, indices, nindices * 4, cudaMemcpyHostToDevice, c->stream));
    SR_TRY(upload_records(c, verts, nverts, nk, &d->indexed));
    d->have_indexed = true;
    d->stage = space ? STAGE_FRAGMENT : STAGE_GEOMETRY;
    *out = d.release();
    return SR_OK;
}
int sr_draw_set_generated(sr_draw *d, int which, const float *verts, uint64_t nverts, uint32_t nk) {
    if (!d || which < 1 || which > 3 || (!verts && nverts)) return sr_fail(SR_ERR_INVALID_ARGUMENT, "bad arguments");
    if (nverts % (uint64_t)which != 0) return sr_fail(SR_ERR_INVALID_ARGUMENT, "vertex count not a multiple of the primitive size");
    if (nk != d->nk) return sr_fail(SR_ERR_INVALID_ARGUMENT, "nk mismatch");
    if (d->stage == STAGE_VERTEX) return sr_fail(SR_ERR_INVALID_STATE, "run the vertex stage first");
    SR_TRY(upload_records(d->pipeline->ctx, verts, nverts, nk, &d->gen[which - 1]));
    if (which == 3) { d->tri_seq.reset(); d->tri_literal_total = 0; d->tri_count_dev.reset(); }
    return SR_OK;
}
int sr_draw_count(sr_draw *d, int which, uint64_t *nverts, uint32_t *nk) {
    if (!d || which < 0 || which > 3) return sr_fail(SR_ERR_INVALID_ARGUMENT, "bad arguments");
    if (which == 3) SR_TRY(resolve_tri_count(d));
    if (nverts) *nverts = which == 0 ? (d->have_indexed ? d->indexed.n : 0) : d->gen[which - 1].n;
    if (nk) *nk = d->nk;
    return SR_OK;
}
int sr_draw_download(sr_draw *d, int which, float *dst, uint64_t capacity_floats) {
    if (!d || which < 0 || which > 3 || !dst) return sr_fail(SR_ERR_INVALID_ARGUMENT, "bad arguments");
    if (which == 3) SR_TRY(resolve_tri_count(d));
    const VertexStream &s = which == 0 ? d->indexed : d->gen[which - 1];
    const uint64_t n = which == 0 ? (d->have_indexed ? s.n : 0) : s.n;
    if (n == 0) return SR_OK;
    if (capacity_floats < n * (4 + d->nk)) return sr_fail(SR_ERR_INVALID_ARGUMENT, "capacity too small");
    sr_context *c = d->pipeline->ctx;
    SR_CUDA(cudaSetDevice(c->device));
    if (which == 0) SR_TRY(materialize_vertices(d));
    Buf tmp;
    SR_TRY(c->alloc(n * (4 + d->nk) * 4, &tmp));
    SR_LAUNCH(c, k_planes_to_records, ceil_div(n, 256), 256, 0, s.pos->as<float4>(), s.attr->as<float4>(), s.np, n, d->nk, tmp->as<float>());
    SR_CUDA(cudaMemcpyAsync(dst, tmp->ptr, n * (4 + d->nk) * 4, cudaMemcpyDeviceToHost, c->stream));
    SR_CUDA(cudaStreamSynchronize(c->stream));
    return SR_OK;
}
int sr_draw_download_sequence(sr_draw *d, uint32_t *dst, uint64_t capacity) {
    if (!d || !dst) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    SR_TRY(resolve_tri_count(d));
    const uint64_t n = d->gen[2].n / 3;
    if (capacity < n) return sr_fail(SR_ERR_INVALID_ARGUMENT, "capacity too small");
    if (n == 0) return SR_OK;
    sr_context *c = d->pipeline->ctx;
    if (d->tri_seq) {
        SR_CUDA(cudaMemcpyAsync(dst, d->tri_seq->ptr, n * 4, cudaMemcpyDeviceToHost, c->stream));
        SR_CUDA(cudaStreamSynchronize(c->stream));
    } else {
        for (uint64_t i = 0; i < n; ++i) dst[i] = (uint32_t)i;
    }
    return SR_OK;
}

// parity introspection: the per-tile triangle lists the HEADLINE path built for the latest opaque draw of this
// context -- k_bin_small's lists (every triangle of a small draw) or k_micro + k_large_fill's (the triangles whose clamped
// bounding box exceeds *micro_area pixels and that the tightened small-triangle path did not take; *micro_area = 0: all).
// CSR offsets (ntiles + 1), ids ascending per tile.  ids may be NULL to query *total.
int sr_context_last_opaque_lists(sr_context *c, uint64_t *offsets, uint32_t *ids, uint64_t ids_capacity, uint64_t *total, uint32_t *micro_area) {
    if (!c || !total) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    SR_TRY(settle(c));  // (an overflowed pass is replayed with a larger arena first)
    if (!c->last_off || !c->last_list) return sr_fail(SR_ERR_INVALID_STATE, "no opaque draw on this context yet");
    SR_CUDA(cudaSetDevice(c->device));
    const uint32_t ntiles = c->last_ntiles;
    std::vector<uint32_t> off(ntiles + 1);
    SR_CUDA(cudaMemcpyAsync(off.data(), c->last_off->ptr, (size_t)(ntiles + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
    SR_CUDA(cudaStreamSynchronize(c->stream));
    const uint32_t n = off[ntiles];
    if (n > c->list_cap) return sr_fail(SR_ERR_INVALID_STATE, "lists (%u entries) exceed the arena (%u)", n, c->list_cap);
    std::vector<uint32_t> list(std::max(n, 1u));
    if (n) SR_CUDA(cudaMemcpyAsync(list.data(), c->list_arena->ptr, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    SR_CUDA(cudaStreamSynchronize(c->stream));
    for (uint32_t t = 0; t < ntiles; ++t) {
        std::sort(list.begin() + off[t], list.begin() + off[t + 1]);  // fill order within a tile is not deterministic (atomic cursors)
        if (offsets) offsets[t] = off[t];
    }
    if (offsets) offsets[ntiles] = n;
    if (ids)
        for (uint64_t i = 0; i < n && i < ids_capacity; ++i) ids[i] = list[i];
    *total = n;
    if (micro_area) *micro_area = c->last_micro_area;
    return SR_OK;
}
int sr_selftest_division(sr_context *c, uint64_t seed, uint64_t count, uint64_t *mismatches) {
    if (!c || !mismatches) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    SR_CUDA(cudaSetDevice(c->device));
    Buf bad;
    SR_TRY(c->alloc(8, &bad));
    SR_CUDA(cudaMemsetAsync(bad->ptr, 0, 8, c->stream));
    const uint32_t threads = 148 * 8 * 256;
    SR_LAUNCH(c, k_selftest_division, 148 * 8, 256, 0, seed, (count + threads - 1) / threads, bad->as<unsigned long long>());
    SR_CUDA(cudaMemcpyAsync(mismatches, bad->ptr, 8, cudaMemcpyDeviceToHost, c->stream));
    SR_CUDA(cudaStreamSynchronize(c->stream));
    return SR_OK;
}

int sr_draw_bins(sr_draw *d, uint64_t *offsets, uint32_t *ids, uint64_t ids_capacity, uint64_t *total) {
    if (!d || !total) return sr_fail(SR_ERR_INVALID_ARGUMENT, "null");
    if (d->stage != STAGE_FRAGMENT) return sr_fail(SR_ERR_INVALID_STATE, "bins need screen-space vertices");
    sr_pipeline *p = d->pipeline;
    sr_context *c = p->ctx;
    sr_framebuffer *fb = p->fb;
    SR_CUDA(cudaSetDevice(c->device));
    SR_TRY(materialize_vertices(d));
    SR_TRY(resolve_tri_count(d));
    const SrPrimSource src = prim_source(d, 3);
    const uint32_t ntris = src.n0 + src.n1, ntiles = fb->ntx * fb->nty;
    Bins b;
    SR_TRY(build_bins<3>(c, fb, src, ntris, d->cull, &b));
    std::vector<uint32_t> rects(std::max(ntris, 1u)), off(ntiles + 1), list(std::max(b.total, 1u)), seq;
    if (ntris) SR_CUDA(cudaMemcpyAsync(rects.data(), b.rects->ptr, (size_t)ntris * 4, cudaMemcpyDeviceToHost, c->stream));
    SR_CUDA(cudaMemcpyAsync(off.data(), b.off->ptr, (size_t)(ntiles + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
    if (b.total) SR_CUDA(cudaMemcpyAsync(list.data(), b.list->ptr, (size_t)b.total * 4, cudaMemcpyDeviceToHost, c->stream));
    if (src.seq1 && src.n1) {
        seq.resize(src.n1);
        SR_CUDA(cudaMemcpyAsync(seq.data(), src.seq1, (size_t)src.n1 * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    SR_CUDA(cudaStreamSynchronize(c->stream));
    // expand the per-tile group lists into exact per-tile triangle lists, exactly as the tile kernels do
    uint64_t n = 0;
    for (uint32_t tile = 0; tile < ntiles; ++tile) {
        if (offsets) offsets[tile] = n;
        std::vector<uint32_t> groups(list.begin() + off[tile], list.begin() + off[tile + 1]);
        std::sort(groups.begin(), groups.end());
        const uint32_t tx = tile % fb->ntx, ty = tile / fb->ntx;
        for (uint32_t g : groups)
            for (uint32_t j = 0; j < SR_GROUP; ++j) {
                const uint32_t t = g * SR_GROUP + j;
                if (t >= ntris) break;
                const uint32_t r = rects[t];
                if (r == SR_RECT_INVALID) continue;
                const uint32_t tx0 = r & 255u, ty0 = (r >> 8) & 255u, tx1 = (r >> 16) & 255u, ty1 = r >> 24;
                if (!(tx0 <= tx && tx <= tx1 && ty0 <= ty && ty <= ty1)) continue;
                if (ids && n < ids_capacity) ids[n] = (t < src.n0 || seq.empty()) ? t : src.n0 + seq[t - src.n0];
                ++n;
            }
    }
    if (offsets) offsets[ntiles] = n;
    *total = n;
    return SR_OK;
}

}  // extern "C"
