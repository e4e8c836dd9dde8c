// sr_raster.cuh -- primitive setup + screen-tile binning, and the tile rasterisers
// (FragmentShader::run, src/pipeline/stages/fragment.rs:168-319; rasterize_triangle / _line / _point,
// src/pipeline/stages/rasterization/{triangle,line,point}.rs).
//
// Canonical semantics: the reference with ONE frame-sized tile (tile_size >= dimensions), i.e. every
// primitive visits the pixels of its integer bounding box clamped to the frame exactly once, in
// submission order.  The GPU partitions the frame into disjoint SR_TILE_W x SR_TILE_H tiles; the union
// over tiles of (clamped bbox intersected with the tile) is that same pixel set, so results do not depend
// on the GPU tile size.
#pragma once

#include <type_traits>

#include "sr_shaders.cuh"

#ifndef SR_RASTER_THREADS
#define SR_RASTER_THREADS 256
#endif
#define SR_RASTER_WARPS (SR_RASTER_THREADS / 32)
#define SR_SMALL_AREA 16          // bbox pixels a single lane rasterises itself; larger boxes go warp-wide
#define SR_MICRO_AREA_DEFAULT 16   // bbox pixels up to which k_micro rasterises a triangle itself
#define SR_MICRO_AREA_MAX 4096
#define SR_DEPTH_FAR_BITS 0xFF7FFFFFu  // f32::MIN, Depth::far() (src/framebuffer/attachments/depth.rs:31)

struct SrBinParams {
    SrPrimSource src;
    uint32_t nprims;
    uint32_t cull;
    uint32_t width, height, ntx, nty;
    uint32_t shard_rank, shard_world;
    uint32_t expand;       // lines: widen the box by one pixel (rounding of the clipped end-points, Wu neighbours)
    uint32_t *rects;       // per primitive: packed tile rectangle or SR_RECT_INVALID
    uint32_t *tile_count;  // count pass: entries per tile; fill pass: running cursor per tile
    const uint32_t *tile_off;
    uint32_t *list;        // fill pass: group ids per tile
    uint32_t capacity;     // fill pass: entries `list` can hold (the pass skips itself if the scanned total exceeds it)
};

// ---- per-primitive tile rectangle ------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ uint32_t sr_prim_rect(const SrBinParams &p, uint32_t t) {
    const SrVertexSet *vs;
    uint32_t vi[NV];
    sr_prim_vertices<NV>(p.src, t, vs, vi);
    float x[NV], y[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const float4 v = __ldg(vs->pos + vi[k]);
        x[k] = v.x;
        y[k] = v.y;
    }
    bool nan = false;
#pragma unroll
    for (int k = 0; k < NV; ++k) nan = nan || isnan(x[k]) || isnan(y[k]);
    if (nan) return SR_RECT_INVALID;  // the reference panics on NaN (cast(..).unwrap()); defined here as "skipped"
    uint32_t minx, miny, maxx, maxy;
    if (NV == 3) {
        if (p.cull != SR_CULL_NONE) {  // triangle.rs:54-61
            const float area = x[0] * y[1] + x[1] * y[2] + x[2] * y[0] - x[1] * y[0] - x[2] * y[1] - x[0] * y[2];
            const uint32_t winding = signbit(area) ? SR_CLOCKWISE : SR_COUNTER_CLOCKWISE;
            if (winding == p.cull) return SR_RECT_INVALID;
        }
        // triangle.rs:74-78 with tile = the whole frame
        minx = sr_clamp_as_int(fminf(fminf(x[0], x[1]), x[2]), 0, p.width - 1);
        miny = sr_clamp_as_int(fminf(fminf(y[0], y[1]), y[2]), 0, p.height - 1);
        maxx = sr_clamp_as_int(fmaxf(fmaxf(x[0], x[1]), x[2]), 0, p.width - 1);
        maxy = sr_clamp_as_int(fmaxf(fmaxf(y[0], y[1]), y[2]), 0, p.height - 1);
    } else if (NV == 2) {
        // conservative: every pixel rasterize_line can plot lies in the end-point box widened by two pixels (Wu plots the
        // minor coordinates trunc(yend) and trunc(yend) + 1 with yend up to half a pixel past the clipped end point, and
        // rounds the major coordinate of the end points, line.rs:181-199)
        minx = sr_clamp_as_int(fminf(x[0], x[NV - 1]), 0, p.width - 1);
        miny = sr_clamp_as_int(fminf(y[0], y[NV - 1]), 0, p.height - 1);
        maxx = sr_clamp_as_int(fmaxf(x[0], x[NV - 1]), 0, p.width - 1);
        maxy = sr_clamp_as_int(fmaxf(y[0], y[NV - 1]), 0, p.height - 1);
        minx = minx > 2 ? minx - 2 : 0;
        miny = miny > 2 ? miny - 2 : 0;
        maxx = maxx + 2 < p.width ? maxx + 2 : p.width - 1;
        maxy = maxy + 2 < p.height ? maxy + 2 : p.height - 1;
    } else {
        // point.rs:46: bounds.0 <= x < bounds.1 with bounds = (0,0)..(w-1,h-1)
        if (!(0.0f <= x[0] && x[0] < (float)(p.width - 1) && 0.0f <= y[0] && y[0] < (float)(p.height - 1))) return SR_RECT_INVALID;
        minx = maxx = __float2uint_rz(x[0]);
        miny = maxy = __float2uint_rz(y[0]);
    }
    return sr_pack_rect(minx / SR_TILE_W, miny / SR_TILE_H, maxx / SR_TILE_W, maxy / SR_TILE_H);
}

// Warp = one group of 32 consecutive primitives.  For every tile touched by at least one primitive of
// the group: FILL ? append the group id to the tile's list : count it.
// Consecutive groups of a coherent mesh land in the same one or two tiles, so per-group global atomics
// would all hit the same few addresses.  Each warp therefore posts up to SR_BIN_SLOTS (tile) entries to
// shared memory and warp 0 merges equal tiles of the whole CTA with __match_any_sync: one atomic per
// distinct tile per CTA.  Groups spanning more tiles (large primitives) use direct atomics.
#define SR_BIN_THREADS 256
#define SR_BIN_WARPS (SR_BIN_THREADS / 32)
#define SR_BIN_SLOTS 4
template <bool FILL>
__device__ __forceinline__ void sr_bin_group(const SrBinParams &p, uint32_t rect, uint32_t group) {
    __shared__ uint32_t s_tile[SR_BIN_WARPS * SR_BIN_SLOTS];
    __shared__ uint32_t s_bits[SR_BIN_WARPS][32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool valid = rect != SR_RECT_INVALID;
    uint32_t gx0 = valid ? (rect & 255u) : 255u, gy0 = valid ? ((rect >> 8) & 255u) : 255u;
    uint32_t gx1 = valid ? ((rect >> 16) & 255u) : 0u, gy1 = valid ? (rect >> 24) : 0u;
    gx0 = __reduce_min_sync(0xffffffffu, gx0);
    gy0 = __reduce_min_sync(0xffffffffu, gy0);
    gx1 = __reduce_max_sync(0xffffffffu, gx1);
    gy1 = __reduce_max_sync(0xffffffffu, gy1);
    const bool any_valid = __any_sync(0xffffffffu, valid);
    const bool compact = any_valid && (gx1 - gx0 + 1) * (gy1 - gy0 + 1) <= SR_BIN_SLOTS;
    uint32_t nslots = 0;
    if (compact) {
        for (uint32_t ty = gy0; ty <= gy1; ++ty)
            for (uint32_t tx = gx0; tx <= gx1; ++tx) {
                const bool hit = valid && sr_rect_hits(rect, tx, ty);
                if (!__any_sync(0xffffffffu, hit)) continue;
                const uint32_t tile = ty * p.ntx + tx;
                if (tile % p.shard_world != p.shard_rank) continue;
                if (lane == 0) s_tile[warp * SR_BIN_SLOTS + nslots] = tile;
                ++nslots;
            }
    } else if (any_valid) {
        // a group that spans more tiles: every primitive marks the tiles of its own rectangle in a bitmap of the group's
        // union rectangle (rectangles of more than 32 tiles are spread over the lanes), then each lane posts the set bits of
        // one bitmap word with its own atomic.  Union rectangles above 1024 tiles: tiles dealt to the lanes and tested
        // against the 32 rectangles handed round with shuffles.
        const uint32_t gw = gx1 - gx0 + 1, n = gw * (gy1 - gy0 + 1);
        if (n <= 1024u) {
            uint32_t *bits = s_bits[warp];
            bits[lane] = 0;
            __syncwarp();
            const uint32_t rx0 = rect & 255u, ry0 = (rect >> 8) & 255u, rx1 = (rect >> 16) & 255u, ry1 = rect >> 24;
            const uint32_t rw = rx1 - rx0 + 1, rn = valid ? rw * (ry1 - ry0 + 1) : 0u;
            const bool big = rn > 32u;
            if (!big)
                for (uint32_t i = 0; i < rn; ++i) {
                    const uint32_t at = (ry0 + i / rw - gy0) * gw + (rx0 + i % rw - gx0);
                    atomicOr(&bits[at >> 5], 1u << (at & 31u));
                }
            for (uint32_t rest = __ballot_sync(0xffffffffu, big); rest; rest &= rest - 1) {
                const uint32_t r = __shfl_sync(0xffffffffu, rect, __ffs(rest) - 1);
                const uint32_t bx0 = r & 255u, by0 = (r >> 8) & 255u, bx1 = (r >> 16) & 255u, by1 = r >> 24;
                const uint32_t bw = bx1 - bx0 + 1, bn = bw * (by1 - by0 + 1);
                for (uint32_t i = lane; i < bn; i += 32) {
                    const uint32_t at = (by0 + i / bw - gy0) * gw + (bx0 + i % bw - gx0);
                    atomicOr(&bits[at >> 5], 1u << (at & 31u));
                }
            }
            __syncwarp();
            for (uint32_t word = bits[lane]; word; word &= word - 1) {
                const uint32_t i = lane * 32u + (uint32_t)(__ffs(word) - 1);
                const uint32_t tile = (gy0 + i / gw) * p.ntx + gx0 + i % gw;
                if (tile % p.shard_world != p.shard_rank) continue;
                const uint32_t at = atomicAdd(p.tile_count + tile, 1u);
                if (FILL) p.list[p.tile_off[tile] + at] = group;
            }
        } else {
            for (uint32_t base = 0; base < n; base += 32) {
                const uint32_t i = base + lane;
                const uint32_t tx = gx0 + i % gw, ty = gy0 + i / gw;
                bool hit = false;
#pragma unroll 4
                for (int l = 0; l < 32; ++l) {
                    const uint32_t r = __shfl_sync(0xffffffffu, rect, l);
                    hit = hit || (r != SR_RECT_INVALID && sr_rect_hits(r, tx, ty));
                }
                const uint32_t tile = ty * p.ntx + tx;
                if (i < n && hit && tile % p.shard_world == p.shard_rank) {
                    const uint32_t at = atomicAdd(p.tile_count + tile, 1u);
                    if (FILL) p.list[p.tile_off[tile] + at] = group;
                }
            }
        }
    }
    if (lane == 0)
        for (uint32_t k = nslots; k < SR_BIN_SLOTS; ++k) s_tile[warp * SR_BIN_SLOTS + k] = 0xFFFFFFFFu;
    __syncthreads();
    if (warp == 0) {
        static_assert(SR_BIN_WARPS * SR_BIN_SLOTS == 32, "one slot per lane of warp 0");
        const uint32_t tile = s_tile[lane];
        const uint32_t peers = __match_any_sync(0xffffffffu, tile);
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (tile != 0xFFFFFFFFu && (int)lane == leader) base = atomicAdd(p.tile_count + tile, (uint32_t)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (FILL && tile != 0xFFFFFFFFu) {
            const uint32_t first_group = group - warp;  // this lane is in warp 0: group of warp w is first_group + w
            p.list[p.tile_off[tile] + base + __popc(peers & ((1u << lane) - 1u))] = first_group + lane / SR_BIN_SLOTS;
        }
    }
}

// pass 1: primitive setup (rect per primitive) + per-tile entry counts
template <int NV>
__global__ void __launch_bounds__(SR_BIN_THREADS) k_bin_setup(const SrBinParams p) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;  // grid covers whole warps
    uint32_t rect = SR_RECT_INVALID;
    if (t < p.nprims) rect = sr_prim_rect<NV>(p, t);
    if (t < ((p.nprims + 31u) & ~31u)) p.rects[t] = rect;  // the rect array is padded to whole groups (a warp reads a group's 32 rectangles with one coalesced load)
    sr_bin_group<false>(p, rect, t >> 5);
}
// pass 2: fill the per-tile group lists (order inside a list is arbitrary; consumers that need
// submission order sort the list, which is tiny because entries are groups of 32 primitives)
__global__ void __launch_bounds__(SR_BIN_THREADS) k_bin_fill(const SrBinParams p) {
    if (p.tile_off[p.ntx * p.nty] > p.capacity) return;  // lists do not fit: the host re-runs the pass with a larger arena
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t rect = t < p.nprims ? __ldg(p.rects + t) : SR_RECT_INVALID;
    sr_bin_group<true>(p, rect, t >> 5);
}

// The ordered path's bins for small draws (<= 8192 primitives) in ONE single-CTA launch: per-primitive rectangles, per-tile
// counts of 32-primitive groups (shared-memory counters), exclusive scan, list fill -- instead of a memset + k_bin_setup +
// k_tile_offsets + k_bin_fill, each of which costs a launch and a few dependent round trips (20 us apiece for a
// 968-triangle model, profiles/r1i_ordered_kernels_summary.txt).  One warp per group; the tiles of the group's union
// rectangle are dealt to the lanes and tested against the 32 rectangles (handed round with shuffles).
#define SR_BIN_SMALL_G_THREADS 1024
template <int NV>
__global__ void __launch_bounds__(SR_BIN_SMALL_G_THREADS) k_bin_small_groups(const SrBinParams p, uint32_t *tile_off) {
    extern __shared__ uint32_t s_cnt[];  // per-tile counters, later the fill cursors
    __shared__ uint32_t s_wsum[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t ntiles = p.ntx * p.nty;
    const uint32_t nprims = min(p.nprims, sr_prim_count(p.src));  // (the true count of a clip stage that did not synchronise)
    const uint32_t ngroups = (nprims + 31u) / 32u;
    for (uint32_t i = tid; i < ntiles; i += SR_BIN_SMALL_G_THREADS) s_cnt[i] = 0;
    __syncthreads();
    __shared__ uint32_t s_bits[SR_BIN_SMALL_G_THREADS / 32][32];  // per warp: which tiles of the group's union rectangle are hit
    auto place = [&](uint32_t rect, uint32_t group, bool fill) {
        const bool valid = rect != SR_RECT_INVALID;
        if (!__any_sync(0xffffffffu, valid)) return;
        uint32_t gx0 = valid ? (rect & 255u) : 255u, gy0 = valid ? ((rect >> 8) & 255u) : 255u;
        uint32_t gx1 = valid ? ((rect >> 16) & 255u) : 0u, gy1 = valid ? (rect >> 24) : 0u;
        gx0 = __reduce_min_sync(0xffffffffu, gx0);
        gy0 = __reduce_min_sync(0xffffffffu, gy0);
        gx1 = __reduce_max_sync(0xffffffffu, gx1);
        gy1 = __reduce_max_sync(0xffffffffu, gy1);
        const uint32_t gw = gx1 - gx0 + 1, n = gw * (gy1 - gy0 + 1);
        if (n <= 1024u) {
            // every primitive marks the tiles of its own rectangle in a bitmap of the union rectangle (rectangles of more
            // than 32 tiles are spread over the lanes), then each lane posts the set bits of one bitmap word
            uint32_t *bits = s_bits[warp];
            bits[lane] = 0;
            __syncwarp();
            const uint32_t rx0 = rect & 255u, ry0 = (rect >> 8) & 255u, rx1 = (rect >> 16) & 255u, ry1 = rect >> 24;
            const uint32_t rw = rx1 - rx0 + 1, rn = valid ? rw * (ry1 - ry0 + 1) : 0u;
            const bool big = rn > 32u;
            if (!big)
                for (uint32_t i = 0; i < rn; ++i) {
                    const uint32_t at = (ry0 + i / rw - gy0) * gw + (rx0 + i % rw - gx0);
                    atomicOr(&bits[at >> 5], 1u << (at & 31u));
                }
            for (uint32_t rest = __ballot_sync(0xffffffffu, big); rest; rest &= rest - 1) {
                const uint32_t r = __shfl_sync(0xffffffffu, rect, __ffs(rest) - 1);
                const uint32_t bx0 = r & 255u, by0 = (r >> 8) & 255u, bx1 = (r >> 16) & 255u, by1 = r >> 24;
                const uint32_t bw = bx1 - bx0 + 1, bn = bw * (by1 - by0 + 1);
                for (uint32_t i = lane; i < bn; i += 32) {
                    const uint32_t at = (by0 + i / bw - gy0) * gw + (bx0 + i % bw - gx0);
                    atomicOr(&bits[at >> 5], 1u << (at & 31u));
                }
            }
            __syncwarp();
            for (uint32_t word = bits[lane]; word; word &= word - 1) {
                const uint32_t i = lane * 32u + (uint32_t)(__ffs(word) - 1);
                const uint32_t tile = (gy0 + i / gw) * p.ntx + gx0 + i % gw;
                if (tile % p.shard_world != p.shard_rank) continue;
                const uint32_t at = atomicAdd(&s_cnt[tile], 1u);
                if (fill) p.list[at] = group;
            }
            __syncwarp();
            return;
        }
        for (uint32_t base = 0; base < n; base += 32) {
            const uint32_t i = base + lane;
            const uint32_t tx = gx0 + i % gw, ty = gy0 + i / gw;
            bool hit = false;
#pragma unroll 4
            for (int l = 0; l < 32; ++l) {
                const uint32_t r = __shfl_sync(0xffffffffu, rect, l);
                hit = hit || (r != SR_RECT_INVALID && sr_rect_hits(r, tx, ty));
            }
            const uint32_t tile = ty * p.ntx + tx;
            if (i < n && hit && tile % p.shard_world == p.shard_rank) {
                const uint32_t at = atomicAdd(&s_cnt[tile], 1u);
                if (fill) p.list[at] = group;
            }
        }
    };
    constexpr uint32_t GW = SR_BIN_SMALL_G_THREADS / 32;  // groups in flight
    for (uint32_t g = warp; g < ngroups; g += GW) {
        const uint32_t t = g * 32 + lane;
        const uint32_t rect = t < nprims ? sr_prim_rect<NV>(p, t) : SR_RECT_INVALID;
        p.rects[t] = rect;  // (the array is padded to whole groups)
        place(rect, g, false);
    }
    __syncthreads();
    const uint32_t chunk = (ntiles + SR_BIN_SMALL_G_THREADS - 1) / SR_BIN_SMALL_G_THREADS;
    const uint32_t b = min(tid * chunk, ntiles), e = min(b + chunk, ntiles);
    uint32_t sum = 0;
    for (uint32_t i = b; i < e; ++i) sum += s_cnt[i];
    uint32_t inc = sum;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += v;
    }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_wsum[lane];
#pragma unroll
        for (uint32_t d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += v;
        }
        s_wsum[lane] = w;
    }
    __syncthreads();
    const uint32_t total = s_wsum[31];
    uint32_t run = (warp ? s_wsum[warp - 1] : 0u) + inc - sum;
    for (uint32_t i = b; i < e; ++i) {
        const uint32_t c = s_cnt[i];
        tile_off[i] = run;
        s_cnt[i] = run;
        run += c;
    }
    if (tid == 0) tile_off[ntiles] = total;
    __syncthreads();
    if (total > p.capacity) return;  // lists do not fit: the host re-runs this launch with a larger arena
    for (uint32_t g = warp; g < ngroups; g += GW) {
        const uint32_t t = g * 32 + lane;
        place(p.rects[t], g, true);
    }
}

// exclusive scan of the per-tile counts (a few thousand tiles: one block of 1024 threads, each thread owning a run of
// consecutive tiles so that one block-wide scan suffices), total to off[ntiles]; the counts are zeroed for re-use as cursors
#define SR_OFFSETS_THREADS 1024
__global__ void __launch_bounds__(SR_OFFSETS_THREADS) k_tile_offsets(const uint32_t *count, uint32_t ntiles, uint32_t *off, uint32_t *cursor) {
    __shared__ uint32_t ws[32];
    const uint32_t per = (ntiles + SR_OFFSETS_THREADS - 1) / SR_OFFSETS_THREADS;
    const uint32_t beg = min(threadIdx.x * per, ntiles), end = min(beg + per, ntiles);
    uint32_t sum = 0;
    for (uint32_t i = beg; i < end; ++i) sum += count[i];
    uint32_t total;
    uint32_t run = sr_block_exclusive_scan(sum, &total, ws);
    for (uint32_t i = beg; i < end; ++i) {
        const uint32_t v = count[i];
        off[i] = run;
        cursor[i] = 0;
        run += v;
    }
    if (threadIdx.x == 0) off[ntiles] = total;
}

// =====================================================================================================
// shared per-triangle arithmetic (rasterization/triangle.rs:64,104-115)
// =====================================================================================================
struct SrTri {
    float x3, y3;
    float a, b, c, d;  // (y2-y3), (x3-x2), (y3-y1), (x1-x3)
    float det;
    float rdet;        // correctly rounded 1/det (valid when `fast`)
    uint32_t dsign;    // sign bit of det
    bool fast;         // |det| in [2^-40, 2^40]: the exact shortcuts below are valid
};
__device__ __forceinline__ SrTri sr_tri_setup(float x1, float y1, float x2, float y2, float x3, float y3) {
    SrTri t;
    t.x3 = x3; t.y3 = y3;
    t.a = y2 - y3; t.b = x3 - x2; t.c = y3 - y1; t.d = x1 - x3;
    t.det = t.a * (x1 - x3) + t.b * (y1 - y3);
    const uint32_t db = __float_as_uint(t.det);
    t.dsign = db & 0x80000000u;
    t.fast = ((db & 0x7FFFFFFFu) - 0x2B800000u) < (0x53800000u - 0x2B800000u);
    t.rdet = __frcp_rn(t.det);
    return t;
}
// numerators of u and v at pixel centre (x,y) (triangle.rs:108-109)
__device__ __forceinline__ void sr_tri_numerators(const SrTri &t, float x, float y, float &nu, float &nv) {
    const float dx = x - t.x3, dy = y - t.y3;
    nu = t.a * dx + t.b * dy;
    nv = t.c * dx + t.d * dy;
}
// u = nu/det, v = nv/det, w = 1-u-v and the inside test, bit-exact with triangle.rs:108-113.
// Exact early-out: for |det| in the fast range and finite |n| >= 2^-60 the quotient n/det is a non-zero float
// whose sign is sign(n)*sign(det), so `n/det < 0` iff the sign bits differ.
__device__ __forceinline__ bool sr_tri_inside(const SrTri &t, float nu, float nv, float &u, float &v, float &w) {
    const uint32_t bu = __float_as_uint(nu), bv = __float_as_uint(nv);
    if (t.fast) {
        if (((bu ^ t.dsign) - 0xA1800000u) < 0x5E000000u) return false;  // negative quotient, magnitude in [2^-60, inf)
        if (((bv ^ t.dsign) - 0xA1800000u) < 0x5E000000u) return false;
        const bool fu = ((bu & 0x7FFFFFFFu) - 0x21800000u) < (0x5D800000u - 0x21800000u);
        const bool fv = ((bv & 0x7FFFFFFFu) - 0x21800000u) < (0x5D800000u - 0x21800000u);
        if (fu && fv) {
            u = sr_div_exact(nu, t.det, t.rdet);
            v = sr_div_exact(nv, t.det, t.rdet);
        } else {
            u = nu / t.det;
            v = nv / t.det;
        }
    } else {
        u = nu / t.det;
        v = nv / t.det;
    }
    w = 1.0f - u - v;
    return !(u < 0.0f || v < 0.0f || w < 0.0f);
}
// Barycentrics of pixel (px,py) exactly as triangle.rs:104-113.  Returns false when the pixel is outside.
__device__ __forceinline__ bool sr_tri_bary(const SrTri &t, uint32_t px, uint32_t py, float &u, float &v, float &w) {
    float nu, nv;
    sr_tri_numerators(t, (float)px + 0.5f, (float)py + 0.5f, nu, nv);
    return sr_tri_inside(t, nu, nv, u, v, w);
}


// =====================================================================================================
// sm_100a async-copy plumbing: mbarrier and cp.async.bulk (TMA bulk copy, SASS UBLKCP)
// =====================================================================================================
__device__ __forceinline__ uint32_t sr_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sr_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sr_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sr_mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void sr_mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sr_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sr_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sr_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void sr_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "SR_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra SR_DONE_%=;\n\t"
        "bra SR_WAIT_%=;\n\t"
        "SR_DONE_%=:\n\t"
        "}" ::"r"(sr_smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (16-byte aligned, size a multiple of 16), completion counted on `bar`
__device__ __forceinline__ void sr_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sr_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(sr_smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy (tile write-back), tracked by the bulk async-group of the issuing thread
__device__ __forceinline__ void sr_bulk_s2g(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(sr_smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sr_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct SrTileParams {
    SrPrimSource tris, lines, points;
    uint32_t ntris, nlines, npoints;
    const uint32_t *tri_rects, *line_rects, *point_rects;
    const uint32_t *tri_off, *line_off, *point_off;  // per-tile CSR offsets (ntiles+1) into the group lists
    uint32_t *tri_list, *line_list, *point_list;
    uint32_t tri_cap, line_cap, point_cap;  // list capacities: the kernel skips itself if a scanned total exceeds its list
    SrFbView fb;
    uint32_t shard_rank, shard_world;
    uint32_t blend, stencil_test, stencil_op, stencil_value, aa_lines;
    uint32_t line_base, point_base;  // canonical index of the first line / point (after all triangles / lines)
    SrFsConst fs;
};

__device__ __forceinline__ float *sr_fb_pixel(const SrFbView &fb, uint32_t px, uint32_t py) {  // (20-byte pixels; 8-byte ones: address below)
    return fb.aos + ((uint64_t)py * fb.width + px) * 5;
}
__device__ __forceinline__ void *sr_fb_pixel_addr(const SrFbView &fb, uint32_t px, uint32_t py) {
    return reinterpret_cast<unsigned char *>(fb.aos) + ((uint64_t)py * fb.width + px) * (fb.u8color ? 8u : 20u);
}

// =====================================================================================================
// Opaque path: Blend = (), stencil test Always / op Keep, shader never discards.
// In that state the reference's in-order result at a pixel is the covering fragment with z<0 that
// maximises (z, submission index) among those with z >= the depth already stored (`d >= dt`, later
// primitives win ties, triangle.rs:120-126).  That is order-independent, so every fragment is reduced
// into one 64-bit key per pixel, (order-preserving depth bits << 32 | primitive+1), with atomicMax; each
// pixel is then shaded ONCE and the tile is written back to HBM once.
//
// Two producers feed the keys:
//  * k_micro: one thread per triangle, in submission order.  A triangle whose frame-clamped bounding box
//    holds at most `micro_area` pixels is rasterised on the spot into the frame's visibility buffer (keys in
//    HBM/L2, row-major with the pitch padded to whole tiles) -- no bin lists, no second visit.
//    Larger triangles are appended to a compact list and counted per tile.
//  * k_tile_opaque: one CTA per tile.  Pulls the tile's keys into shared memory with TMA bulk copies (one 512 B
//    row each), sweeps the tile's (large) triangle list warp-cooperatively against them, then resolves.
// =====================================================================================================
#define SR_VIS_FAR_KEY ((unsigned long long)(~SR_DEPTH_FAR_BITS) << 32)  // sr_depth_key(f32::MIN) << 32, primitive 0
static_assert((SR_DEPTH_FAR_BITS & 0x80000000u) != 0, "far depth is negative");

// visibility buffer: (nty * SR_TILE_H) rows of pitch = ntx * SR_TILE_W keys, so every tile row is a full 512 B run
__device__ __forceinline__ uint32_t sr_vis_index(uint32_t px, uint32_t py, uint32_t ntx) { return py * (ntx * SR_TILE_W) + px; }

struct SrMicroParams {
    SrPrimSource src;
    uint32_t ntris, cull;
    uint32_t tri_begin, tri_end;  // k_micro walks [tri_begin, tri_end) (the whole draw unless the front end is range-sharded, section 6)
    uint32_t width, height, ntx, nty;
    uint32_t shard_rank, shard_world;
    uint32_t micro_area;          // bbox pixels up to which a triangle is rasterised by k_micro (0: none)
    unsigned long long *vis;      // visibility buffer (ntiles * SR_TILE_PIXELS keys, see sr_vis_index)
    uint32_t *large_count;        // number of entries in large_ids
    uint32_t *large_ids;          // triangles left to the tile kernel ...
    uint32_t *large_rects;        // ... and their packed tile rectangles
    uint32_t *tile_count;         // per tile: number of large triangles touching it
};

// keys of the tiles this rank owns := far (pending clear) or the depth already in the framebuffer
__global__ void __launch_bounds__(256) k_vis_init(unsigned long long *vis, const SrFbView fb, uint32_t shard_rank, uint32_t shard_world) {
    const uint32_t tile = shard_rank + blockIdx.x * shard_world;
    const uint32_t x0 = (tile % fb.ntx) * SR_TILE_W, y0 = (tile / fb.ntx) * SR_TILE_H;
    for (uint32_t i = threadIdx.x * 2; i < SR_TILE_PIXELS; i += 512) {  // two keys (16 B) per thread and step
        const uint32_t px = x0 + i % SR_TILE_W, py = y0 + i / SR_TILE_W;
        uint32_t dk0 = ~SR_DEPTH_FAR_BITS, dk1 = ~SR_DEPTH_FAR_BITS;
        if (!fb.pending_clear && py < fb.height) {
            if (px < fb.width) dk0 = sr_depth_key(sr_fb_load_depth(fb, (uint64_t)py * fb.width + px));
            if (px + 1 < fb.width) dk1 = sr_depth_key(sr_fb_load_depth(fb, (uint64_t)py * fb.width + px + 1));
        }
        *reinterpret_cast<ulonglong2 *>(vis + sr_vis_index(px, py, fb.ntx)) =
            make_ulonglong2((unsigned long long)dk0 << 32, (unsigned long long)dk1 << 32);
    }
}

// ---- range-sharded front end: every rank rasterises its own triangle range into a full-frame
// key buffer of its own; the tile's owner pulls the other ranks' keys over NVLink and max-merges them (k_tile_opaque,
// PHASE 2).  Keys of tiles the rank does not own are handed back "far" here (its own tiles are reset by its resolve).
// Tile ownership of a range-sharded frame: tile t belongs to rank owner[t % period].  Rank 0 -- whose write-back is local
// while every other rank's crosses NVLink, but whose tiles cost (world-1) key pulls -- may hold a different share than the rest.
#define SR_OWNER_PERIOD_MAX 64
#define SR_SHARD_MAX_WORLD 8
struct SrTileOwners {
    uint8_t owner[SR_OWNER_PERIOD_MAX];
    uint32_t period;
    uint32_t by_rows, ntx, world;  // by_rows: tile ROW y belongs to rank y % world (frames whose front end is culled by screen rows, below)
};
__device__ __forceinline__ uint32_t sr_tile_owner(const SrTileOwners &o, uint32_t tile) {
    return o.by_rows ? (tile / o.ntx) % o.world : (uint32_t)o.owner[tile % o.period];
}
__device__ __forceinline__ void sr_fill_tile_clear(const SrFbView &fb, uint32_t x0, uint32_t y0) {
    if (fb.soa) {
        const float o[9] = {fb.clear[0], fb.clear[1], fb.clear[2], fb.clear[3], __uint_as_float(SR_DEPTH_FAR_BITS),
                            fb.clear1[0], fb.clear1[1], fb.clear1[2], fb.clear1[3]};
        const uint32_t cols = min(x0 + SR_TILE_W, fb.width) - x0;
        for (uint32_t i = threadIdx.x; i < cols * SR_TILE_H; i += 256) {
            const uint32_t py = y0 + i / cols;
            if (py >= fb.height) continue;
            if (fb.soa == 2) sr_fb_store_pixel2(fb, (uint64_t)py * fb.width + x0 + i % cols, o);
            else sr_fb_store_pixel(fb, (uint64_t)py * fb.width + x0 + i % cols, o);
        }
        return;
    }
    if (fb.u8color) {
        float q[4] = {fb.clear[0], fb.clear[1], fb.clear[2], fb.clear[3]};
        sr_quantise_u8(q, false, 0);
        const uint32_t pat[2] = {sr_pack_u8(q), SR_DEPTH_FAR_BITS};
        const uint32_t run = (min(x0 + SR_TILE_W, fb.width) - x0) * 2;  // words per tile row inside the frame
        for (uint32_t r = 0; r < SR_TILE_H && y0 + r < fb.height; ++r) {
            uint32_t *row = reinterpret_cast<uint32_t *>(fb.aos) + ((uint64_t)(y0 + r) * fb.width + x0) * 2;
            for (uint32_t i = threadIdx.x; i < run; i += 256) row[i] = pat[i & 1u];
        }
        return;
    }
    const uint32_t run = (min(x0 + SR_TILE_W, fb.width) - x0) * 5;  // floats per tile row inside the frame
    const float pat[5] = {fb.clear[0], fb.clear[1], fb.clear[2], fb.clear[3], __uint_as_float(SR_DEPTH_FAR_BITS)};
    if ((fb.width & 3u) == 0 && (reinterpret_cast<uintptr_t>(fb.aos) & 15u) == 0 && x0 + SR_TILE_W <= fb.width && y0 + SR_TILE_H <= fb.height) {
        // A whole tile (the usual case; config 3's frame is 56 % background and every empty tile comes through here): a tile row is
        // 80 float4, the 20-byte pixel pattern repeats every five of them.  Thread t < 240 owns column t % 80 of the rows t / 80,
        // t / 80 + 3, ...: its float4 is the same in every row, so it is formed once and the loop is eleven 16-byte stores.
        static_assert(SR_TILE_W * 5 / 4 == 80 && SR_TILE_H == 32, "column / row ownership below");
        const uint32_t t = threadIdx.x;
        if (t < 240u) {
            const uint32_t q = t % 80u, f = (q % 5u) * 4u;  // first float of this float4 within the pattern of four pixels
            auto at = [&](uint32_t i) { const uint32_t m = i % 5u; return m == 0 ? pat[0] : m == 1 ? pat[1] : m == 2 ? pat[2] : m == 3 ? pat[3] : pat[4]; };
            const float4 v = make_float4(at(f), at(f + 1u), at(f + 2u), at(f + 3u));
            const uint64_t pitch4 = (uint64_t)fb.width * 5u / 4u;  // float4 per framebuffer row
            float4 *dst = reinterpret_cast<float4 *>(fb.aos + ((uint64_t)y0 * fb.width + x0) * 5) + (uint64_t)(t / 80u) * pitch4 + q;
#pragma unroll 1
            for (uint32_t r = t / 80u; r < SR_TILE_H; r += 3u, dst += 3u * pitch4) *dst = v;
        }
        return;
    }
    if ((fb.width & 3u) == 0 && (run & 3u) == 0 && (reinterpret_cast<uintptr_t>(fb.aos) & 15u) == 0) {
        // 16-byte stores: the 20-byte pixel pattern repeats every five float4 (x0 is a multiple of 64 pixels = 1280 bytes)
        const uint32_t run4 = run / 4, rows = min(SR_TILE_H, fb.height - y0);
        for (uint32_t i = threadIdx.x; i < run4 * rows; i += 256) {
            const uint32_t r = i / run4, q = i % run4, k = (q % 5u) * 4u;
            float4 *row = reinterpret_cast<float4 *>(fb.aos + ((uint64_t)(y0 + r) * fb.width + x0) * 5);
            row[q] = make_float4(pat[k % 5u], pat[(k + 1u) % 5u], pat[(k + 2u) % 5u], pat[(k + 3u) % 5u]);
        }
        return;
    }
    for (uint32_t r = 0; r < SR_TILE_H && y0 + r < fb.height; ++r) {
        float *row = fb.aos + ((uint64_t)(y0 + r) * fb.width + x0) * 5;
        for (uint32_t i = threadIdx.x; i < run; i += 256) row[i] = pat[i % 5];
    }
}
// `fill`: rank 0 also writes the clear colour + far depth into the framebuffer pixels of the tiles it does NOT own, so that their
// owners need not send pixels nothing was drawn on over NVLink (the frame starts from a clear: that is what selects this path).
__global__ void __launch_bounds__(256) k_vis_clear_foreign(unsigned long long *vis, const SrFbView fb, const SrTileOwners own, uint32_t rank,
                                                           uint32_t all, uint32_t fill) {
    const uint32_t tile = blockIdx.x;
    const bool mine = sr_tile_owner(own, tile) == rank;
    if (!all && mine) return;
    const uint32_t x0 = (tile % fb.ntx) * SR_TILE_W, y0 = (tile / fb.ntx) * SR_TILE_H;
    for (uint32_t i = threadIdx.x * 2; i < SR_TILE_PIXELS; i += 512)
        *reinterpret_cast<ulonglong2 *>(vis + sr_vis_index(x0 + i % SR_TILE_W, y0 + i / SR_TILE_W, fb.ntx)) =
            make_ulonglong2(SR_VIS_FAR_KEY, SR_VIS_FAR_KEY);
    if (!fill || mine || x0 >= fb.width) return;
    sr_fill_tile_clear(fb, x0, y0);
}
// the framebuffer half of the above on its own (rank 0 runs it on a second stream, beside its k_micro)
__global__ void __launch_bounds__(256) k_fb_fill_foreign(const SrFbView fb, const SrTileOwners own, uint32_t rank) {
    const uint32_t tile = blockIdx.x;
    if (sr_tile_owner(own, tile) == rank) return;
    const uint32_t x0 = (tile % fb.ntx) * SR_TILE_W, y0 = (tile / fb.ntx) * SR_TILE_H;
    if (x0 < fb.width) sr_fill_tile_clear(fb, x0, y0);
}
// The merge of a range-sharded frame as a streaming kernel of its own (the alternative to merging inside the resolve,
// k_tile_opaque PHASE 2 with npeers > 0): one light CTA per owned tile pulls the peers' keys of that tile straight from
// peer-mapped memory (16-byte coalesced loads over NVLink, all peers' loads of a thread in flight together), max-merges them
// into the rank's own buffer and marks the vertices the winners reference in `mark` (one bit per vertex), so that the vertex
// stage can afterwards shade exactly those -- the rank's vertex stage no longer runs over the whole mesh.
// Per foreign tile: which of its 32 key rows hold anything but "far" (bit r = row r).  The owner of the tile reads the word
// first and pulls only those rows -- a rank whose triangles leave part of the frame untouched sends nothing for it.
__global__ void __launch_bounds__(256) k_vis_rows_touched(const unsigned long long *vis, uint32_t ntx, const SrTileOwners own, uint32_t rank,
                                                          uint32_t *touched) {
    const uint32_t tile = blockIdx.x;
    if (sr_tile_owner(own, tile) == rank) return;
    __shared__ uint32_t bits;
    if (threadIdx.x == 0) bits = 0;
    __syncthreads();
    const uint32_t x0 = (tile % ntx) * SR_TILE_W, y0 = (tile / ntx) * SR_TILE_H;
    uint32_t mine = 0;
#pragma unroll
    for (uint32_t j = 0; j < SR_TILE_PIXELS / 2 / 256; ++j) {
        const uint32_t i = (j * 256 + threadIdx.x) * 2;
        const ulonglong2 k = *reinterpret_cast<const ulonglong2 *>(vis + sr_vis_index(x0 + i % SR_TILE_W, y0 + i / SR_TILE_W, ntx));
        if (k.x != SR_VIS_FAR_KEY || k.y != SR_VIS_FAR_KEY) mine |= 1u << (i / SR_TILE_W);
    }
    static_assert(SR_TILE_H <= 32, "one bit per tile row");
    mine = __reduce_or_sync(0xffffffffu, mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicOr(&bits, mine);
    __syncthreads();
    if (threadIdx.x == 0) touched[tile] = bits;
}
struct SrMergeParams {
    unsigned long long *vis;
    const unsigned long long *peer_vis[SR_SHARD_MAX_WORLD - 1];
    const uint32_t *peer_touched[SR_SHARD_MAX_WORLD - 1];  // the peers' row masks (k_vis_rows_touched)
    uint32_t npeers, ntx, rank;
    SrTileOwners owners;
    const uint32_t *indices;  // winners' vertex indices (indexed triangles only)
    uint32_t ntris;
    uint32_t *mark;           // or null
    uint32_t skip_lo, skip_hi;  // vertices in [skip_lo, skip_hi) are shaded already
    const uint8_t *shaded_blocks;  // or: vertices of the 256-vertex blocks flagged here are shaded already (chunk-culled front end)
};
__global__ void __launch_bounds__(256) k_shard_merge(const __grid_constant__ SrMergeParams p) {
    const uint32_t tile = blockIdx.x;
    if (sr_tile_owner(p.owners, tile) != p.rank) return;
    const uint32_t x0 = (tile % p.ntx) * SR_TILE_W, y0 = (tile / p.ntx) * SR_TILE_H;
    __shared__ uint32_t rows[SR_SHARD_MAX_WORLD - 1];
    if (threadIdx.x < p.npeers) rows[threadIdx.x] = __ldcv(p.peer_touched[threadIdx.x] + tile);
    __syncthreads();
#pragma unroll
    for (uint32_t j = 0; j < SR_TILE_PIXELS / 2 / 256; ++j) {
        const uint32_t i = (j * 256 + threadIdx.x) * 2;
        const uint32_t row = i / SR_TILE_W;
        const uint32_t at = sr_vis_index(x0 + i % SR_TILE_W, y0 + row, p.ntx);
        ulonglong2 k = *reinterpret_cast<const ulonglong2 *>(p.vis + at);
        ulonglong2 r[SR_SHARD_MAX_WORLD - 1];
        bool any = false;
#pragma unroll
        for (uint32_t q = 0; q < SR_SHARD_MAX_WORLD - 1; ++q) {
            r[q] = make_ulonglong2(0ull, 0ull);
            if (q < p.npeers && ((rows[q] >> row) & 1u)) {
                r[q] = __ldcv(reinterpret_cast<const ulonglong2 *>(p.peer_vis[q] + at));
                any = true;
            }
        }
#pragma unroll
        for (uint32_t q = 0; q < SR_SHARD_MAX_WORLD - 1; ++q) {
            k.x = r[q].x > k.x ? r[q].x : k.x;
            k.y = r[q].y > k.y ? r[q].y : k.y;
        }
        if (any) *reinterpret_cast<ulonglong2 *>(p.vis + at) = k;
        if (p.mark != nullptr) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t id = (uint32_t)(h ? k.y : k.x);
                if (id == 0 || id - 1 >= p.ntris) continue;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const uint32_t v = __ldg(p.indices + (uint64_t)(id - 1) * 3 + c);
                    if (v >= p.skip_lo && v < p.skip_hi) continue;
                    if (p.shaded_blocks != nullptr && p.shaded_blocks[v >> 8]) continue;
                    const uint32_t bit = 1u << (v & 31u);
                    if ((p.mark[v >> 5] & bit) == 0) atomicOr(p.mark + (v >> 5), bit);
                }
            }
        }
    }
}

// Cross-rank progress words live in every rank's exchange block; a rank publishes its frame number by storing it into
// slot `me` of every peer's block (system-scope release after a system fence: everything earlier kernels of this stream
// wrote is visible to a peer that acquires the word), and waits by polling its OWN block.
struct SrShardPeers {
    uint32_t *word[SR_SHARD_MAX_WORLD];  // peer p's copy of the word array (null for p == me / absent)
};
__global__ void k_shard_signal(const SrShardPeers peers, uint32_t me, uint32_t value) {
    const uint32_t p = threadIdx.x;
    if (p >= SR_SHARD_MAX_WORLD || peers.word[p] == nullptr) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peers.word[p] + me), "r"(value) : "memory");
}
// spins until every peer's word is >= value; gives up after `timeout_ns` (a peer died) and raises *error so that the host
// reports a failure instead of hanging the GPU
__global__ void k_shard_wait(const uint32_t *words, uint32_t world, uint32_t me, uint32_t value, unsigned long long timeout_ns,
                             uint32_t *error) {
    const uint32_t p = threadIdx.x;
    if (p >= world || p == me) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(words + p) : "memory");
        if ((int32_t)(v - value) >= 0) return;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) {
            atomicExch(error, 1u + p);
            return;
        }
        __nanosleep(200);
    }
}

// ---- candidate pixels of a small triangle -----------------------------------------------------------------
// The reference tests every pixel of the integer bounding box [trunc(min), trunc(max)] (triangle.rs:74-85).
// For a small, well-conditioned triangle most of those pixels are provably rejected by the reference's own f32
// inside test, so they need not be visited (DESIGN.md "candidate tightening" carries the proof):
//   let hx = xmax-xmin < 5, hy = ymax-ymin < 5 and |det| >= 1.  A pixel centre c = px+0.5 with c <= xmin - m
//   (m = 1/16 - rounding slack) has an exact barycentric weight lambda_k <= -m/(2 hx) (the centre is a convex
//   combination of the vertices only if it lies in their bounding box).  The f32 values u, v and w = (1-u)-v differ
//   from the exact weights by at most 2(E + L*Ed)/|det| + 3 eps (1 + 2L) with E <= 4.0001 eps (2 hx hy + (hx+hy)/2),
//   Ed <= 8.0002 eps hx hy, L = |lambda| <= 55/|det|: below 1.4e-3 < m/(2 hx) = 6.2e-3.  So the most negative of
//   u, v, w IS negative in f32 and the reference rejects the pixel.  Same on the other three sides.
// Pixels inside the kept range run the reference's exact test; results are bit-identical either way.
__device__ __forceinline__ bool sr_tightening_applies(float xmin, float xmax, float ymin, float ymax, float det) {
    return xmax - xmin < 4.99f && ymax - ymin < 4.99f && fabsf(det) >= 1.0f;  // NaN-safe: any NaN coordinate gives a NaN det -> false
}
// keep px iff xmin - 1/16 < px + 0.5 < xmax + 1/16 (the rounding of the subtraction, <= 2^-8 for |x| < 2^16, is inside the slack)
__device__ __forceinline__ int sr_tight_lo(float vmin) { return __float2int_ru(vmin - 0.5625f); }
__device__ __forceinline__ int sr_tight_hi(float vmax) { return __float2int_rd(vmax - 0.4375f); }

// One lane walks a small pixel box [minx, minx+bw) x [miny, miny+bh) of its triangle and calls
// emit(px, py, key) for every fragment that passes coverage and z<0 (triangle.rs:104-120).
// Pixel centres are generated incrementally: (float)px + 0.5f is exact and so is adding 1.0f to it, so xf/yf carry
// exactly the values triangle.rs:104-105 computes; the row terms b*dy and d*dy are hoisted (same roundings).
// The body is written for SIMT: lanes of a warp sit on different triangles, so the warp executes the union of all
// paths anyway -- u and v are therefore always computed (exact-division shortcut when valid) instead of branching
// on the sign of the numerators first.
// BOUNDED: the caller guarantees |det| in [1, 2^40] and numerators below 2^60 (k_micro's tightened path: extents
// below 5 pixels), so only the lower validity bound of the exact-division shortcut has to be checked per pixel.
template <bool BOUNDED, class Emit>
__device__ __forceinline__ void sr_raster_box(const SrTri &tr, float z1, float z2, float z3, uint32_t minx, uint32_t miny,
                                              uint32_t bw, uint32_t bh, uint32_t id, Emit emit) {
    const float xf0 = (float)minx + 0.5f;
    float xf = xf0, yf = (float)miny + 0.5f;
    float dy = yf - tr.y3, bdy = tr.b * dy, ddy = tr.d * dy;
    uint32_t px = minx, py = miny;
    const uint32_t n = bw * bh, maxx = minx + bw - 1;
    // validity range of sr_div_exact folded into two constants (an invalid det makes the range empty)
    const float lo = (BOUNDED || tr.fast) ? 0x1p-60f : __int_as_float(0x7f800000), hi = 0x1p60f;
#pragma unroll 1
    for (uint32_t i = 0; i < n; ++i) {
        const float dx = xf - tr.x3;
        const float nu = tr.a * dx + bdy, nv = tr.c * dx + ddy;
        float u, v;
        const float au = fabsf(nu), av = fabsf(nv);
        if (BOUNDED ? (au >= lo && av >= lo) : (au >= lo && au < hi && av >= lo && av < hi)) {
            u = sr_div_exact(nu, tr.det, tr.rdet);
            v = sr_div_exact(nv, tr.det, tr.rdet);
        } else {
            u = nu / tr.det;
            v = nv / tr.det;
        }
        const float w = 1.0f - u - v;
        if (!(u < 0.0f || v < 0.0f || w < 0.0f)) {
            const float z = (z1 * u + z2 * v) + z3 * w;
            if (z < 0.0f)  // triangle.rs:120; for negative z the order-preserving key is ~bits
                emit(px, py, ((unsigned long long)(~__float_as_uint(z)) << 32) | (unsigned long long)(id + 1u));
        }
        ++px; xf += 1.0f;
        if (px > maxx) {
            px = minx; xf = xf0; ++py;
            yf += 1.0f; dy = yf - tr.y3; bdy = tr.b * dy; ddy = tr.d * dy;
        }
    }
}

__device__ __forceinline__ void sr_red_max_u64(unsigned long long *addr, unsigned long long key) {
    asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(addr), "l"(key) : "memory");
}
__device__ __forceinline__ unsigned long long sr_ld_relaxed_u64(const unsigned long long *addr) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(addr) : "memory");
    return v;
}

// ---- early depth rejection of a whole small triangle ---------------------------------------------------------
// A fragment only changes the visibility buffer if its key exceeds the stored one, i.e. if its depth key is >= the
// stored depth key (triangle.rs:126, `d >= dt`).  Keys only grow while a frame is drawn, so a (possibly stale)
// read of the candidate pixels' depth keys is a lower bound of what the fragment will meet.  If an upper bound of
// every depth the triangle can produce is strictly below all of them, none of its fragments can win and the
// triangle is skipped before any coverage arithmetic.  Upper bound: emitted fragments have finite u, v, w in [0, 1]
// with |u + v + w - 1| <= 2^-23 (w = fl(fl(1-u)-v)); for z1, z2, z3 < 0 and M = max|zi| the f32 value
// fl(fl(fl(z1 u) + fl(z2 v)) + fl(z3 w)) is at most zmax + M (2^-23 + 3 * 2^-24) * 1.001 < zmax + M 2^-21; the bound used
// is fl(zmax + M 2^-20).  Triangles with a vertex at z >= 0 are never rejected here.  Results are bit-identical with
// or without the rule (tests: test_opaque_path_split_is_invisible, full-size order/idempotence properties).
__device__ __forceinline__ uint32_t sr_ld_depth_key(const unsigned long long *slot) {
    uint32_t v;  // high word of the little-endian 64-bit key; L2-coherent load (L1 may hold the previous frame's keys)
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1+4];" : "=r"(v) : "l"(slot) : "memory");
    return v;
}
__device__ __forceinline__ bool sr_micro_occluded(const unsigned long long *vis, uint32_t pitch, int lx, int ly, int hx, int hy, float z1,
                                                  float z2, float z3) {
    const float zmax = fmaxf(fmaxf(z1, z2), z3), zmin = fminf(fminf(z1, z2), z3);
    const float zb = zmax + zmin * -0x1p-20f;
    if (!(zb < 0.0f)) return false;  // also NaN
    const uint32_t kb = ~__float_as_uint(zb);  // depth key of a negative value
    const uint32_t cw = (uint32_t)(hx - lx + 1), ch = (uint32_t)(hy - ly + 1);
    const unsigned long long *row = vis + (uint32_t)ly * pitch + (uint32_t)lx;
    uint32_t kmin = 0xFFFFFFFFu;
    if (cw <= 3 && ch <= 3) {  // the common box: all loads in flight at once
#pragma unroll
        for (uint32_t r = 0; r < 3; ++r)
#pragma unroll
            for (uint32_t c = 0; c < 3; ++c) {
                uint32_t k = 0xFFFFFFFFu;
                if (r < ch && c < cw) k = sr_ld_depth_key(row + r * pitch + c);
                kmin = min(kmin, k);
            }
    } else {
        for (uint32_t r = 0; r < ch && kmin > kb; ++r, row += pitch)
            for (uint32_t c = 0; c < cw; ++c) kmin = min(kmin, sr_ld_depth_key(row + c));
    }
    return kb < kmin;
}

// The tightened path of k_micro: |det| in [1, 50], numerators below 2^7, no NaN.  Straight-line body (no divergent
// branch inside the loop; lanes sit on different triangles, so every branch would be taken by somebody anyway):
// u, v by the exact-division shortcut, the reduction as a predicated red.global.max.u64.  A numerator below the
// shortcut's validity bound (|n| < 2^-60, e.g. a pixel centre exactly on an edge) makes the lane redo its box with
// IEEE divisions afterwards (returns true); re-emitting a fragment is harmless because max is idempotent.
__device__ __forceinline__ bool sr_micro_box(const SrTri &tr, float z1, float z2, float z3, int lx, int ly, int hx, int hy, uint32_t id,
                                             unsigned long long *vis, uint32_t pitch) {
    const float xf0 = (float)lx + 0.5f, xlast = (float)hx + 0.5f;
    float xf = xf0, yf = (float)ly + 0.5f;
    float dy = yf - tr.y3, bdy = tr.b * dy, ddy = tr.d * dy;
    const uint32_t cw = (uint32_t)(hx - lx + 1), n = cw * (uint32_t)(hy - ly + 1);
    uint32_t idx = (uint32_t)ly * pitch + (uint32_t)lx;
    const uint32_t key_lo = id + 1u, row_skip = pitch - cw;
    bool redo = false;
#pragma unroll 1
    for (uint32_t i = 0; i < n; ++i) {
        const float dx = xf - tr.x3;
        const float nu = tr.a * dx + bdy, nv = tr.c * dx + ddy;
        const bool ok = fabsf(nu) >= 0x1p-60f && fabsf(nv) >= 0x1p-60f;
        const float u = sr_div_exact(nu, tr.det, tr.rdet), v = sr_div_exact(nv, tr.det, tr.rdet);
        const float w = 1.0f - u - v;
        const float z = (z1 * u + z2 * v) + z3 * w;
        const bool pass = ok && !(u < 0.0f || v < 0.0f || w < 0.0f) && z < 0.0f;  // triangle.rs:113,120
        redo = redo || !ok;
        // for negative z the order-preserving depth key is ~bits
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            ".reg .b64 k;\n\t"
            "setp.ne.u32 p, %3, 0;\n\t"
            "mov.b64 k, {%1, %2};\n\t"
            "@p red.relaxed.gpu.global.max.u64 [%0], k;\n\t"
            "}" ::"l"(vis + idx), "r"(key_lo), "r"(~__float_as_uint(z)), "r"((uint32_t)pass)
            : "memory");
        ++idx;
        const bool wrap = xf >= xlast;
        xf += 1.0f;
        if (wrap) {
            xf = xf0; idx += row_skip;
            yf += 1.0f; dy = yf - tr.y3; bdy = tr.b * dy; ddy = tr.d * dy;
        }
    }
    return redo;
}

#ifndef SR_MICRO_PF_DIST
#define SR_MICRO_PF_DIST 2048  // CTAs ahead whose index lines a k_micro CTA prefetches into L2 (0: off).  Measured on config 3: micro 128.8 ->
                                // 123.2 us, frame 0.278 -> 0.271 ms; 1024 / 4096 / 8192 ahead: 0.2725 / 0.2725 / 0.276.  Prefetching the POSITIONS of a
                                // later CTA as well (its indices loaded here) costs more than it hides: 0.288 ms
#endif
// 128-thread CTAs, 12 per SM (40 registers): the finer CTA granularity keeps ~4 more warps resident than 6 x 256
#define SR_MICRO_THREADS 128
#define SR_MICRO_MIN_BLOCKS 12
// Everything k_micro does for one triangle once its three screen positions are known (warp-collective: every lane of
// the warp calls it, `valid` = the lane holds a triangle).
template <bool PRECHECK, bool EARLYZ>
__device__ __forceinline__ void sr_micro_triangle(const SrMicroParams &p, uint32_t t, bool valid, const float4 &A, const float4 &B,
                                                  const float4 &C, uint32_t lane) {
    bool large = false;
    uint32_t rect = SR_RECT_INVALID;
    if (valid) {
        bool culled = false;
        if (p.cull != SR_CULL_NONE) {  // triangle.rs:54-61 (a NaN area is "not negative", like is_sign_negative of the reference's NaN)
            const float area = A.x * B.y + B.x * C.y + C.x * A.y - B.x * A.y - C.x * B.y - A.x * C.y;
            culled = (signbit(area) ? SR_CLOCKWISE : SR_COUNTER_CLOCKWISE) == p.cull;
        }
        if (!culled) {
            const SrTri tr = sr_tri_setup(A.x, A.y, B.x, B.y, C.x, C.y);
            const float xmin = fminf(fminf(A.x, B.x), C.x), xmax = fmaxf(fmaxf(A.x, B.x), C.x);
            const float ymin = fminf(fminf(A.y, B.y), C.y), ymax = fmaxf(fmaxf(A.y, B.y), C.y);
            const bool sharded = p.shard_world > 1;
            // walks the candidate box, reducing every fragment into the visibility buffer (tiles of other ranks are skipped
            // when sharded); `bounded` marks the tightened path
            auto raster = [&](auto bounded, int lx, int ly, int hx, int hy) {
                if (lx > hx || ly > hy) return;
                const uint32_t tx0 = (uint32_t)lx / SR_TILE_W, ty0 = (uint32_t)ly / SR_TILE_H;
                const bool one_tile = tx0 == (uint32_t)hx / SR_TILE_W && ty0 == (uint32_t)hy / SR_TILE_H;
                if (sharded && one_tile && (ty0 * p.ntx + tx0) % p.shard_world != p.shard_rank) return;
                const bool per_pixel_owner = sharded && !one_tile;
                if (EARLYZ && !per_pixel_owner && sr_micro_occluded(p.vis, p.ntx * SR_TILE_W, lx, ly, hx, hy, A.z, B.z, C.z)) return;
                sr_raster_box<decltype(bounded)::value>(
                    tr, A.z, B.z, C.z, (uint32_t)lx, (uint32_t)ly, (uint32_t)(hx - lx + 1), (uint32_t)(hy - ly + 1), t,
                    [&](uint32_t px, uint32_t py, unsigned long long key) {
                        if (per_pixel_owner && ((py / SR_TILE_H) * p.ntx + px / SR_TILE_W) % p.shard_world != p.shard_rank) return;
                        unsigned long long *slot = p.vis + sr_vis_index(px, py, p.ntx);
                        if (PRECHECK && !(key > sr_ld_relaxed_u64(slot))) return;
                        sr_red_max_u64(slot, key);
                    });
            };
            if (p.micro_area != 0 && sr_tightening_applies(xmin, xmax, ymin, ymax, tr.det)) {
                // small and well conditioned (this also implies six finite coordinates): the candidate pixels are the
                // reference's clamped bounding box intersected with the tightened range (rule above;
                // for x >= 0 trunc(xmin) <= ceil(xmin - 0.5625) and floor(xmax - 0.4375) <= trunc(xmax), so the
                // intersection is just the tightened range clamped to the frame)
                const int lx = max(0, sr_tight_lo(xmin)), ly = max(0, sr_tight_lo(ymin));
                const int hx = min((int)p.width - 1, sr_tight_hi(xmax)), hy = min((int)p.height - 1, sr_tight_hi(ymax));
                if (lx <= hx && ly <= hy) {
                    if (sharded) raster(std::true_type(), lx, ly, hx, hy);
                    else if (EARLYZ && sr_micro_occluded(p.vis, p.ntx * SR_TILE_W, lx, ly, hx, hy, A.z, B.z, C.z)) {}
                    else if (sr_micro_box(tr, A.z, B.z, C.z, lx, ly, hx, hy, t, p.vis, p.ntx * SR_TILE_W)) raster(std::false_type(), lx, ly, hx, hy);
                }
            } else if (!(isnan(A.x) || isnan(A.y) || isnan(B.x) || isnan(B.y) || isnan(C.x) || isnan(C.y))) {
                // (the reference panics on NaN coordinates, cast(..).unwrap(); defined here as "skipped")
                // triangle.rs:74-78 with tile = the whole frame
                const uint32_t minx = sr_clamp_as_int(xmin, 0, p.width - 1), maxx = sr_clamp_as_int(xmax, 0, p.width - 1);
                const uint32_t miny = sr_clamp_as_int(ymin, 0, p.height - 1), maxy = sr_clamp_as_int(ymax, 0, p.height - 1);
                if (minx <= maxx && miny <= maxy) {
                    if ((maxx - minx + 1) * (maxy - miny + 1) <= p.micro_area) {
                        raster(std::false_type(), (int)minx, (int)miny, (int)maxx, (int)maxy);
                    } else {
                        const uint32_t tx0 = minx / SR_TILE_W, ty0 = miny / SR_TILE_H, tx1 = maxx / SR_TILE_W, ty1 = maxy / SR_TILE_H;
                        rect = sr_pack_rect(tx0, ty0, tx1, ty1);
                        if (!sharded) {
                            large = true;
                        } else {
                            for (uint32_t ty = ty0; ty <= ty1 && !large; ++ty)
                                for (uint32_t tx = tx0; tx <= tx1; ++tx)
                                    if ((ty * p.ntx + tx) % p.shard_world == p.shard_rank) { large = true; break; }
                        }
                    }
                }
            }
        }
    }
    // large triangles: order-preserving (within the warp) append to the compact list + per-tile counts
    const uint32_t m = __ballot_sync(0xffffffffu, large);
    if (m == 0) return;
    uint32_t base = 0;
    if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(p.large_count, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (large) {
        const uint32_t slot = base + __popc(m & ((1u << lane) - 1u));
        p.large_ids[slot] = t;
        p.large_rects[slot] = rect;
    }
    // per-tile counts: the warp walks its large triangles one at a time and spreads each tile rectangle over the lanes
    // (a big triangle touches hundreds of tiles; one lane issuing that many dependent atomics would be the critical path)
    for (uint32_t rest = m; rest; rest &= rest - 1) {
        const uint32_t r = __shfl_sync(0xffffffffu, rect, __ffs(rest) - 1);
        const uint32_t tx0 = r & 255u, ty0 = (r >> 8) & 255u, tx1 = (r >> 16) & 255u, ty1 = r >> 24;
        const uint32_t rw = tx1 - tx0 + 1, n = rw * (ty1 - ty0 + 1);
        for (uint32_t i = lane; i < n; i += 32) {
            const uint32_t tile = (ty0 + i / rw) * p.ntx + tx0 + i % rw;
            if (tile % p.shard_world == p.shard_rank) atomicAdd(p.tile_count + tile, 1u);
        }
    }
}

// one thread per triangle, in submission order
template <bool PRECHECK, bool EARLYZ>
__global__ void __launch_bounds__(SR_MICRO_THREADS, SR_MICRO_MIN_BLOCKS) k_micro(const __grid_constant__ SrMicroParams p) {
    const uint32_t t = p.tri_begin + blockIdx.x * SR_MICRO_THREADS + threadIdx.x;
#if SR_MICRO_PF_DIST > 0
    // The kernel is bound by a chain of three dependent round trips per triangle (indices -> positions -> depth keys).  The first
    // one is the only access that misses L2 by construction (the index buffer is streamed once): every CTA asks L2 for the index
    // lines of the CTA SR_MICRO_PF_DIST launches behind it (128 triangles x 12 B = twelve 128-byte lines, one per thread).
    if (threadIdx.x < SR_MICRO_THREADS * 12 / 128) {
        const uint64_t tp = (uint64_t)p.tri_begin + (uint64_t)(blockIdx.x + SR_MICRO_PF_DIST) * SR_MICRO_THREADS;
        if (tp + SR_MICRO_THREADS <= p.tri_end && tp + SR_MICRO_THREADS <= p.src.n0)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.src.indices + tp * 3 + threadIdx.x * 32));
    }
#endif
    float4 A = make_float4(0, 0, 0, 0), B = A, C = A;
    if (t < p.tri_end) {
        const SrVertexSet *vs;
        uint32_t vi[3];
        sr_prim_vertices<3>(p.src, t, vs, vi);
        A = __ldg(vs->pos + vi[0]); B = __ldg(vs->pos + vi[1]); C = __ldg(vs->pos + vi[2]);
    }
    sr_micro_triangle<PRECHECK, EARLYZ>(p, t, t < p.tri_end, A, B, C, threadIdx.x & 31);
}

// The same over a LIST of 1024-triangle chunks (chunk-culled front end of a sharded frame): eight CTAs per listed
// chunk; *count chunks are listed, the launch covers the worst case and the CTAs beyond the count return at once.
#define SR_CHUNK_TRIS 1024
template <bool PRECHECK, bool EARLYZ>
__global__ void __launch_bounds__(SR_MICRO_THREADS, SR_MICRO_MIN_BLOCKS) k_micro_chunks(const __grid_constant__ SrMicroParams p, const uint32_t *list,
                                                                                       const uint32_t *count) {
    constexpr uint32_t PER = SR_CHUNK_TRIS / SR_MICRO_THREADS;
    const uint32_t ci = blockIdx.x / PER;
    if (ci >= *count) return;
    const uint32_t t = list[ci] * SR_CHUNK_TRIS + (blockIdx.x % PER) * SR_MICRO_THREADS + threadIdx.x;
    float4 A = make_float4(0, 0, 0, 0), B = A, C = A;
    if (t < p.tri_end) {
        const SrVertexSet *vs;
        uint32_t vi[3];
        sr_prim_vertices<3>(p.src, t, vs, vi);
        A = __ldg(vs->pos + vi[0]); B = __ldg(vs->pos + vi[1]); C = __ldg(vs->pos + vi[2]);
    }
    sr_micro_triangle<PRECHECK, EARLYZ>(p, t, t < p.tri_end, A, B, C, threadIdx.x & 31);
}

// Small draws (a model of a few thousand triangles): the whole front end -- triangle fetch, cull, tile rectangles,
// per-tile counts, the exclusive scan of the counts and the list fill -- in ONE single-CTA launch with the counters
// in shared memory, instead of two memsets + k_micro + k_tile_offsets + k_large_fill.  At this size the frame is
// bound by the latency of that chain of tiny dependent launches, not by throughput.  Every triangle goes to the
// tile lists (the tile kernel's short-list sweep); the visibility buffer is not used.
#define SR_BIN_SMALL_THREADS 1024
#define SR_BIN_SMALL_MAX_TRIS 8192
#define SR_BIN_SMALL_MAX_TRIS_DEV (34 * 1024)  // host-side bound of a draw whose true count (<= 1024 x the clipper's 34) lives on the device
#define SR_BIN_SMALL_MAX_TILES 8192
#define SR_BIN_SMALL_HUGE 256      // tiles
#define SR_BIN_SMALL_MAX_HUGE 64
__global__ void __launch_bounds__(SR_BIN_SMALL_THREADS) k_bin_small(const __grid_constant__ SrMicroParams p, uint32_t *tile_off,
                                                                    uint32_t *list, uint32_t capacity) {
    extern __shared__ uint32_t s_cnt[];  // per-tile counters, later the fill cursors
    __shared__ uint32_t s_wsum[32];
    // triangles covering more than SR_BIN_SMALL_HUGE tiles (a full-screen quad of a post-processing pass: every tile of the
    // frame, twice) are walked by the whole CTA instead of one warp
    __shared__ uint32_t s_huge_rect[SR_BIN_SMALL_MAX_HUGE], s_huge_tri[SR_BIN_SMALL_MAX_HUGE], s_nhuge;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t ntiles = p.ntx * p.nty;
    for (uint32_t i = tid; i < ntiles; i += SR_BIN_SMALL_THREADS) s_cnt[i] = 0;
    if (tid == 0) s_nhuge = 0;
    __syncthreads();
    // one round of SR_BIN_SMALL_THREADS triangles at a time, NOT unrolled: this kernel runs once, on one CTA, with a cold
    // instruction cache -- unrolled eight times it was 11.6 k SASS instructions and most of its time was instruction fetch
    // (ncu: 24 of 30 stall cycles per issue `no_instruction`).  The rectangles wait for the fill phase in shared memory.
    uint32_t *s_rect = s_cnt + ntiles;  // [rounds * SR_BIN_SMALL_THREADS], each thread reads back only its own entries
    const uint32_t ntris = min(p.ntris, sr_prim_count(p.src));  // (the true count of a clip stage that did not synchronise)
    const uint32_t rounds = (ntris + SR_BIN_SMALL_THREADS - 1) / SR_BIN_SMALL_THREADS;
    // A lane walks its own triangle's tile rectangle when it is small (the usual case: a handful of tiles); a rectangle of
    // more than 16 tiles is spread over the lanes of the warp (a big triangle touches hundreds of tiles).
    const bool sharded = p.shard_world > 1;
    auto spread = [&](uint32_t mine, uint32_t t, bool fill) {
        bool big = false;
        if (mine != SR_RECT_INVALID) {
            const uint32_t tx0 = mine & 255u, ty0 = (mine >> 8) & 255u, tx1 = (mine >> 16) & 255u, ty1 = mine >> 24;
            big = (tx1 - tx0 + 1) * (ty1 - ty0 + 1) > 16u;
            if (!big) {
                for (uint32_t ty = ty0; ty <= ty1; ++ty)
                    for (uint32_t tx = tx0; tx <= tx1; ++tx) {
                        const uint32_t tile = ty * p.ntx + tx;
                        if (sharded && tile % p.shard_world != p.shard_rank) continue;
                        const uint32_t at = atomicAdd(&s_cnt[tile], 1u);
                        if (fill) list[at] = t;
                    }
            }
        }
        for (uint32_t rest = __ballot_sync(0xffffffffu, big); rest; rest &= rest - 1) {
            const int src = __ffs(rest) - 1;
            const uint32_t r = __shfl_sync(0xffffffffu, mine, src), st = __shfl_sync(0xffffffffu, t, src);
            const uint32_t tx0 = r & 255u, ty0 = (r >> 8) & 255u, tx1 = (r >> 16) & 255u, ty1 = r >> 24;
            const uint32_t rw = tx1 - tx0 + 1, n = rw * (ty1 - ty0 + 1);
            for (uint32_t i = lane; i < n; i += 32) {
                const uint32_t tile = (ty0 + i / rw) * p.ntx + tx0 + i % rw;
                if (sharded && tile % p.shard_world != p.shard_rank) continue;
                const uint32_t at = atomicAdd(&s_cnt[tile], 1u);
                if (fill) list[at] = st;
            }
        }
    };
#pragma unroll 1
    for (uint32_t k = 0; k < rounds; ++k) {
        const uint32_t t = k * SR_BIN_SMALL_THREADS + tid;
        uint32_t rect = SR_RECT_INVALID;
        if (t < ntris) {
            const SrVertexSet *vs;
            uint32_t vi[3];
            sr_prim_vertices<3>(p.src, t, vs, vi);
            const float4 A = __ldg(vs->pos + vi[0]), B = __ldg(vs->pos + vi[1]), C = __ldg(vs->pos + vi[2]);
            bool skip = isnan(A.x) || isnan(A.y) || isnan(B.x) || isnan(B.y) || isnan(C.x) || isnan(C.y);  // (the reference panics)
            if (p.cull != SR_CULL_NONE) {  // triangle.rs:54-61
                const float area = A.x * B.y + B.x * C.y + C.x * A.y - B.x * A.y - C.x * B.y - A.x * C.y;
                skip = skip || (signbit(area) ? SR_CLOCKWISE : SR_COUNTER_CLOCKWISE) == p.cull;
            }
            if (!skip) {  // triangle.rs:74-78 with tile = the whole frame
                const uint32_t minx = sr_clamp_as_int(fminf(fminf(A.x, B.x), C.x), 0, p.width - 1);
                const uint32_t maxx = sr_clamp_as_int(fmaxf(fmaxf(A.x, B.x), C.x), 0, p.width - 1);
                const uint32_t miny = sr_clamp_as_int(fminf(fminf(A.y, B.y), C.y), 0, p.height - 1);
                const uint32_t maxy = sr_clamp_as_int(fmaxf(fmaxf(A.y, B.y), C.y), 0, p.height - 1);
                if (minx <= maxx && miny <= maxy) rect = sr_pack_rect(minx / SR_TILE_W, miny / SR_TILE_H, maxx / SR_TILE_W, maxy / SR_TILE_H);
            }
        }
        if (rect != SR_RECT_INVALID && ((rect >> 16 & 255u) - (rect & 255u) + 1) * ((rect >> 24) - (rect >> 8 & 255u) + 1) > SR_BIN_SMALL_HUGE) {
            const uint32_t slot = atomicAdd(&s_nhuge, 1u);
            if (slot < SR_BIN_SMALL_MAX_HUGE) {  // (a full table leaves the triangle to its warp)
                s_huge_rect[slot] = rect;
                s_huge_tri[slot] = t;
                rect = SR_RECT_INVALID;
            }
        }
        s_rect[t] = rect;
        spread(rect, t, false);
    }
    __syncthreads();
    const uint32_t nhuge = min(s_nhuge, (uint32_t)SR_BIN_SMALL_MAX_HUGE);
    auto spread_huge = [&](bool fill) {
        for (uint32_t h = 0; h < nhuge; ++h) {
            const uint32_t r = s_huge_rect[h], st = s_huge_tri[h];
            const uint32_t tx0 = r & 255u, ty0 = (r >> 8) & 255u, tx1 = (r >> 16) & 255u, ty1 = r >> 24;
            const uint32_t rw = tx1 - tx0 + 1, n = rw * (ty1 - ty0 + 1);
            for (uint32_t i = tid; i < n; i += SR_BIN_SMALL_THREADS) {
                const uint32_t tile = (ty0 + i / rw) * p.ntx + tx0 + i % rw;
                if (sharded && tile % p.shard_world != p.shard_rank) continue;
                const uint32_t at = atomicAdd(&s_cnt[tile], 1u);
                if (fill) list[at] = st;
            }
        }
    };
    spread_huge(false);
    __syncthreads();
    // exclusive scan of the counters: a run of consecutive tiles per thread, warp scan, scan of the warp totals
    const uint32_t chunk = (ntiles + SR_BIN_SMALL_THREADS - 1) / SR_BIN_SMALL_THREADS;
    const uint32_t b = min(tid * chunk, ntiles), e = min(b + chunk, ntiles);
    uint32_t sum = 0;
    for (uint32_t i = b; i < e; ++i) sum += s_cnt[i];
    uint32_t inc = sum;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += v;
    }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_wsum[lane];
#pragma unroll
        for (uint32_t d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += v;
        }
        s_wsum[lane] = w;  // inclusive totals of the warps
    }
    __syncthreads();
    const uint32_t total = s_wsum[31];
    uint32_t run = (warp ? s_wsum[warp - 1] : 0u) + inc - sum;
    for (uint32_t i = b; i < e; ++i) {
        const uint32_t c = s_cnt[i];
        tile_off[i] = run;
        s_cnt[i] = run;
        run += c;
    }
    if (tid == 0) tile_off[ntiles] = total;
    __syncthreads();
    if (total > capacity) return;  // the lists do not fit: the host re-runs this launch with a larger arena
#pragma unroll 1
    for (uint32_t k = 0; k < rounds; ++k) spread(s_rect[k * SR_BIN_SMALL_THREADS + tid], k * SR_BIN_SMALL_THREADS + tid, true);
    spread_huge(true);
}

struct SrLineSetup {
    float4 ps, pe;   // unclipped end-point positions: interpolation runs over these (line.rs:75-77)
    float cl[4];     // end points clipped to the frame (line.rs:51)
    float d;         // clipped length (line.rs:52)
    uint32_t vi0, vi1, second;
    bool valid;
};
__device__ __forceinline__ bool sr_liang_barsky(float x1, float y1, float x2, float y2, float xmin, float ymin, float xmax, float ymax, float *o);
// fetch + Liang-Barsky clip against the one frame-sized tile ((0,0),(w-1,h-1)) (fragment.rs:255-258, line.rs:51-52)
__device__ __forceinline__ void sr_line_setup(const SrPrimSource &lines, uint32_t width, uint32_t height, uint32_t t, SrLineSetup &r) {
    const SrVertexSet *vs;
    uint32_t vi[2];
    sr_prim_vertices<2>(lines, t, vs, vi);
    r.ps = __ldg(vs->pos + vi[0]);
    r.pe = __ldg(vs->pos + vi[1]);
    r.vi0 = vi[0]; r.vi1 = vi[1];
    r.second = t < lines.n0 ? 0u : 1u;
    r.valid = false;
    if (!sr_liang_barsky(r.ps.x, r.ps.y, r.pe.x, r.pe.y, 0.0f, 0.0f, (float)(width - 1), (float)(height - 1), r.cl)) return;
    if (!isfinite(r.cl[0]) || !isfinite(r.cl[1]) || !isfinite(r.cl[2]) || !isfinite(r.cl[3])) return;  // reference would panic
    r.d = sr_hypot32(r.cl[0] - r.cl[2], r.cl[1] - r.cl[3]);
    r.valid = true;
}

// Non-antialiased lines and points of an opaque draw (Blend = (), stencil Always/Keep, non-discarding shader).  Like the
// triangles of that state, the in-order result at a pixel is the fragment maximising (z, submission index) -- a Bresenham
// line plots every pixel once with alpha 1 (line.rs:125-151), a point plots one pixel -- so they are reduced into the
// same visibility buffer, one thread per primitive, with key ids that sort after every triangle (fragment.rs:268-311:
// triangles, then lines, then points).  The resolve of k_tile_opaque<FS, true> shades the winners.
struct SrExtraParams {
    SrPrimSource lines, points;
    uint32_t nlines, npoints, ntris;
    uint32_t width, height, ntx;
    uint32_t shard_rank, shard_world;
    unsigned long long *vis;
};
__global__ void __launch_bounds__(128) k_lines_vis(const __grid_constant__ SrExtraParams p) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.nlines) return;
    SrLineSetup L;
    sr_line_setup(p.lines, p.width, p.height, t, L);
    if (!L.valid) return;
    const unsigned long long id = (unsigned long long)(p.ntris + t + 1u);
    // draw_line_bresenham (line.rs:125-151); the clipped end points lie inside the frame, 32 bits hold the reference's i64 walk
    int bx0 = (int)L.cl[0], by0 = (int)L.cl[1];
    const int bx1 = (int)L.cl[2], by1 = (int)L.cl[3];
    const int dx = abs(bx1 - bx0), dy = -abs(by1 - by0);
    const int sx = bx0 < bx1 ? 1 : -1, sy = by0 < by1 ? 1 : -1;
    int err = dx + dy;
    while (true) {
        if (bx0 >= 0 && by0 >= 0) {  // line.rs:56
            const float tt = sr_hypot32(L.cl[0] - ((float)bx0 + 0.5f), L.cl[1] - ((float)by0 + 0.5f)) / L.d;  // line.rs:75
            const float z = sr_lerp(tt, L.ps.z, L.pe.z);
            const uint32_t px = (uint32_t)bx0, py = (uint32_t)by0;
            if (z < 0.0f && (p.shard_world == 1 || ((py / SR_TILE_H) * p.ntx + px / SR_TILE_W) % p.shard_world == p.shard_rank))
                sr_red_max_u64(p.vis + sr_vis_index(px, py, p.ntx), ((unsigned long long)(~__float_as_uint(z)) << 32) | id);
        }
        if (bx0 == bx1 && by0 == by1) break;
        const int e2 = 2 * err;
        if (e2 >= dy) { err += dy; bx0 += sx; }
        if (e2 <= dx) { err += dx; by0 += sy; }
    }
}
__global__ void __launch_bounds__(128) k_points_vis(const __grid_constant__ SrExtraParams p) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.npoints) return;
    const SrVertexSet *vs;
    uint32_t vi[1];
    sr_prim_vertices<1>(p.points, t, vs, vi);
    const float4 P = __ldg(vs->pos + vi[0]);
    // point.rs:46: bounds.0 <= x < bounds.1 with bounds = (0,0)..(w-1,h-1) (NaN fails every comparison)
    if (!(0.0f <= P.x && P.x < (float)(p.width - 1) && 0.0f <= P.y && P.y < (float)(p.height - 1))) return;
    const uint32_t px = __float2uint_rz(P.x), py = __float2uint_rz(P.y);
    if (!(P.z < 0.0f)) return;
    if (p.shard_world > 1 && ((py / SR_TILE_H) * p.ntx + px / SR_TILE_W) % p.shard_world != p.shard_rank) return;
    sr_red_max_u64(p.vis + sr_vis_index(px, py, p.ntx),
                   ((unsigned long long)(~__float_as_uint(P.z)) << 32) | (unsigned long long)(p.ntris + p.nlines + t + 1u));
}

// second pass over the large triangles only: write their ids into the per-tile lists
__global__ void __launch_bounds__(256) k_large_fill(const uint32_t *large_count, const uint32_t *large_ids, const uint32_t *large_rects,
                                                    uint32_t ntx, uint32_t ntiles, uint32_t shard_rank, uint32_t shard_world,
                                                    const uint32_t *tile_off, uint32_t *tile_cursor, uint32_t *list, uint32_t capacity) {
    if (tile_off[ntiles] > capacity) return;  // the lists do not fit: the host re-runs this pass with a larger arena
    const uint32_t n = *large_count;
    // one warp per large triangle, its tile rectangle spread over the lanes
    const uint32_t lane = threadIdx.x & 31, nwarps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < n; e += nwarps) {
        const uint32_t t = large_ids[e], rect = large_rects[e];
        const uint32_t tx0 = rect & 255u, ty0 = (rect >> 8) & 255u, tx1 = (rect >> 16) & 255u, ty1 = rect >> 24;
        const uint32_t rw = tx1 - tx0 + 1, cnt = rw * (ty1 - ty0 + 1);
        for (uint32_t i = lane; i < cnt; i += 32) {
            const uint32_t tile = (ty0 + i / rw) * ntx + tx0 + i % rw;
            if (tile % shard_world != shard_rank) continue;
            list[tile_off[tile] + atomicAdd(tile_cursor + tile, 1u)] = t;
        }
    }
}

#ifndef SR_OPQ_THREADS
#define SR_OPQ_THREADS 256
#endif
#ifndef SR_OPQ_SKIP_EMPTY
#define SR_OPQ_SKIP_EMPTY 1
#endif
#ifndef SR_OPQ_MIN_CTAS
#define SR_OPQ_MIN_CTAS 4
#endif
#define SR_OPQ_WARPS (SR_OPQ_THREADS / 32)
#define SR_OPQ_STAGE_FLOATS (32 * 5)  // one warp's 32 finished pixels, AoS {r,g,b,a,depth}
// the short-list sweep's triangle records and the resolve's staging buffers share one region
#define SR_OPQ_SWEEP_BYTES (SR_OPQ_WARPS * 32 * 64)
#define SR_OPQ_STAGE_BYTES (SR_OPQ_WARPS * 2 * SR_OPQ_STAGE_FLOATS * 4)
#define SR_OPQ_REGION_BYTES (SR_OPQ_SWEEP_BYTES > SR_OPQ_STAGE_BYTES ? SR_OPQ_SWEEP_BYTES : SR_OPQ_STAGE_BYTES)
#define SR_OPQ_SMEM_BYTES (SR_TILE_PIXELS * 8 + SR_OPQ_REGION_BYTES + SR_TILE_W * 8 + 16)
// PHASE 2 (merge + resolve of a range-sharded frame): a second 16 KB staging buffer for a peer's keys + two mbarriers
#define SR_OPQ_MERGE_OFFSET ((SR_OPQ_SMEM_BYTES + 127) / 128 * 128)
#define SR_OPQ_MERGE_SMEM_BYTES (SR_OPQ_MERGE_OFFSET + SR_TILE_PIXELS * 8 + 16)
static_assert(SR_OPQ_REGION_BYTES >= SR_TILE_PIXELS * 8, "the sweep/staging region doubles as the first peer-key buffer");
static_assert(SR_TILE_W % 32 == 0, "a warp resolves 32 consecutive pixels of one tile row");

struct SrOpaqueParams {
    SrPrimSource tris;
    uint32_t ntris;
    unsigned long long *vis;        // null: keys start from the framebuffer depth / the pending clear
    uint32_t reset_vis;             // write far keys back once the tile's keys are on chip (the next frame skips k_vis_init)
    const uint32_t *tile_off;       // per-tile CSR offsets into `list` (large triangles)
    const uint32_t *list;
    uint32_t list_capacity, ntiles;  // the pass skips itself when tile_off[ntiles] > list_capacity (see k_large_fill)
    SrFbView fb;
    uint32_t shard_rank, shard_world;
    SrFsConst fs;
    // EXTRA instantiation only: non-antialiased lines and points of the same draw were reduced into the visibility buffer
    // by k_lines_vis / k_points_vis with key ids ntris + 1 + i (lines), ntris + nlines + 1 + i (points)
    SrPrimSource lines, points;
    uint32_t nlines, npoints;
    uint32_t line_base, point_base;  // canonical numbers of the first line / point (winner plane)
    // PHASE 2 only: the other ranks' key buffers (peer-mapped over NVLink, same layout as `vis`)
    const unsigned long long *peer_vis[SR_SHARD_MAX_WORLD - 1];
    uint32_t npeers;
    SrTileOwners owners;     // PHASE 2: the launch covers every tile, a CTA whose tile belongs to another rank returns
    uint32_t elide_clear;    // PHASE 2, ranks other than 0: 32-pixel runs nothing was drawn on are not stored (rank 0 pre-filled them)
};

// PHASE 0: the whole pass (keys -> list sweep -> resolve -> write-back).
// PHASE 1: range-sharded front end, list sweep only: the tile's keys (this rank's own buffer) take the rank's large
//          triangles and go back to the buffer; no resolve.  CTAs of tiles with an empty list return at once.
// PHASE 2: range-sharded merge + resolve: the owner of the tile loads its own keys and, double-buffered, every peer's
//          keys of the same tile with TMA bulk copies from peer-mapped memory, max-merges them in shared memory and
//          resolves -- the exchange is fused into the resolve, there is no staging copy in HBM and no separate collective.
//          The merged key is exactly the key one GPU would have reduced: max over all fragments of (depth key, primitive+1).
#ifndef SR_OPQ_RESOLVE_CTAS
#define SR_OPQ_RESOLVE_CTAS 5  // the resolve without the list sweep needs 51 registers without spilling.  (Measured on config 3, one GPU:
                               // sweep and resolve as two launches with 6 resolve CTAs per SM = 0.1268 ms against 0.1264 ms for the
                               // combined kernel at 4 -- occupancy is not what bounds the resolve; the combined kernel stays.)
#endif
template <int FS, bool EXTRA, int PHASE = 0>
__global__ void __launch_bounds__(SR_OPQ_THREADS, PHASE == 2 ? SR_OPQ_RESOLVE_CTAS : SR_OPQ_MIN_CTAS) k_tile_opaque(const __grid_constant__ SrOpaqueParams p) {
    extern __shared__ __align__(128) unsigned char sr_smem[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(sr_smem);
    float *stage_all = reinterpret_cast<float *>(keys + SR_TILE_PIXELS);
    unsigned long long *far_row = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(stage_all) + SR_OPQ_REGION_BYTES);
    uint64_t *bar = reinterpret_cast<uint64_t *>(far_row + SR_TILE_W);

    const uint32_t tile = PHASE == 2 ? blockIdx.x : p.shard_rank + blockIdx.x * p.shard_world;
    if (PHASE == 2 && sr_tile_owner(p.owners, tile) != p.shard_rank) return;
    const uint32_t tx = tile % p.fb.ntx, ty = tile / p.fb.ntx;
    const uint32_t x0 = tx * SR_TILE_W, y0 = ty * SR_TILE_H;
    if (p.tile_off[p.ntiles] > p.list_capacity) return;
    const uint32_t lbeg = p.tile_off[tile];
    uint32_t L = p.tile_off[tile + 1] - lbeg;
    if (PHASE == 2) L = 0;  // the lists were swept into the keys by every rank's PHASE 1
    if (PHASE == 1 && L == 0) return;
    if (L == 0 && !p.fb.pending_clear && p.vis == nullptr) return;  // nothing to draw, contents already in HBM
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t W = p.fb.width, H = p.fb.height;

    if (p.vis != nullptr) {
        // the tile's keys: SR_TILE_H rows of 512 B in the visibility buffer, one TMA bulk copy per row
        if (tid == 0) {
            sr_mbar_init(bar, 1);
            sr_mbar_init_fence();
            sr_mbar_arrive_expect_tx(bar, SR_TILE_PIXELS * 8);
        }
        __syncthreads();
        if (tid < SR_TILE_H) sr_bulk_g2s(keys + tid * SR_TILE_W, p.vis + sr_vis_index(x0, y0 + tid, p.fb.ntx), SR_TILE_W * 8, bar);
        if (p.reset_vis && tid < SR_TILE_W) far_row[tid] = SR_VIS_FAR_KEY;
        sr_mbar_wait(bar, 0);
        if (SR_OPQ_SKIP_EMPTY && PHASE != 1 && L == 0 && (PHASE != 2 || p.npeers == 0)) {
            // Nothing was drawn on this tile (config 3 covers 44 % of the frame) and its keys are "far" already: no keys to hand
            // back, no per-pixel pass -- the clear goes out with 16-byte stores, or nothing at all when the contents stay.
            uint32_t any = 0;
            for (uint32_t i = tid; i < SR_TILE_PIXELS; i += SR_OPQ_THREADS) any |= keys[i] != SR_VIS_FAR_KEY ? 1u : 0u;
            if (__syncthreads_or((int)any) == 0) {
                if (p.fb.pending_clear && !(PHASE == 2 && p.elide_clear)) sr_fill_tile_clear(p.fb, x0, y0);
                return;
            }
        }
        if (p.reset_vis) {
            // the keys are on chip: hand the visibility buffer back all-far, so the next cleared frame needs no init pass
            sr_fence_proxy_async();
            __syncthreads();
            if (tid < SR_TILE_H) {
                sr_bulk_s2g(p.vis + sr_vis_index(x0, y0 + tid, p.fb.ntx), far_row, SR_TILE_W * 8);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    } else {
        for (uint32_t i = tid; i < SR_TILE_PIXELS; i += SR_OPQ_THREADS) {
            const uint32_t px = x0 + i % SR_TILE_W, py = y0 + i / SR_TILE_W;
            uint32_t dk = ~SR_DEPTH_FAR_BITS;
            if (!p.fb.pending_clear && px < W && py < H) dk = sr_depth_key(sr_fb_load_depth(p.fb, (uint64_t)py * W + px));
            keys[i] = (unsigned long long)dk << 32;
        }
        __syncthreads();
    }

    if (PHASE == 2 && p.npeers > 0) {  // (npeers == 0: the keys are already merged -- k_shard_merge -- or there is one GPU)
        unsigned long long *pbuf[2] = {reinterpret_cast<unsigned long long *>(stage_all),
                                       reinterpret_cast<unsigned long long *>(sr_smem + SR_OPQ_MERGE_OFFSET)};
        uint64_t *mb = reinterpret_cast<uint64_t *>(sr_smem + SR_OPQ_MERGE_OFFSET + SR_TILE_PIXELS * 8);
        if (tid == 0) {
            sr_mbar_init(mb + 0, 1);
            sr_mbar_init(mb + 1, 1);
            sr_mbar_init_fence();
        }
        __syncthreads();
        // warp 0 issues: one 512 B bulk copy per tile row straight from the peer's HBM over NVLink
        auto issue = [&](uint32_t k) {
            if (warp != 0) return;
            if (lane == 0) sr_mbar_arrive_expect_tx(mb + (k & 1u), SR_TILE_PIXELS * 8);
            __syncwarp();
            static_assert(SR_TILE_H <= 32, "one lane per tile row");
            if (lane < SR_TILE_H)
                sr_bulk_g2s(pbuf[k & 1u] + lane * SR_TILE_W, p.peer_vis[k] + sr_vis_index(x0, y0 + lane, p.fb.ntx), SR_TILE_W * 8, mb + (k & 1u));
        };
        for (uint32_t k = 0; k < p.npeers && k < 2; ++k) issue(k);
        for (uint32_t k = 0; k < p.npeers; ++k) {
            sr_mbar_wait(mb + (k & 1u), (k >> 1) & 1u);
            const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(pbuf[k & 1u]);
            ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(keys);
            for (uint32_t i = tid; i < SR_TILE_PIXELS / 2; i += SR_OPQ_THREADS) {
                const ulonglong2 a = src[i];
                ulonglong2 b = dst[i];
                b.x = a.x > b.x ? a.x : b.x;
                b.y = a.y > b.y ? a.y : b.y;
                dst[i] = b;
            }
            sr_fence_proxy_async();
            __syncthreads();
            if (k + 2 < p.npeers) issue(k + 2);
        }
    }

    const uint32_t xe = min(x0 + SR_TILE_W, W) - 1, ye = min(y0 + SR_TILE_H, H) - 1;  // last pixel of the tile in the frame

    // ---------------- the tile's triangle list: 32 triangles per warp step, setup per lane ----------------
    if (PHASE != 2 && L > 0) {
        auto emit = [&](uint32_t px, uint32_t py, unsigned long long key) {
            unsigned long long *slot = keys + (py - y0) * SR_TILE_W + (px - x0);
            if (key > *reinterpret_cast<volatile unsigned long long *>(slot)) atomicMax(slot, key);
        };
        // Short lists (few, typically big triangles -- e.g. a 1k-triangle model filling the frame): every warp walks the
        // whole list but sweeps only its own band of tile rows, so the work is balanced whatever the triangle sizes are
        // and a pixel's key is only ever touched by one warp, in list order per lane: a plain compare-and-store suffices.
        const bool per_warp = L < SR_OPQ_WARPS * 32;
        static_assert(SR_TILE_H % SR_OPQ_WARPS == 0, "tile rows split evenly over the warps");
        constexpr uint32_t RH = SR_TILE_H / SR_OPQ_WARPS;
        const uint32_t band_lo = y0 + warp * RH, band_hi = min(band_lo + RH - 1, ye);
        if (per_warp) {
            // Setup once per triangle (one thread each: the chain of dependent gathers runs once per list, not per warp),
            // published as a 64-byte record; a triangle that misses the tile gets an empty row range.
            float4 *s_rec = reinterpret_cast<float4 *>(stage_all);  // aliases the resolve's staging buffers (used after the sweep)
            if (tid < L) {
                const uint32_t t = __ldg(p.list + lbeg + tid);
                const SrVertexSet *vs;
                uint32_t vi[3];
                sr_prim_vertices<3>(p.tris, t, vs, vi);
                const float4 A = __ldg(vs->pos + vi[0]), B = __ldg(vs->pos + vi[1]), C = __ldg(vs->pos + vi[2]);
                uint32_t minx = max(sr_clamp_as_int(fminf(fminf(A.x, B.x), C.x), 0, W - 1), x0);
                uint32_t miny = max(sr_clamp_as_int(fminf(fminf(A.y, B.y), C.y), 0, H - 1), y0);
                const uint32_t maxx = min(sr_clamp_as_int(fmaxf(fmaxf(A.x, B.x), C.x), 0, W - 1), xe);
                uint32_t maxy = min(sr_clamp_as_int(fmaxf(fmaxf(A.y, B.y), C.y), 0, H - 1), ye);
                if (minx > maxx) { miny = 1; maxy = 0; }
                const SrTri tr = sr_tri_setup(A.x, A.y, B.x, B.y, C.x, C.y);
                s_rec[tid * 4 + 0] = make_float4(tr.a, tr.b, tr.c, tr.d);
                s_rec[tid * 4 + 1] = make_float4(tr.x3, tr.y3, tr.det, tr.rdet);
                s_rec[tid * 4 + 2] = make_float4(A.z, B.z, C.z, __uint_as_float(t + 1u));
                s_rec[tid * 4 + 3] = make_float4(__uint_as_float(minx | (maxx << 16)), __uint_as_float(miny | (maxy << 16)), 0.0f, 0.0f);
            }
            __syncthreads();
            // Fixed pixel ownership inside the band: lane -> row lane / LW, columns congruent to lane % LW.  A pixel's key is
            // therefore only ever read and written by one lane, in list order: no atomics and no warp syncs.
            constexpr uint32_t LW = 32 / RH;
            static_assert(32 % RH == 0, "band rows divide the warp");
            const uint32_t py = band_lo + lane / LW;
            const float yf = (float)py + 0.5f;
            unsigned long long *krow = keys + (py - y0) * SR_TILE_W - x0;
            for (uint32_t l = 0; l < L; ++l) {
                const float4 r3 = s_rec[l * 4 + 3];
                const uint32_t bx = __float_as_uint(r3.x), by = __float_as_uint(r3.y);
                if (max(by & 0xffffu, band_lo) > min(by >> 16, band_hi)) continue;  // (warp-uniform)
                if (py < (by & 0xffffu) || py > (by >> 16)) continue;
                const float4 r0 = s_rec[l * 4 + 0], r1 = s_rec[l * 4 + 1], r2 = s_rec[l * 4 + 2];
                SrTri s;
                s.a = r0.x; s.b = r0.y; s.c = r0.z; s.d = r0.w;
                s.x3 = r1.x; s.y3 = r1.y; s.det = r1.z; s.rdet = r1.w;
                s.dsign = __float_as_uint(s.det) & 0x80000000u;
                s.fast = ((__float_as_uint(s.det) & 0x7FFFFFFFu) - 0x2B800000u) < (0x53800000u - 0x2B800000u);
                const uint32_t minx = bx & 0xffffu, maxx = bx >> 16;
                const float dy = yf - s.y3, bdy = s.b * dy, ddy = s.d * dy;  // triangle.rs:105,108-109 (row terms hoisted, same roundings)
                const unsigned long long id = (unsigned long long)__float_as_uint(r2.w);
                for (uint32_t px = x0 + ((minx - x0) / LW) * LW + lane % LW; px <= maxx; px += LW) {
                    if (px < minx) continue;
                    const float dx = ((float)px + 0.5f) - s.x3;
                    const float nu = s.a * dx + bdy, nv = s.c * dx + ddy;
                    float u, v, w;
                    if (!sr_tri_inside(s, nu, nv, u, v, w)) continue;
                    const float z = (r2.x * u + r2.y * v) + r2.z * w;
                    if (!(z < 0.0f)) continue;  // triangle.rs:120
                    const unsigned long long key = ((unsigned long long)sr_depth_key(z) << 32) | id;
                    if (key > krow[px]) krow[px] = key;
                }
            }
        }
        __syncwarp();
        for (uint32_t gb = warp * 32; !per_warp && gb < L; gb += SR_OPQ_WARPS * 32) {
            const bool have = gb + lane < L;
            uint32_t t = 0;
            SrTri tr;
            float z1 = 0, z2 = 0, z3 = 0;
            uint32_t minx = 1, maxx = 0, miny = 1, maxy = 0;
            if (have) {
                t = __ldg(p.list + lbeg + gb + lane);
                const SrVertexSet *vs;
                uint32_t vi[3];
                sr_prim_vertices<3>(p.tris, t, vs, vi);
                const float4 A = __ldg(vs->pos + vi[0]), B = __ldg(vs->pos + vi[1]), C = __ldg(vs->pos + vi[2]);
                tr = sr_tri_setup(A.x, A.y, B.x, B.y, C.x, C.y);
                z1 = A.z; z2 = B.z; z3 = C.z;
                minx = max(sr_clamp_as_int(fminf(fminf(A.x, B.x), C.x), 0, W - 1), x0);
                miny = max(sr_clamp_as_int(fminf(fminf(A.y, B.y), C.y), 0, H - 1), y0);
                maxx = min(sr_clamp_as_int(fmaxf(fmaxf(A.x, B.x), C.x), 0, W - 1), xe);
                maxy = min(sr_clamp_as_int(fmaxf(fmaxf(A.y, B.y), C.y), 0, H - 1), ye);
            }
            const bool nonempty = have && minx <= maxx && miny <= maxy;
            const uint32_t bw = nonempty ? maxx - minx + 1 : 0, bh = nonempty ? maxy - miny + 1 : 0;
            const bool small = nonempty && bw * bh <= SR_SMALL_AREA;
            if (small) sr_raster_box<false>(tr, z1, z2, z3, minx, miny, bw, bh, t, emit);
            uint32_t big = __ballot_sync(0xffffffffu, nonempty && !small);
            while (big) {  // warp-cooperative sweep of one large box at a time
                const int l = __ffs(big) - 1;
                big &= big - 1;
                SrTri s;
                s.x3 = __shfl_sync(0xffffffffu, tr.x3, l); s.y3 = __shfl_sync(0xffffffffu, tr.y3, l);
                s.a = __shfl_sync(0xffffffffu, tr.a, l); s.b = __shfl_sync(0xffffffffu, tr.b, l);
                s.c = __shfl_sync(0xffffffffu, tr.c, l); s.d = __shfl_sync(0xffffffffu, tr.d, l);
                s.det = __shfl_sync(0xffffffffu, tr.det, l); s.rdet = __shfl_sync(0xffffffffu, tr.rdet, l);
                s.dsign = __float_as_uint(s.det) & 0x80000000u;
                s.fast = ((__float_as_uint(s.det) & 0x7FFFFFFFu) - 0x2B800000u) < (0x53800000u - 0x2B800000u);
                const float sz1 = __shfl_sync(0xffffffffu, z1, l), sz2 = __shfl_sync(0xffffffffu, z2, l), sz3 = __shfl_sync(0xffffffffu, z3, l);
                const uint32_t sminx = __shfl_sync(0xffffffffu, minx, l), sminy = __shfl_sync(0xffffffffu, miny, l);
                const uint32_t sbw = __shfl_sync(0xffffffffu, bw, l), sbh = __shfl_sync(0xffffffffu, bh, l);
                const uint32_t st = __shfl_sync(0xffffffffu, t, l);
                for (uint32_t i = lane; i < sbw * sbh; i += 32) {
                    const uint32_t px = sminx + i % sbw, py = sminy + i / sbw;
                    float nu, nv, u, v, w;
                    sr_tri_numerators(s, (float)px + 0.5f, (float)py + 0.5f, nu, nv);
                    if (!sr_tri_inside(s, nu, nv, u, v, w)) continue;
                    const float z = (sz1 * u + sz2 * v) + sz3 * w;
                    if (!(z < 0.0f)) continue;  // triangle.rs:120
                    emit(px, py, ((unsigned long long)sr_depth_key(z) << 32) | (unsigned long long)(st + 1u));
                }
            }
        }
        __syncthreads();
    }

    if (PHASE == 1) {
        // the swept keys go back to this rank's buffer (the sweep ended with a CTA barrier); a full wait, not .read: the
        // stores must have landed before the stream's next kernel publishes the frame number to the peers
        sr_fence_proxy_async();
        __syncthreads();
        if (tid < SR_TILE_H) {
            sr_bulk_s2g(p.vis + sr_vis_index(x0, y0 + tid, p.fb.ntx), keys + tid * SR_TILE_W, SR_TILE_W * 8);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        return;
    }
    // ---------------- resolve: shade every pixel once, write colour + depth (+winner) to HBM once ----------------
    // A warp finishes 32 consecutive pixels of a tile row at a time.  When the whole frame is being (re)written
    // (pending clear) the 640 B of AoS pixels are staged in shared memory and leave with one TMA bulk store.
    constexpr int NK = SrFsInfo<FS>::NK, NP = (NK + 3) / 4;
    float *stage = stage_all + warp * 2 * SR_OPQ_STAGE_FLOATS;
    const bool u8c = p.fb.u8color != 0;  // RGBAu8Color target: 8-byte pixels, colours quantised when they are produced
    const bool soa = p.fb.soa != 0;      // texture-buffer storage: colour plane + depth plane, two bulk stores per 32-pixel run
    constexpr int NOUT = SrFsOutputs<FS>::N;  // 2: the shader returns a tuple of colours for a two-plane texture buffer (plain stores)
    const bool row_aligned = (W % (u8c ? 2u : 4u)) == 0 && (reinterpret_cast<uintptr_t>(p.fb.aos) & 15u) == 0;
    uint32_t nbulk = 0;
    // The winner's vertex indices are fetched one chunk ahead, so the two dependent gathers (indices, then vertices) of
    // consecutive chunks overlap instead of adding up.
    const uint32_t n0 = p.tris.n0;
    auto fetch = [&](uint32_t chunk, uint32_t &id, uint32_t (&vi)[3]) {
        id = 0;
        if (chunk >= SR_TILE_PIXELS / 32) return;
        const uint32_t i = chunk * 32 + lane;
        if (x0 + i % SR_TILE_W >= W || y0 + i / SR_TILE_W >= H) return;
        id = (uint32_t)keys[i];
        if (id == 0) return;
        const uint32_t t = id - 1;
        if (EXTRA && t >= p.ntris) return;  // a line or a point won: nothing to prefetch
        if (t < n0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) vi[k] = __ldg(p.tris.indices + (uint64_t)t * 3 + k);
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) vi[k] = (t - n0) * 3 + k;
        }
    };
    uint32_t id = 0, vi[3] = {0, 0, 0};
    fetch(warp, id, vi);
    for (uint32_t chunk = warp; chunk < SR_TILE_PIXELS / 32; chunk += SR_OPQ_WARPS) {
        const uint32_t id_cur = id, vi0 = vi[0], vi1 = vi[1], vi2 = vi[2];
        fetch(chunk + SR_OPQ_WARPS, id, vi);
        const uint32_t i = chunk * 32 + lane;
        const uint32_t px = x0 + i % SR_TILE_W, py = y0 + i / SR_TILE_W;
        const uint32_t cx0 = x0 + (chunk * 32) % SR_TILE_W;  // first pixel of the chunk (warp-uniform)
        if (cx0 >= W || py >= H) continue;
        const bool bulk = NOUT == 1 && p.fb.pending_clear && row_aligned && cx0 + 32 <= W;
        const bool in_frame = px < W;
        float o[5 + 4 * (NOUT - 1)];
        bool write = false;
        if (in_frame) {
            if (id_cur == 0) {
                if (p.fb.pending_clear) {
                    o[0] = p.fb.clear[0]; o[1] = p.fb.clear[1]; o[2] = p.fb.clear[2]; o[3] = p.fb.clear[3];
                    if (u8c) sr_quantise_u8(o, false, 0);
                    o[4] = __uint_as_float(SR_DEPTH_FAR_BITS);
                    if constexpr (NOUT == 2) { o[5] = p.fb.clear1[0]; o[6] = p.fb.clear1[1]; o[7] = p.fb.clear1[2]; o[8] = p.fb.clear1[3]; }
                    write = true;
                }
            } else if (EXTRA && id_cur - 1 >= p.ntris) {
                // the fragment of a line (line.rs:70-100) or a point (point.rs:62-82) at this pixel, recomputed exactly as
                // k_lines_vis / k_points_vis computed it
                const uint32_t e = id_cur - 1 - p.ntris;
                float sv[4 + NP * 4 + 1];
                uint32_t canonical;
                if (e < p.nlines) {
                    SrLineSetup L;
                    sr_line_setup(p.lines, W, H, e, L);
                    const float t = sr_hypot32(L.cl[0] - ((float)px + 0.5f), L.cl[1] - ((float)py + 0.5f)) / L.d;
                    sv[0] = sr_lerp(t, L.ps.x, L.pe.x);
                    sv[1] = sr_lerp(t, L.ps.y, L.pe.y);
                    sv[2] = sr_lerp(t, L.ps.z, L.pe.z);
                    sv[3] = sr_lerp(t, L.ps.w, L.pe.w);
                    const SrVertexSet *vs = L.second ? &p.lines.vs1 : &p.lines.vs0;
#pragma unroll
                    for (int pl = 0; pl < NP; ++pl) {
                        const float4 ka = __ldg(vs->attr + sr_attr_at(vs->np, L.vi0, pl));
                        const float4 kb = __ldg(vs->attr + sr_attr_at(vs->np, L.vi1, pl));
                        sv[4 + pl * 4 + 0] = sr_lerp(t, ka.x, kb.x);
                        sv[4 + pl * 4 + 1] = sr_lerp(t, ka.y, kb.y);
                        sv[4 + pl * 4 + 2] = sr_lerp(t, ka.z, kb.z);
                        sv[4 + pl * 4 + 3] = sr_lerp(t, ka.w, kb.w);
                    }
                    canonical = p.line_base + sr_prim_canonical(p.lines, e, 0);
                } else {
                    const uint32_t pt = e - p.nlines;
                    const SrVertexSet *vs;
                    uint32_t vi[1];
                    sr_prim_vertices<1>(p.points, pt, vs, vi);
                    const float4 P = __ldg(vs->pos + vi[0]);
                    sv[0] = P.x; sv[1] = P.y; sv[2] = P.z; sv[3] = P.w;
#pragma unroll
                    for (int pl = 0; pl < NP; ++pl) {
                        const float4 k = __ldg(vs->attr + sr_attr_at(vs->np, vi[0], pl));
                        sv[4 + pl * 4 + 0] = k.x; sv[4 + pl * 4 + 1] = k.y; sv[4 + pl * 4 + 2] = k.z; sv[4 + pl * 4 + 3] = k.w;
                    }
                    canonical = p.point_base + sr_prim_canonical(p.points, pt, 0);
                }
                sr_fragment_shader<FS>(p.fs, sv, o);  // (a line's colour alpha is scaled by its coverage, 1.0 without antialiasing)
                if (u8c) sr_quantise_u8(o, e < p.nlines, 1u);  // (u8 colours: mul_alpha by the coverage 1.0 is NOT the identity, helper.rs:36-42)
                o[4] = sv[2];
                write = true;
                if (p.fb.winner) p.fb.winner[(uint64_t)py * W + px] = canonical + 1;
            } else {
                const uint32_t t = id_cur - 1;
                const SrVertexSet *vs = t < n0 ? &p.tris.vs0 : &p.tris.vs1;
                const float4 A = __ldg(vs->pos + vi0), B = __ldg(vs->pos + vi1), C = __ldg(vs->pos + vi2);
                const SrTri tr = sr_tri_setup(A.x, A.y, B.x, B.y, C.x, C.y);
                float u, v, w;
                sr_tri_bary(tr, px, py, u, v, w);  // same arithmetic as the coverage pass: identical u,v,w
                float sv[4 + NP * 4 + 1];
                sv[0] = sr_bary(u, A.x, v, B.x, w, C.x);
                sv[1] = sr_bary(u, A.y, v, B.y, w, C.y);
                sv[2] = sr_bary(u, A.z, v, B.z, w, C.z);
                sv[3] = sr_bary(u, A.w, v, B.w, w, C.w);
                auto KB = [](float u_, float ux, float v_, float vx, float w_, float wx) {
                    return SrFsInfo<FS>::LIT ? sr_bary_fast(u_, ux, v_, vx, w_, wx) : sr_bary(u_, ux, v_, vx, w_, wx);
                };
                if (NP == 2 && (vs->np & 1u) == 0) {
                    // the record's first two float4 are one aligned 32-byte sector: one 256-bit load per vertex
                    float4 a0, a1, b0, b1, c0, c1;
                    sr_ldg_record2(vs->attr + sr_attr_at(vs->np, vi0, 0), a0, a1);
                    sr_ldg_record2(vs->attr + sr_attr_at(vs->np, vi1, 0), b0, b1);
                    sr_ldg_record2(vs->attr + sr_attr_at(vs->np, vi2, 0), c0, c1);
                    sv[4] = KB(u, a0.x, v, b0.x, w, c0.x); sv[5] = KB(u, a0.y, v, b0.y, w, c0.y);
                    sv[6] = KB(u, a0.z, v, b0.z, w, c0.z); sv[7] = KB(u, a0.w, v, b0.w, w, c0.w);
                    sv[8] = KB(u, a1.x, v, b1.x, w, c1.x); sv[9] = KB(u, a1.y, v, b1.y, w, c1.y);
                    sv[10] = KB(u, a1.z, v, b1.z, w, c1.z); sv[11] = KB(u, a1.w, v, b1.w, w, c1.w);
                } else {
#pragma unroll
                    for (int pl = 0; pl < NP; ++pl) {
                        const float4 ka = __ldg(vs->attr + sr_attr_at(vs->np, vi0, pl));
                        const float4 kb = __ldg(vs->attr + sr_attr_at(vs->np, vi1, pl));
                        const float4 kc = __ldg(vs->attr + sr_attr_at(vs->np, vi2, pl));
                        if (pl >= 2) {
                            // planes past position + normal carry texture coordinates: exact, because a Nearest sampler is a
                            // step function of them (one ulp of contraction can select the neighbouring texel)
                            sv[4 + pl * 4 + 0] = sr_bary(u, ka.x, v, kb.x, w, kc.x);
                            sv[4 + pl * 4 + 1] = sr_bary(u, ka.y, v, kb.y, w, kc.y);
                            sv[4 + pl * 4 + 2] = sr_bary(u, ka.z, v, kb.z, w, kc.z);
                            sv[4 + pl * 4 + 3] = sr_bary(u, ka.w, v, kb.w, w, kc.w);
                            continue;
                        }
                        sv[4 + pl * 4 + 0] = KB(u, ka.x, v, kb.x, w, kc.x);
                        sv[4 + pl * 4 + 1] = KB(u, ka.y, v, kb.y, w, kc.y);
                        sv[4 + pl * 4 + 2] = KB(u, ka.z, v, kb.z, w, kc.z);
                        sv[4 + pl * 4 + 3] = KB(u, ka.w, v, kb.w, w, kc.w);
                    }
                }
                sr_fragment_shader<FS>(p.fs, sv, o);
                if (u8c) sr_quantise_u8(o, false, 0);
                o[4] = sv[2];
                write = true;
                if (p.fb.winner) p.fb.winner[(uint64_t)py * W + px] = sr_prim_canonical(p.tris, t, 0) + 1;
            }
        }
        if (PHASE == 2 && p.elide_clear && __all_sync(0xffffffffu, id_cur == 0 || !in_frame)) continue;  // rank 0 already holds the clear
        if (bulk) {
            float *sb = stage + (nbulk & 1u) * SR_OPQ_STAGE_FLOATS;
            if (nbulk >= 2) {  // the bulk store that last read this buffer must have finished reading it
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
            }
            if (soa) {  // 32 colours (512 B), then 32 depths (128 B): the same 640 bytes of staging
                reinterpret_cast<float4 *>(sb)[lane] = make_float4(o[0], o[1], o[2], o[3]);
                sb[128 + lane] = o[4];
            } else if (u8c) {
                reinterpret_cast<uint2 *>(sb)[lane] = make_uint2(sr_pack_u8(o), __float_as_uint(o[4]));
            } else {
#pragma unroll
                for (int k = 0; k < 5; ++k) sb[lane * 5 + k] = o[k];
            }
            sr_fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                if (soa) {
                    const uint64_t i0 = (uint64_t)py * W + cx0;
                    sr_bulk_s2g(reinterpret_cast<float4 *>(p.fb.aos) + i0, sb, 32 * 16);
                    sr_bulk_s2g(sr_fb_depth_plane(p.fb) + i0, sb + 128, 32 * 4);
                } else {
                    sr_bulk_s2g(sr_fb_pixel_addr(p.fb, cx0, py), sb, u8c ? 32 * 8 : 32 * 20);
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            ++nbulk;
        } else if (write) {
            if constexpr (NOUT == 2) sr_fb_store_pixel2(p.fb, (uint64_t)py * W + px, o);
            else sr_fb_store_pixel(p.fb, (uint64_t)py * W + px, o);
        }
    }
    if ((nbulk && lane == 0) || (p.vis != nullptr && p.reset_vis && tid < SR_TILE_H)) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}


// =====================================================================================================
// Draws of a HANDFUL of triangles (a full-screen pass of two, a few UI rectangles): no binning, no keys, no lists.
// The general opaque path is built for meshes -- bin the triangles, reduce keys per pixel, then look every winner up
// again (indices, vertices, setup) -- and a full-screen quad pays all of it per pixel (433 warp-instructions per 32
// pixels, profiles/r2b_*).  Here every CTA (one tile) sets the <= SR_FEW_MAX triangles up once into shared-memory
// records; every pixel walks the records, keeps the fragment with the largest (depth key, primitive + 1) among those
// with z < 0 and depth >= the stored one (triangle.rs:104-126 -- the same arithmetic, so the same bits, as the general
// path) together with its barycentrics, and is shaded from the winner's record at once.  RenderBuffer targets with
// RGBAf32 colour on one GPU; every other case keeps the general path.
// =====================================================================================================
#define SR_FEW_MAX 8
#define SR_FEW_THREADS 256
struct SrFewParams {
    SrPrimSource tris;
    uint32_t ntris;
    uint32_t cull;
    SrFbView fb;
    SrFsConst fs;
};
template <int FS>
__global__ void __launch_bounds__(SR_FEW_THREADS) k_tile_few(const __grid_constant__ SrFewParams p) {
    constexpr int NK = SrFsInfo<FS>::NK, NP = (NK + 3) / 4;
    constexpr int RS = 8 + 3 * NP;  // float4 per record: setup, box, positions, flags, then the three vertices' attribute planes
    __shared__ float4 s_rec[SR_FEW_MAX * RS];
    __shared__ __align__(16) float s_stage[SR_FEW_THREADS / 32][32 * 5];
    __shared__ uint32_t s_any;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t W = p.fb.width, H = p.fb.height;
    const uint32_t x0 = (blockIdx.x % p.fb.ntx) * SR_TILE_W, y0 = (blockIdx.x / p.fb.ntx) * SR_TILE_H;
    const uint32_t xe = min(x0 + SR_TILE_W, W) - 1, ye = min(y0 + SR_TILE_H, H) - 1;
    const uint32_t L = p.ntris;
    if (tid == 0) s_any = 0;
    __syncthreads();
    if (tid < L) {
        const SrVertexSet *vs;
        uint32_t vi[3];
        sr_prim_vertices<3>(p.tris, tid, vs, vi);
        const float4 A = __ldg(vs->pos + vi[0]), B = __ldg(vs->pos + vi[1]), C = __ldg(vs->pos + vi[2]);
        bool skip = isnan(A.x) || isnan(A.y) || isnan(B.x) || isnan(B.y) || isnan(C.x) || isnan(C.y);  // (the reference panics)
        if (p.cull != SR_CULL_NONE) {  // triangle.rs:54-61
            const float area = A.x * B.y + B.x * C.y + C.x * A.y - B.x * A.y - C.x * B.y - A.x * C.y;
            skip = skip || (signbit(area) ? SR_CLOCKWISE : SR_COUNTER_CLOCKWISE) == p.cull;
        }
        // triangle.rs:74-78 (bounding box clamped to the frame), intersected with this tile
        uint32_t minx = max(sr_clamp_as_int(fminf(fminf(A.x, B.x), C.x), 0, W - 1), x0);
        uint32_t miny = max(sr_clamp_as_int(fminf(fminf(A.y, B.y), C.y), 0, H - 1), y0);
        const uint32_t maxx = min(sr_clamp_as_int(fmaxf(fmaxf(A.x, B.x), C.x), 0, W - 1), xe);
        uint32_t maxy = min(sr_clamp_as_int(fmaxf(fmaxf(A.y, B.y), C.y), 0, H - 1), ye);
        if (skip || minx > maxx || miny > maxy) { minx = 1; miny = 1; maxy = 0; }
        else s_any = 1;
        const SrTri tr = sr_tri_setup(A.x, A.y, B.x, B.y, C.x, C.y);
        float4 *r = s_rec + tid * RS;
        r[0] = make_float4(tr.a, tr.b, tr.c, tr.d);
        r[1] = make_float4(tr.x3, tr.y3, tr.det, tr.rdet);
        r[2] = make_float4(A.z, B.z, C.z, __uint_as_float(tid + 1u));
        r[3] = make_float4(__uint_as_float(minx), __uint_as_float(maxx), __uint_as_float(miny), __uint_as_float(maxy));
#pragma unroll
        for (int pl = 0; pl < NP; ++pl) {  // the attributes travel with the record: the pixels read shared memory, not three gathers each
            r[8 + pl * 3 + 0] = __ldg(vs->attr + sr_attr_at(vs->np, vi[0], pl));
            r[8 + pl * 3 + 1] = __ldg(vs->attr + sr_attr_at(vs->np, vi[1], pl));
            r[8 + pl * 3 + 2] = __ldg(vs->attr + sr_attr_at(vs->np, vi[2], pl));
        }
        r[4] = make_float4(A.x, A.y, A.w, B.x);
        r[5] = make_float4(B.y, B.w, C.x, C.y);
        r[6] = make_float4(C.w, __uint_as_float(vi[0]), __uint_as_float(vi[1]), __uint_as_float(vi[2]));
        r[7] = make_float4(__uint_as_float(tr.dsign), __uint_as_float(tr.fast ? 1u : 0u), __uint_as_float(tid < p.tris.n0 ? 0u : 1u), 0.0f);
    }
    __syncthreads();
    if (s_any == 0) {  // no triangle reaches the tile
        if (p.fb.pending_clear) sr_fill_tile_clear(p.fb, x0, y0);
        return;
    }
    const bool vec_ok = (W & 3u) == 0 && (reinterpret_cast<uintptr_t>(p.fb.aos) & 15u) == 0;
    float *sb = s_stage[warp];
    for (uint32_t chunk = warp; chunk < SR_TILE_PIXELS / 32; chunk += SR_FEW_THREADS / 32) {
        const uint32_t cx0 = x0 + (chunk * 32) % SR_TILE_W, py = y0 + (chunk * 32) / SR_TILE_W, px = cx0 + lane;
        if (cx0 >= W || py >= H) continue;  // (warp-uniform)
        const bool in_frame = px < W;
        const uint64_t index = (uint64_t)py * W + px;
        float o[5];
        bool write = false;
        if (in_frame) {
            uint32_t best_hi = ~SR_DEPTH_FAR_BITS, best_l = 0xFFFFFFFFu;
            if (!p.fb.pending_clear) best_hi = sr_depth_key(p.fb.aos[index * 5 + 4]);
            float u = 0.0f, v = 0.0f, w = 0.0f;
            const float xf = (float)px + 0.5f, yf = (float)py + 0.5f;
            for (uint32_t l = 0; l < L; ++l) {
                const float4 *r = s_rec + l * RS;
                const uint4 box = *reinterpret_cast<const uint4 *>(r + 3);
                if (px < box.x || px > box.y || py < box.z || py > box.w) continue;
                const float4 r0 = r[0], r1 = r[1], r2 = r[2], r7 = r[7];
                SrTri s;
                s.a = r0.x; s.b = r0.y; s.c = r0.z; s.d = r0.w;
                s.x3 = r1.x; s.y3 = r1.y; s.det = r1.z; s.rdet = r1.w;
                s.dsign = __float_as_uint(r7.x);
                s.fast = __float_as_uint(r7.y) != 0u;
                float nu, nv, uu, vv, ww;
                sr_tri_numerators(s, xf, yf, nu, nv);
                if (!sr_tri_inside(s, nu, nv, uu, vv, ww)) continue;
                const float z = (r2.x * uu + r2.y * vv) + r2.z * ww;
                if (!(z < 0.0f)) continue;  // triangle.rs:120
                const uint32_t kh = sr_depth_key(z);
                if (kh >= best_hi) { best_hi = kh; best_l = l; u = uu; v = vv; w = ww; }  // d >= dt, later primitive wins ties (triangle.rs:126)
            }
            if (best_l != 0xFFFFFFFFu) {
                const float4 *r = s_rec + best_l * RS;
                const float4 r2 = r[2], r4 = r[4], r5 = r[5], r6 = r[6];
                const float4 A = make_float4(r4.x, r4.y, r2.x, r4.z), B = make_float4(r4.w, r5.x, r2.y, r5.y), C = make_float4(r5.z, r5.w, r2.z, r6.x);
                float sv[4 + NP * 4 + 1];
                sv[0] = sr_bary(u, A.x, v, B.x, w, C.x);
                sv[1] = sr_bary(u, A.y, v, B.y, w, C.y);
                sv[2] = sr_bary(u, A.z, v, B.z, w, C.z);
                sv[3] = sr_bary(u, A.w, v, B.w, w, C.w);
#pragma unroll
                for (int pl = 0; pl < NP; ++pl) {
                    const float4 ka = r[8 + pl * 3 + 0], kb = r[8 + pl * 3 + 1], kc = r[8 + pl * 3 + 2];
                    // (same choice as the general resolve: contraction only for values that feed lit shading, never for texture coordinates)
                    if (SrFsInfo<FS>::LIT && pl < 2) {
                        sv[4 + pl * 4 + 0] = sr_bary_fast(u, ka.x, v, kb.x, w, kc.x); sv[4 + pl * 4 + 1] = sr_bary_fast(u, ka.y, v, kb.y, w, kc.y);
                        sv[4 + pl * 4 + 2] = sr_bary_fast(u, ka.z, v, kb.z, w, kc.z); sv[4 + pl * 4 + 3] = sr_bary_fast(u, ka.w, v, kb.w, w, kc.w);
                    } else {
                        sv[4 + pl * 4 + 0] = sr_bary(u, ka.x, v, kb.x, w, kc.x); sv[4 + pl * 4 + 1] = sr_bary(u, ka.y, v, kb.y, w, kc.y);
                        sv[4 + pl * 4 + 2] = sr_bary(u, ka.z, v, kb.z, w, kc.z); sv[4 + pl * 4 + 3] = sr_bary(u, ka.w, v, kb.w, w, kc.w);
                    }
                }
                sr_fragment_shader<FS>(p.fs, sv, o);
                o[4] = sv[2];
                write = true;
            } else if (p.fb.pending_clear) {
                o[0] = p.fb.clear[0]; o[1] = p.fb.clear[1]; o[2] = p.fb.clear[2]; o[3] = p.fb.clear[3];
                o[4] = __uint_as_float(SR_DEPTH_FAR_BITS);
                write = true;
            }
        }
        // a whole 32-pixel run that is written leaves as forty 16-byte stores (640 contiguous bytes), anything else pixel by pixel
        if (vec_ok && cx0 + 32 <= W && __all_sync(0xffffffffu, write)) {
#pragma unroll
            for (int k = 0; k < 5; ++k) sb[lane * 5 + k] = o[k];
            __syncwarp();
            float4 *dst = reinterpret_cast<float4 *>(p.fb.aos + ((uint64_t)py * W + cx0) * 5);
            dst[lane] = reinterpret_cast<const float4 *>(sb)[lane];
            if (lane < 8) dst[32 + lane] = reinterpret_cast<const float4 *>(sb)[32 + lane];
            __syncwarp();
        } else if (write) {
#pragma unroll
            for (int k = 0; k < 5; ++k) p.fb.aos[index * 5 + k] = o[k];
        }
    }
}

// =====================================================================================================
// Ordered tile rasteriser: any blend, stencil, discarding shaders, lines and points.  Strictly in
// submission order per pixel.  Tile colour, depth, stencil (and winner) live in shared memory; each
// warp owns a fixed set of tile rows, so all operations on a pixel are issued by one warp in program order
// (triangles: row r belongs to warp r % 8; lines and points, which run after a CTA barrier: bands of 4 consecutive rows).
// =====================================================================================================
// Everything the per-pixel loop needs of one triangle, produced once (one thread per triangle, 256 at a time) so that the
// strictly ordered sweep -- a serial chain by construction -- contains no dependent global loads and no reciprocal.
struct SrOrdSetup {
    float4 e;         // a, b, c, d           (triangle.rs:108-109 edge terms)
    float4 f;         // x3, y3, det, 1/det
    float4 A, B, C;   // screen positions {x, y, z, 1/w}
    uint32_t bx, by;  // minx | maxx<<16, miny | maxy<<16 (frame-clamped bbox intersected with the tile)
    uint32_t canonical, second;  // canonical primitive number; vertices come from the generated stream
    uint32_t vi[3], pad;
};
static_assert(sizeof(SrOrdSetup) == 112, "seven float4");
#define SR_ORD_LIST_CAP 2048  // group ids sorted in shared memory; longer lists are sorted in place in HBM
#ifndef SR_ORD_LANE_BOX
#define SR_ORD_LANE_BOX 20  // a band's batch takes the lane-per-triangle sweep when 7 of 8 of its (untightened, band-clipped) boxes are at most this
#endif
#ifndef SR_ORD_IL_ROWS
#define SR_ORD_IL_ROWS 5  // a batch whose (tile-clipped) boxes average at most this many rows interleaves the row ownership (k_tile_ordered)
#endif
#ifndef SR_ORD_LANE_PIX
#define SR_ORD_LANE_PIX 32  // largest (tightened, band-clipped) box a single lane sweeps in k_tile_ordered's small-triangle runs
#endif
static_assert(SR_ORD_LANE_PIX <= 32, "box positions are kept in a 32-bit mask");
#define SR_ORD_RING 64  // fragments a warp can hold between coverage and shading (at most 31 carried over + 32 new)
#define SR_ORD_SMEM_BYTES (SR_TILE_PIXELS * (16 + 4 + 4 + 1) + SR_RASTER_THREADS * sizeof(SrOrdSetup) + SR_ORD_LIST_CAP * 4 + \
                           SR_RASTER_WARPS * SR_ORD_RING * 16)

struct SrOrdCtx {
    float4 *color;
    float *depth;
    uint32_t *winner;
    uint8_t *stencil;
    const SrTileParams *p;
    uint32_t x0, y0, xe, ye;  // tile pixel rectangle inside the frame (inclusive)
    bool has_stencil;
    uint32_t sbytes, smax;   // stencil element size and the type's MAX
    uint32_t mesh_stencil;
};

// Lines and points: the warps of the CTA walk the lines / points of the tile in submission order, but only the thread that
// owns a pixel plots it -- warp = band of SR_TILE_H / SR_RASTER_WARPS tile rows, lane = column mod 32.  A pixel therefore
// still sees its fragments in order (one thread, program order) while the fragments of one line are spread over the lanes,
// and a warp skips every line that does not cross its band.  Walking is a few instructions per step; plotting
// (interpolation, shader, blend) is the expensive part.
#define SR_ORD_BAND (SR_TILE_H / SR_RASTER_WARPS)
__device__ __forceinline__ bool sr_ord_owns(const SrOrdCtx &c, uint32_t x, uint32_t y) {  // (x, y) inside the tile
    return (y - c.y0) / SR_ORD_BAND == (threadIdx.x >> 5) && (x & 31u) == (threadIdx.x & 31u);
}

// stencil step (triangle.rs:91-99, line.rs:58-66, point.rs:52-60)
__device__ __forceinline__ bool sr_ord_stencil_step(const SrOrdCtx &c, uint32_t li) {
    if (!c.has_stencil) return true;  // stencil type (): Always / Keep
    const uint32_t sval = sr_stencil_load(c.stencil, c.sbytes, li);
    if (!sr_stencil_test_fn(c.p->stencil_test, sval, c.mesh_stencil)) return false;
    sr_stencil_store(c.stencil, c.sbytes, li, sr_stencil_op_fn(c.p->stencil_op, sval, c.mesh_stencil, c.smax));
    return true;
}
// everything after the stencil step of one fragment (triangle.rs:117-143, line.rs:81-106, point.rs:62-82)
template <int FS>
__device__ __forceinline__ void sr_ord_shade_write(const SrOrdCtx &c, uint32_t li, const float *sv, bool use_alpha, float alpha,
                                                   uint32_t canonical, uint32_t alpha_u8 = 1u) {
    const float z = sv[2];
    if (!(z < 0.0f)) return;
    if (!(z >= c.depth[li])) return;
    float col[4];
    if (!sr_fragment_shader<FS>(c.p->fs, sv, col)) return;  // Fragment::Discard
    if (c.p->fb.u8color) {  // RGBAu8Color target (Blend = () only): the tile holds channel values 0..255
        sr_quantise_u8(col, use_alpha, alpha_u8);
        c.color[li] = make_float4(col[0], col[1], col[2], col[3]);
        c.depth[li] = z;
        c.winner[li] = canonical + 1;
        return;
    }
    if (use_alpha) col[3] = col[3] * alpha;                 // Color::mul_alpha (src/color/predefined.rs:82-86)
    const float4 old = c.color[li];
    const float dstc[4] = {old.x, old.y, old.z, old.w};
    float outc[4];
    sr_blend(c.p->blend, col, dstc, outc);
    c.color[li] = make_float4(outc[0], outc[1], outc[2], outc[3]);
    c.depth[li] = z;
    c.winner[li] = canonical + 1;
}

// direction-free bitonic sort (ascending); elements at index >= n are virtual +inf, so n need not be a power of two
__device__ __forceinline__ void sr_block_sort(uint32_t *a, uint32_t n) {
    uint32_t n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (uint32_t k = 2; k <= n2; k <<= 1) {
        for (uint32_t i = threadIdx.x; i < n2 / 2; i += blockDim.x) {
            const uint32_t h = k >> 1, blk = i / h, wi = i % h;
            const uint32_t lo = blk * k + wi, hi = blk * k + k - 1 - wi;
            if (hi < n) {
                const uint32_t va = a[lo], vb = a[hi];
                if (va > vb) { a[lo] = vb; a[hi] = va; }
            }
        }
        __syncthreads();
        for (uint32_t j = k >> 2; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < n2 / 2; i += blockDim.x) {
                const uint32_t lo = (i / j) * 2 * j + i % j, hi = lo + j;
                if (hi < n) {
                    const uint32_t va = a[lo], vb = a[hi];
                    if (va > vb) { a[lo] = vb; a[hi] = va; }
                }
            }
            __syncthreads();
        }
    }
}

// liang_barsky_iterative (src/geometry/line.rs:6-54)
__device__ __forceinline__ bool sr_liang_barsky(float x1, float y1, float x2, float y2, float xmin, float ymin, float xmax,
                                                float ymax, float *o) {
    float t0 = 0.0f, t1 = 1.0f;
    const float dx = x2 - x1, dy = y2 - y1;
    for (int edge = 0; edge < 4; ++edge) {
        float pp, q;
        switch (edge) {
            case 0: pp = -dx; q = x1 - xmin; break;
            case 1: pp = dx; q = xmax - x1; break;
            case 2: pp = -dy; q = y1 - ymin; break;
            default: pp = dy; q = ymax - y1; break;
        }
        if (pp == 0.0f && q < 0.0f) return false;
        const float r = q / pp;
        if (pp < 0.0f) {
            if (r > t1) return false;
            else if (r > t0) t0 = r;
        } else if (pp > 0.0f) {
            if (r < t0) return false;
            else if (r < t1) t1 = r;
        }
    }
    o[0] = x1 + t0 * dx; o[1] = y1 + t0 * dy; o[2] = x1 + t1 * dx; o[3] = y1 + t1 * dy;
    return true;
}

struct SrLineCtx {
    float x1, y1, d;       // clipped start point and clipped length (line.rs:51-52)
    float4 ps, pe;         // unclipped end-point positions: interpolation runs over these (line.rs:75-77)
    const SrVertexSet *vs;
    uint32_t vi[2];
    uint32_t canonical;
};

// the rasterize_fragment closure of rasterize_line (line.rs:54-109); one non-inlined copy per shader
template <int FS>
__device__ __noinline__ void sr_ord_plot_line(const SrOrdCtx &c, const SrLineCtx &L, long long x, long long y, double alpha) {
    constexpr int NK = SrFsInfo<FS>::NK, NP = (NK + 3) / 4;
    if (x < 0 || y < 0) return;
    // only this tile's pixels (the frame test also covers Wu's +1 neighbour past the last row/column,
    // where the reference would index out of bounds)
    if (x < (long long)c.x0 || x > (long long)c.xe || y < (long long)c.y0 || y > (long long)c.ye) return;
    const uint32_t li = ((uint32_t)y - c.y0) * SR_TILE_W + ((uint32_t)x - c.x0);
    if (!sr_ord_stencil_step(c, li)) return;
    const float xf = (float)x + 0.5f, yf = (float)y + 0.5f;
    const float t = sr_hypot32(L.x1 - xf, L.y1 - yf) / L.d;
    float sv[4 + NP * 4 + 1];
    sv[0] = sr_lerp(t, L.ps.x, L.pe.x);
    sv[1] = sr_lerp(t, L.ps.y, L.pe.y);
    sv[2] = sr_lerp(t, L.ps.z, L.pe.z);
    sv[3] = sr_lerp(t, L.ps.w, L.pe.w);
#pragma unroll
    for (int pl = 0; pl < NP; ++pl) {
        const float4 ka = __ldg(L.vs->attr + sr_attr_at(L.vs->np, L.vi[0], pl));
        const float4 kb = __ldg(L.vs->attr + sr_attr_at(L.vs->np, L.vi[1], pl));
        sv[4 + pl * 4 + 0] = sr_lerp(t, ka.x, kb.x);
        sv[4 + pl * 4 + 1] = sr_lerp(t, ka.y, kb.y);
        sv[4 + pl * 4 + 2] = sr_lerp(t, ka.z, kb.z);
        sv[4 + pl * 4 + 3] = sr_lerp(t, ka.w, kb.w);
    }
    sr_ord_shade_write<FS>(c, li, sv, true, (float)alpha, L.canonical, (uint32_t)alpha /* NumCast f64 -> u8 */);
}

__device__ __forceinline__ double sr_fract64(double x) { return x - trunc(x); }

// rasterize_line (src/pipeline/stages/rasterization/line.rs:22-119).  Setup (fetch, Liang-Barsky clip against the frame,
// clipped length) runs once per line, one thread each, 256 lines at a time; the record is what the walk needs.
struct SrOrdLineRec {
    float4 ps, pe;   // unclipped end-point positions
    float cl[4];     // clipped end points (line.rs:51)
    float d;         // clipped length (line.rs:52)
    uint32_t vi0, vi1, canonical, second, valid;
    uint32_t pad[2];
};
static_assert(sizeof(SrOrdLineRec) <= sizeof(SrOrdSetup), "line records reuse the triangle records' shared memory");
__device__ __forceinline__ void sr_ord_line_setup(const SrTileParams &p, uint32_t t, SrOrdLineRec &r) {
    SrLineSetup L;
    sr_line_setup(p.lines, p.fb.width, p.fb.height, t, L);
    r.ps = L.ps; r.pe = L.pe;
    r.cl[0] = L.cl[0]; r.cl[1] = L.cl[1]; r.cl[2] = L.cl[2]; r.cl[3] = L.cl[3];
    r.d = L.d;
    r.vi0 = L.vi0; r.vi1 = L.vi1; r.second = L.second;
    r.canonical = p.line_base + sr_prim_canonical(p.lines, t, 0);
    r.valid = L.valid ? 1u : 0u;
}
// The walk of one line (line.rs:125-239), restricted to the pixels of one band of tile rows: calls plot(x, y, alpha) for
// every plotted pixel with band_lo <= y <= band_hi and tx0 <= x <= tx1, in the reference's plotting order.
template <class Plot>
__device__ __forceinline__ void sr_line_walk(const SrOrdLineRec &r, bool aa, int band_lo, int band_hi, int tx0, int tx1, Plot plot) {
    const float cl[4] = {r.cl[0], r.cl[1], r.cl[2], r.cl[3]};
    auto inside = [&](int x, int y) { return y >= band_lo && y <= band_hi && x >= tx0 && x <= tx1; };
    if (!aa) {
        // draw_line_bresenham (line.rs:125-151).  The reference walks in i64; the clipped end points lie inside the frame
        // (< 2^16), so every quantity below fits 32 bits with the same decisions.
        int bx0 = (int)cl[0], by0 = (int)cl[1];
        const int bx1 = (int)cl[2], by1 = (int)cl[3];
        const int dx = abs(bx1 - bx0), dy = -abs(by1 - by0);
        const int sx = bx0 < bx1 ? 1 : -1, sy = by0 < by1 ? 1 : -1;
        int err = dx + dy;
        while (true) {
            if (inside(bx0, by0)) plot(bx0, by0, 1.0);
            if (bx0 == bx1 && by0 == by1) break;
            if (sy > 0 ? by0 > band_hi : by0 < band_lo) break;  // y is monotonic: the walk has left the band for good
            const int e2 = 2 * err;
            if (e2 >= dy) { err += dy; bx0 += sx; }
            if (e2 <= dx) { err += dx; by0 += sy; }
        }
    } else {
        // draw_line_xiaolin_wu (line.rs:159-239), f64 throughout
        double wx0 = (double)cl[0], wy0 = (double)cl[1], wx1 = (double)cl[2], wy1 = (double)cl[3];
        const bool steep = fabs(wy1 - wy0) > fabs(wx1 - wx0);
        if (steep) { double tmp = wx0; wx0 = wy0; wy0 = tmp; tmp = wx1; wx1 = wy1; wy1 = tmp; }
        if (wx0 > wx1) { double tmp = wx0; wx0 = wx1; wx1 = tmp; tmp = wy0; wy0 = wy1; wy1 = tmp; }
        const double dx = wx1 - wx0, dy = wy1 - wy0;
        const double gradient = dx < 0.0001 ? 1.0 : dy / dx;
        auto plot_float = [&](double a, double b, double opacity) {
            // (the coordinates are within a pixel of the frame: the i64 casts of the reference fit 32 bits; negative
            // coordinates are not plotted, line.rs:56)
            const int x = steep ? (int)b : (int)a, y = steep ? (int)a : (int)b;
            if (inside(x, y)) plot(x, y, opacity);
        };
        // both arms of the reference's `if steep` plot (x, y) / (y, x); plot_float takes (x_major, y_minor)
        double xend = round(wx0);
        double yend = wy0 + gradient * (xend - wx0);
        double xgap = 1.0 - sr_fract64(wx0 + 0.5);
        const double xpxl1 = xend, ypxl1 = trunc(yend);
        plot_float(xpxl1, ypxl1, (1.0 - sr_fract64(yend)) * xgap);
        plot_float(xpxl1, ypxl1 + 1.0, sr_fract64(yend) * xgap);
        double intery = yend + gradient;
        xend = round(wx1);
        yend = wy1 + gradient * (xend - wx1);
        xgap = sr_fract64(wx1 + 0.5);
        const double xpxl2 = xend, ypxl2 = trunc(yend);
        plot_float(xpxl2, ypxl2, (1.0 - sr_fract64(yend)) * xgap);
        plot_float(xpxl2, ypxl2 + 1.0, sr_fract64(yend) * xgap);
        for (double x = xpxl1 + 1.0; x <= (xpxl2 - 1.0); x += 1.0) {
            const double y = trunc(intery);
            plot_float(x, y, 1.0 - sr_fract64(intery));
            plot_float(x, y + 1.0, sr_fract64(intery));
            intery += gradient;
        }
    }
}

// Up to 32 consecutive line records, one per lane, rasterised by ONE warp into its band of tile rows with every pixel still
// seeing its fragments in submission order.  Rounds: (A) every lane walks its line and bids for the pixels of its
// not-yet-plotted fragments with atomicMax of (round << 8 | 255 - lane) -- the earliest line of the chunk wins a pixel;
// (B) every lane walks again and plots its fragments in order for as long as it holds the pixel, stopping at the first
// pixel an earlier line still has to plot.  A lane's plotted fragments are therefore a prefix of its walk, the earliest
// unfinished line always completes, and a later line can never overtake an earlier one on a pixel.  Lines seldom share
// pixels inside a band, so almost every chunk finishes in one round with all its lines walked in parallel.
template <int FS>
__device__ void sr_ord_lines_chunk(const SrOrdCtx &c, const SrOrdLineRec *recs, uint32_t n, uint32_t *claim, uint32_t &round) {
    const SrTileParams &p = *c.p;
    const uint32_t lane = threadIdx.x & 31u;
    const int tx0 = (int)c.x0, tx1 = (int)c.xe;
    const int band_lo = (int)c.y0 + (int)(threadIdx.x >> 5) * SR_ORD_BAND, band_hi = min(band_lo + SR_ORD_BAND - 1, (int)c.ye);
    bool active = false;
    if (lane < n) {
        // the whole line misses the band: nothing to walk.  Wu plots rows trunc(yend) and trunc(yend) + 1 with yend up to
        // half a pixel past the clipped end point (line.rs:181-199), i.e. up to two rows beyond trunc(y) of the end points
        const int ya = (int)recs[lane].cl[1], yb = (int)recs[lane].cl[3];
        active = !(max(ya, yb) + 2 < band_lo || min(ya, yb) - 2 > band_hi);
    }
    if (!__any_sync(0xffffffffu, active)) return;
    const SrOrdLineRec &r = recs[min(lane, n - 1)];
    SrLineCtx L;
    L.vs = r.second ? &p.lines.vs1 : &p.lines.vs0;
    L.vi[0] = r.vi0; L.vi[1] = r.vi1;
    L.ps = r.ps; L.pe = r.pe;
    L.canonical = r.canonical;
    L.x1 = r.cl[0]; L.y1 = r.cl[1];
    L.d = r.d;
    const bool aa = p.aa_lines != 0;
    uint32_t ndone = 0;  // fragments of this lane's walk (inside the band) already plotted
    while (__any_sync(0xffffffffu, active)) {
        ++round;
        const uint32_t bid = (round << 8) | (255u - lane);
        if (active) {
            uint32_t j = 0;
            sr_line_walk(r, aa, band_lo, band_hi, tx0, tx1, [&](int x, int y, double) {
                if (j++ >= ndone) atomicMax(&claim[(y - band_lo) * SR_TILE_W + (x - tx0)], bid);
            });
        }
        __syncwarp();
        bool more = false;
        if (active) {
            uint32_t j = 0;
            bool blocked = false;
            sr_line_walk(r, aa, band_lo, band_hi, tx0, tx1, [&](int x, int y, double alpha) {
                if (j++ < ndone || blocked) return;
                if (claim[(y - band_lo) * SR_TILE_W + (x - tx0)] != bid) { blocked = true; more = true; return; }
                sr_ord_plot_line<FS>(c, L, x, y, alpha);
                ++ndone;
            });
        }
        __syncwarp();
        active = more;
    }
}

// rasterize_point (src/pipeline/stages/rasterization/point.rs:21-86); membership in this tile was decided by the rect.
// Fetched one thread per point, 256 at a time; every thread then scans the records and plots the pixels it owns.
struct SrOrdPointRec {
    float4 P;
    uint32_t vi, canonical, second, valid;
};
template <int FS>
__device__ __forceinline__ void sr_ord_point(const SrOrdCtx &c, const SrOrdPointRec &r) {
    constexpr int NK = SrFsInfo<FS>::NK, NP = (NK + 3) / 4;
    const SrTileParams &p = *c.p;
    const float4 P = r.P;
    const uint32_t px = __float2uint_rz(P.x), py = __float2uint_rz(P.y);
    if (px < c.x0 || px > c.xe || py < c.y0 || py > c.ye) return;
    if (!sr_ord_owns(c, px, py)) return;
    const uint32_t li = (py - c.y0) * SR_TILE_W + (px - c.x0);
    if (!sr_ord_stencil_step(c, li)) return;
    const SrVertexSet *vs = r.second ? &p.points.vs1 : &p.points.vs0;
    float sv[4 + NP * 4 + 1];
    sv[0] = P.x; sv[1] = P.y; sv[2] = P.z; sv[3] = P.w;
#pragma unroll
    for (int pl = 0; pl < NP; ++pl) {
        const float4 k = __ldg(vs->attr + sr_attr_at(vs->np, r.vi, pl));
        sv[4 + pl * 4 + 0] = k.x; sv[4 + pl * 4 + 1] = k.y; sv[4 + pl * 4 + 2] = k.z; sv[4 + pl * 4 + 3] = k.w;
    }
    sr_ord_shade_write<FS>(c, li, sv, false, 1.0f, r.canonical);
}

template <int FS>
__global__ void __launch_bounds__(SR_RASTER_THREADS) k_tile_ordered(const __grid_constant__ SrTileParams p) {
    extern __shared__ __align__(128) unsigned char sr_smem[];
    float4 *s_color = reinterpret_cast<float4 *>(sr_smem);
    float *s_depth = reinterpret_cast<float *>(s_color + SR_TILE_PIXELS);
    uint32_t *s_winner = reinterpret_cast<uint32_t *>(s_depth + SR_TILE_PIXELS);
    SrOrdSetup *s_setup = reinterpret_cast<SrOrdSetup *>(s_winner + SR_TILE_PIXELS);
    uint32_t *s_list = reinterpret_cast<uint32_t *>(s_setup + SR_RASTER_THREADS);
    uint4 *s_ring = reinterpret_cast<uint4 *>(s_list + SR_ORD_LIST_CAP) + (threadIdx.x >> 5) * SR_ORD_RING;
    uint8_t *s_stencil = reinterpret_cast<uint8_t *>(reinterpret_cast<uint4 *>(s_list + SR_ORD_LIST_CAP) + SR_RASTER_WARPS * SR_ORD_RING);
    __shared__ uint32_t s_wcount[SR_RASTER_WARPS], s_wrows[SR_RASTER_WARPS];
    __shared__ uint32_t s_claim[SR_RASTER_WARPS][SR_ORD_BAND * SR_TILE_W];  // per warp: bids for the pixels of its band (sr_ord_lines_chunk)
    __shared__ uint8_t s_wband[SR_RASTER_WARPS][SR_RASTER_WARPS];     // [warp][band]: hits of that warp's group crossing the band
    __shared__ uint8_t s_band[SR_RASTER_WARPS][SR_RASTER_THREADS];  // per band: compacted record indices, in list order

    constexpr int NK = SrFsInfo<FS>::NK, NP = (NK + 3) / 4;
    constexpr uint32_t RH = SR_TILE_H / SR_RASTER_WARPS;  // tile rows owned by each warp
    constexpr uint32_t NW = SR_RASTER_WARPS;
    static_assert((NW & (NW - 1)) == 0 && RH * SR_TILE_W <= SR_ORD_BAND * SR_TILE_W, "triangles: row r of the tile belongs to warp r % NW");
    static_assert(SR_TILE_H % SR_RASTER_WARPS == 0, "tile height must split evenly over the warps");

    const uint32_t tile = p.shard_rank + blockIdx.x * p.shard_world;
    const uint32_t tx = tile % p.fb.ntx, ty = tile / p.fb.ntx;
    const uint32_t x0 = tx * SR_TILE_W, y0 = ty * SR_TILE_H;
    {   // optimistic list sizing (see build_bins): nothing may be touched if a list did not fit
        const uint32_t nt = p.fb.ntx * p.fb.nty;
        if (p.tri_off[nt] > p.tri_cap || p.line_off[nt] > p.line_cap || p.point_off[nt] > p.point_cap) return;
    }
    const uint32_t tbeg = p.tri_off[tile], LT = p.tri_off[tile + 1] - tbeg;
    const uint32_t lbeg = p.line_off[tile], LL = p.line_off[tile + 1] - lbeg;
    const uint32_t pbeg = p.point_off[tile], LP = p.point_off[tile + 1] - pbeg;
    if (LT + LL + LP == 0 && !p.fb.pending_clear) return;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t W = p.fb.width, H = p.fb.height;

    SrOrdCtx c;
    c.color = s_color; c.depth = s_depth; c.winner = s_winner; c.stencil = s_stencil; c.p = &p;
    c.x0 = x0; c.y0 = y0;
    c.xe = min(x0 + SR_TILE_W, W) - 1; c.ye = min(y0 + SR_TILE_H, H) - 1;
    c.has_stencil = p.fb.stencil != nullptr;
    c.sbytes = c.has_stencil ? p.fb.stencil_bytes : 1u;  // (the tile's stencil values sit at the end of the dynamic shared memory: sbytes per pixel)
    c.smax = c.sbytes == 1 ? 0xFFu : c.sbytes == 2 ? 0xFFFFu : 0xFFFFFFFFu;
    c.mesh_stencil = p.stencil_value & c.smax;  // the mesh's stencil value is of the buffer's type

    for (uint32_t i = lane; i < SR_ORD_BAND * SR_TILE_W; i += 32) s_claim[warp][i] = 0;
    uint32_t claim_round = 0;  // (warp-uniform) bidding round of sr_ord_lines_chunk
    // load the tile (or generate the pending clear on chip)
    for (uint32_t i = tid; i < SR_TILE_PIXELS; i += SR_RASTER_THREADS) {
        const uint32_t px = x0 + i % SR_TILE_W, py = y0 + i / SR_TILE_W;
        float cl[4] = {p.fb.clear[0], p.fb.clear[1], p.fb.clear[2], p.fb.clear[3]};
        if (p.fb.u8color) sr_quantise_u8(cl, false, 0);
        float4 col = make_float4(cl[0], cl[1], cl[2], cl[3]);
        float d = __uint_as_float(SR_DEPTH_FAR_BITS);
        uint32_t s = 0;
        if (!p.fb.pending_clear && px < W && py < H) {
            float src[5];
            sr_fb_load_pixel(p.fb, (uint64_t)py * W + px, src);
            col = make_float4(src[0], src[1], src[2], src[3]);
            d = src[4];
            if (c.has_stencil) s = sr_stencil_load(p.fb.stencil, c.sbytes, (uint64_t)py * W + px);
        }
        s_color[i] = col;
        s_depth[i] = d;
        sr_stencil_store(s_stencil, c.sbytes, i, s);
        s_winner[i] = 0;
    }
    __syncthreads();

    // Dense batches.  A tile's lists hold GROUPS of 32 consecutive primitives; when the submission order is not spatially
    // coherent most primitives of a listed group miss the tile, and a batch of 8 groups would set up only a handful of
    // primitives between its barriers.  `gather` therefore walks the group list with the cheap test only (packed tile
    // rectangle against this tile) and queues the ids of the primitives that hit, in list order, until 256 are waiting
    // (or the list ends); the expensive part -- dependent fetches, clipping, edge setup, the in-order sweep -- always runs
    // on full batches.  Ids beyond the batch are carried over to the next one.
    __shared__ uint32_t s_hits[2 * SR_RASTER_THREADS];
    auto gather = [&](const uint32_t *lst, uint32_t Ln, const uint32_t *rects, uint32_t nprim, uint32_t &gpos, uint32_t nh) -> uint32_t {
        while (nh < SR_RASTER_THREADS && gpos < Ln) {
            const uint32_t gi = gpos + warp;
            bool hit = false;
            uint32_t t = 0;
            if (gi < Ln) {
                t = lst[gi] * SR_GROUP + lane;
                const uint32_t rect = t < nprim ? __ldg(rects + t) : SR_RECT_INVALID;
                hit = rect != SR_RECT_INVALID && sr_rect_hits(rect, tx, ty);
            }
            const uint32_t mask = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) s_wcount[warp] = __popc(mask);
            __syncthreads();
            uint32_t base = 0, total = 0;
#pragma unroll
            for (uint32_t w2 = 0; w2 < SR_RASTER_WARPS; ++w2) {
                const uint32_t cnt = s_wcount[w2];
                if (w2 < warp) base += cnt;
                total += cnt;
            }
            if (hit) s_hits[nh + base + __popc(mask & ((1u << lane) - 1u))] = t;
            __syncthreads();
            nh += total;
            gpos += SR_RASTER_WARPS;
        }
        return nh;
    };
    // the ids beyond the batch just processed move to the front of the queue
    auto carry_over = [&](uint32_t nh, uint32_t nb) -> uint32_t {
        const uint32_t rest = nh - nb;
        const uint32_t v = tid < rest ? s_hits[nb + tid] : 0u;
        __syncthreads();
        if (tid < rest) s_hits[tid] = v;
        __syncthreads();
        return rest;
    };

    // ---------------- triangles ----------------
    if (LT > 0) {
        uint32_t *lst = p.tri_list + tbeg;
        if (LT <= SR_ORD_LIST_CAP) {
            for (uint32_t i = tid; i < LT; i += SR_RASTER_THREADS) s_list[i] = lst[i];
            __syncthreads();
            lst = s_list;
        }
        sr_block_sort(lst, LT);
        __syncthreads();
        uint32_t gpos = 0, queued = 0;
        while (gpos < LT || queued > 0) {
            const uint32_t nh = gather(lst, LT, p.tri_rects, p.ntris, gpos, queued);
            const uint32_t nb = min(nh, (uint32_t)SR_RASTER_THREADS);
            bool hit = tid < nb;
            SrOrdSetup su;
            {
                const uint32_t t = hit ? s_hits[tid] : 0u;
                if (hit) {
                    const SrVertexSet *vs;
                    uint32_t vi[3];
                    sr_prim_vertices<3>(p.tris, t, vs, vi);
                    const float4 A = __ldg(vs->pos + vi[0]), B = __ldg(vs->pos + vi[1]), C = __ldg(vs->pos + vi[2]);
                    const uint32_t minx = max(sr_clamp_as_int(fminf(fminf(A.x, B.x), C.x), 0, W - 1), x0);
                    const uint32_t miny = max(sr_clamp_as_int(fminf(fminf(A.y, B.y), C.y), 0, H - 1), y0);
                    const uint32_t maxx = min(sr_clamp_as_int(fmaxf(fmaxf(A.x, B.x), C.x), 0, W - 1), c.xe);
                    const uint32_t maxy = min(sr_clamp_as_int(fmaxf(fmaxf(A.y, B.y), C.y), 0, H - 1), c.ye);
                    hit = minx <= maxx && miny <= maxy;
                    if (hit) {
                        const SrTri tr = sr_tri_setup(A.x, A.y, B.x, B.y, C.x, C.y);
                        su.e = make_float4(tr.a, tr.b, tr.c, tr.d);
                        su.f = make_float4(tr.x3, tr.y3, tr.det, tr.rdet);
                        su.A = A; su.B = B; su.C = C;
                        su.bx = minx | (maxx << 16);
                        su.by = miny | (maxy << 16);
                        su.canonical = sr_prim_canonical(p.tris, t, 0);
                        su.second = t < p.tris.n0 ? 0u : 1u;
                        su.vi[0] = vi[0]; su.vi[1] = vi[1]; su.vi[2] = vi[2];
                        su.pad = tr.fast ? 1u : 0u;  // (validity of the exact-division shortcut: decided once, at setup)
                    }
                }
            }
            // Compaction in list order (ballot + warp counts) and, per band of rows, the ordered list of the hits that cross
            // it: a warp then visits only the triangles of its own band instead of scanning every record of the batch.
            const uint32_t mask = __ballot_sync(0xffffffffu, hit);
            const uint32_t rows_w = __reduce_add_sync(0xffffffffu, hit ? (su.by >> 16) - (su.by & 0xffffu) + 1u : 0u);
            if (lane == 0) { s_wcount[warp] = __popc(mask); s_wrows[warp] = rows_w; }
            __syncthreads();
            uint32_t base = 0, total = 0, rows_total = 0;
#pragma unroll
            for (uint32_t w2 = 0; w2 < SR_RASTER_WARPS; ++w2) {
                const uint32_t cnt = s_wcount[w2];
                if (w2 < warp) base += cnt;
                total += cnt;
                rows_total += s_wrows[w2];
            }
            // Row ownership of this batch (CTA-uniform; batches are separated by barriers, so it may change from one to the next):
            // contiguous bands of RH rows, or -- when the batch's triangles are only a few rows high -- row r belongs to warp
            // r % NW.  The consecutive small triangles of a batch lie in a few neighbouring rows, which contiguous bands would hand
            // to one or two warps while the others wait at the barrier; tall triangles would be visited by every warp instead.
            const bool interleave = rows_total <= SR_ORD_IL_ROWS * total;
            const uint32_t rstep = interleave ? NW : 1u;
            uint32_t bandmask = 0;  // warps whose rows this triangle's box crosses
            if (hit) {
                const uint32_t lo = (su.by & 0xffffu) - y0, hi = (su.by >> 16) - y0;
                if (interleave) {
                    const uint32_t nrow = hi - lo + 1, m = nrow >= NW ? (1u << NW) - 1u : (1u << nrow) - 1u, sh = lo % NW;
                    bandmask = ((m << sh) | (m >> (NW - sh))) & ((1u << NW) - 1u);
                } else {
                    bandmask = ((2u << (hi / RH)) - 1u) & ~((1u << (lo / RH)) - 1u);
                }
            }
            uint32_t bm[SR_RASTER_WARPS];
#pragma unroll
            for (uint32_t b = 0; b < SR_RASTER_WARPS; ++b) bm[b] = __ballot_sync(0xffffffffu, (bandmask >> b) & 1u);
            if (lane < SR_RASTER_WARPS) {
                uint32_t mine = 0;
#pragma unroll
                for (uint32_t b = 0; b < SR_RASTER_WARPS; ++b) mine = lane == b ? bm[b] : mine;
                s_wband[warp][lane] = (uint8_t)__popc(mine);
            }
            __syncthreads();
            const uint32_t at = base + __popc(mask & ((1u << lane) - 1u));
            // entries the earlier warps put on band b's list: summed once per warp (lane b), handed to the lanes by shuffle
            uint32_t bprefix = 0;
            if (lane < SR_RASTER_WARPS)
                for (uint32_t w2 = 0; w2 < warp; ++w2) bprefix += s_wband[w2][lane];
#pragma unroll
            for (uint32_t b = 0; b < SR_RASTER_WARPS; ++b) {
                const uint32_t bbase = __shfl_sync(0xffffffffu, bprefix, b);
                if (hit && ((bandmask >> b) & 1u)) s_band[b][bbase + __popc(bm[b] & ((1u << lane) - 1u))] = (uint8_t)at;
            }
            if (hit) s_setup[at] = su;
            uint32_t nband = 0;
#pragma unroll
            for (uint32_t w2 = 0; w2 < SR_RASTER_WARPS; ++w2) nband += s_wband[w2][warp];
            __syncthreads();
            // the rows of [ylo, yhi] (frame coordinates, inside the tile) this warp owns: first, first + rstep, ...; returns their number
            auto own_rows = [&](uint32_t ylo, uint32_t yhi, uint32_t &first) -> uint32_t {
                if (interleave) {
                    first = ylo + ((warp - (ylo - y0)) & (NW - 1u));
                    return first <= yhi ? (yhi - first) / NW + 1u : 0u;
                }
                first = max(ylo, y0 + warp * RH);
                const uint32_t last = min(yhi, y0 + warp * RH + RH - 1);
                return first <= last ? last - first + 1u : 0u;
            };
            // index of an owned row among the warp's RH rows (the warp's bidding words, sr_ord_lines_chunk's layout)
            auto own_row_index = [&](uint32_t py) -> uint32_t { return interleave ? (py - y0) / NW : (py - y0) - warp * RH; };
            if (!SrFsInfo<FS>::DISCARDS) {
                // Deferred shading.  Whether a fragment passes depends only on coverage, z and the depth left by earlier
                // fragments of its pixel (the shader cannot discard), so the in-order sweep only runs the cheap part --
                // stencil step, exact barycentrics, depth test and depth update -- and queues every passing fragment
                // {pixel, triangle, u, v} in the warp's ring.  As soon as 32 are queued the warp shades them together
                // (all lanes busy whatever the triangle shapes are) and blends them into the tile in queue order; fragments
                // of one pixel are queued in submission order because the pixel belongs to exactly one warp.
                uint32_t head = 0, count = 0;  // (warp-uniform)
                auto shade_and_blend = [&](uint32_t n) {
                    float col[4];
                    uint32_t li = 0xFFFFFFFFu - lane;  // distinct dummies for idle lanes
                    if (lane < n) {
                        const uint4 rec = s_ring[(head + lane) % SR_ORD_RING];
                        li = rec.x & 0xffffu;
                        const SrOrdSetup &q = s_setup[rec.x >> 16];
                        const float u = __uint_as_float(rec.y), v = __uint_as_float(rec.z), w = 1.0f - u - v;  // triangle.rs:110
                        const float4 A = q.A, B = q.B, C = q.C;
                        float sv[4 + NP * 4 + 1];
                        sv[0] = sr_bary(u, A.x, v, B.x, w, C.x);
                        sv[1] = sr_bary(u, A.y, v, B.y, w, C.y);
                        sv[2] = sr_bary(u, A.z, v, B.z, w, C.z);
                        sv[3] = sr_bary(u, A.w, v, B.w, w, C.w);
                        const SrVertexSet *vs = q.second ? &p.tris.vs1 : &p.tris.vs0;
                        const uint32_t vi0 = q.vi[0], vi1 = q.vi[1], vi2 = q.vi[2];
#pragma unroll
                        for (int pl = 0; pl < NP; ++pl) {
                            const float4 ka = __ldg(vs->attr + sr_attr_at(vs->np, vi0, pl));
                            const float4 kb = __ldg(vs->attr + sr_attr_at(vs->np, vi1, pl));
                            const float4 kc = __ldg(vs->attr + sr_attr_at(vs->np, vi2, pl));
                            sv[4 + pl * 4 + 0] = sr_bary(u, ka.x, v, kb.x, w, kc.x);
                            sv[4 + pl * 4 + 1] = sr_bary(u, ka.y, v, kb.y, w, kc.y);
                            sv[4 + pl * 4 + 2] = sr_bary(u, ka.z, v, kb.z, w, kc.z);
                            sv[4 + pl * 4 + 3] = sr_bary(u, ka.w, v, kb.w, w, kc.w);
                        }
                        sr_fragment_shader<FS>(p.fs, sv, col);
                    }
                    // blend in queue order: lanes that hold fragments of the same pixel take turns, lowest lane first
                    const uint32_t peers = __match_any_sync(0xffffffffu, li);
                    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
                    const uint32_t rounds = __reduce_max_sync(0xffffffffu, lane < n ? (uint32_t)__popc(peers) : 0u);
                    for (uint32_t r = 0; r < rounds; ++r) {
                        if (lane < n && rank == r) {
                            const float4 old = s_color[li];
                            const float dstc[4] = {old.x, old.y, old.z, old.w};
                            float outc[4];
                            sr_blend(p.blend, col, dstc, outc);
                            s_color[li] = make_float4(outc[0], outc[1], outc[2], outc[3]);
                        }
                        __syncwarp();
                    }
                    head = (head + n) % SR_ORD_RING;
                    count -= n;
                };
                // Runs of SMALL triangles (box inside the band of at most 32 pixels) are swept one triangle per LANE: a mesh of
                // pixel-sized triangles would otherwise keep 3-9 of 32 lanes busy for a whole warp step per triangle.  (A) every
                // lane finds the covered pixels of its own triangle (a bit mask of box positions, triangle.rs:104-120); (B) bidding
                // rounds keep every pixel's fragments in submission order: each lane bids (round, 255 - lane) with atomicMax for
                // the pixels of its not-yet-applied fragments, then applies -- depth test, depth update, queue for shading --
                // exactly those fragments whose pixel it holds.  Lanes hold consecutive list entries, so the lowest lane wins
                // every contested pixel: a later triangle never overtakes an earlier one, and the earliest unfinished lane always
                // completes.  A triangle covers a pixel at most once, so the order among its own fragments is free.  Triangles
                // seldom overlap inside a 32-triangle run (shared edges only), so nearly every run takes one round.
                // Active stencil configurations touch every pixel of the box before coverage (triangle.rs:91-99) and keep the
                // cooperative sweep below.
                const bool lane_ok = !c.has_stencil || (p.stencil_test == SR_STENCIL_ALWAYS && p.stencil_op == SR_STENCIL_KEEP);
                uint32_t *claim = s_claim[warp];
                // The mode is chosen per batch and band: one lane per triangle only pays when the runs are long, i.e. when (nearly)
                // all the triangles of the band are small; a mixed population keeps the cooperative sweep throughout.
                bool lane_mode = false;
                if (lane_ok && nband >= 8) {
                    uint32_t nsmall = 0;
                    for (uint32_t k = lane; k < nband; k += 32) {
                        const uint2 box = *reinterpret_cast<const uint2 *>(&s_setup[s_band[warp][k]].bx);
                        const uint32_t bw = (box.x >> 16) - (box.x & 0xffffu) + 1;
                        uint32_t first;
                        const uint32_t bh = own_rows(box.y & 0xffffu, box.y >> 16, first);
                        nsmall += bw * bh <= SR_ORD_LANE_BOX ? 1u : 0u;
                    }
                    nsmall = __reduce_add_sync(0xffffffffu, nsmall);
                    lane_mode = nsmall * 8 >= nband * 7;
                }
                uint32_t ncoop = lane_mode ? 0u : nband;  // entries from k on that take the cooperative sweep without another look
                for (uint32_t k = 0; k < nband;) {
                    uint32_t run = 0;
                    if (ncoop == 0) {
                        bool small = false;
                        uint32_t s = 0, minx = 0, r0 = 0, bw = 0, bh = 0;
                        if (k + lane < nband) {
                            s = s_band[warp][k + lane];
                            const uint2 box = *reinterpret_cast<const uint2 *>(&s_setup[s].bx);
                            minx = box.x & 0xffffu;
                            uint32_t maxx = box.x >> 16, ylo = box.y & 0xffffu, yhi = box.y >> 16;
                            // candidate tightening (proof above sr_tightening_applies): the pixels outside the tightened range
                            // fail the reference's own coverage test, and without an active stencil nothing else looks at them
                            const SrOrdSetup &q = s_setup[s];
                            const float xmin = fminf(fminf(q.A.x, q.B.x), q.C.x), xmax = fmaxf(fmaxf(q.A.x, q.B.x), q.C.x);
                            const float ymin = fminf(fminf(q.A.y, q.B.y), q.C.y), ymax = fmaxf(fmaxf(q.A.y, q.B.y), q.C.y);
                            bool empty = false;
                            if (sr_tightening_applies(xmin, xmax, ymin, ymax, q.f.z)) {
                                const int lx = max((int)minx, sr_tight_lo(xmin)), hx = min((int)maxx, sr_tight_hi(xmax));
                                const int ly = max((int)ylo, sr_tight_lo(ymin)), hy = min((int)yhi, sr_tight_hi(ymax));
                                empty = lx > hx || ly > hy;
                                if (!empty) { minx = (uint32_t)lx; maxx = (uint32_t)hx; ylo = (uint32_t)ly; yhi = (uint32_t)hy; }
                            }
                            bh = empty ? 0u : own_rows(ylo, yhi, r0);  // (r0: the first owned row; box row j is frame row r0 + j * rstep)
                            bw = bh == 0 ? 0u : maxx - minx + 1;
                            small = bw * bh <= SR_ORD_LANE_PIX;
                        }
                        const uint32_t smallmask = __ballot_sync(0xffffffffu, small);
                        run = ~smallmask ? (uint32_t)__ffs(~smallmask) - 1u : 32u;  // leading run of small triangles
                        if (run == 0) ncoop = min(smallmask ? (uint32_t)__ffs(smallmask) - 1u : 32u, nband - k);  // leading run of the others
                        if (run > 0) {
                            const bool mine = lane < run;
                            const SrOrdSetup &q = s_setup[s];
                            SrTri tr;
                            tr.a = q.e.x; tr.b = q.e.y; tr.c = q.e.z; tr.d = q.e.w;
                            tr.x3 = q.f.x; tr.y3 = q.f.y; tr.det = q.f.z; tr.rdet = q.f.w;
                            tr.dsign = __float_as_uint(tr.det) & 0x80000000u;
                            tr.fast = q.pad != 0u;
                            const float z1 = q.A.z, z2 = q.B.z, z3 = q.C.z;
                            const uint32_t canonical = q.canonical;
                            uint32_t rem = 0;  // box positions (row-major) whose fragment is still to be applied
                            if (mine)
                                for (uint32_t j = 0; j < bh; ++j)  // (one row per call: the owned rows may be NW apart)
                                    sr_raster_box<false>(tr, z1, z2, z3, minx, r0 + j * rstep, bw, 1u, 0u, [&](uint32_t px, uint32_t, unsigned long long) {
                                        rem |= 1u << (j * bw + (px - minx));
                                    });
                            const uint32_t cbase = own_row_index(r0) * SR_TILE_W + (minx - x0);  // box origin among the warp's rows
                            // box position i < 32 -> (row, column) without an integer division: (i * m) >> 16 == i / bw for every i < 32, bw <= 32
                            const uint32_t bwm = (1u << 16) / max(bw, 1u) + 1u;
                            auto row_of = [&](uint32_t i) { return (i * bwm) >> 16; };
                            while (__any_sync(0xffffffffu, rem != 0)) {
                                ++claim_round;
                                const uint32_t bid = (claim_round << 8) | (255u - lane);
                                for (uint32_t m = rem; m; m &= m - 1) {
                                    const uint32_t i = (uint32_t)__ffs(m) - 1u;
                                    const uint32_t rr = row_of(i);
                                    atomicMax(&claim[cbase + rr * SR_TILE_W + (i - rr * bw)], bid);
                                }
                                __syncwarp();
                                uint32_t todo = 0;  // the fragments this lane applies in this round
                                for (uint32_t m = rem; m; m &= m - 1) {
                                    const uint32_t i = (uint32_t)__ffs(m) - 1u;
                                    const uint32_t rr = row_of(i);
                                    if (claim[cbase + rr * SR_TILE_W + (i - rr * bw)] == bid) todo |= 1u << i;
                                }
                                rem &= ~todo;
                                while (__any_sync(0xffffffffu, todo != 0)) {
                                    bool pass = false;
                                    uint32_t li = 0;
                                    float u = 0.0f, v = 0.0f, w;
                                    if (todo) {
                                        const uint32_t i = (uint32_t)__ffs(todo) - 1u;
                                        todo &= todo - 1;
                                        const uint32_t rr = row_of(i);
                                        const uint32_t px = minx + (i - rr * bw), py = r0 + rr * rstep;
                                        li = (py - y0) * SR_TILE_W + (px - x0);
                                        sr_tri_bary(tr, px, py, u, v, w);  // (covered: the same arithmetic as pass A)
                                        const float z = sr_bary(u, z1, v, z2, w, z3);
                                        if (z >= s_depth[li]) {  // triangle.rs:126 (z < 0 was part of pass A)
                                            pass = true;
                                            s_depth[li] = z;
                                            s_winner[li] = canonical + 1;
                                        }
                                    }
                                    const uint32_t m = __ballot_sync(0xffffffffu, pass);
                                    if (pass)
                                        s_ring[(head + count + __popc(m & ((1u << lane) - 1u))) % SR_ORD_RING] =
                                            make_uint4(li | (s << 16), __float_as_uint(u), __float_as_uint(v), 0u);
                                    count += __popc(m);
                                    __syncwarp();
                                    if (count >= 32) shade_and_blend(32);
                                }
                            }
                            k += run;
                            continue;
                        }
                    }
                    const uint32_t s = s_band[warp][k];
                    ++k;
                    --ncoop;
                    const uint2 box = *reinterpret_cast<const uint2 *>(&s_setup[s].bx);
                    const uint32_t minx = box.x & 0xffffu, maxx = box.x >> 16;
                    uint32_t r0;
                    const uint32_t nrows = own_rows(box.y & 0xffffu, box.y >> 16, r0);
                    const SrOrdSetup &q = s_setup[s];
                    SrTri tr;
                    tr.a = q.e.x; tr.b = q.e.y; tr.c = q.e.z; tr.d = q.e.w;
                    tr.x3 = q.f.x; tr.y3 = q.f.y; tr.det = q.f.z; tr.rdet = q.f.w;
                    tr.dsign = __float_as_uint(tr.det) & 0x80000000u;
                    tr.fast = q.pad != 0u;
                    const float z1 = q.A.z, z2 = q.B.z, z3 = q.C.z;
                    const uint32_t bw = maxx - minx + 1, npix = bw * nrows, canonical = q.canonical;
                    const uint32_t bwm = 0xFFFFFFFFu / bw + 1u;  // __umulhi(i, bwm) == i / bw for i, bw < 2^16 (one division per triangle, none per pixel)
                    for (uint32_t base = 0; base < npix; base += 32) {
                        const uint32_t i = base + lane;
                        bool pass = false;
                        uint32_t li = 0;
                        float u = 0.0f, v = 0.0f, w;
                        if (i < npix) {
                            const uint32_t rr = bw == 1u ? i : __umulhi(i, bwm);  // (bw == 1: the multiplier 2^32 does not fit)
                            const uint32_t px = minx + (i - rr * bw), py = r0 + rr * rstep;
                            li = (py - y0) * SR_TILE_W + (px - x0);
                            if (sr_ord_stencil_step(c, li) && sr_tri_bary(tr, px, py, u, v, w)) {
                                const float z = sr_bary(u, z1, v, z2, w, z3);
                                if (z < 0.0f && z >= s_depth[li]) {  // triangle.rs:120,126
                                    pass = true;
                                    s_depth[li] = z;
                                    s_winner[li] = canonical + 1;
                                }
                            }
                        }
                        const uint32_t m = __ballot_sync(0xffffffffu, pass);
                        if (pass)
                            s_ring[(head + count + __popc(m & ((1u << lane) - 1u))) % SR_ORD_RING] =
                                make_uint4(li | (s << 16), __float_as_uint(u), __float_as_uint(v), 0u);
                        count += __popc(m);
                        __syncwarp();
                        if (count >= 32) shade_and_blend(32);
                    }
                }
                if (count) shade_and_blend(count);  // the records refer to this batch's setups: drain before the next batch
            } else {
            for (uint32_t k = 0; k < nband; ++k) {
                const uint32_t s = s_band[warp][k];
                const uint2 box = *reinterpret_cast<const uint2 *>(&s_setup[s].bx);
                const uint32_t minx = box.x & 0xffffu, maxx = box.x >> 16;
                uint32_t r0;
                const uint32_t nrows = own_rows(box.y & 0xffffu, box.y >> 16, r0);
                const SrOrdSetup &q = s_setup[s];
                SrTri tr;
                tr.a = q.e.x; tr.b = q.e.y; tr.c = q.e.z; tr.d = q.e.w;
                tr.x3 = q.f.x; tr.y3 = q.f.y; tr.det = q.f.z; tr.rdet = q.f.w;
                tr.dsign = __float_as_uint(tr.det) & 0x80000000u;
                tr.fast = q.pad != 0u;
                const float4 A = q.A, B = q.B, C = q.C;
                const uint32_t bw = maxx - minx + 1, npix = bw * nrows;
                const uint32_t bwm = 0xFFFFFFFFu / bw + 1u;  // __umulhi(i, bwm) == i / bw
                const SrVertexSet *vs = q.second ? &p.tris.vs1 : &p.tris.vs0;
                const uint32_t vi0 = q.vi[0], vi1 = q.vi[1], vi2 = q.vi[2], canonical = q.canonical;
                for (uint32_t i = lane; i < npix; i += 32) {
                    const uint32_t rr = bw == 1u ? i : __umulhi(i, bwm);  // (bw == 1: the multiplier 2^32 does not fit)
                    const uint32_t px = minx + (i - rr * bw), py = r0 + rr * rstep;
                    const uint32_t li = (py - y0) * SR_TILE_W + (px - x0);
                    if (!sr_ord_stencil_step(c, li)) continue;
                    float u, v, w;
                    if (!sr_tri_bary(tr, px, py, u, v, w)) continue;
                    float sv[4 + NP * 4 + 1];
                    sv[2] = sr_bary(u, A.z, v, B.z, w, C.z);
                    if (!(sv[2] < 0.0f) || !(sv[2] >= s_depth[li])) continue;
                    sv[0] = sr_bary(u, A.x, v, B.x, w, C.x);
                    sv[1] = sr_bary(u, A.y, v, B.y, w, C.y);
                    sv[3] = sr_bary(u, A.w, v, B.w, w, C.w);
#pragma unroll
                    for (int pl = 0; pl < NP; ++pl) {
                        const float4 ka = __ldg(vs->attr + sr_attr_at(vs->np, vi0, pl));
                        const float4 kb = __ldg(vs->attr + sr_attr_at(vs->np, vi1, pl));
                        const float4 kc = __ldg(vs->attr + sr_attr_at(vs->np, vi2, pl));
                        sv[4 + pl * 4 + 0] = sr_bary(u, ka.x, v, kb.x, w, kc.x);
                        sv[4 + pl * 4 + 1] = sr_bary(u, ka.y, v, kb.y, w, kc.y);
                        sv[4 + pl * 4 + 2] = sr_bary(u, ka.z, v, kb.z, w, kc.z);
                        sv[4 + pl * 4 + 3] = sr_bary(u, ka.w, v, kb.w, w, kc.w);
                    }
                    sr_ord_shade_write<FS>(c, li, sv, false, 1.0f, canonical);
                }
                __syncwarp();
            }
            }
            __syncthreads();
            queued = carry_over(nh, nb);
        }
    }

    // ---------------- lines, then points (fragment.rs:284-311): walked in order, plotted by pixel owner (sr_ord_owns) ----------------
    for (int kind = 2; kind >= 1; --kind) {
        const uint32_t Ln = kind == 2 ? LL : LP;
        if (Ln == 0) continue;
        uint32_t *lst = kind == 2 ? p.line_list + lbeg : p.point_list + pbeg;
        if (Ln <= SR_ORD_LIST_CAP) {
            for (uint32_t i = tid; i < Ln; i += SR_RASTER_THREADS) s_list[i] = lst[i];
            __syncthreads();
            lst = s_list;
        }
        sr_block_sort(lst, Ln);
        __syncthreads();
        {
            const uint32_t nprim = kind == 2 ? p.nlines : p.npoints;
            const uint32_t *rects = kind == 2 ? p.line_rects : p.point_rects;
            SrOrdLineRec *s_line = reinterpret_cast<SrOrdLineRec *>(s_setup);
            SrOrdPointRec *s_point = reinterpret_cast<SrOrdPointRec *>(s_setup);
            // 8 groups = 256 primitives per batch: fetched and set up one thread each (the dependent gathers of a whole batch
            // overlap), then walked in list order -- thread tid holds primitive lst[gb + tid/32] * 32 + tid%32 and the list
            // is sorted, so record order is submission order
            uint32_t gpos = 0, queued = 0;
            while (gpos < Ln || queued > 0) {
                const uint32_t nh = gather(lst, Ln, rects, nprim, gpos, queued);
                const uint32_t nb = min(nh, (uint32_t)SR_RASTER_THREADS);
                bool hit = tid < nb;
                const uint32_t t = hit ? s_hits[tid] : 0u;
                // records of the primitives that touch this tile, compacted in list order (ballot + warp counts)
                SrOrdLineRec lr;
                SrOrdPointRec pr;
                if (kind == 2) {
                    lr.valid = 0;
                    if (hit) sr_ord_line_setup(p, t, lr);
                    hit = hit && lr.valid;
                } else if (hit) {
                    const SrVertexSet *vs;
                    uint32_t vi[1];
                    sr_prim_vertices<1>(p.points, t, vs, vi);
                    pr.P = __ldg(vs->pos + vi[0]);
                    pr.vi = vi[0];
                    pr.second = t < p.points.n0 ? 0u : 1u;
                    pr.canonical = p.point_base + sr_prim_canonical(p.points, t, 0);
                    pr.valid = 1;
                }
                const uint32_t mask = __ballot_sync(0xffffffffu, hit);
                if (lane == 0) s_wcount[warp] = __popc(mask);
                __syncthreads();
                uint32_t base = 0, total = 0;
#pragma unroll
                for (uint32_t w2 = 0; w2 < SR_RASTER_WARPS; ++w2) {
                    const uint32_t cnt = s_wcount[w2];
                    if (w2 < warp) base += cnt;
                    total += cnt;
                }
                if (hit) {
                    const uint32_t at = base + __popc(mask & ((1u << lane) - 1u));
                    if (kind == 2) s_line[at] = lr;
                    else s_point[at] = pr;
                }
                __syncthreads();
                if (kind == 2) {
                    for (uint32_t k = 0; k < total; k += 32) sr_ord_lines_chunk<FS>(c, s_line + k, min(32u, total - k), s_claim[warp], claim_round);
                } else {
                    for (uint32_t k = 0; k < total; ++k) sr_ord_point<FS>(c, s_point[k]);
                }
                __syncthreads();
                queued = carry_over(nh, nb);
            }
        }
        __syncthreads();
    }

    // write the tile back once
    for (uint32_t i = tid; i < SR_TILE_PIXELS; i += SR_RASTER_THREADS) {
        const uint32_t px = x0 + i % SR_TILE_W, py = y0 + i / SR_TILE_W;
        if (px >= W || py >= H) continue;
        const float4 col = s_color[i];
        const float o5[5] = {col.x, col.y, col.z, col.w, s_depth[i]};
        sr_fb_store_pixel(p.fb, (uint64_t)py * W + px, o5);
        if (c.has_stencil) sr_stencil_store(p.fb.stencil, c.sbytes, (uint64_t)py * W + px, sr_stencil_load(s_stencil, c.sbytes, i));
        if (p.fb.winner && s_winner[i]) p.fb.winner[(uint64_t)py * W + px] = s_winner[i];  // plane is zeroed per draw
    }
}

// self-test of sr_div_exact against the IEEE division it replaces (parity evidence for the coverage shortcut):
// random sign/exponent/mantissa patterns over the whole validity range, plus all-ones / sparse mantissas.
__global__ void __launch_bounds__(256) k_selftest_division(uint64_t seed, uint64_t per_thread, unsigned long long *mismatches) {
    uint64_t state = seed + ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull;
    auto next = [&]() {
        state += 0x9E3779B97F4A7C15ull;
        uint64_t z = state;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    };
    auto make = [&](uint64_t r, uint32_t elo, uint32_t ehi) {
        uint32_t mant = (uint32_t)(r >> 8) & 0x7FFFFFu;
        const uint32_t mode = (uint32_t)(r >> 40) & 15u;
        if (mode == 0) mant = 0x7FFFFFu;
        else if (mode == 1) mant = 0;
        else if (mode == 2) mant &= 0x7FF000u;
        else if (mode == 3) mant |= 0x7FF000u;
        else if (mode == 4) mant = 1u << ((r >> 44) % 23);
        const uint32_t e = elo + (uint32_t)((r >> 48) % (ehi - elo));
        return __uint_as_float(((uint32_t)(r & 1) << 31) | (e << 23) | mant);
    };
    unsigned long long bad = 0;
    for (uint64_t i = 0; i < per_thread; ++i) {
        const float det = make(next(), 87, 167);  // |det| in [2^-40, 2^40)
        const float n = make(next(), 67, 187);    // |n| in [2^-60, 2^60)
        const float q = sr_div_exact(n, det, __frcp_rn(det));
        if (__float_as_uint(q) != __float_as_uint(__fdiv_rn(n, det))) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// materialise a pending clear (RenderBuffer::clear, renderbuffer/mod.rs:126-133) when no draw did it
__global__ void __launch_bounds__(256) k_fb_fill(SrFbView fb) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)fb.width * fb.height) return;
    float o[9] = {fb.clear[0], fb.clear[1], fb.clear[2], fb.clear[3], __uint_as_float(SR_DEPTH_FAR_BITS), fb.clear1[0], fb.clear1[1], fb.clear1[2],
                  fb.clear1[3]};
    if (fb.u8color) sr_quantise_u8(o, false, 0);
    if (fb.soa == 2) sr_fb_store_pixel2(fb, i, o);
    else sr_fb_store_pixel(fb, i, o);
    if (fb.stencil) sr_stencil_store(fb.stencil, fb.stencil_bytes, i, 0u);
    if (fb.winner) fb.winner[i] = 0;
}
// realtime_example/src/main.rs:100-116: the presentation loop `(c.x * 255.0) as u8` per channel.  Rust's float -> u8 `as`
// cast truncates toward zero and saturates (NaN -> 0), which is what cvt.rzi.u32.f32 + min does.  order 0: bytes r,g,b,a;
// order 1: a,b,g,r (what the example writes into SDL's RGBA8888 streaming texture).
// Four pixels per thread: 80 contiguous bytes in (five 128-bit loads), 16 bytes out.
__global__ void __launch_bounds__(256) k_fb_to_rgba8(const float *aos, uint64_t n, uint32_t order, uint32_t *out) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;  // pixel quad
    if (q * 4 >= n) return;
    auto pack = [&](float r, float g, float b, float a) {
        const uint32_t R = sr_as_u8(r), G = sr_as_u8(g), B = sr_as_u8(b), A = sr_as_u8(a);
        return order ? (A | (B << 8) | (G << 16) | (R << 24)) : (R | (G << 8) | (B << 16) | (A << 24));
    };
    if (q * 4 + 4 <= n && (reinterpret_cast<uintptr_t>(aos) & 15u) == 0) {
        const float4 *src = reinterpret_cast<const float4 *>(aos + q * 20);
        const float4 v0 = __ldcs(src), v1 = __ldcs(src + 1), v2 = __ldcs(src + 2), v3 = __ldcs(src + 3), v4 = __ldcs(src + 4);
        // pixels: {v0.x v0.y v0.z v0.w | v1.x} {v1.y v1.z v1.w v2.x | v2.y} {v2.z v2.w v3.x v3.y | v3.z} {v3.w v4.x v4.y v4.z | v4.w}
        const uint4 o = make_uint4(pack(v0.x, v0.y, v0.z, v0.w), pack(v1.y, v1.z, v1.w, v2.x), pack(v2.z, v2.w, v3.x, v3.y),
                                   pack(v3.w, v4.x, v4.y, v4.z));
        *reinterpret_cast<uint4 *>(out + q * 4) = o;
    } else {
        for (uint64_t i = q * 4; i < n && i < q * 4 + 4; ++i) out[i] = pack(aos[i * 5], aos[i * 5 + 1], aos[i * 5 + 2], aos[i * 5 + 3]);
    }
}
__global__ void __launch_bounds__(256) k_fb_split(const float *aos, uint64_t n, float *color, float *depth) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (color) { color[i * 4] = aos[i * 5]; color[i * 4 + 1] = aos[i * 5 + 1]; color[i * 4 + 2] = aos[i * 5 + 2]; color[i * 4 + 3] = aos[i * 5 + 3]; }
    if (depth) depth[i] = aos[i * 5 + 4];
}
__global__ void __launch_bounds__(256) k_fb_merge(float *aos, uint64_t n, const float *color, const float *depth) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (color) { aos[i * 5] = color[i * 4]; aos[i * 5 + 1] = color[i * 4 + 1]; aos[i * 5 + 2] = color[i * 4 + 2]; aos[i * 5 + 3] = color[i * 4 + 3]; }
    if (depth) aos[i * 5 + 4] = depth[i];
}
// texture-buffer storage -> the 20-byte AoS pixels sr_framebuffer_download returns, and its presentation read-back
__global__ void __launch_bounds__(256) k_soa_to_aos(const SrFbView fb, float *out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)fb.width * fb.height) return;
    float o[5];
    sr_fb_load_pixel(fb, i, o);
#pragma unroll
    for (int k = 0; k < 5; ++k) out[i * 5 + k] = o[k];
}
__global__ void __launch_bounds__(256) k_soa_to_rgba8(const SrFbView fb, uint32_t order, uint32_t *out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)fb.width * fb.height) return;
    float o[5];
    sr_fb_load_pixel(fb, i, o);
    const uint32_t R = sr_as_u8(o[0]), G = sr_as_u8(o[1]), B = sr_as_u8(o[2]), A = sr_as_u8(o[3]);
    out[i] = order ? (A | (B << 8) | (G << 16) | (R << 24)) : (R | (G << 8) | (B << 16) | (A << 24));
}
// the same three for the 8-byte pixels of an RGBAu8Color target (the colour plane is u8 x 4 per pixel, stored as is)
__global__ void __launch_bounds__(256) k_fb8_to_rgba8(const uint2 *aos, uint64_t n, uint32_t order, uint32_t *out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t v = aos[i].x;
    out[i] = order ? __byte_perm(v, 0, 0x0123) : v;
}
__global__ void __launch_bounds__(256) k_fb8_split(const uint2 *aos, uint64_t n, uint32_t *color, float *depth) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 v = aos[i];
    if (color) color[i] = v.x;
    if (depth) depth[i] = __uint_as_float(v.y);
}
__global__ void __launch_bounds__(256) k_fb8_merge(uint2 *aos, uint64_t n, const uint32_t *color, const float *depth) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint2 v = aos[i];
    if (color) v.x = color[i];
    if (depth) v.y = __float_as_uint(depth[i]);
    aos[i] = v;
}
