// sr_stages.cuh -- vertex stage, geometry stage (clipper + registered geometry shaders), finish.
#pragma once

#include "sr_shaders.cuh"

// =====================================================================================================
// exclusive scan of u32 (three small kernels; inputs here are <= a few hundred MB)
// =====================================================================================================
#define SR_SCAN_THREADS 256
#define SR_SCAN_ITEMS 8
#define SR_SCAN_BLOCK (SR_SCAN_THREADS * SR_SCAN_ITEMS)

__device__ __forceinline__ uint32_t sr_block_exclusive_scan(uint32_t v, uint32_t *total, uint32_t *warp_sums) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += n;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const uint32_t nw = blockDim.x >> 5;
        uint32_t s = lane < nw ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= (uint32_t)o) s += n;
        }
        if (lane < nw) warp_sums[lane] = s;  // inclusive over warps
    }
    __syncthreads();
    const uint32_t warp_base = wid ? warp_sums[wid - 1] : 0;
    if (total) *total = warp_sums[(blockDim.x >> 5) - 1];
    __syncthreads();
    return warp_base + inc - v;
}

__global__ void __launch_bounds__(SR_SCAN_THREADS) k_scan_reduce(const uint32_t *in, uint64_t n, uint32_t *block_sums) {
    __shared__ uint32_t ws[32];
    const uint64_t base = (uint64_t)blockIdx.x * SR_SCAN_BLOCK + (uint64_t)threadIdx.x * SR_SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SR_SCAN_ITEMS; ++i)
        if (base + i < n) s += in[base + i];
    uint32_t total;
    sr_block_exclusive_scan(s, &total, ws);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
// single block: exclusive scan of block_sums in place, grand total to *total
__global__ void __launch_bounds__(SR_SCAN_THREADS) k_scan_sums(uint32_t *block_sums, uint32_t nblocks, uint32_t *total) {
    __shared__ uint32_t ws[32];
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nblocks; base += SR_SCAN_THREADS) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nblocks ? block_sums[i] : 0;
        uint32_t t;
        const uint32_t ex = sr_block_exclusive_scan(v, &t, ws);
        if (i < nblocks) block_sums[i] = carry + ex;
        carry += t;
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(SR_SCAN_THREADS) k_scan_apply(const uint32_t *in, uint64_t n, const uint32_t *block_sums,
                                                                uint32_t *out) {
    __shared__ uint32_t ws[32];
    const uint64_t base = (uint64_t)blockIdx.x * SR_SCAN_BLOCK + (uint64_t)threadIdx.x * SR_SCAN_ITEMS;
    uint32_t v[SR_SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SR_SCAN_ITEMS; ++i) {
        v[i] = base + i < n ? in[base + i] : 0;
        s += v[i];
    }
    uint32_t ex = sr_block_exclusive_scan(s, nullptr, ws) + block_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SR_SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
}

// two independent exclusive scans of at most SR_SCAN_BLOCK elements each in ONE single-block launch (small meshes:
// the clipper's kept/literal counts; launch latency, not work, dominates there)
__global__ void __launch_bounds__(SR_SCAN_THREADS) k_scan_pair_small(const uint32_t *a, const uint32_t *b, uint32_t n, uint32_t *oa,
                                                                     uint32_t *ob, uint32_t *totals) {
    __shared__ uint32_t ws[32];
    const uint32_t base = threadIdx.x * SR_SCAN_ITEMS;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
        const uint32_t *in = which ? b : a;
        uint32_t *out = which ? ob : oa;
        uint32_t v[SR_SCAN_ITEMS], s = 0;
#pragma unroll
        for (int i = 0; i < SR_SCAN_ITEMS; ++i) {
            v[i] = base + i < n ? in[base + i] : 0;
            s += v[i];
        }
        uint32_t total;
        uint32_t ex = sr_block_exclusive_scan(s, &total, ws);
#pragma unroll
        for (int i = 0; i < SR_SCAN_ITEMS; ++i) {
            if (base + i < n) out[base + i] = ex;
            ex += v[i];
        }
        if (threadIdx.x == 0) totals[which] = total;
    }
}

// =====================================================================================================
// a2: vertex stage (VertexShader::run / run_to_fragment, src/pipeline/stages/vertex.rs:87-160)
// One thread shades one vertex (see k_vertex).
// =====================================================================================================
struct SrMeshView {
    const float *planes;  // plane c starts at planes + c*pstride; pstride is a multiple of 4
    uint64_t pstride;
    uint64_t nverts;
    uint32_t vin;
};

// Which vertices a k_vertex launch shades: [begin, end), optionally only those whose bit is set in `mask` and that lie outside
// [skip_lo, skip_hi).  The whole mesh is {0, nverts, null, 0, 0}; a range-sharded frame shades the vertex
// range of the rank's own triangles first and, after the key merge, only the vertices the winners of its tiles reference.
struct SrVertexSpan {
    uint64_t begin, end;
    const uint32_t *mask;
    uint64_t skip_lo, skip_hi;
    const uint8_t *blocks;  // k_vertex: shade only the 256-vertex blocks flagged here; k_vertex_marked: skip them (shaded already)
};
#define SR_VERTEX_BLOCK 256
template <int VS>
__device__ __forceinline__ void sr_vertex_one(const SrVsConst &c, const SrMeshView &m, float4 *pos, float4 *attr, const uint64_t onp,
                                              const uint64_t v) {
    constexpr int VIN = SrVsInfo<VS>::VIN, NK = SrVsInfo<VS>::NK, NP = (NK + 3) / 4;
    float in[VIN];
#pragma unroll
    for (int ch = 0; ch < VIN; ++ch) in[ch] = __ldg(m.planes + (uint64_t)ch * m.pstride + v);
    float out[4 + NP * 4];
#pragma unroll
    for (int i = 4 + NK; i < 4 + NP * 4; ++i) out[i] = 0.0f;
    sr_vertex_shader<VS>(c, in, out);
    if (c.normalize) sr_normalize_vertex(c.vpm, out);
    pos[v] = make_float4(out[0], out[1], out[2], out[3]);
    if (NP == 2) {
        sr_stg_record2(attr + sr_attr_at(onp, v, 0), make_float4(out[4], out[5], out[6], out[7]), make_float4(out[8], out[9], out[10], out[11]));
    } else {
#pragma unroll
        for (int p = 0; p < NP; ++p)
            attr[sr_attr_at(onp, v, p)] = make_float4(out[4 + 4 * p], out[5 + 4 * p], out[6 + 4 * p], out[7 + 4 * p]);
    }
}
template <int VS>
__global__ void __launch_bounds__(256) k_vertex(const __grid_constant__ SrVsConst c, const SrMeshView m, float4 *pos,
                                                float4 *attr, const uint64_t onp, const SrVertexSpan span) {
    // One vertex per thread: a warp reads 128 contiguous bytes of every SoA input plane and writes 512 contiguous
    // bytes of positions plus 32 contiguous attribute records, so every sector that moves is fully used both ways.
    static_assert(SR_VERTEX_BLOCK == 256, "one CTA per vertex block");
    if (span.blocks != nullptr && !span.blocks[(span.begin >> 8) + blockIdx.x]) return;  // (span.begin is a multiple of 256 then)
    const uint64_t v = span.begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= span.end) return;
    sr_vertex_one<VS>(c, m, pos, attr, onp, v);
}
// The marked vertices of [span.begin, span.end) outside [skip_lo, skip_hi): a warp takes 1024 consecutive vertices, reads their
// 32 mask words with one coalesced load and walks only the non-zero words (32 consecutive vertices each, one per lane) -- a
// sparse mask costs one 128-byte load per 1024 vertices instead of a thread per vertex.  (One warp per 128 vertices was measured
// too: better balanced where the marks are dense, but the scan of an empty mask went from 0.010 to 0.050 ms on config 4's
// 50 M vertices and the dense rank did not gain, 0.127 -> 0.154 ms: kept coarse.)
template <int VS>
__global__ void __launch_bounds__(256) k_vertex_marked(const __grid_constant__ SrVsConst c, const SrMeshView m, float4 *pos,
                                                       float4 *attr, const uint64_t onp, const SrVertexSpan span) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t base = span.begin + (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 1024;  // span.begin is a multiple of 32
    if (base >= span.end) return;
    const uint64_t wi = (base >> 5) + lane;
    uint32_t word = wi * 32 < span.end ? span.mask[wi] : 0u;
    if (__ballot_sync(0xffffffffu, word != 0u) == 0u) return;
#pragma unroll 1
    for (uint32_t j = 0; j < 32; ++j) {
        const uint32_t wj = __shfl_sync(0xffffffffu, word, j);
        if (wj == 0u) continue;
        const uint64_t v = base + j * 32 + lane;
        if (((wj >> lane) & 1u) == 0u || v >= span.end || (v >= span.skip_lo && v < span.skip_hi)) continue;
        if (span.blocks != nullptr && span.blocks[v >> 8]) continue;
        sr_vertex_one<VS>(c, m, pos, attr, onp, v);
    }
}

// SR_VS_PASSTHROUGH (test shader): Vin = {x,y,z,w,k...}, any nk <= SR_MAX_NK; one vertex per thread.
__global__ void __launch_bounds__(128) k_vertex_passthrough(const __grid_constant__ SrVsConst c, const SrMeshView m,
                                                            float4 *pos, float4 *attr, const uint64_t onp) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.nverts) return;
    float p[4];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) p[ch] = m.planes[(uint64_t)ch * m.pstride + i];
    if (c.normalize) sr_normalize_vertex(c.vpm, p);
    pos[i] = make_float4(p[0], p[1], p[2], p[3]);
    const uint32_t nk = m.vin - 4, np = (nk + 3) / 4;
    for (uint32_t pl = 0; pl < np; ++pl) {
        float k[4];
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t ch = 4 + pl * 4 + j;
            k[j] = ch < m.vin ? m.planes[(uint64_t)ch * m.pstride + i] : 0.0f;
        }
        attr[sr_attr_at(onp, i, pl)] = make_float4(k[0], k[1], k[2], k[3]);
    }
}

// a3: GeometryShader::finish (src/pipeline/stages/geometry.rs:60-129): normalize every position in place
__global__ void __launch_bounds__(256) k_normalize(float4 *pos, uint64_t n, const __grid_constant__ SrVsConst c) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = pos[i];
    float p[4] = {v.x, v.y, v.z, v.w};
    sr_normalize_vertex(c.vpm, p);
    pos[i] = make_float4(p[0], p[1], p[2], p[3]);
}

// smallest and one-past-largest vertex index a run of indices references (out[0] = min, out[1] = max + 1; out preset to ~0, 0)
__global__ void __launch_bounds__(256) k_index_minmax(const uint32_t *idx, uint64_t n, uint32_t *out) {
    uint32_t lo = 0xFFFFFFFFu, hi = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t v = idx[i];
        lo = min(lo, v);
        hi = max(hi, v + 1u);
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out, lo);
        atomicMax(out + 1, hi);
    }
}

// ---- chunk-culled front end of a sharded frame ------------------------------------------------------------
// Static per mesh: the object-space bounding box of every block of 256 consecutive vertices, and the vertex range every chunk of
// 1024 consecutive triangles references.  Per frame: the conservative range of screen TILE ROWS each block can reach under the draw's
// transform, and from that which chunks can touch the rows of a rank.
__global__ void __launch_bounds__(SR_VERTEX_BLOCK) k_block_aabb(const float *planes, uint64_t pstride, uint64_t nverts, float *aabb /* [nblk][6] */) {
    const uint64_t v = (uint64_t)blockIdx.x * SR_VERTEX_BLOCK + threadIdx.x;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (v < nverts) {
#pragma unroll
        for (int k = 0; k < 3; ++k) lo[k] = hi[k] = planes[(uint64_t)k * pstride + v];
    }
    __shared__ float s[SR_VERTEX_BLOCK / 32][6];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { s[threadIdx.x >> 5][k] = lo[k]; s[threadIdx.x >> 5][3 + k] = hi[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float r = s[0][threadIdx.x];
        for (int wgt = 1; wgt < SR_VERTEX_BLOCK / 32; ++wgt) r = threadIdx.x < 3 ? fminf(r, s[wgt][threadIdx.x]) : fmaxf(r, s[wgt][threadIdx.x]);
        aabb[(uint64_t)blockIdx.x * 6 + threadIdx.x] = r;  // a NaN coordinate is dropped by fminf/fmaxf: such triangles draw nothing anyway
    }
}
// vertex range of each chunk of `chunk_tris` consecutive triangles: out[c] = {min index, max index + 1}
__global__ void __launch_bounds__(256) k_chunk_vrange(const uint32_t *idx, uint32_t ntris, uint32_t chunk_tris, uint2 *out) {
    const uint32_t c = blockIdx.x;
    const uint64_t i0 = (uint64_t)c * chunk_tris * 3, i1 = min((uint64_t)ntris * 3, i0 + (uint64_t)chunk_tris * 3);
    uint32_t lo = 0xFFFFFFFFu, hi = 0;
    for (uint64_t i = i0 + threadIdx.x; i < i1; i += 256) {
        const uint32_t v = idx[i];
        lo = min(lo, v);
        hi = max(hi, v + 1u);
    }
    __shared__ uint32_t slo[8], shi[8];
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int wgt = 1; wgt < 8; ++wgt) { lo = min(lo, slo[wgt]); hi = max(hi, shi[wgt]); }
        out[c] = make_uint2(lo, hi);
    }
}
// Conservative tile-row range of every vertex block under clip = pvm * (x, y, z, 1), screen y = vp_sy * (clip.y / clip.w) + vp_ty
// (the viewport matrix of ClipVertex::normalize).  A projective map with w > 0 on all eight corners maps the box into the convex hull
// of the corners' images, so the corners' screen y bound every vertex's; the float evaluation of the real vertex shader differs from
// this one by rounding only, which the margin of two pixels (plus 1e-4 relative) covers by orders of magnitude.  A corner with
// w <= 0 (or a non-finite value) makes the block "every row".  rows[b] = lo | hi << 16.
__global__ void __launch_bounds__(256) k_block_rows(const float *aabb, uint32_t nblk, const __grid_constant__ SrVsConst c, uint32_t height, uint32_t nty,
                                                    uint32_t tile_h, uint32_t *rows) {
    const uint32_t b = blockIdx.x * 256 + threadIdx.x;
    if (b >= nblk) return;
    const float *bb = aabb + (uint64_t)b * 6;
    float ymin = INFINITY, ymax = -INFINITY;
    bool bounded = true;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float x = bb[(k & 1) ? 3 : 0], y = bb[(k & 2) ? 4 : 1], z = bb[(k & 4) ? 5 : 2];
        const float cy = c.pvm[0 * 4 + 1] * x + c.pvm[1 * 4 + 1] * y + c.pvm[2 * 4 + 1] * z + c.pvm[3 * 4 + 1];
        const float cw = c.pvm[0 * 4 + 3] * x + c.pvm[1 * 4 + 3] * y + c.pvm[2 * 4 + 3] * z + c.pvm[3 * 4 + 3];
        if (!(cw > 1e-6f) || !isfinite(cy)) { bounded = false; continue; }
        const float sy = c.vpm[1 * 4 + 1] * (cy / cw) + c.vpm[3 * 4 + 1];
        ymin = fminf(ymin, sy);
        ymax = fmaxf(ymax, sy);
    }
    uint32_t lo = 0, hi = nty - 1;
    if (bounded && isfinite(ymin) && isfinite(ymax)) {
        const float m = 2.0f + 1e-4f * fmaxf(fabsf(ymin), fabsf(ymax));
        const float a = fminf(fmaxf(ymin - m, 0.0f), (float)(height - 1)), z = fminf(fmaxf(ymax + m, 0.0f), (float)(height - 1));
        lo = (uint32_t)a / tile_h;
        hi = (uint32_t)z / tile_h;  // (triangle.rs:66-78 clamps a bounding box into the frame the same way: off-screen boxes land on the border rows)
    }
    rows[b] = lo | (hi << 16);
}
// flag[c] = 1 iff chunk c can reach a tile row of `rank` (row y belongs to rank y % world); the vertex blocks of flagged chunks are
// flagged in `blocks` (the rank shades exactly those before its k_micro)
__global__ void __launch_bounds__(256) k_chunk_select(const uint2 *vrange, const uint32_t *rows, uint32_t nchunk, uint32_t rank, uint32_t world,
                                                      uint32_t *flag, uint8_t *blocks) {
    const uint32_t c = blockIdx.x * 256 + threadIdx.x;
    if (c >= nchunk) return;
    const uint2 vr = vrange[c];
    uint32_t lo = 0xFFFFu, hi = 0, sel = 0;
    if (vr.y > vr.x) {
        const uint32_t b0 = vr.x >> 8, b1 = (vr.y - 1u) >> 8;
        for (uint32_t b = b0; b <= b1; ++b) {
            const uint32_t r = rows[b];
            lo = min(lo, r & 0xFFFFu);
            hi = max(hi, r >> 16);
        }
        if (hi - lo + 1u >= world) sel = 1;
        else
            for (uint32_t y = lo; y <= hi; ++y) sel |= (y % world == rank) ? 1u : 0u;
        if (sel)
            for (uint32_t b = b0; b <= b1; ++b) blocks[b] = 1;
    }
    flag[c] = sel;
}
// ordered compaction of the flagged chunks (pos = exclusive scan of flag); the count lands in *count
__global__ void __launch_bounds__(256) k_chunk_compact(const uint32_t *flag, const uint32_t *pos, uint32_t nchunk, uint32_t *list, uint32_t *count) {
    const uint32_t c = blockIdx.x * 256 + threadIdx.x;
    if (c >= nchunk) return;
    if (flag[c]) list[pos[c]] = c;
    if (c == nchunk - 1) *count = pos[c] + flag[c];
}

// AoS -> SoA plane transpose used by mesh upload and vertex injection
__global__ void __launch_bounds__(256) k_aos_to_planes(const float *aos, uint64_t n, uint32_t nfloats, float *planes,
                                                       uint64_t pstride) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * nfloats) return;
    const uint64_t v = i / nfloats;
    const uint32_t ch = (uint32_t)(i % nfloats);
    planes[(uint64_t)ch * pstride + v] = aos[i];
}
// records {pos4, k[nk]} (AoS) -> position array + attribute records
__global__ void __launch_bounds__(256) k_records_to_planes(const float *rec, uint64_t n, uint32_t nk, float4 *pos, float4 *attr,
                                                           uint64_t onp) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *r = rec + i * (4 + nk);
    pos[i] = make_float4(r[0], r[1], r[2], r[3]);
    const uint32_t np = (nk + 3) / 4;
    for (uint32_t p = 0; p < np; ++p) {
        float k[4];
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) k[j] = (p * 4 + j) < nk ? r[4 + p * 4 + j] : 0.0f;
        attr[sr_attr_at(onp, i, p)] = make_float4(k[0], k[1], k[2], k[3]);
    }
}
__global__ void __launch_bounds__(256) k_planes_to_records(const float4 *pos, const float4 *attr, uint64_t anp, uint64_t n,
                                                           uint32_t nk, float *rec) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float *r = rec + i * (4 + nk);
    const float4 p = pos[i];
    r[0] = p.x; r[1] = p.y; r[2] = p.z; r[3] = p.w;
    const uint32_t np = (nk + 3) / 4;
    for (uint32_t pl = 0; pl < np; ++pl) {
        const float4 a = attr[sr_attr_at(anp, i, pl)];
        const float k[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j)
            if (pl * 4 + j < nk) r[4 + pl * 4 + j] = k[j];
    }
}
// largest index of an index buffer (range check on upload; the reference would panic on a bad index)
__global__ void __launch_bounds__(256) k_index_max(const uint32_t *idx, uint64_t n, uint32_t *out) {
    uint32_t m = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) m = max(m, __ldg(idx + i));
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}
// usize (u64) indices -> u32
__global__ void __launch_bounds__(256) k_narrow_indices(const uint64_t *in, uint64_t n, uint32_t *out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)in[i];
}

// =====================================================================================================
// a4: geometry stage (GeometryShader::run / clip_primitives, src/pipeline/stages/geometry.rs:132-336)
// Input primitives of one kind in the reference's visit order: generated first, then the indexed mesh.
// =====================================================================================================
struct SrGeoIn {
    SrVertexSet gen;
    uint32_t ngen;  // primitives
    SrVertexSet idx;
    const uint32_t *indices;
    uint32_t nidx;  // primitives
    uint32_t nplanes;
};
struct SrGeoOut {
    float4 *pos;
    float4 *attr;
    uint64_t np;
};

template <int NV>
__device__ __forceinline__ void sr_geo_load(const SrGeoIn &in, uint32_t prim, float rec[NV][4 + SR_MAX_NK]) {
    const SrVertexSet &vs = prim < in.ngen ? in.gen : in.idx;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const uint32_t vi = prim < in.ngen ? prim * NV + k : __ldg(in.indices + (uint64_t)(prim - in.ngen) * NV + k);
        const float4 p = vs.pos[vi];
        rec[k][0] = p.x; rec[k][1] = p.y; rec[k][2] = p.z; rec[k][3] = p.w;
        for (uint32_t pl = 0; pl < in.nplanes; ++pl) {
            const float4 a = vs.attr[sr_attr_at(vs.np, vi, pl)];
            rec[k][4 + pl * 4] = a.x; rec[k][5 + pl * 4] = a.y; rec[k][6 + pl * 4] = a.z; rec[k][7 + pl * 4] = a.w;
        }
    }
}
__device__ __forceinline__ void sr_geo_store(const SrGeoOut &out, uint64_t at, const float *rec, uint32_t nplanes) {
    out.pos[at] = make_float4(rec[0], rec[1], rec[2], rec[3]);
    for (uint32_t pl = 0; pl < nplanes; ++pl)
        out.attr[sr_attr_at(out.np, at, pl)] = make_float4(rec[4 + pl * 4], rec[5 + pl * 4], rec[6 + pl * 4], rec[7 + pl * 4]);
}

// ClippingPlane::has_inside / intersect (src/geometry/clip.rs:33-63)
__device__ __forceinline__ bool sr_has_inside(int plane, const float *v) {
    const float x = v[0], y = v[1], z = v[2], w = v[3];
    switch (plane) {
        case 0: return x >= -w;
        case 1: return x <= w;
        case 2: return y >= -w;
        case 3: return y <= w;
        case 4: return z >= 0.0f;
        default: return z <= w;
    }
}
__device__ __forceinline__ void sr_intersect(int plane, const float *v1, const float *v2, uint32_t nfloats, float *out) {
    float a, b;
    switch (plane) {
        case 0: a = v1[3] + v1[0]; b = v2[3] + v2[0]; break;
        case 1: a = v1[3] - v1[0]; b = v2[3] - v2[0]; break;
        case 2: a = v1[3] + v1[1]; b = v2[3] + v2[1]; break;
        case 3: a = v1[3] - v1[1]; b = v2[3] - v2[1]; break;
        case 4: a = v1[2]; b = v2[2]; break;
        default: a = v1[3] - v1[2]; b = v2[3] - v2[2]; break;
    }
    const float t = a / (a - b);
    for (uint32_t i = 0; i < nfloats; ++i) out[i] = sr_lerp(t, v1[i], v2[i]);
}

// The triangle clipper, literally (geometry.rs:265-298), but symbolic: polygon entries are ids
// 0..2 = the input vertices a,b,c and 3 + edge*6 + plane = intersect(plane, edge.s, edge.p).
// Two entries with the same id are bit-identical vertices.
__device__ __forceinline__ int sr_clip_polygon(const float rec[3][4 + SR_MAX_NK], uint8_t *poly) {
    int n = 0;
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        const int s = e, p = e == 2 ? 0 : e + 1;
#pragma unroll
        for (int plane = 0; plane < 6; ++plane) {
            const bool s_in = sr_has_inside(plane, rec[s]);
            const bool p_in = sr_has_inside(plane, rec[p]);
            if (s_in != p_in) poly[n++] = (uint8_t)(3 + e * 6 + plane);
            if (p_in) poly[n++] = (uint8_t)p;
        }
    }
    return n;
}
__device__ __forceinline__ int sr_clip_tri_count(int n) { return n == 3 ? 1 : (n > 3 ? n - 2 : 0); }
__device__ __forceinline__ void sr_clip_tri_ids(const uint8_t *poly, int n, int i, uint8_t ids[3]) {
    if (n == 3) { ids[0] = poly[0]; ids[1] = poly[1]; ids[2] = poly[2]; }
    else { ids[0] = poly[n - 1]; ids[1] = poly[i]; ids[2] = poly[i + 1]; }
}
// A triangle with two bit-identical vertices has det == +-0 exactly and can never produce a
// fragment (triangle.rs:64,108-113,120); with stencil op Keep it has no observable effect at all.
__device__ __forceinline__ bool sr_clip_tri_degenerate(const uint8_t ids[3]) {
    return ids[0] == ids[1] || ids[0] == ids[2] || ids[1] == ids[2];
}

// pass 1: per input triangle, number of kept and of literal output triangles
__global__ void __launch_bounds__(128) k_clip_tri_count(const SrGeoIn in, uint32_t drop_degenerate, uint32_t *kept,
                                                        uint32_t *literal) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= in.ngen + in.nidx) return;
    float rec[3][4 + SR_MAX_NK];
    sr_geo_load<3>(in, t, rec);
    uint8_t poly[36];
    const int n = sr_clip_polygon(rec, poly);
    const int nt = sr_clip_tri_count(n);
    int k = 0;
    for (int i = 0; i < nt; ++i) {
        uint8_t ids[3];
        sr_clip_tri_ids(poly, n, i, ids);
        if (!(drop_degenerate && sr_clip_tri_degenerate(ids))) ++k;
    }
    kept[t] = (uint32_t)k;
    literal[t] = (uint32_t)nt;
}
// pass 2: materialise the kept triangles at their scanned offsets (order-preserving compaction)
__global__ void __launch_bounds__(128) k_clip_tri_emit(const SrGeoIn in, uint32_t drop_degenerate, const uint32_t *kept_off,
                                                       const uint32_t *literal_off, SrGeoOut out, uint32_t *seq) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= in.ngen + in.nidx) return;
    float rec[3][4 + SR_MAX_NK];
    sr_geo_load<3>(in, t, rec);
    uint8_t poly[36];
    const int n = sr_clip_polygon(rec, poly);
    const int nt = sr_clip_tri_count(n);
    const uint32_t nfloats = 4 + in.nplanes * 4;
    uint32_t o = kept_off[t];
    const uint32_t lbase = literal_off[t];
    float tmp[4 + SR_MAX_NK];
    for (int i = 0; i < nt; ++i) {
        uint8_t ids[3];
        sr_clip_tri_ids(poly, n, i, ids);
        if (drop_degenerate && sr_clip_tri_degenerate(ids)) continue;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int id = ids[k];
            if (id < 3) {
                sr_geo_store(out, (uint64_t)o * 3 + k, rec[id], in.nplanes);
            } else {
                const int e = (id - 3) / 6, plane = (id - 3) % 6;
                sr_intersect(plane, rec[e], rec[e == 2 ? 0 : e + 1], nfloats, tmp);
                sr_geo_store(out, (uint64_t)o * 3 + k, tmp, in.nplanes);
            }
        }
        seq[o] = lbase + (uint32_t)i;
        ++o;
    }
}

// Both passes and the scans between them for at most 1024 input triangles in ONE single-CTA launch (a model of a thousand
// triangles: the chain of tiny dependent launches, not the work, is what such a frame costs).  totals = {kept, literal}.
#define SR_CLIP_SMALL_MAX 1024
__global__ void __launch_bounds__(SR_CLIP_SMALL_MAX) k_clip_tri_small(const SrGeoIn in, uint32_t drop_degenerate, SrGeoOut out, uint32_t *seq,
                                                                      uint32_t *totals) {
    __shared__ uint32_t ws[32];
    const uint32_t t = threadIdx.x;
    const bool have = t < in.ngen + in.nidx;
    float rec[3][4 + SR_MAX_NK];
    uint8_t poly[36];
    int n = 0, nt = 0, k = 0;
    if (have) {
        sr_geo_load<3>(in, t, rec);
        n = sr_clip_polygon(rec, poly);
        nt = sr_clip_tri_count(n);
        for (int i = 0; i < nt; ++i) {
            uint8_t ids[3];
            sr_clip_tri_ids(poly, n, i, ids);
            if (!(drop_degenerate && sr_clip_tri_degenerate(ids))) ++k;
        }
    }
    uint32_t total_kept, total_literal;
    uint32_t o = sr_block_exclusive_scan((uint32_t)k, &total_kept, ws);
    const uint32_t lbase = sr_block_exclusive_scan((uint32_t)nt, &total_literal, ws);
    if (t == 0) { totals[0] = total_kept; totals[1] = total_literal; }
    if (!have) return;
    const uint32_t nfloats = 4 + in.nplanes * 4;
    float tmp[4 + SR_MAX_NK];
    for (int i = 0; i < nt; ++i) {
        uint8_t ids[3];
        sr_clip_tri_ids(poly, n, i, ids);
        if (drop_degenerate && sr_clip_tri_degenerate(ids)) continue;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int id = ids[c];
            if (id < 3) {
                sr_geo_store(out, (uint64_t)o * 3 + c, rec[id], in.nplanes);
            } else {
                const int e = (id - 3) / 6, plane = (id - 3) % 6;
                sr_intersect(plane, rec[e], rec[e == 2 ? 0 : e + 1], nfloats, tmp);
                sr_geo_store(out, (uint64_t)o * 3 + c, tmp, in.nplanes);
            }
        }
        seq[o] = lbase + (uint32_t)i;
        ++o;
    }
}

// SR_GS_CLIP_SH: Sutherland-Hodgman against the same six planes, one plane after the other; an edge runs from the
// previous vertex s to the current vertex p and crossings are intersect(plane, s, p) (clip.rs:47-63).  A triangle
// clipped by six planes has at most nine vertices; the result is fanned around its first vertex.  MODE 0 = count, 1 = emit.
#define SR_SH_MAX 9
template <int MODE>
__global__ void __launch_bounds__(128) k_clip_tri_sh(const SrGeoIn in, uint32_t *count, const uint32_t *off, SrGeoOut out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= in.ngen + in.nidx) return;
    float rec[3][4 + SR_MAX_NK];
    sr_geo_load<3>(in, t, rec);
    const uint32_t nfloats = 4 + in.nplanes * 4;
    float poly[2][SR_SH_MAX][4 + SR_MAX_NK];
    int n = 3, cur = 0;
    for (int k = 0; k < 3; ++k)
        for (uint32_t i = 0; i < nfloats; ++i) poly[0][k][i] = rec[k][i];
    for (int plane = 0; plane < 6 && n > 0; ++plane) {
        int m = 0;
        float(*src)[4 + SR_MAX_NK] = poly[cur], (*dst)[4 + SR_MAX_NK] = poly[cur ^ 1];
        for (int i = 0; i < n; ++i) {
            const float *s = src[(i + n - 1) % n], *p = src[i];
            const bool s_in = sr_has_inside(plane, s), p_in = sr_has_inside(plane, p);
            if (p_in) {
                if (!s_in && m < SR_SH_MAX) sr_intersect(plane, s, p, nfloats, dst[m++]);
                if (m < SR_SH_MAX) { for (uint32_t j = 0; j < nfloats; ++j) dst[m][j] = p[j]; ++m; }
            } else if (s_in && m < SR_SH_MAX) {
                sr_intersect(plane, s, p, nfloats, dst[m++]);
            }
        }
        n = m;
        cur ^= 1;
    }
    const int nt = n >= 3 ? n - 2 : 0;
    if (MODE == 0) {
        count[t] = (uint32_t)nt;
    } else {
        uint64_t o = (uint64_t)off[t] * 3;
        for (int i = 1; i + 1 < n; ++i) {
            sr_geo_store(out, o, poly[cur][0], in.nplanes);
            sr_geo_store(out, o + 1, poly[cur][i], in.nplanes);
            sr_geo_store(out, o + 2, poly[cur][i + 1], in.nplanes);
            o += 3;
        }
    }
}

// line clipper (geometry.rs:300-327): MODE 0 = count, 1 = emit
template <int MODE>
__global__ void __launch_bounds__(128) k_clip_line(const SrGeoIn in, uint32_t *count, const uint32_t *off, SrGeoOut out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= in.ngen + in.nidx) return;
    float rec[2][4 + SR_MAX_NK];
    sr_geo_load<2>(in, t, rec);
    const uint32_t nfloats = 4 + in.nplanes * 4;
    float tmp[4 + SR_MAX_NK];
    int intersections = 0;
    bool emit = true;
    for (int plane = 0; plane < 6; ++plane) {
        const bool s_in = sr_has_inside(plane, rec[0]);
        const bool p_in = sr_has_inside(plane, rec[1]);
        if (s_in != p_in) {
            sr_intersect(plane, rec[0], rec[1], nfloats, tmp);
            float *dst = s_in ? rec[1] : rec[0];  // `if s_in { end = .. } else if p_in { start = .. }`
            for (uint32_t i = 0; i < nfloats; ++i) dst[i] = tmp[i];
            intersections += 1;
        } else if (!s_in) {
            emit = false;
            break;
        }
        if (intersections > 2) break;
    }
    if (MODE == 0) {
        count[t] = emit ? 1u : 0u;
    } else if (emit) {
        const uint32_t o = off[t];
        sr_geo_store(out, (uint64_t)o * 2, rec[0], in.nplanes);
        sr_geo_store(out, (uint64_t)o * 2 + 1, rec[1], in.nplanes);
    }
}
// point clipper (geometry.rs:329-333)
template <int MODE>
__global__ void __launch_bounds__(128) k_clip_point(const SrGeoIn in, uint32_t *count, const uint32_t *off, SrGeoOut out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= in.ngen + in.nidx) return;
    float rec[1][4 + SR_MAX_NK];
    sr_geo_load<1>(in, t, rec);
    bool inside = true;
#pragma unroll
    for (int plane = 0; plane < 6; ++plane) inside = inside && sr_has_inside(plane, rec[0]);
    if (MODE == 0) count[t] = inside ? 1u : 0u;
    else if (inside) sr_geo_store(out, off[t], rec[0], in.nplanes);
}

// `_ => storage.re_emit(primitive)`: copy a primitive stream through unchanged
template <int NV>
__global__ void __launch_bounds__(128) k_geo_reemit(const SrGeoIn in, SrGeoOut out, uint64_t out_base) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= in.ngen + in.nidx) return;
    float rec[NV][4 + SR_MAX_NK];
    sr_geo_load<NV>(in, t, rec);
#pragma unroll
    for (int k = 0; k < NV; ++k) sr_geo_store(out, out_base + (uint64_t)t * NV + k, rec[k], in.nplanes);
}

// geometry_shader_visualize_{face,vertex}_normals (full_example/src/shaders.rs:35-89): triangle -> lines
#define SR_NORMAL_LENGTH 0.05f
template <int GS>
__global__ void __launch_bounds__(128) k_geo_normals(const SrGeoIn in, const __grid_constant__ SrVsConst c, SrGeoOut out,
                                                     uint64_t out_base) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= in.ngen + in.nidx) return;
    float rec[3][4 + SR_MAX_NK];
    sr_geo_load<3>(in, t, rec);
    const uint32_t nk = in.nplanes * 4;
    float o[4 + SR_MAX_NK];
    if (GS == SR_GS_FACE_NORMALS) {
        const float third = 1.0f / 3.0f;
        float center[SR_MAX_NK];
        for (uint32_t i = 0; i < nk; ++i) center[i] = sr_bary(third, rec[0][4 + i], third, rec[1][4 + i], third, rec[2][4 + i]);
        float nn[4], tip[4];
        sr_normalize4(center + 4, nn);
#pragma unroll
        for (int i = 0; i < 4; ++i) tip[i] = center[i] + nn[i] * SR_NORMAL_LENGTH;
        for (uint32_t i = 0; i < nk; ++i) o[4 + i] = center[i];
        sr_mat_vec(c.pv, center, o);
        sr_geo_store(out, out_base + (uint64_t)t * 2, o, in.nplanes);
        sr_mat_vec(c.pv, tip, o);
        sr_geo_store(out, out_base + (uint64_t)t * 2 + 1, o, in.nplanes);
    } else {
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            float tip[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) tip[i] = rec[v][4 + i] + rec[v][8 + i] * SR_NORMAL_LENGTH;
            for (uint32_t i = 0; i < nk; ++i) o[4 + i] = rec[v][4 + i];
            sr_mat_vec(c.pv, rec[v] + 4, o);
            sr_geo_store(out, out_base + (uint64_t)t * 6 + v * 2, o, in.nplanes);
            sr_mat_vec(c.pv, tip, o);
            sr_geo_store(out, out_base + (uint64_t)t * 6 + v * 2 + 1, o, in.nplanes);
        }
    }
}
