// ball_query.cu -- fixed-radius neighbour query of the set-abstraction path.
//
// Semantics (ball_query_gpu.cu:15-51): for each centre scan the points in ascending
// index k, keep the first `nsample` with d2 < radius^2 (d2 rounded as sqdist_ref with
// dx = new_x - x, radius^2 = rn(radius*radius) in fp32), pad the row with the first hit, and
// leave a row with no hit untouched.  Equivalently: the row holds the `nsample` SMALLEST
// indices of the hit set in ascending order -- which does not depend on scan order.
//
// Grid kernels (default)
// ----------------------
// The reference tests every centre against every point (M*N distance evaluations per frame).
// Here the points of a frame are first binned into a uniform grid whose cells are at least
// 1.01 * radius wide (bq_build_kernel: frame box, histogram, exclusive scan, scatter of
// (x,y,z,k) records -- one thread-block CLUSTER of up to 8 CTAs per frame, the phases separated by
// cluster barriers, so a batch of 16 frames builds on 128 SMs in one launch).  A centre then only
// meets the 3x3x3 cells around its own: fp32 rounding of the cell coordinate is ~1e-4 of a cell, far
// inside the 1% margin, so no point with d2 < r^2 can lie outside that neighbourhood.
// bq_query_topk_kernel gives one warp to each centre: the lanes stride over the candidate records
// (coalesced 16-byte loads), evaluate the exact reference distance, and keep the nsample SMALLEST hit
// indices in a sorted list held one entry per lane (nsample <= 32) or two (<= 64): a hit below the
// list's current threshold is inserted with ballot + popc + shfl_up -- no shared memory, no
// dependence on the frame size, and the list is already in the order the row wants.
// nsample > 64 keeps the older bitmap kernel (bit k of a per-warp shared-memory bitmap per hit,
// read back in index order).
//
// Tiled kernel (fallback, PDM_BQ_KERNEL=tiled): one thread per centre like the reference,
// points streamed through shared memory; identical results by construction.
#include <stdlib.h>

#include <map>
#include <mutex>

#include "common.cuh"

namespace pdm {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kBQTile = 1024;

__global__ void __launch_bounds__(128)
ball_query_tiled_kernel(int n, int m, float radius2, int nsample,
                        const float *__restrict__ new_xyz, const float *__restrict__ xyz,
                        int *__restrict__ idx) {
    __shared__ float tile[kBQTile * 3];
    const int bi = blockIdx.y;
    const int pi = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = pi < m;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (live) {
        const float *q = new_xyz + ((size_t)bi * m + pi) * 3;
        qx = __ldg(q); qy = __ldg(q + 1); qz = __ldg(q + 2);
    }
    const float *pts = xyz + (size_t)bi * n * 3;
    int *row = idx + ((size_t)bi * m + (live ? pi : 0)) * nsample;
    int cnt = live ? 0 : nsample;  // dead lanes count as finished
    for (int base = 0; base < n; base += kBQTile) {
        const int tcnt = min(kBQTile, n - base);
        if (__syncthreads_and(cnt >= nsample)) break;  // also orders tile reuse
        for (int t = threadIdx.x; t < tcnt * 3; t += blockDim.x) tile[t] = __ldg(pts + (size_t)base * 3 + t);
        __syncthreads();
        if (cnt < nsample) {
            for (int k = 0; k < tcnt; ++k) {
                const float d2 = sqdist_ref(__fsub_rn(qx, tile[k * 3 + 0]), __fsub_rn(qy, tile[k * 3 + 1]),
                                            __fsub_rn(qz, tile[k * 3 + 2]));
                if (d2 < radius2) {
                    const int kk = base + k;
                    if (cnt == 0)
                        for (int l = 1; l < nsample; ++l) row[l] = kk;
                    row[cnt] = kk;
                    if (++cnt >= nsample) break;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// grid build: one CTA per frame
// ---------------------------------------------------------------------------------------
struct BQGrid {        // per frame, written by the build kernel, read by the query kernel
    float ox, oy, oz;  // origin (frame minimum over finite coordinates)
    float ix, iy, iz;  // 1 / cell size per axis
    int gx, gy, gz;    // cells per axis (z fastest in the linear index)
    int ncell;
};

__device__ __forceinline__ int bq_cell_coord(float v, float o, float inv, int g) {
    const float f = __fmul_rn(__fsub_rn(v, o), inv);
    int c = (int)f;  // NaN -> 0, +-inf saturate
    return max(0, min(g - 1, c));
}

constexpr int kBuildThreads = 1024;
constexpr int kBuildMaxCl = 8;      // portable cluster size
constexpr int kBuildAux = 16;       // ints of per-frame build state next to the cell counters (zero-filled by the caller)

// order-preserving float <-> uint map (never 0 for a finite float: 0 means "no value" in the atomics below)
__device__ __forceinline__ unsigned bq_f2ord(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float bq_ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
// all threads of all CTAs of the cluster; orders global-memory traffic (atomics, stores) before it against loads after it
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Stacked (ragged) frames, pointnet2_stack/src/ball_query_gpu.cu:15-70: frame f owns rows [sum xyz_cnt[:f], +xyz_cnt[f])
// of xyz and the centres [sum new_cnt[:f], +new_cnt[f]); indices are LOCAL to the frame; a centre without a hit gets
// idx[0] = -1 (:69).  xyz_cnt == nullptr: uniform (B,N,3) / (B,M,3) frames of the batch API.
struct BQRagged {
    const int *xyz_cnt = nullptr, *new_cnt = nullptr;
    int nframes = 0;
};

// One cluster of `cl` CTAs (1..8) per frame; CTA `rank` owns a contiguous slice of the frame's points and of its cells.
// The caller zero-fills the cell counters and the per-slice totals (one memset).  Phases: frame box (every CTA reads
// the whole frame: cheaper than an exchange + two more cluster barriers) -> histogram of my points (global atomics;
// totals per cell slice through shared-memory counters) | cluster barrier | exclusive scan of my cell slice | cluster
// barrier | scatter of my points' (x,y,z,k) records.  No distributed shared memory: the cluster is there for the barriers.
__global__ void __launch_bounds__(kBuildThreads)
bq_build_kernel(int n, float radius, int cmax, const float *__restrict__ xyz, BQGrid *__restrict__ grids,
                int *__restrict__ cellend, int *__restrict__ parts, float4 *__restrict__ sorted, BQRagged rg) {
    __shared__ float red[6][kBuildThreads / 32];
    __shared__ int spart[kBuildMaxCl];
    __shared__ BQGrid sg;
    __shared__ int wsum[kBuildThreads / 32];
    __shared__ int carry, tile_total;
    const int cl = (int)cluster_nctarank(), rank = (int)cluster_ctarank();
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int bi = blockIdx.x / cl;
    size_t pstart = (size_t)bi * n;
    if (rg.xyz_cnt) {
        int ps = 0;
        for (int k = 0; k < bi; ++k) ps += __ldg(rg.xyz_cnt + k);
        pstart = (size_t)ps;
        n = __ldg(rg.xyz_cnt + bi);
    }
    const float *pts = xyz + pstart * 3;
    int *cend = cellend + (size_t)bi * cmax;
    int *part = parts + (size_t)bi * kBuildAux;                       // [0, 8): slice totals
    unsigned *fbox = reinterpret_cast<unsigned *>(part + kBuildMaxCl);    // [8, 14): frame box
    float4 *srt = sorted + pstart;
    const int per = (n + cl - 1) / cl;
    const int s0 = min(n, rank * per), s1 = min(n, s0 + per);     // my points

    // 1. frame box over finite coordinates: my slice, merged through six global atomics (order-preserving uint images;
    //    the minima are kept as maxima of the complement, so that the caller's zero fill means "nothing yet")
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = s0 + tid; k < s1; k += kBuildThreads) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = __ldg(pts + (size_t)k * 3 + a);
            if (fabsf(v) < INFINITY) {  // false for NaN and +-inf
                lo[a] = fminf(lo[a], v);
                hi[a] = fmaxf(hi[a], v);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(kFullMask, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(kFullMask, hi[a], o));
        }
        if (lane == 0) {
            red[a][w] = lo[a];
            red[3 + a][w] = hi[a];
        }
    }
    __syncthreads();
    if (tid < 6) {
        float v = tid < 3 ? INFINITY : -INFINITY;
        for (int i = 0; i < kBuildThreads / 32; ++i) v = tid < 3 ? fminf(v, red[tid][i]) : fmaxf(v, red[tid][i]);
        if (fabsf(v) < INFINITY) atomicMax(fbox + tid, tid < 3 ? ~bq_f2ord(v) : bq_f2ord(v));
    }
    cluster_sync_all();
    if (tid < kBuildMaxCl) spart[tid] = 0;
    if (tid == 0) {
        float l[3], h[3];
        for (int a = 0; a < 3; ++a) {
            const unsigned wl = __ldcg(fbox + a), wh = __ldcg(fbox + 3 + a);
            l[a] = wl ? bq_ord2f(~wl) : 0.f;     // no finite coordinate on this axis: 0
            h[a] = wh ? bq_ord2f(wh) : 0.f;
        }
        // cells at least 1.01 r wide; fewer, wider cells when the frame would need > cmax of them
        const float hmin = radius * 1.01f;
        int g[3];
        float ext[3];
        for (int a = 0; a < 3; ++a) {
            ext[a] = h[a] - l[a];
            float cells = (hmin > 0.f && hmin < INFINITY) ? floorf(ext[a] / hmin) + 1.f : 1.f;
            if (!(cells >= 1.f)) cells = 1.f;
            g[a] = (int)fminf(cells, 1024.f);
        }
        while ((long long)g[0] * g[1] * g[2] > (long long)cmax) {
            int a = g[0] >= g[1] ? (g[0] >= g[2] ? 0 : 2) : (g[1] >= g[2] ? 1 : 2);
            g[a] = (g[a] + 1) / 2;
        }
        float inv[3];
        for (int a = 0; a < 3; ++a) {
            float cs = fmaxf(hmin, ext[a] / (float)g[a] * 1.0001f);
            inv[a] = (cs > 0.f && cs < INFINITY) ? 1.0f / cs : 0.f;
        }
        sg.ox = l[0]; sg.oy = l[1]; sg.oz = l[2];
        sg.ix = inv[0]; sg.iy = inv[1]; sg.iz = inv[2];
        sg.gx = g[0]; sg.gy = g[1]; sg.gz = g[2];
        sg.ncell = g[0] * g[1] * g[2];
        if (rank == 0) grids[bi] = sg;
    }
    __syncthreads();
    const BQGrid G = sg;
    const int cpr = ((G.ncell + cl - 1) / cl + 3) / 4 * 4;       // cells per slice
    const int c_lo = min(G.ncell, rank * cpr), c_hi = min(G.ncell, c_lo + cpr);

    // 2. histogram of my points (the counters live in cellend); totals per cell slice
    for (int k = s0 + tid; k < s1; k += kBuildThreads) {
        const int cx = bq_cell_coord(__ldg(pts + (size_t)k * 3 + 0), G.ox, G.ix, G.gx);
        const int cy = bq_cell_coord(__ldg(pts + (size_t)k * 3 + 1), G.oy, G.iy, G.gy);
        const int cz = bq_cell_coord(__ldg(pts + (size_t)k * 3 + 2), G.oz, G.iz, G.gz);
        const int c = (cx * G.gy + cy) * G.gz + cz;
        atomicAdd(&cend[c], 1);
        atomicAdd(&spart[c / cpr], 1);
    }
    __syncthreads();
    if (tid < cl && spart[tid]) atomicAdd(&part[tid], spart[tid]);
    cluster_sync_all();
    // 3. exclusive scan of my cell slice, in place
    if (tid == 0) {
        int base = 0;
        for (int i = 0; i < rank; ++i) base += __ldcg(part + i);
        carry = base;
    }
    __syncthreads();
    for (int base = c_lo; base < c_hi; base += 4 * kBuildThreads) {
        const int c0 = base + tid * 4;
        int v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = (c0 + q < c_hi) ? __ldcg(cend + c0 + q) : 0;
        const int tsum = v[0] + v[1] + v[2] + v[3];
        int incl = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        if (w == 0) {
            int sv = wsum[lane];
            int si = sv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(kFullMask, si, o);
                if (lane >= o) si += y;
            }
            wsum[lane] = si - sv;  // exclusive warp offsets
            if (lane == 31) tile_total = si;
        }
        __syncthreads();
        int run = carry + wsum[w] + incl - tsum;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (c0 + q < c_hi) cend[c0 + q] = run;
            run += v[q];
        }
        __syncthreads();
        if (tid == 0) carry += tile_total;
        __syncthreads();
    }
    cluster_sync_all();
    // 4. scatter my points: the start cursors advance to the (exclusive) ends
    for (int k = s0 + tid; k < s1; k += kBuildThreads) {
        const float x = __ldg(pts + (size_t)k * 3 + 0), y = __ldg(pts + (size_t)k * 3 + 1), z = __ldg(pts + (size_t)k * 3 + 2);
        const int cx = bq_cell_coord(x, G.ox, G.ix, G.gx), cy = bq_cell_coord(y, G.oy, G.iy, G.gy), cz = bq_cell_coord(z, G.oz, G.iz, G.gz);
        const int slot = atomicAdd(&cend[(cx * G.gy + cy) * G.gz + cz], 1);
        srt[slot] = make_float4(x, y, z, __int_as_float(k));
    }
}

// ---------------------------------------------------------------------------------------
// grid query: one warp per centre
// ---------------------------------------------------------------------------------------
// Crowded centre, one warp: bit k of a per-warp bitmap in shared memory for every hit, read back in
// index order ("first nsample hits in ascending k" whatever order the candidates were visited in).
// Lanes 0..8 pass the candidate range [rs, re) of their (dx,dy) column.  bm: this warp's bitmap area, 32 * ((1 << wpl_log2) | 1) words.
__device__ __forceinline__ void bq_warp_bitmap_pass(int lane, int rs, int re, float qx, float qy, float qz,
                                                    float radius2, int nsample, int wpl_log2, unsigned *bm,
                                                    const float4 *__restrict__ srt, int *__restrict__ row) {
    const int wpl = 1 << wpl_log2;
    const int stride = wpl | 1;  // odd stride: lane-contiguous ownership without bank conflicts
    unsigned *mine = bm + lane * stride;  // words [lane*wpl, lane*wpl + wpl) of the frame's bitmap
    __syncwarp();
    for (int j = 0; j < wpl; ++j) mine[j] = 0u;
    __syncwarp();
    // a crowded centre has long ranges: walk them one after the other, lanes striding inside a range
    // (the flat index space of pass 1 costs ~110 instructions per 32 candidates in index arithmetic)
#pragma unroll 1
    for (int r = 0; r < 9; ++r) {
        const int r0 = __shfl_sync(kFullMask, rs, r), r1 = __shfl_sync(kFullMask, re, r);
        for (int i = r0 + lane; i < r1; i += 64) {   // two records in flight per lane (ncu: the distance
            const bool two = i + 32 < r1;             // computation waited for its load: 20 % of the samples)
            const float4 pa = __ldg(srt + i);
            const float4 pb = __ldg(srt + (two ? i + 32 : i));
            const float da = sqdist_ref(__fsub_rn(qx, pa.x), __fsub_rn(qy, pa.y), __fsub_rn(qz, pa.z));
            const float db = sqdist_ref(__fsub_rn(qx, pb.x), __fsub_rn(qy, pb.y), __fsub_rn(qz, pb.z));
            if (da < radius2) {
                const unsigned k = (unsigned)__float_as_int(pa.w);
                const unsigned word = k >> 5;
                atomicOr(&bm[(word >> wpl_log2) * stride + (word & (wpl - 1))], 1u << (k & 31u));
            }
            if (two && db < radius2) {
                const unsigned k = (unsigned)__float_as_int(pb.w);
                const unsigned word = k >> 5;
                atomicOr(&bm[(word >> wpl_log2) * stride + (word & (wpl - 1))], 1u << (k & 31u));
            }
        }
    }
    __syncwarp();
    // (cnt > 32 >= ... the row is filled completely when nsample <= cnt; pad otherwise)
    int c = 0;
    for (int j = 0; j < wpl; ++j) c += __popc(mine[j]);
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o) incl += y;
    }
    const int total = __shfl_sync(kFullMask, incl, 31);
    int pos = incl - c;
    int first = 0x7fffffff;
    if (c > 0 && (pos < nsample || pos == 0)) {
        for (int j = 0; j < wpl && pos < nsample; ++j) {
            unsigned bits = mine[j];
            while (bits && pos < nsample) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1u;
                const int k = ((lane * wpl + j) << 5) + b;
                if (first == 0x7fffffff) first = k;
                row[pos++] = k;
            }
        }
    }
    if (total < nsample) {  // pad with the first (smallest) hit
        const int f = __reduce_min_sync(kFullMask, first);
        for (int l = total + lane; l < nsample; l += 32) row[l] = f;
    }
}

constexpr int kQueryWarps = 8;

__global__ void __launch_bounds__(kQueryWarps * 32)
bq_query_kernel(int n, int m, float radius2, int nsample, int cmax, int wpl_log2 /*bitmap words per lane = 1 << wpl_log2*/,
                const float *__restrict__ new_xyz, const BQGrid *__restrict__ grids,
                const int *__restrict__ cellend, const float4 *__restrict__ sorted,
                int *__restrict__ idx, BQRagged rg = BQRagged{}) {
    extern __shared__ unsigned bitmap_all[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int bi = blockIdx.y;
    const int qi = blockIdx.x * (blockDim.x >> 5) + w;   // 8 warps per CTA, fewer when the bitmaps of a large frame need the room
    if (qi >= m) return;  // whole warp
    size_t xrow0 = (size_t)bi * n, qrow = (size_t)bi * m + qi;
    if (rg.xyz_cnt) {     // stacked: m = total centres, qi = global centre row; find its frame
        int acc = 0, ps = 0;
        bi = 0;
        for (;;) {
            const int c = __ldg(rg.new_cnt + bi);
            if (qi < acc + c || bi == rg.nframes - 1) break;
            acc += c;
            ps += __ldg(rg.xyz_cnt + bi);
            ++bi;
        }
        xrow0 = (size_t)ps;
        qrow = (size_t)qi;
        n = __ldg(rg.xyz_cnt + bi);
        if (n <= 0) {     // a frame without points: the empty-ball flag of ball_query_gpu.cu:69
            if (lane == 0) idx[qrow * nsample] = -1;
            return;
        }
    }
    const int wpl = 1 << wpl_log2;
    const int stride = wpl | 1;  // odd stride: lane-contiguous ownership without bank conflicts
    unsigned *bm = bitmap_all + (size_t)w * 32 * stride;
    unsigned *mine = bm + lane * stride;  // words [lane*wpl, lane*wpl + wpl) of the frame's bitmap

    const BQGrid G = grids[bi];
    const float *q = new_xyz + qrow * 3;
    const float qx = __ldg(q), qy = __ldg(q + 1), qz = __ldg(q + 2);
    const int cx = bq_cell_coord(qx, G.ox, G.ix, G.gx);
    const int cy = bq_cell_coord(qy, G.oy, G.iy, G.gy);
    const int cz = bq_cell_coord(qz, G.oz, G.iz, G.gz);
    const int *cend = cellend + (size_t)bi * cmax;
    const float4 *srt = sorted + xrow0;
    const int z0 = max(cz - 1, 0), z1 = min(cz + 1, G.gz - 1);
    int *row = idx + qrow * nsample;

    // lanes 0..8 fetch the candidate range of their (dx,dy) column
    int rs = 0, re = 0;
    if (lane < 9) {
        const int x = cx + lane / 3 - 1, y = cy + lane % 3 - 1;
        if (x >= 0 && x < G.gx && y >= 0 && y < G.gy) {
            const int c0 = (x * G.gy + y) * G.gz + z0;
            const int c1 = (x * G.gy + y) * G.gz + z1;
            rs = c0 == 0 ? 0 : __ldg(cend + c0 - 1);
            re = __ldg(cend + c1);
        }
    }
    // The nine ranges are walked as ONE flat index space (prefix sums broadcast to registers), so
    // that a typical centre (20-60 candidates in total, a handful per range) costs one or two
    // rounds of loads instead of nine dependent ones.
    int pre = re - rs;  // lengths -> inclusive prefix over lanes 0..8
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const int y = __shfl_up_sync(kFullMask, pre, o);
        if (lane >= o) pre += y;
    }
    const int total_cand = __shfl_sync(kFullMask, pre, 8);
    int pstart[9], rstart[9];   // exclusive prefix and first record of every range
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        pstart[r] = r == 0 ? 0 : __shfl_sync(kFullMask, pre, r - 1);
        rstart[r] = __shfl_sync(kFullMask, rs, r);
    }
    auto record_of = [&](int f) -> int {   // flat candidate number -> position in the sorted records
        int i = rstart[0] + f;
#pragma unroll
        for (int r = 1; r < 9; ++r) i = f >= pstart[r] ? rstart[r] + (f - pstart[r]) : i;
        return i;
    };

    // Pass 1 -- most centres have few neighbours (KITTI SA1: median 9, 87 % at most 32): collect the
    // hits in a 32-entry list (the first words of this warp's bitmap area), then sort them with a
    // bitonic network, one per lane.  Abandoned as soon as a 33rd hit shows up.
    int cnt = 0;
    for (int base = 0; base < total_cand && cnt <= 32; base += 32) {
        const int f = base + lane;
        bool hit = false;
        int k = 0;
        if (f < total_cand) {
            const float4 pt = __ldg(srt + record_of(f));
            const float d2 = sqdist_ref(__fsub_rn(qx, pt.x), __fsub_rn(qy, pt.y), __fsub_rn(qz, pt.z));
            hit = d2 < radius2;
            k = __float_as_int(pt.w);
        }
        const unsigned ball = __ballot_sync(kFullMask, hit);
        const int slot = cnt + __popc(ball & ((1u << lane) - 1u));
        if (hit && slot < 32) bm[slot] = (unsigned)k;
        cnt += __popc(ball);
    }
    if (cnt == 0) {        // no hit: the row stays as the caller left it (stacked API: empty-ball flag)
        if (rg.xyz_cnt && lane == 0) row[0] = -1;
        return;
    }
    if (cnt <= 32) {
        __syncwarp();
        int v = lane < cnt ? (int)bm[lane] : 0x7fffffff;
        for (int k = 2; (k >> 1) < cnt; k <<= 1) {   // only the stages a list of cnt needs (padding sorts last)
            for (int j = k >> 1; j > 0; j >>= 1) {
                const int o = __shfl_xor_sync(kFullMask, v, j);
                const bool keep_min = ((lane & k) == 0) == ((lane & j) == 0);
                v = keep_min ? min(v, o) : max(v, o);
            }
        }
        const int first = __shfl_sync(kFullMask, v, 0);
        for (int l = lane; l < nsample; l += 32) row[l] = (l < cnt) ? v : first;   // l < cnt implies l == lane
        return;
    }

    // Pass 2 -- crowded centre: bitmap pass (shared with the thread-per-centre kernel)
    bq_warp_bitmap_pass(lane, rs, re, qx, qy, qz, radius2, nsample, wpl_log2, bm, srt, row);
}

// ---------------------------------------------------------------------------------------
// grid query, nsample <= 64: one warp per centre, the nsample smallest hit indices in registers
// ---------------------------------------------------------------------------------------
// L[r] of lane l = entry 32 r + l of the ascending list (padded with INT_MAX).  Inserting x: its rank is the number of
// entries below it (ballot + popc); lanes above the rank take their left neighbour's entry (shfl_up), the lane at the
// rank takes x; the entry pushed out of a full register moves on to the next one.
template <int R>
__device__ __forceinline__ void bq_list_insert(int (&L)[R], int x, int lane) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int pos = __popc(__ballot_sync(kFullMask, L[r] < x));
        if (pos < 32) {                                       // warp-uniform
            const int last = __shfl_sync(kFullMask, L[r], 31);
            const int up = __shfl_up_sync(kFullMask, L[r], 1);
            L[r] = lane < pos ? L[r] : (lane == pos ? x : up);
            x = last;                                         // moves on (INT_MAX when the register was not full)
        }
    }
}

template <int R>
__global__ void __launch_bounds__(256)
bq_query_topk_kernel(int n, int m, float radius2, int nsample, int cmax, const float *__restrict__ new_xyz,
                     const BQGrid *__restrict__ grids, const int *__restrict__ cellend,
                     const float4 *__restrict__ sorted, int *__restrict__ idx, BQRagged rg) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int bi = blockIdx.y;
    const int qi = blockIdx.x * (blockDim.x >> 5) + w;
    if (qi >= m) return;  // whole warp
    size_t xrow0 = (size_t)bi * n, qrow = (size_t)bi * m + qi;
    if (rg.xyz_cnt) {     // stacked: m = total centres, qi = global centre row; find its frame
        int acc = 0, ps = 0;
        bi = 0;
        for (;;) {
            const int c = __ldg(rg.new_cnt + bi);
            if (qi < acc + c || bi == rg.nframes - 1) break;
            acc += c;
            ps += __ldg(rg.xyz_cnt + bi);
            ++bi;
        }
        xrow0 = (size_t)ps;
        qrow = (size_t)qi;
        if (__ldg(rg.xyz_cnt + bi) <= 0) {     // a frame without points: the empty-ball flag of ball_query_gpu.cu:69
            if (lane == 0) idx[qrow * nsample] = -1;
            return;
        }
    }
    const BQGrid G = grids[bi];
    const float *q = new_xyz + qrow * 3;
    const float qx = __ldg(q), qy = __ldg(q + 1), qz = __ldg(q + 2);
    const int cx = bq_cell_coord(qx, G.ox, G.ix, G.gx);
    const int cy = bq_cell_coord(qy, G.oy, G.iy, G.gy);
    const int cz = bq_cell_coord(qz, G.oz, G.iz, G.gz);
    const int *cend = cellend + (size_t)bi * cmax;
    const float4 *srt = sorted + xrow0;
    const int z0 = max(cz - 1, 0), z1 = min(cz + 1, G.gz - 1);
    int *row = idx + qrow * nsample;

    // lanes 0..8 fetch the candidate range of their (dx,dy) column; the nine ranges are walked as ONE flat index space
    int rs = 0, re = 0;
    if (lane < 9) {
        const int x = cx + lane / 3 - 1, y = cy + lane % 3 - 1;
        if (x >= 0 && x < G.gx && y >= 0 && y < G.gy) {
            const int c0 = (x * G.gy + y) * G.gz + z0;
            const int c1 = (x * G.gy + y) * G.gz + z1;
            rs = c0 == 0 ? 0 : __ldg(cend + c0 - 1);
            re = __ldg(cend + c1);
        }
    }
    int pre = re - rs;  // lengths -> inclusive prefix over lanes 0..8
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const int y = __shfl_up_sync(kFullMask, pre, o);
        if (lane >= o) pre += y;
    }
    const int total_cand = __shfl_sync(kFullMask, pre, 8);
    int pstart[9], rstart[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        pstart[r] = r == 0 ? 0 : __shfl_sync(kFullMask, pre, r - 1);
        rstart[r] = __shfl_sync(kFullMask, rs, r);
    }
    auto record_of = [&](int f) -> int {
        int i = rstart[0] + f;
#pragma unroll
        for (int r = 1; r < 9; ++r) i = f >= pstart[r] ? rstart[r] + (f - pstart[r]) : i;
        return i;
    };

    int L[R];
#pragma unroll
    for (int r = 0; r < R; ++r) L[r] = 0x7fffffff;
    int hits = 0, tau = 0x7fffffff;     // tau: only indices below it can still enter the list
    const int tr = (nsample - 1) >> 5, tl = (nsample - 1) & 31;     // where entry nsample - 1 lives
    for (int base = 0; base < total_cand; base += 64) {              // two records in flight per lane
        const int fa = base + lane, fb = base + 32 + lane;
        const bool la = fa < total_cand, lb = fb < total_cand;
        float4 pa = make_float4(0.f, 0.f, 0.f, 0.f), pb = pa;
        if (la) pa = __ldg(srt + record_of(fa));
        if (lb) pb = __ldg(srt + record_of(fb));
        const bool ha = la && sqdist_ref(__fsub_rn(qx, pa.x), __fsub_rn(qy, pa.y), __fsub_rn(qz, pa.z)) < radius2;
        const bool hb = lb && sqdist_ref(__fsub_rn(qx, pb.x), __fsub_rn(qy, pb.y), __fsub_rn(qz, pb.z)) < radius2;
        const int ka = __float_as_int(pa.w), kb = __float_as_int(pb.w);
        const unsigned balla = __ballot_sync(kFullMask, ha), ballb = __ballot_sync(kFullMask, hb);
        hits += __popc(balla) + __popc(ballb);
        unsigned qa = __ballot_sync(kFullMask, ha && ka < tau);
        while (qa) {
            const int src = __ffs(qa) - 1;
            qa &= qa - 1u;
            bq_list_insert<R>(L, __shfl_sync(kFullMask, ka, src), lane);
        }
        unsigned qb = __ballot_sync(kFullMask, hb && kb < tau);
        while (qb) {
            const int src = __ffs(qb) - 1;
            qb &= qb - 1u;
            bq_list_insert<R>(L, __shfl_sync(kFullMask, kb, src), lane);
        }
        if (hits >= nsample) {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (r == tr) tau = __shfl_sync(kFullMask, L[r], tl);
        }
    }
    if (hits == 0) {        // no hit: the row stays as the caller left it (stacked API: empty-ball flag)
        if (rg.xyz_cnt && lane == 0) row[0] = -1;
        return;
    }
    const int have = min(hits, nsample);
    const int first = __shfl_sync(kFullMask, L[0], 0);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int l = 32 * r + lane;
        if (l < nsample) row[l] = l < have ? L[r] : first;
    }
}

// clusters of `cl` build CTAs the device keeps resident at once (cached per device and size)
static int build_active_clusters(int cl) {
    static std::mutex mu;
    static std::map<int, int> cache;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(dev * 16 + cl);
    if (it != cache.end()) return it->second;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)cl);
    cfg.blockDim = dim3(kBuildThreads);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int v = 0;
    if (cudaOccupancyMaxActiveClusters(&v, (const void *)bq_build_kernel, &cfg) != cudaSuccess) {
        (void)cudaGetLastError();
        v = cl == 1 ? kNumSMs : 0;
    }
    cache[dev * 16 + cl] = v;
    return v;
}

// rg.xyz_cnt != nullptr (stacked API): n / m are the TOTAL row counts of xyz / new_xyz over the b frames
static int ball_query_grid(int b, int n, int m, float radius, float radius2, int nsample,
                           const float *new_xyz, const float *xyz, int *idx, cudaStream_t st,
                           BQRagged rg = BQRagged{}, int mode = PDM_FPS_MODE_AUTO) {
    const bool ragged = rg.xyz_cnt != nullptr;
    // cells per frame: ~4 per point, power of two, bounded (stacked: sized for twice the mean frame; a larger frame
    // simply gets wider cells, bq_build_kernel)
    const long long npf = ragged ? 2LL * (n / b + 1) : n;
    int cmax = 4096;
    while (cmax < 4 * npf && cmax < 262144) cmax <<= 1;
    // lookup kernel: per-warp bitmaps are the cheaper way to order the hits while they are small (n <= 32768: at most 32
    // words per lane); beyond that -- or never, with PDM_BQ_KERNEL=bitmap / always, with =topk -- the register list
    const char *kenv = getenv("PDM_BQ_KERNEL");
    const bool topk = nsample <= 64 && (kenv && kenv[0] == 'b' ? false : (kenv && kenv[0] == 't' && kenv[1] == 'o') ? true : n > 32768);
    // bitmap kernel only: a warp's bitmap has n bits, large frames get fewer warps per CTA
    const int words = (n + 31) / 32;
    int wpl_log2 = 0;
    while ((32 << wpl_log2) < words) ++wpl_log2;
    const int wpl = 1 << wpl_log2;
    int qwarps = kQueryWarps;
    while (qwarps > 1 && (size_t)qwarps * 32 * (wpl | 1) * sizeof(unsigned) > 200 * 1024) qwarps >>= 1;
    const size_t smem = (size_t)qwarps * 32 * (wpl | 1) * sizeof(unsigned);
    if (!topk && smem > 200 * 1024) return PDM_ERR_UNSUPPORTED;  // caller falls back to the tiled kernel

    const size_t rows = ragged ? (size_t)n : (size_t)b * n;
    const size_t sz_grid = ((sizeof(BQGrid) * b + 255) / 256) * 256;
    const size_t sz_cend = ((size_t)b * (cmax + kBuildAux) * sizeof(int) + 255) / 256 * 256;   // counters + per-frame build state
    const size_t sz_sorted = rows * sizeof(float4);
    char *scratch = static_cast<char *>(stream_scratch(st, sz_grid + sz_cend + sz_sorted));
    if (!scratch) return PDM_ERR_INVALID_ARG;  // message recorded by stream_scratch
    BQGrid *grids = reinterpret_cast<BQGrid *>(scratch);
    int *cend = reinterpret_cast<int *>(scratch + sz_grid);
    int *parts = cend + (size_t)b * cmax;
    float4 *sorted = reinterpret_cast<float4 *>(scratch + sz_grid + sz_cend);
    PDM_CHECK_CUDA(cudaMemsetAsync(cend, 0, (size_t)b * (cmax + kBuildAux) * sizeof(int), st));

    prefer_max_smem((const void *)bq_build_kernel);     // same L1 / shared split as the rest of the chain (common.cuh)
    prefer_max_smem((const void *)bq_query_topk_kernel<1>);
    prefer_max_smem((const void *)bq_query_topk_kernel<2>);
    // build: a cluster of 1..8 CTAs per frame (>= ~1024 points per CTA): the largest size whose b clusters are all
    // resident at once (this B200 keeps 15 clusters of 8 such CTAs: a 16-frame batch would run in two waves), else the
    // one with the fewest waves
    int cl_hi = (int)((npf + 1023) / 1024);
    cl_hi = cl_hi < 1 ? 1 : (cl_hi > kBuildMaxCl ? kBuildMaxCl : cl_hi);
    // throughput mode (many batches in flight on several streams): one CTA per frame.  A cluster needs that many SMs of
    // one GPC free at the same moment, which a busy GPU rarely offers: measured 41.2k frames/s with clusters of 6 against
    // 47.7k with single CTAs in the pipelined SA chain, and the reverse for a batch alone (2.79 vs 2.86 ms per step).
    if (mode == PDM_FPS_MODE_THROUGHPUT) cl_hi = 1;
    if (const char *ce = getenv("PDM_BQ_BUILD_CL")) {      // A/B knob
        const int f = atoi(ce);
        if (f >= 1 && f < cl_hi) cl_hi = f;
    }
    int cl = 1, best_waves = 1 << 30;
    for (int c = cl_hi; c >= 1; --c) {
        const int act = build_active_clusters(c);
        if (act < 1) continue;
        const int waves = (b + act - 1) / act;
        if (waves < best_waves) { best_waves = waves; cl = c; }
        if (waves == 1) break;
    }
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(b * cl));
        cfg.blockDim = dim3(kBuildThreads);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cl;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, bq_build_kernel, n, radius, cmax, xyz, grids, cend, parts, sorted, rg);
        if (e != cudaSuccess) return fail((int)e, "ball_query(grid build): %s", cudaGetErrorString(e));
        count_launch();
    }
    if (topk) {
        dim3 grid((m + 7) / 8, ragged ? 1 : b);
        if (nsample <= 32)
            bq_query_topk_kernel<1><<<grid, 256, 0, st>>>(n, m, radius2, nsample, cmax, new_xyz, grids, cend, sorted, idx, rg);
        else
            bq_query_topk_kernel<2><<<grid, 256, 0, st>>>(n, m, radius2, nsample, cmax, new_xyz, grids, cend, sorted, idx, rg);
    } else {
        prefer_max_smem((const void *)bq_query_kernel);
        if (smem > 48 * 1024 && ensure_dynamic_smem((const void *)bq_query_kernel, smem) != PDM_OK)
            return PDM_ERR_UNSUPPORTED;  // message already recorded; caller falls back to the tiled kernel
        dim3 grid((m + qwarps - 1) / qwarps, ragged ? 1 : b);
        bq_query_kernel<<<grid, qwarps * 32, smem, st>>>(n, m, radius2, nsample, cmax, wpl_log2, new_xyz, grids, cend, sorted,
                                                         idx, rg);
    }
    count_launch();
    const cudaError_t e1 = cudaGetLastError();
    if (e1 != cudaSuccess) return fail((int)e1, "ball_query(grid): %s", cudaGetErrorString(e1));
    return PDM_OK;
}

// Stacked fallback (unusable radius, or a stack whose bitmaps do not fit): the reference loop itself, one thread per centre.
__global__ void __launch_bounds__(128)
ball_query_stack_scan_kernel(int nframes, int m_total, float radius2, int nsample, const float *__restrict__ new_xyz,
                             const int *__restrict__ new_cnt, const float *__restrict__ xyz,
                             const int *__restrict__ xyz_cnt, int *__restrict__ idx) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= m_total) return;
    int bi = 0, acc = 0, ps = 0;
    for (;;) {
        const int c = __ldg(new_cnt + bi);
        if (qi < acc + c || bi == nframes - 1) break;
        acc += c;
        ps += __ldg(xyz_cnt + bi);
        ++bi;
    }
    const int n = __ldg(xyz_cnt + bi);
    const float *pts = xyz + (size_t)ps * 3;
    const float qx = __ldg(new_xyz + (size_t)qi * 3), qy = __ldg(new_xyz + (size_t)qi * 3 + 1), qz = __ldg(new_xyz + (size_t)qi * 3 + 2);
    int *row = idx + (size_t)qi * nsample;
    int cnt = 0;
    for (int k = 0; k < n; ++k) {
        const float d2 = sqdist_ref(__fsub_rn(qx, __ldg(pts + (size_t)k * 3)), __fsub_rn(qy, __ldg(pts + (size_t)k * 3 + 1)),
                                    __fsub_rn(qz, __ldg(pts + (size_t)k * 3 + 2)));
        if (d2 < radius2) {
            if (cnt == 0)
                for (int l = 1; l < nsample; ++l) row[l] = k;
            row[cnt] = k;
            if (++cnt >= nsample) break;
        }
    }
    if (cnt == 0) row[0] = -1;
}

}  // namespace pdm

extern "C" int pdm_stack_ball_query(int b, int m_total, int n_total, float radius, int nsample, const float *new_xyz,
                                    const int *new_xyz_batch_cnt, const float *xyz, const int *xyz_batch_cnt, int *idx,
                                    void *stream) {
    using namespace pdm;
    if (b < 0 || m_total < 0 || n_total < 0 || nsample < 0) return fail(PDM_ERR_INVALID_ARG, "stack_ball_query: negative size");
    if (b == 0 || m_total == 0 || nsample == 0) return PDM_OK;
    if (!new_xyz || !idx || !new_xyz_batch_cnt || !xyz_batch_cnt || (n_total > 0 && !xyz))
        return fail(PDM_ERR_INVALID_ARG, "stack_ball_query: null pointer");
    const float radius2 = radius * radius;  // fp32, as pointnet2_stack ball_query_gpu.cu:43
    cudaStream_t st = (cudaStream_t)stream;
    const char *force = getenv("PDM_BQ_KERNEL");
    const bool scan = force && force[0] == 't' && force[1] == 'i';
    if (!scan && n_total > 0 && radius > 0.f && radius < INFINITY) {
        BQRagged rg;
        rg.xyz_cnt = xyz_batch_cnt;
        rg.new_cnt = new_xyz_batch_cnt;
        rg.nframes = b;
        const int rc = ball_query_grid(b, n_total, m_total, radius, radius2, nsample, new_xyz, xyz, idx, st, rg);
        if (rc != PDM_ERR_UNSUPPORTED) return rc;
    }
    ball_query_stack_scan_kernel<<<(m_total + 127) / 128, 128, 0, st>>>(b, m_total, radius2, nsample, new_xyz,
                                                                       new_xyz_batch_cnt, xyz, xyz_batch_cnt, idx);
    count_launch();
    PDM_CHECK_LAUNCH("stack_ball_query");
    return PDM_OK;
}

extern "C" int pdm_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                              const float *xyz, int *idx, void *stream) {
    return pdm_ball_query_ex(b, n, m, radius, nsample, new_xyz, xyz, idx, PDM_FPS_MODE_AUTO, stream);
}

extern "C" int pdm_ball_query_ex(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                                 const float *xyz, int *idx, int mode, void *stream) {
    using namespace pdm;
    if (b < 0 || n < 0 || m < 0 || nsample < 0) return fail(PDM_ERR_INVALID_ARG, "ball_query: negative size");
    if (b == 0 || m == 0 || nsample == 0 || n == 0) return PDM_OK;
    if (!new_xyz || !xyz || !idx) return fail(PDM_ERR_INVALID_ARG, "ball_query: null pointer");
    if (b > 65535) return fail(PDM_ERR_UNSUPPORTED, "ball_query: batch %d > 65535", b);
    const float radius2 = radius * radius;  // fp32, as ball_query_gpu.cu:29
    cudaStream_t st = (cudaStream_t)stream;
    const char *force = getenv("PDM_BQ_KERNEL");  // "tiled" | "thread" | unset (debug/testing knob)
    const bool tiled = force && force[0] == 't' && force[1] == 'i';
    // the grid needs a usable radius; NaN / non-positive radii have no hits at all or are
    // handled by the scan kernel with the reference's exact comparison
    if (!tiled && radius > 0.f && radius < INFINITY) {
        const int rc = ball_query_grid(b, n, m, radius, radius2, nsample, new_xyz, xyz, idx, st, BQRagged{}, mode);
        if (rc != PDM_ERR_UNSUPPORTED) return rc;
    }
    dim3 grid((m + 127) / 128, b);
    ball_query_tiled_kernel<<<grid, 128, 0, st>>>(n, m, radius2, nsample, new_xyz, xyz, idx);
    count_launch();
    PDM_CHECK_LAUNCH("ball_query");
    return PDM_OK;
}
