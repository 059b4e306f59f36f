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
// (x,y,z,k) records -- one CTA per frame).  A centre then only meets the 3x3x3 cells around
// its own: fp32 rounding of the cell coordinate is ~1e-4 of a cell, far inside the 1% margin,
// so no point with d2 < r^2 can lie outside that neighbourhood.  bq_query_kernel gives one
// warp to each centre: the lanes stride over the candidate records (coalesced 16-byte loads),
// evaluate the exact reference distance, and set bit k of a per-warp bitmap in shared memory
// for every hit.  The bitmap is then read back in index order, which yields "first nsample
// hits in ascending k" directly, whatever order the candidates were visited in.
//
// Tiled kernel (fallback, PDM_BQ_KERNEL=tiled): one thread per centre like the reference,
// points streamed through shared memory; identical results by construction.
#include <stdlib.h>

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

// Stacked (ragged) frames, pointnet2_stack/src/ball_query_gpu.cu:15-70: frame f owns rows [sum xyz_cnt[:f], +xyz_cnt[f])
// of xyz and the centres [sum new_cnt[:f], +new_cnt[f]); indices are LOCAL to the frame; a centre without a hit gets
// idx[0] = -1 (:69).  xyz_cnt == nullptr: uniform (B,N,3) / (B,M,3) frames of the batch API.
struct BQRagged {
    const int *xyz_cnt = nullptr, *new_cnt = nullptr;
    int nframes = 0;
};

__global__ void __launch_bounds__(kBuildThreads)
bq_build_kernel(int n, float radius, int cmax, const float *__restrict__ xyz, BQGrid *__restrict__ grids,
                int *__restrict__ cellid, int *__restrict__ cellend, float4 *__restrict__ sorted,
                BQRagged rg = BQRagged{}) {
    __shared__ float red[6][kBuildThreads / 32];
    __shared__ BQGrid sg;
    __shared__ int wsum[kBuildThreads / 32];
    __shared__ int carry, tile_total;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int bi = blockIdx.x;
    size_t pstart = (size_t)bi * n;
    if (rg.xyz_cnt) {
        int ps = 0;
        for (int k = 0; k < bi; ++k) ps += __ldg(rg.xyz_cnt + k);
        pstart = (size_t)ps;
        n = __ldg(rg.xyz_cnt + bi);
    }
    const float *pts = xyz + pstart * 3;
    int *cid = cellid + pstart;
    int *cend = cellend + (size_t)bi * cmax;
    float4 *srt = sorted + pstart;

    // 1. frame box over finite coordinates
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = tid; k < n; k += kBuildThreads) {
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
    if (tid == 0) {
        float l[3], h[3];
        for (int a = 0; a < 3; ++a) {
            l[a] = INFINITY;
            h[a] = -INFINITY;
            for (int i = 0; i < kBuildThreads / 32; ++i) {
                l[a] = fminf(l[a], red[a][i]);
                h[a] = fmaxf(h[a], red[3 + a][i]);
            }
            if (!(l[a] <= h[a])) l[a] = h[a] = 0.f;  // no finite coordinate on this axis
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
        grids[bi] = sg;
        carry = 0;
    }
    __syncthreads();
    const BQGrid G = sg;

    // 2. histogram (global atomics; the counters live in cellend)
    for (int c = tid; c < G.ncell; c += kBuildThreads) cend[c] = 0;
    __syncthreads();
    for (int k = tid; k < n; k += kBuildThreads) {
        const int cx = bq_cell_coord(__ldg(pts + (size_t)k * 3 + 0), G.ox, G.ix, G.gx);
        const int cy = bq_cell_coord(__ldg(pts + (size_t)k * 3 + 1), G.oy, G.iy, G.gy);
        const int cz = bq_cell_coord(__ldg(pts + (size_t)k * 3 + 2), G.oz, G.iz, G.gz);
        const int c = (cx * G.gy + cy) * G.gz + cz;
        cid[k] = c;
        atomicAdd(&cend[c], 1);
    }
    __syncthreads();

    // 3. exclusive scan of the counts, in place (tiles of 4 * kBuildThreads cells)
    for (int base = 0; base < G.ncell; base += 4 * kBuildThreads) {
        const int c0 = base + tid * 4;
        int v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = (c0 + q < G.ncell) ? cend[c0 + q] : 0;
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
            int s = wsum[lane];
            int si = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(kFullMask, si, o);
                if (lane >= o) si += y;
            }
            wsum[lane] = si - s;  // exclusive warp offsets
            if (lane == 31) tile_total = si;
        }
        __syncthreads();
        int run = carry + wsum[w] + incl - tsum;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (c0 + q < G.ncell) cend[c0 + q] = run;
            run += v[q];
        }
        __syncthreads();
        if (tid == 0) carry += tile_total;
        __syncthreads();
    }

    // 4. scatter: the start cursors advance to the (exclusive) ends
    for (int k = tid; k < n; k += kBuildThreads) {
        const int slot = atomicAdd(&cend[cid[k]], 1);
        srt[slot] = make_float4(__ldg(pts + (size_t)k * 3 + 0), __ldg(pts + (size_t)k * 3 + 1),
                                __ldg(pts + (size_t)k * 3 + 2), __int_as_float(k));
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
// grid query: one THREAD per centre, crowded centres finished by the warp
// ---------------------------------------------------------------------------------------
// (Opt-in experiment, see ball_query_grid.)  The warp-per-centre kernel above spends ~900 warp instructions on a centre (prefix sums, flat
// candidate walk, sort network -- executed by 32 lanes for ~40 candidates and ~9 hits) and is bound
// by instruction issue (ncu: 78 % of the issue slots).  Here a thread walks the nine candidate
// ranges of its own centre (ranges kept in shared memory, records read 16 bytes at a time; a lane's
// consecutive records share cache lines) and appends the hits to a 32-entry list in shared memory;
// 87 % of KITTI-shaped SA1 centres have at most 32 hits, which is then the complete hit set: an
// insertion sort puts it in ascending index order and the warp writes the rows of its 32 centres
// with coalesced 128-byte stores.  A centre with more than 32 hits stops scanning and is finished
// by its whole warp with the bitmap pass.  Same results by construction: identical distance
// arithmetic, "the nsample smallest hit indices in ascending order, padded with the first".
constexpr int kTpcThreads = 128;
constexpr int kTpcList = 32;
constexpr int kTpcStride = kTpcThreads + 1;   // odd row stride: per-thread and per-warp accesses both conflict-free

__global__ void __launch_bounds__(kTpcThreads)
bq_query_tpc_kernel(int n, int m, float radius2, int nsample, int cmax, int wpl_log2,
                    const float *__restrict__ new_xyz, const BQGrid *__restrict__ grids,
                    const int *__restrict__ cellend, const float4 *__restrict__ sorted,
                    int *__restrict__ idx) {
    extern __shared__ unsigned tpc_smem[];
    int *list = reinterpret_cast<int *>(tpc_smem);              // [kTpcList][kTpcStride]
    int *rs_s = list + kTpcList * kTpcStride;                    // [9][kTpcStride]
    int *re_s = rs_s + 9 * kTpcStride;                           // [9][kTpcStride]
    unsigned *bitmap_all = reinterpret_cast<unsigned *>(re_s + 9 * kTpcStride);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int bi = blockIdx.y;
    const int qi = blockIdx.x * kTpcThreads + tid;
    const bool live = qi < m;
    const BQGrid G = grids[bi];
    const float *q = new_xyz + ((size_t)bi * m + (live ? qi : 0)) * 3;
    const float qx = __ldg(q), qy = __ldg(q + 1), qz = __ldg(q + 2);
    const int cx = bq_cell_coord(qx, G.ox, G.ix, G.gx);
    const int cy = bq_cell_coord(qy, G.oy, G.iy, G.gy);
    const int cz = bq_cell_coord(qz, G.oz, G.iz, G.gz);
    const int *cend = cellend + (size_t)bi * cmax;
    const float4 *srt = sorted + (size_t)bi * n;
    const int z0 = max(cz - 1, 0), z1 = min(cz + 1, G.gz - 1);

    // candidate ranges of the 3x3 columns (z0..z1 is contiguous in the cell order)
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        const int x = cx + r / 3 - 1, y = cy + r % 3 - 1;
        int s = 0, e = 0;
        if (live && x >= 0 && x < G.gx && y >= 0 && y < G.gy) {
            const int c0 = (x * G.gy + y) * G.gz + z0, c1 = (x * G.gy + y) * G.gz + z1;
            s = c0 == 0 ? 0 : __ldg(cend + c0 - 1);
            e = __ldg(cend + c1);
        }
        rs_s[r * kTpcStride + tid] = s;
        re_s[r * kTpcStride + tid] = e;
    }
    int cnt = 0;
    {
        // four records in flight per thread: the walk is a chain of dependent L2 reads otherwise
        constexpr int PF = 4;
        int r = 0, i = rs_s[tid], end = re_s[tid];
        bool more = true;
        while (more) {
            int ii[PF];
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                while (i >= end && r < 8) {
                    ++r;
                    i = rs_s[r * kTpcStride + tid];
                    end = re_s[r * kTpcStride + tid];
                }
                ii[u] = i < end ? i++ : -1;
            }
            if (ii[0] < 0) break;
            float4 pt[PF];
#pragma unroll
            for (int u = 0; u < PF; ++u) pt[u] = __ldg(srt + max(ii[u], 0));
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const float d2 = sqdist_ref(__fsub_rn(qx, pt[u].x), __fsub_rn(qy, pt[u].y), __fsub_rn(qz, pt[u].z));
                if (ii[u] >= 0 && d2 < radius2 && cnt <= kTpcList) {
                    if (cnt < kTpcList) list[cnt * kTpcStride + tid] = __float_as_int(pt[u].w);
                    ++cnt;
                }
            }
            more = ii[PF - 1] >= 0 && cnt <= kTpcList;   // crowded (> 32 hits): the warp takes over below
        }
    }
    const bool crowded = cnt > kTpcList;
    if (!crowded) {   // insertion sort of my hit list (ascending point index)
        for (int a = 1; a < cnt; ++a) {
            const int v = list[a * kTpcStride + tid];
            int b = a - 1;
            while (b >= 0) {
                const int u = list[b * kTpcStride + tid];
                if (u <= v) break;
                list[(b + 1) * kTpcStride + tid] = u;
                --b;
            }
            list[(b + 1) * kTpcStride + tid] = v;
        }
    }
    __syncwarp();
    const int wbase = tid - lane;                 // first thread of my warp
    int *rows = idx + ((size_t)bi * m + (qi - lane)) * nsample;   // row of lane 0's centre
    // rows of the complete lists: one coalesced store per 32 slots
    unsigned wmask = __ballot_sync(kFullMask, live && !crowded && cnt > 0);
    while (wmask) {
        const int c = __ffs(wmask) - 1;
        wmask &= wmask - 1u;
        const int cc = __shfl_sync(kFullMask, cnt, c);
        const int first = list[wbase + c];
        const int v = lane < cc ? list[lane * kTpcStride + wbase + c] : first;
        int *row = rows + (size_t)c * nsample;
        for (int l = lane; l < nsample; l += 32) row[l] = l < cc ? v : first;   // l < cc implies l == lane
    }
    // crowded centres: bitmap pass by the whole warp
    unsigned cmask = __ballot_sync(kFullMask, live && crowded);
    if (cmask) {
        unsigned *bm = bitmap_all + (size_t)w * 32 * ((1 << wpl_log2) | 1);
        while (cmask) {
            const int c = __ffs(cmask) - 1;
            cmask &= cmask - 1u;
            const float cqx = __shfl_sync(kFullMask, qx, c), cqy = __shfl_sync(kFullMask, qy, c), cqz = __shfl_sync(kFullMask, qz, c);
            const int rs = lane < 9 ? rs_s[lane * kTpcStride + wbase + c] : 0;
            const int re = lane < 9 ? re_s[lane * kTpcStride + wbase + c] : 0;
            bq_warp_bitmap_pass(lane, rs, re, cqx, cqy, cqz, radius2, nsample, wpl_log2, bm, srt, rows + (size_t)c * nsample);
            __syncwarp();
        }
    }
}

// rg.xyz_cnt != nullptr (stacked API): n / m are the TOTAL row counts of xyz / new_xyz over the b frames
static int ball_query_grid(int b, int n, int m, float radius, float radius2, int nsample,
                           const float *new_xyz, const float *xyz, int *idx, cudaStream_t st,
                           BQRagged rg = BQRagged{}) {
    const bool ragged = rg.xyz_cnt != nullptr;
    // cells per frame: ~4 per point, power of two, bounded (stacked: sized for twice the mean frame; a larger frame
    // simply gets wider cells, bq_build_kernel)
    const long long npf = ragged ? 2LL * (n / b + 1) : n;
    int cmax = 4096;
    while (cmax < 4 * npf && cmax < 262144) cmax <<= 1;
    const int words = (n + 31) / 32;
    int wpl_log2 = 0;  // bitmap words per lane, rounded up to a power of two
    while ((32 << wpl_log2) < words) ++wpl_log2;
    const int wpl = 1 << wpl_log2;
    int qwarps = kQueryWarps;     // a warp's bitmap has n bits: large frames get fewer warps per CTA
    while (qwarps > 1 && (size_t)qwarps * 32 * (wpl | 1) * sizeof(unsigned) > 200 * 1024) qwarps >>= 1;
    const size_t smem = (size_t)qwarps * 32 * (wpl | 1) * sizeof(unsigned);
    if (smem > 200 * 1024) return PDM_ERR_UNSUPPORTED;  // caller falls back to the tiled kernel

    const size_t sz_grid = ((sizeof(BQGrid) * b + 255) / 256) * 256;
    const size_t rows = ragged ? (size_t)n : (size_t)b * n;
    const size_t sz_cid = ((rows * sizeof(int) + 255) / 256) * 256;
    const size_t sz_cend = (size_t)b * cmax * sizeof(int);
    const size_t sz_sorted = rows * sizeof(float4);
    char *scratch = static_cast<char *>(stream_scratch(st, sz_grid + sz_cid + sz_cend + sz_sorted));
    if (!scratch) return PDM_ERR_INVALID_ARG;  // message recorded by stream_scratch
    BQGrid *grids = reinterpret_cast<BQGrid *>(scratch);
    int *cid = reinterpret_cast<int *>(scratch + sz_grid);
    int *cend = reinterpret_cast<int *>(scratch + sz_grid + sz_cid);
    float4 *sorted = reinterpret_cast<float4 *>(scratch + sz_grid + sz_cid + sz_cend);

    prefer_max_smem((const void *)bq_build_kernel);
    prefer_max_smem((const void *)bq_query_kernel);
    bq_build_kernel<<<b, kBuildThreads, 0, st>>>(n, radius, cmax, xyz, grids, cid, cend, sorted, rg);
    count_launch();
    cudaError_t e1 = cudaGetLastError();
    if (e1 == cudaSuccess) {
        // thread-per-centre kernel: opt-in (PDM_BQ_KERNEL=thread).  Measured on B200 (SA1, batch 16): 386 us
        // alone vs 181 us for the warp-per-centre kernel (a thread's walk is a chain of dependent L2 reads
        // and every warp serialises the bitmap passes of its ~4 crowded centres); the pipelined chain runs
        // at the same rate with either (44.3k vs 45.7k frames/s), so the warp kernel stays the default.
        static const bool warp_kernel = [] { const char *e = getenv("PDM_BQ_KERNEL"); return !(e && e[0] == 't' && e[1] == 'h'); }();
        const size_t smem_tpc = (size_t)(kTpcList + 18) * kTpcStride * sizeof(int) +
                                (size_t)(kTpcThreads / 32) * 32 * (wpl | 1) * sizeof(unsigned);
        if (!warp_kernel && !ragged && smem_tpc <= 100 * 1024 &&
            ensure_dynamic_smem((const void *)bq_query_tpc_kernel, smem_tpc) == PDM_OK) {
            dim3 grid((m + kTpcThreads - 1) / kTpcThreads, b);
            bq_query_tpc_kernel<<<grid, kTpcThreads, smem_tpc, st>>>(n, m, radius2, nsample, cmax, wpl_log2, new_xyz,
                                                                    grids, cend, sorted, idx);
            count_launch();
            e1 = cudaGetLastError();
        } else {
            if (smem > 48 * 1024 && ensure_dynamic_smem((const void *)bq_query_kernel, smem) != PDM_OK)
                return PDM_ERR_UNSUPPORTED;  // message already recorded; caller falls back to the tiled kernel
            dim3 grid((m + qwarps - 1) / qwarps, ragged ? 1 : b);
            bq_query_kernel<<<grid, qwarps * 32, smem, st>>>(n, m, radius2, nsample, cmax, wpl_log2, new_xyz,
                                                                 grids, cend, sorted, idx, rg);
            count_launch();
            e1 = cudaGetLastError();
        }
    }
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
        const int rc = ball_query_grid(b, n, m, radius, radius2, nsample, new_xyz, xyz, idx, st);
        if (rc != PDM_ERR_UNSUPPORTED) return rc;
    }
    dim3 grid((m + 127) / 128, b);
    ball_query_tiled_kernel<<<grid, 128, 0, st>>>(n, m, radius2, nsample, new_xyz, xyz, idx);
    count_launch();
    PDM_CHECK_LAUNCH("ball_query");
    return PDM_OK;
}
