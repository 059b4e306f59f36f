// stack_ops.cu -- the stacked (ragged-batch) operator family of pcdet/ops/pointnet2/pointnet2_stack.
//
// Stacked tensors hold the frames of a batch back to back: xyz (N1+N2+.., 3) with xyz_batch_cnt = [N1, N2, ..] as a
// DEVICE int tensor, features (N1+N2+.., C) channel-last.  Nothing here reads a count on the host (no sync): kernels
// find the frame of a row from the count arrays themselves.  Ball query and farthest point sampling reuse the grid /
// bucket kernels of the batch family through their ragged descriptors (ball_query.cu: BQRagged, fps.cu: FpsRagged);
// this file holds the rest:
//   group_points(+grad)        pointnet2_stack/src/group_points_gpu.cu:14-125
//   three_nn / interpolate     pointnet2_stack/src/interpolate_gpu.cu:17-194
//   voxel_query                pointnet2_stack/src/voxel_query_gpu.cu:10-113
//   vector-pool family         pointnet2_stack/src/vector_pool_gpu.cu:19-485
// Arithmetic follows the reference SASS (nvcc 12.9, default contraction): squared distances as sqdist_ref
// (FMUL y, FFMA x, FFMA z), interpolation as fma(w2,f2, fma(w0,f0, w1*f1)), IEEE division for the vector-pool cell.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace pdm {

constexpr unsigned kAll = 0xffffffffu;

// frame of row `r` of a stacked tensor with per-frame counts cnt[0..b): returns the frame, its first row in `start`;
// rows beyond the total fall into the last frame (reference loop: group_points_gpu.cu:30-35)
__device__ __forceinline__ int stack_frame_of(const int *__restrict__ cnt, int b, int r, int &start) {
    int f = 0, acc = 0;
    for (;;) {
        const int c = __ldg(cnt + f);
        if (r < acc + c || f == b - 1) break;
        acc += c;
        ++f;
    }
    start = acc;
    return f;
}
__device__ __forceinline__ int stack_start_of(const int *__restrict__ cnt, int f) {
    int s = 0;
    for (int k = 0; k < f; ++k) s += __ldg(cnt + k);
    return s;
}

// ---------------------------------------------------------------------------------------------------------------
// grouping: out[pt, c, s] = features[fstart(frame of pt) + idx[pt, s], c]        (group_points_gpu.cu:67-101)
// One warp per centre; lane = neighbour slot (chunks of 32), 4 channels per 16-byte row read, stores coalesced
// along s.  The reference re-derives the frame and re-reads idx for every (c, s) element.
// ---------------------------------------------------------------------------------------------------------------
template <bool VEC4>
__global__ void __launch_bounds__(256)
stack_group_points_kernel(int b, int m, int c, int nsample, const float *__restrict__ features,
                          const int *__restrict__ features_cnt, const int *__restrict__ idx,
                          const int *__restrict__ idx_cnt, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int pt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pt >= m) return;
    int istart;
    const int f = stack_frame_of(idx_cnt, b, pt, istart);
    const size_t fstart = (size_t)stack_start_of(features_cnt, f);
    float *o = out + (size_t)pt * c * nsample;
    for (int s0 = 0; s0 < nsample; s0 += 32) {
        const int s = s0 + lane;
        const bool live = s < nsample;
        const int k = live ? __ldg(idx + (size_t)pt * nsample + s) : 0;
        const float *row = features + (fstart + (size_t)k) * c;
        if (VEC4) {
            for (int ch = 0; ch < c; ch += 4) {
                if (live) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(row + ch));
                    st_cs_f1(o + (size_t)(ch + 0) * nsample + s, v.x);
                    st_cs_f1(o + (size_t)(ch + 1) * nsample + s, v.y);
                    st_cs_f1(o + (size_t)(ch + 2) * nsample + s, v.z);
                    st_cs_f1(o + (size_t)(ch + 3) * nsample + s, v.w);
                }
            }
        } else {
            for (int ch = 0; ch < c; ++ch)
                if (live) st_cs_f1(o + (size_t)ch * nsample + s, __ldg(row + ch));
        }
    }
}

// grad_features[fstart + idx[pt, s], c] += grad_out[pt, c, s]     (atomic, like group_points_gpu.cu:14-44; thread =
// (pt, s, 4 channels) with the channel fastest so the atomics of a warp fall into few rows)
__global__ void __launch_bounds__(256)
stack_group_points_grad_kernel(int b, int m, int c, int nsample, const float *__restrict__ grad_out,
                               const int *__restrict__ idx, const int *__restrict__ idx_cnt,
                               const int *__restrict__ features_cnt, float *__restrict__ grad_features) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)m * nsample * c;
    if (e >= total) return;
    const int ch = (int)(e % c);
    const long long ps = e / c;
    const int s = (int)(ps % nsample), pt = (int)(ps / nsample);
    int istart;
    const int f = stack_frame_of(idx_cnt, b, pt, istart);
    const size_t fstart = (size_t)stack_start_of(features_cnt, f);
    const int k = __ldg(idx + (size_t)pt * nsample + s);
    atomicAdd(grad_features + (fstart + (size_t)k) * c + ch, __ldg(grad_out + ((size_t)pt * c + ch) * nsample + s));
}

// ---------------------------------------------------------------------------------------------------------------
// three nearest known points of the same frame                                      (interpolate_gpu.cu:17-75)
// A CTA takes 256 consecutive unknown rows (they span one or two frames, rarely more); for every frame it spans,
// the frame's known points stream through shared memory in tiles and the threads of that frame scan the tile --
// the reference lets every thread walk global memory on its own.  double bests initialised to 1e40 are equivalent
// to float bests initialised to +inf.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kNNThreads = 256, kNNTile = 1024;

__global__ void __launch_bounds__(kNNThreads)
stack_three_nn_kernel(int b, int n, const float *__restrict__ unknown, const int *__restrict__ unknown_cnt,
                      const float *__restrict__ known, const int *__restrict__ known_cnt, float *__restrict__ dist2,
                      int *__restrict__ idx) {
    __shared__ float tile[kNNTile * 3];
    __shared__ int s_first, s_last;
    const int pt = blockIdx.x * kNNThreads + threadIdx.x;
    const bool live = pt < n;
    int ustart = 0;
    const int myf = stack_frame_of(unknown_cnt, b, live ? pt : n - 1, ustart);
    if (threadIdx.x == 0) s_first = myf;
    if (threadIdx.x == kNNThreads - 1) s_last = myf;   // dead threads report the frame of the last row
    __syncthreads();
    const int f0 = s_first, f1 = s_last;
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (live) { ux = __ldg(unknown + (size_t)pt * 3); uy = __ldg(unknown + (size_t)pt * 3 + 1); uz = __ldg(unknown + (size_t)pt * 3 + 2); }
    float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
    int i1 = 0, i2 = 0, i3 = 0, mystart = 0;
    int kstart = stack_start_of(known_cnt, f0);
    for (int f = f0; f <= f1; ++f) {
        const int kn = __ldg(known_cnt + f);
        const float *kp = known + (size_t)kstart * 3;
        const bool mine = live && f == myf;
        if (mine) mystart = kstart;
        for (int base = 0; base < kn; base += kNNTile) {
            const int tcnt = min(kNNTile, kn - base);
            __syncthreads();
            for (int t = threadIdx.x; t < tcnt * 3; t += kNNThreads) tile[t] = __ldg(kp + (size_t)base * 3 + t);
            __syncthreads();
            if (mine) {
                for (int k = 0; k < tcnt; ++k) {
                    const float d = sqdist_ref(__fsub_rn(ux, tile[k * 3]), __fsub_rn(uy, tile[k * 3 + 1]), __fsub_rn(uz, tile[k * 3 + 2]));
                    const int kk = base + k;
                    if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = kk; }
                    else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = kk; }
                    else if (d < b3) { b3 = d; i3 = kk; }
                }
            }
        }
        kstart += kn;
    }
    if (live) {
        dist2[(size_t)pt * 3] = b1; dist2[(size_t)pt * 3 + 1] = b2; dist2[(size_t)pt * 3 + 2] = b3;
        idx[(size_t)pt * 3] = i1 + mystart; idx[(size_t)pt * 3 + 1] = i2 + mystart; idx[(size_t)pt * 3 + 2] = i3 + mystart;
    }
}

// out[pt, c] = w0 f[idx0, c] + w1 f[idx1, c] + w2 f[idx2, c]                       (interpolate_gpu.cu:100-120)
template <bool VEC4>
__global__ void __launch_bounds__(256)
stack_three_interpolate_kernel(int n, int c, const float *__restrict__ features, const int *__restrict__ idx,
                               const float *__restrict__ weight, float *__restrict__ out) {
    const int per = VEC4 ? c / 4 : c;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n * per) return;
    const int pt = (int)(e / per), q = (int)(e % per);
    const int k0 = __ldg(idx + (size_t)pt * 3), k1 = __ldg(idx + (size_t)pt * 3 + 1), k2 = __ldg(idx + (size_t)pt * 3 + 2);
    const float w0 = __ldg(weight + (size_t)pt * 3), w1 = __ldg(weight + (size_t)pt * 3 + 1), w2 = __ldg(weight + (size_t)pt * 3 + 2);
    if (VEC4) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(features + (size_t)k0 * c) + q);
        const float4 bq = __ldg(reinterpret_cast<const float4 *>(features + (size_t)k1 * c) + q);
        const float4 cq = __ldg(reinterpret_cast<const float4 *>(features + (size_t)k2 * c) + q);
        float4 r;
        r.x = __fmaf_rn(w2, cq.x, __fmaf_rn(w0, a.x, __fmul_rn(w1, bq.x)));
        r.y = __fmaf_rn(w2, cq.y, __fmaf_rn(w0, a.y, __fmul_rn(w1, bq.y)));
        r.z = __fmaf_rn(w2, cq.z, __fmaf_rn(w0, a.z, __fmul_rn(w1, bq.z)));
        r.w = __fmaf_rn(w2, cq.w, __fmaf_rn(w0, a.w, __fmul_rn(w1, bq.w)));
        reinterpret_cast<float4 *>(out + (size_t)pt * c)[q] = r;
    } else {
        out[(size_t)pt * c + q] = __fmaf_rn(w2, __ldg(features + (size_t)k2 * c + q),
                                            __fmaf_rn(w0, __ldg(features + (size_t)k0 * c + q),
                                                      __fmul_rn(w1, __ldg(features + (size_t)k1 * c + q))));
    }
}

__global__ void __launch_bounds__(256)
stack_three_interpolate_grad_kernel(int n, int c, const float *__restrict__ grad_out, const int *__restrict__ idx,
                                    const float *__restrict__ weight, float *__restrict__ grad_features) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n * c) return;
    const int pt = (int)(e / c), ch = (int)(e % c);
    const float g = __ldg(grad_out + e);
#pragma unroll
    for (int k = 0; k < 3; ++k)
        atomicAdd(grad_features + (size_t)__ldg(idx + (size_t)pt * 3 + k) * c + ch, __fmul_rn(g, __ldg(weight + (size_t)pt * 3 + k)));
}

// ---------------------------------------------------------------------------------------------------------------
// deterministic gradients: sort (target row, source element) pairs, one thread per (target row, 4 channels) sums its
// segment in ascending source order (same scheme as det_backward.cu, channel-last layout).
//   mode 0 (grouping):     source e = pt * nsample + s, value grad_out[pt, ch, s]
//   mode 1 (interpolate):  source e = pt * 3 + k,       value grad_out[pt, ch] * weight[e]
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
stack_det_keys_kernel(long long total, int per, int b, const int *__restrict__ idx, const int *__restrict__ idx_cnt,
                      const int *__restrict__ features_cnt, unsigned *__restrict__ keys, unsigned *__restrict__ vals) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    unsigned base = 0u;
    if (idx_cnt) {     // grouping: local index -> global feature row
        int istart;
        const int f = stack_frame_of(idx_cnt, b, (int)(e / per), istart);
        base = (unsigned)stack_start_of(features_cnt, f);
    }
    keys[e] = base + (unsigned)__ldg(idx + e);
    vals[e] = (unsigned)e;
}

__global__ void __launch_bounds__(256)
stack_det_sum_kernel(long long total, int rows, int c, int per, int mode, const unsigned *__restrict__ keys,
                     const unsigned *__restrict__ vals, const float *__restrict__ grad_out,
                     const float *__restrict__ weight, float *__restrict__ grad_features) {
    const int cq = (c + 3) / 4;
    const long long tq = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tq >= (long long)rows * cq) return;
    const long long t = tq / cq;
    const int c0 = (int)(tq % cq) * 4;
    long long lo = 0, hi = total;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if ((long long)__ldg(keys + mid) < t) lo = mid + 1; else hi = mid;
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    bool any = false;
    for (long long e = lo; e < total && (long long)__ldg(keys + e) == t; ++e) {
        const long long src = __ldg(vals + e);
        const long long pt = src / per;
        const int s = (int)(src % per);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (c0 + k < c) {
                const float g = mode == 0 ? __ldg(grad_out + ((size_t)pt * c + c0 + k) * per + s)
                                          : __fmul_rn(__ldg(grad_out + (size_t)pt * c + c0 + k), __ldg(weight + src));
                acc[k] = __fadd_rn(acc[k], g);
            }
        }
        any = true;
    }
    if (!any) return;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (c0 + k < c) grad_features[(size_t)t * c + c0 + k] = __fadd_rn(grad_features[(size_t)t * c + c0 + k], acc[k]);
}

static int stack_scatter_add_det(long long total, int per, int rows, int c, int mode, int b, const float *grad_out,
                                 const int *idx, const int *idx_cnt, const int *features_cnt, const float *weight,
                                 float *grad_features, cudaStream_t st, const char *what) {
    if (total >= 0xffffffffLL || (long long)rows >= 0xffffffffLL) return fail(PDM_ERR_UNSUPPORTED, "%s: more than 2^32 elements", what);
    int bits = 1;
    while ((1LL << bits) < rows) ++bits;
    size_t temp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp, (const unsigned *)nullptr, (unsigned *)nullptr, (const unsigned *)nullptr,
                                    (unsigned *)nullptr, (int)total, 0, bits, st);
    auto align = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t arr = align((size_t)total * 4);
    char *ws = static_cast<char *>(stream_scratch(st, 4 * arr + align(temp)));
    if (!ws) return PDM_ERR_INVALID_ARG;
    unsigned *k_in = (unsigned *)ws, *k_out = (unsigned *)(ws + arr), *v_in = (unsigned *)(ws + 2 * arr), *v_out = (unsigned *)(ws + 3 * arr);
    stack_det_keys_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, per, b, idx, idx_cnt, features_cnt, k_in, v_in);
    count_launch();
    PDM_CHECK_LAUNCH(what);
    PDM_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(ws + 4 * arr, temp, k_in, k_out, v_in, v_out, (int)total, 0, bits, st));
    const long long work = (long long)rows * ((c + 3) / 4);
    stack_det_sum_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(total, rows, c, per, mode, k_out, v_out, grad_out, weight, grad_features);
    count_launch();
    PDM_CHECK_LAUNCH(what);
    return PDM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// voxel query                                                                         (voxel_query_gpu.cu:10-88)
// the (2 z_range + 1)(2 y_range + 1)(2 x_range + 1) voxels around a centre's own in the reference's z, y, x order; a
// voxel holds one point index (or -1); keep the first nsample with d2 <= r^2.
// ---------------------------------------------------------------------------------------------------------------
// One warp per centre: the lanes take 32 consecutive voxels of the neighbourhood in the reference's (z, y, x) loop order, a
// ballot keeps the hits in that order (the reference walks up to (2 z_range + 1)(2 y_range + 1)(2 x_range + 1) voxels with one
// thread: a chain of dependent loads).
__global__ void __launch_bounds__(256)
stack_voxel_query_kernel(int m, int r1, int r2, int r3, int nsample, float radius2, int z_range, int y_range, int x_range,
                         const float *__restrict__ new_xyz, const float *__restrict__ xyz, const int *__restrict__ new_coords,
                         const int *__restrict__ point_indices, int *__restrict__ idx) {
    const int lane = threadIdx.x & 31;
    const int pt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pt >= m) return;
    const float nx = __ldg(new_xyz + (size_t)pt * 3), ny = __ldg(new_xyz + (size_t)pt * 3 + 1), nz = __ldg(new_xyz + (size_t)pt * 3 + 2);
    const int4 co = __ldg(reinterpret_cast<const int4 *>(new_coords) + pt);   // [batch, z, y, x]
    int *row = idx + (size_t)pt * nsample;
    const int wy = 2 * y_range + 1, wx = 2 * x_range + 1;
    const int total = (2 * z_range + 1) * wy * wx;
    int cnt = 0, first = 0;
    for (int base = 0; base < total && cnt < nsample; base += 32) {
        const int v = base + lane;
        int k = -1;
        bool hit = false;
        if (v < total) {
            const int dz = v / (wy * wx) - z_range, rem = v % (wy * wx);
            const int dy = rem / wx - y_range, dx = rem % wx - x_range;
            const int zc = co.y + dz, yc = co.z + dy, xc = co.w + dx;
            if (zc >= 0 && zc < r1 && yc >= 0 && yc < r2 && xc >= 0 && xc < r3) {
                k = __ldg(point_indices + (((size_t)co.x * r1 + zc) * r2 + yc) * r3 + xc);
                if (k >= 0) {
                    const float d2 = sqdist_ref(__fsub_rn(__ldg(xyz + (size_t)k * 3), nx), __fsub_rn(__ldg(xyz + (size_t)k * 3 + 1), ny),
                                                __fsub_rn(__ldg(xyz + (size_t)k * 3 + 2), nz));
                    hit = !(d2 > radius2);
                }
            }
        }
        const unsigned ball = __ballot_sync(kAll, hit);
        if (ball) {
            if (cnt == 0) first = __shfl_sync(kAll, k, __ffs(ball) - 1);
            const int slot = cnt + __popc(ball & ((1u << lane) - 1u));
            if (hit && slot < nsample) row[slot] = k;
            cnt += __popc(ball);
        }
    }
    if (cnt == 0) {
        if (lane == 0) row[0] = -1;
    } else {
        for (int l = min(cnt, nsample) + lane; l < nsample; l += 32) row[l] = first;     // pad with the first hit
    }
}

// ---------------------------------------------------------------------------------------------------------------
// vector-pool family                                                                  (vector_pool_gpu.cu)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool vp_in_range(float lx, float ly, float lz, float dmax, float r2, int neighbor_type) {
    if (neighbor_type == 1) return !(sqdist_ref(lx, ly, lz) > r2);
    return !((fabsf(lx) > dmax) | (fabsf(ly) > dmax) | (fabsf(lz) > dmax));
}

// query_stacked_local_neighbor_idxs_kernel (:98-160): the first min(1000, nsample > 0 ? nsample : inf) support points
// (ascending index) inside the ball / cube around a centre, appended to one stacked list; start_len = (start, count).
// One warp per centre: 32 candidates per step, hits kept in order with a ballot; two passes (count, then write) instead
// of the reference's 1000-entry per-thread array.  The list order across centres is whatever the atomics give (as in the
// reference); inside a centre it is ascending.
__global__ void __launch_bounds__(256)
stack_local_neighbor_idxs_kernel(int b, int m, const float *__restrict__ support_xyz, const int *__restrict__ xyz_cnt,
                                 const float *__restrict__ new_xyz, const int *__restrict__ new_cnt,
                                 int *__restrict__ stack_neighbor_idxs, int *__restrict__ start_len, int *__restrict__ cumsum,
                                 int avg_length, float dmax, int nsample, int neighbor_type) {
    const int lane = threadIdx.x & 31;
    const int pt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pt >= m) return;
    int nstart;
    const int f = stack_frame_of(new_cnt, b, pt, nstart);
    const int xstart = stack_start_of(xyz_cnt, f);
    const int n = __ldg(xyz_cnt + f);
    const float *pts = support_xyz + (size_t)xstart * 3;
    const float nx = __ldg(new_xyz + (size_t)pt * 3), ny = __ldg(new_xyz + (size_t)pt * 3 + 1), nz = __ldg(new_xyz + (size_t)pt * 3 + 2);
    const float r2 = __fmul_rn(dmax, dmax);
    const int limit = nsample > 0 ? min(nsample, 1000) : 1000;
    int cnt = 0;
    for (int base = 0; base < n && cnt < limit; base += 32) {
        const int k = base + lane;
        bool hit = false;
        if (k < n)
            hit = vp_in_range(__fsub_rn(__ldg(pts + (size_t)k * 3), nx), __fsub_rn(__ldg(pts + (size_t)k * 3 + 1), ny),
                              __fsub_rn(__ldg(pts + (size_t)k * 3 + 2), nz), dmax, r2, neighbor_type);
        cnt += __popc(__ballot_sync(kAll, hit));
    }
    cnt = min(cnt, limit);
    int start = 0;
    if (lane == 0) {
        start = atomicAdd(cumsum, cnt);
        start_len[(size_t)pt * 2] = start;
        start_len[(size_t)pt * 2 + 1] = cnt;
    }
    start = __shfl_sync(kAll, start, 0);
    const long long max_thresh = (long long)avg_length * m;
    if (start >= max_thresh) return;
    int wcnt = cnt;
    if ((long long)start + cnt >= max_thresh) wcnt = (int)(max_thresh - start);
    int done = 0;
    for (int base = 0; base < n && done < wcnt; base += 32) {
        const int k = base + lane;
        bool hit = false;
        if (k < n)
            hit = vp_in_range(__fsub_rn(__ldg(pts + (size_t)k * 3), nx), __fsub_rn(__ldg(pts + (size_t)k * 3 + 1), ny),
                              __fsub_rn(__ldg(pts + (size_t)k * 3 + 2), nz), dmax, r2, neighbor_type);
        const unsigned ball = __ballot_sync(kAll, hit);
        const int slot = done + __popc(ball & ((1u << lane) - 1u));
        if (hit && slot < wcnt) stack_neighbor_idxs[(size_t)start + slot] = k + xstart;
        done += __popc(ball);
    }
}

// query_three_nn_by_stacked_local_idxs_kernel (:19-77): per (centre, grid cell) the three nearest of the centre's
// neighbour list to the cell centre; missing second / third neighbours repeat the first.
__global__ void __launch_bounds__(256)
stack_three_nn_local_kernel(int m, int g, const float *__restrict__ support_xyz, const float *__restrict__ grid_centers,
                            int *__restrict__ grid_idxs, float *__restrict__ grid_dist2,
                            const int *__restrict__ stack_neighbor_idxs, const int *__restrict__ start_len) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)m * g) return;
    const int pt = (int)(e / g);
    const float cx = __ldg(grid_centers + e * 3), cy = __ldg(grid_centers + e * 3 + 1), cz = __ldg(grid_centers + e * 3 + 2);
    const int *nb = stack_neighbor_idxs + __ldg(start_len + (size_t)pt * 2);
    const int len = __ldg(start_len + (size_t)pt * 2 + 1);
    float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
    int i1 = -1, i2 = -1, i3 = -1;
    for (int k = 0; k < len; ++k) {
        const int q = __ldg(nb + k);
        const float d = sqdist_ref(__fsub_rn(cx, __ldg(support_xyz + (size_t)q * 3)), __fsub_rn(cy, __ldg(support_xyz + (size_t)q * 3 + 1)),
                                   __fsub_rn(cz, __ldg(support_xyz + (size_t)q * 3 + 2)));
        if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = q; }
        else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = q; }
        else if (d < b3) { b3 = d; i3 = q; }
    }
    if (i2 == -1) { i2 = i1; b2 = b1; }
    if (i3 == -1) { i3 = i1; b3 = b1; }
    grid_dist2[e * 3] = b1; grid_dist2[e * 3 + 1] = b2; grid_dist2[e * 3 + 2] = b3;
    grid_idxs[e * 3] = i1; grid_idxs[e * 3 + 1] = i2; grid_idxs[e * 3 + 2] = i3;
}

// vector_pool_kernel_stack (:183-299).  One warp per centre; its output row (num_c_out sums, 3 G local-xyz sums,
// G counters) lives in shared memory while the warp walks the frame's support points 32 at a time; hits are applied
// in ascending index order -- the reference's serial order, so the fp32 sums are bit-identical -- with the lanes spread
// over the channels of a hit.  grouped_idxs slots are handed out per 32-point step with one atomic.
// The caller zero-fills new_features / new_local_xyz / point_cnt_of_grid (pointnet2_utils.py:397-399); rows are stored.
__global__ void __launch_bounds__(128)
stack_vector_pool_kernel(int b, int m, const float *__restrict__ support_xyz, const float *__restrict__ support_features,
                         const int *__restrict__ xyz_cnt, const float *__restrict__ new_xyz, float *__restrict__ new_features,
                         float *__restrict__ new_local_xyz, const int *__restrict__ new_cnt, int ngx, int ngy, int ngz,
                         float dmax, int c_in, int c_out, int ceg, int g, int *__restrict__ point_cnt_of_grid,
                         int *__restrict__ grouped_idxs, int use_xyz, float gsx, float gsy, float gsz,
                         int *__restrict__ cum_sum, int num_max_sum_points, int nsample, int neighbor_type, int pooling_type) {
    extern __shared__ float vp_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int pt = blockIdx.x * (blockDim.x >> 5) + w;
    if (pt >= m) return;
    const int per_warp = c_out + 4 * g;
    float *acc = vp_smem + (size_t)w * per_warp;      // [c_out]
    float *lxyz = acc + c_out;                          // [3 g]
    int *cntg = reinterpret_cast<int *>(lxyz + 3 * g);  // [g]
    for (int i = lane; i < per_warp; i += 32) acc[i] = 0.f;   // (int 0 == float 0 bit pattern)
    __syncwarp();
    int nstart;
    const int f = stack_frame_of(new_cnt, b, pt, nstart);
    const int xstart = stack_start_of(xyz_cnt, f);
    const int n = __ldg(xyz_cnt + f);
    const float *pts = support_xyz + (size_t)xstart * 3;
    const float *feat = support_features + (size_t)xstart * c_in;
    const float nx = __ldg(new_xyz + (size_t)pt * 3), ny = __ldg(new_xyz + (size_t)pt * 3 + 1), nz = __ldg(new_xyz + (size_t)pt * 3 + 2);
    const float r2 = __fmul_rn(dmax, dmax);
    int sample_cnt = 0;
    bool stop = false;
    for (int base = 0; base < n && !stop; base += 32) {
        const int k = base + lane;
        float lx = 0.f, ly = 0.f, lz = 0.f;
        bool hit = false;
        if (k < n) {
            lx = __fsub_rn(__ldg(pts + (size_t)k * 3), nx);
            ly = __fsub_rn(__ldg(pts + (size_t)k * 3 + 1), ny);
            lz = __fsub_rn(__ldg(pts + (size_t)k * 3 + 2), nz);
            hit = vp_in_range(lx, ly, lz, dmax, r2, neighbor_type);
        }
        int gi = 0;
        if (hit) {
            const int gxi = (int)floorf(__fdiv_rn(__fadd_rn(lx, dmax), gsx));
            const int gyi = (int)floorf(__fdiv_rn(__fadd_rn(ly, dmax), gsy));
            const int gzi = (int)floorf(__fdiv_rn(__fadd_rn(lz, dmax), gsz));
            gi = gxi * ngy * ngz + gyi * ngz + gzi;
            gi = min(max(gi, 0), g - 1);
        }
        unsigned ball = __ballot_sync(kAll, hit);
        // which of the hits are taken (serial rules of the reference), in ascending k
        unsigned take = 0u;
        {
            unsigned rest = ball;
            while (rest && !stop) {
                const int src = __ffs(rest) - 1;
                rest &= rest - 1u;
                const int sg = __shfl_sync(kAll, gi, src);
                if (pooling_type == 0) {
                    take |= 1u << src;
                    if (lane == 0) cntg[sg] += 1;
                    ++sample_cnt;
                    if (nsample > 0 && sample_cnt >= nsample) stop = true;
                } else {
                    __syncwarp();
                    const int have = cntg[sg];
                    __syncwarp();
                    if (have == 0) {
                        take |= 1u << src;
                        if (lane == 0) cntg[sg] = 1;
                        ++sample_cnt;
                        if ((nsample > 0 && sample_cnt >= nsample) || sample_cnt >= g) stop = true;
                    }
                }
            }
            __syncwarp();
        }
        // grouped_idxs slots for the taken hits of this step
        const int ntake = __popc(take);
        int slot0 = 0;
        if (ntake) {
            if (lane == 0) slot0 = atomicAdd(cum_sum, ntake);
            slot0 = __shfl_sync(kAll, slot0, 0);
        }
        if (take & (1u << lane)) {
            const int slot = slot0 + __popc(take & ((1u << lane) - 1u));
            if (slot < num_max_sum_points) {
                grouped_idxs[(size_t)slot * 3] = xstart + k;
                grouped_idxs[(size_t)slot * 3 + 1] = pt;
                grouped_idxs[(size_t)slot * 3 + 2] = gi;
            }
        }
        // apply the taken hits in ascending k; lanes = channels of the cell
        unsigned rest = take;
        while (rest) {
            const int src = __ffs(rest) - 1;
            rest &= rest - 1u;
            const int sg = __shfl_sync(kAll, gi, src);
            const float sx = __shfl_sync(kAll, lx, src), sy = __shfl_sync(kAll, ly, src), sz = __shfl_sync(kAll, lz, src);
            const float *frow = feat + (size_t)(base + src) * c_in;
            float *arow = acc + (size_t)sg * ceg;
            // channel i of the hit goes to slot i % ceg of its cell, in ascending i (the reference's serial loop).  One coalesced
            // load of 32 channels per step; lanes < ceg own a slot and take the values of the lanes i = slot, slot + ceg, ...
            // through shuffles, in that order (ceg >= 32: the 32 channels of a step fall into distinct slots).
            for (int i0 = 0; i0 < c_in; i0 += 32) {
                const int i = i0 + lane;
                const float v = i < c_in ? __ldg(frow + i) : 0.f;
                if (ceg >= 32) {
                    if (i < c_in) { const int sl = i % ceg; arow[sl] = pooling_type == 0 ? __fadd_rn(arow[sl], v) : v; }
                    __syncwarp();
                } else {
                    const int sl = (i0 + lane) % ceg;                    // slot of lane's channel; owner lanes: lane < ceg
                    const int own = (i0 % ceg + lane) % ceg;             // slot owned by this lane in this step (lane < ceg)
                    float a = lane < ceg ? arow[own] : 0.f;
                    (void)sl;
                    for (int q = 0; q * ceg < 32; ++q) {
                        const int src = lane + q * ceg;                  // lane holding the q-th channel of my slot in this step
                        const float o = __shfl_sync(kAll, v, src & 31);
                        if (lane < ceg && src < 32 && i0 + src < c_in) a = pooling_type == 0 ? __fadd_rn(a, o) : o;
                    }
                    if (lane < ceg) arow[own] = a;
                    __syncwarp();
                }
            }
            if (use_xyz && lane < 3) {
                const float v = lane == 0 ? sx : (lane == 1 ? sy : sz);
                lxyz[sg * 3 + lane] = pooling_type == 0 ? __fadd_rn(lxyz[sg * 3 + lane], v) : v;
            }
        }
        __syncwarp();
    }
    __syncwarp();
    for (int i = lane; i < c_out; i += 32) new_features[(size_t)pt * c_out + i] = acc[i];
    if (use_xyz)
        for (int i = lane; i < 3 * g; i += 32) new_local_xyz[(size_t)pt * 3 * g + i] = lxyz[i];
    for (int i = lane; i < g; i += 32) point_cnt_of_grid[(size_t)pt * g + i] = cntg[i];
}

// vector_pool_grad_kernel_stack (:376-401): thread = (list entry, input channel), channel fastest
__global__ void __launch_bounds__(256)
stack_vector_pool_grad_kernel(const float *__restrict__ grad_new_features, const int *__restrict__ point_cnt_of_grid,
                              const int *__restrict__ grouped_idxs, float *__restrict__ grad_support_features,
                              int c_out, int c_in, int ceg, int g, int entries) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)entries * c_in) return;
    const int ent = (int)(e / c_in), ch = (int)(e % c_in);
    const int is = __ldg(grouped_idxs + (size_t)ent * 3), in = __ldg(grouped_idxs + (size_t)ent * 3 + 1), ig = __ldg(grouped_idxs + (size_t)ent * 3 + 2);
    const int npts = __ldg(point_cnt_of_grid + (size_t)in * g + ig);
    const float cur = __frcp_rn(fmaxf((float)npts, 1.0f));
    atomicAdd(grad_support_features + (size_t)is * c_in + ch,
              __fmul_rn(__ldg(grad_new_features + (size_t)in * c_out + (size_t)ig * ceg + ch % ceg), cur));
}

}  // namespace pdm

using namespace pdm;

#define PDM_NEG(cond, what) if (cond) return fail(PDM_ERR_INVALID_ARG, "%s: negative size", what)

extern "C" int pdm_stack_group_points(int b, int m, int c, int nsample, const float *features, const int *features_batch_cnt,
                                      const int *idx, const int *idx_batch_cnt, float *out, void *stream) {
    PDM_NEG(b < 0 || m < 0 || c < 0 || nsample < 0, "stack_group_points");
    if (b == 0 || m == 0 || c == 0 || nsample == 0) return PDM_OK;
    if (!features || !features_batch_cnt || !idx || !idx_batch_cnt || !out) return fail(PDM_ERR_INVALID_ARG, "stack_group_points: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((m + 7) / 8);
    if (c % 4 == 0 && ((uintptr_t)features & 15) == 0)
        stack_group_points_kernel<true><<<grid, 256, 0, st>>>(b, m, c, nsample, features, features_batch_cnt, idx, idx_batch_cnt, out);
    else
        stack_group_points_kernel<false><<<grid, 256, 0, st>>>(b, m, c, nsample, features, features_batch_cnt, idx, idx_batch_cnt, out);
    count_launch();
    PDM_CHECK_LAUNCH("stack_group_points");
    return PDM_OK;
}

extern "C" int pdm_stack_group_points_grad(int b, int m, int c, int n, int nsample, const float *grad_out, const int *idx,
                                           const int *idx_batch_cnt, const int *features_batch_cnt, float *grad_features,
                                           int deterministic, void *stream) {
    PDM_NEG(b < 0 || m < 0 || c < 0 || n < 0 || nsample < 0, "stack_group_points_grad");
    if (b == 0 || m == 0 || c == 0 || nsample == 0 || n == 0) return PDM_OK;
    if (!grad_out || !idx || !idx_batch_cnt || !features_batch_cnt || !grad_features)
        return fail(PDM_ERR_INVALID_ARG, "stack_group_points_grad: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)m * nsample * c;
    if (deterministic)
        return stack_scatter_add_det((long long)m * nsample, nsample, n, c, 0, b, grad_out, idx, idx_batch_cnt, features_batch_cnt,
                                     nullptr, grad_features, st, "stack_group_points_grad_det");
    if ((total + 255) / 256 > 0x7fffffffLL) return fail(PDM_ERR_UNSUPPORTED, "stack_group_points_grad: too large");
    stack_group_points_grad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(b, m, c, nsample, grad_out, idx, idx_batch_cnt,
                                                                                  features_batch_cnt, grad_features);
    count_launch();
    PDM_CHECK_LAUNCH("stack_group_points_grad");
    return PDM_OK;
}

extern "C" int pdm_stack_three_nn(int b, int n, int m, const float *unknown, const int *unknown_batch_cnt, const float *known,
                                  const int *known_batch_cnt, float *dist2, int *idx, void *stream) {
    PDM_NEG(b < 0 || n < 0 || m < 0, "stack_three_nn");
    if (b == 0 || n == 0) return PDM_OK;
    if (!unknown || !unknown_batch_cnt || !known_batch_cnt || !dist2 || !idx || (m > 0 && !known))
        return fail(PDM_ERR_INVALID_ARG, "stack_three_nn: null pointer");
    stack_three_nn_kernel<<<(n + kNNThreads - 1) / kNNThreads, kNNThreads, 0, (cudaStream_t)stream>>>(
        b, n, unknown, unknown_batch_cnt, known, known_batch_cnt, dist2, idx);
    count_launch();
    PDM_CHECK_LAUNCH("stack_three_nn");
    return PDM_OK;
}

extern "C" int pdm_stack_three_interpolate(int n, int c, const float *features, const int *idx, const float *weight, float *out,
                                           void *stream) {
    PDM_NEG(n < 0 || c < 0, "stack_three_interpolate");
    if (n == 0 || c == 0) return PDM_OK;
    if (!features || !idx || !weight || !out) return fail(PDM_ERR_INVALID_ARG, "stack_three_interpolate: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = c % 4 == 0 && (((uintptr_t)features | (uintptr_t)out) & 15) == 0;
    const long long work = (long long)n * (v4 ? c / 4 : c);
    if (v4) stack_three_interpolate_kernel<true><<<(unsigned)((work + 255) / 256), 256, 0, st>>>(n, c, features, idx, weight, out);
    else stack_three_interpolate_kernel<false><<<(unsigned)((work + 255) / 256), 256, 0, st>>>(n, c, features, idx, weight, out);
    count_launch();
    PDM_CHECK_LAUNCH("stack_three_interpolate");
    return PDM_OK;
}

extern "C" int pdm_stack_three_interpolate_grad(int n, int c, int m, const float *grad_out, const int *idx, const float *weight,
                                                float *grad_features, int deterministic, void *stream) {
    PDM_NEG(n < 0 || c < 0 || m < 0, "stack_three_interpolate_grad");
    if (n == 0 || c == 0 || m == 0) return PDM_OK;
    if (!grad_out || !idx || !weight || !grad_features) return fail(PDM_ERR_INVALID_ARG, "stack_three_interpolate_grad: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (deterministic)
        return stack_scatter_add_det((long long)n * 3, 3, m, c, 1, 0, grad_out, idx, nullptr, nullptr, weight, grad_features, st,
                                     "stack_three_interpolate_grad_det");
    const long long work = (long long)n * c;
    stack_three_interpolate_grad_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(n, c, grad_out, idx, weight, grad_features);
    count_launch();
    PDM_CHECK_LAUNCH("stack_three_interpolate_grad");
    return PDM_OK;
}

extern "C" int pdm_stack_voxel_query(int m, int r1, int r2, int r3, int nsample, float radius, int z_range, int y_range,
                                     int x_range, const float *new_xyz, const float *xyz, const int *new_coords,
                                     const int *point_indices, int *idx, void *stream) {
    PDM_NEG(m < 0 || r1 < 0 || r2 < 0 || r3 < 0 || nsample < 0 || z_range < 0 || y_range < 0 || x_range < 0, "stack_voxel_query");
    if (m == 0 || nsample == 0) return PDM_OK;
    if (!new_xyz || !xyz || !new_coords || !point_indices || !idx) return fail(PDM_ERR_INVALID_ARG, "stack_voxel_query: null pointer");
    stack_voxel_query_kernel<<<(m + 7) / 8, 256, 0, (cudaStream_t)stream>>>(m, r1, r2, r3, nsample, radius * radius, z_range,
                                                                              y_range, x_range, new_xyz, xyz, new_coords,
                                                                              point_indices, idx);
    count_launch();
    PDM_CHECK_LAUNCH("stack_voxel_query");
    return PDM_OK;
}

extern "C" int pdm_stack_query_local_neighbor_idxs(int b, int m, const float *support_xyz, const int *xyz_batch_cnt,
                                                   const float *new_xyz, const int *new_xyz_batch_cnt, int *stack_neighbor_idxs,
                                                   int *start_len, int *cumsum, int avg_length_of_neighbor_idxs,
                                                   float max_neighbour_distance, int nsample, int neighbor_type, void *stream) {
    PDM_NEG(b < 0 || m < 0 || avg_length_of_neighbor_idxs < 0, "stack_query_local_neighbor_idxs");
    if (b == 0 || m == 0) return PDM_OK;
    if (!support_xyz || !xyz_batch_cnt || !new_xyz || !new_xyz_batch_cnt || !start_len || !cumsum ||
        (avg_length_of_neighbor_idxs > 0 && !stack_neighbor_idxs))
        return fail(PDM_ERR_INVALID_ARG, "stack_query_local_neighbor_idxs: null pointer");
    stack_local_neighbor_idxs_kernel<<<(m + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
        b, m, support_xyz, xyz_batch_cnt, new_xyz, new_xyz_batch_cnt, stack_neighbor_idxs, start_len, cumsum,
        avg_length_of_neighbor_idxs, max_neighbour_distance, nsample, neighbor_type);
    count_launch();
    PDM_CHECK_LAUNCH("stack_query_local_neighbor_idxs");
    return PDM_OK;
}

extern "C" int pdm_stack_query_three_nn_by_local_idxs(int m, int num_total_grids, const float *support_xyz,
                                                      const float *new_xyz_grid_centers, int *new_xyz_grid_idxs,
                                                      float *new_xyz_grid_dist2, const int *stack_neighbor_idxs,
                                                      const int *start_len, void *stream) {
    PDM_NEG(m < 0 || num_total_grids < 0, "stack_query_three_nn_by_local_idxs");
    if (m == 0 || num_total_grids == 0) return PDM_OK;
    if (!support_xyz || !new_xyz_grid_centers || !new_xyz_grid_idxs || !new_xyz_grid_dist2 || !start_len)
        return fail(PDM_ERR_INVALID_ARG, "stack_query_three_nn_by_local_idxs: null pointer");
    const long long work = (long long)m * num_total_grids;
    stack_three_nn_local_kernel<<<(unsigned)((work + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        m, num_total_grids, support_xyz, new_xyz_grid_centers, new_xyz_grid_idxs, new_xyz_grid_dist2, stack_neighbor_idxs, start_len);
    count_launch();
    PDM_CHECK_LAUNCH("stack_query_three_nn_by_local_idxs");
    return PDM_OK;
}

extern "C" int pdm_stack_vector_pool(int b, int n, int m, int num_c_in, int num_c_out, int num_total_grids,
                                     const float *support_xyz, const int *xyz_batch_cnt, const float *support_features,
                                     const float *new_xyz, const int *new_xyz_batch_cnt, float *new_features,
                                     float *new_local_xyz, int *point_cnt_of_grid, int *grouped_idxs, int num_grid_x,
                                     int num_grid_y, int num_grid_z, float max_neighbour_distance, int use_xyz,
                                     int num_max_sum_points, int nsample, int neighbor_type, int pooling_type,
                                     int *cum_sum_device, void *stream) {
    PDM_NEG(b < 0 || n < 0 || m < 0 || num_c_in < 0 || num_c_out < 0 || num_total_grids <= 0 || num_max_sum_points < 0, "stack_vector_pool");
    if (!cum_sum_device) return fail(PDM_ERR_INVALID_ARG, "stack_vector_pool: null counter");
    cudaStream_t st = (cudaStream_t)stream;
    PDM_CHECK_CUDA(cudaMemsetAsync(cum_sum_device, 0, sizeof(int), st));
    if (b == 0 || m == 0) return PDM_OK;
    if (!support_xyz || !xyz_batch_cnt || !support_features || !new_xyz || !new_xyz_batch_cnt || !new_features || !new_local_xyz ||
        !point_cnt_of_grid || (num_max_sum_points > 0 && !grouped_idxs))
        return fail(PDM_ERR_INVALID_ARG, "stack_vector_pool: null pointer");
    if (num_total_grids != num_grid_x * num_grid_y * num_grid_z || num_c_out % num_total_grids != 0)
        return fail(PDM_ERR_INVALID_ARG, "stack_vector_pool: inconsistent grid / channel counts");
    const int ceg = num_c_out / num_total_grids;                   // vector_pool_gpu.cu:311
    if (ceg == 0 || num_c_in % ceg != 0) return fail(PDM_ERR_INVALID_ARG, "stack_vector_pool: c_in must be a multiple of c_out / grids");
    const float gsx = max_neighbour_distance * 2 / num_grid_x;      // :312-314 (host fp32)
    const float gsy = max_neighbour_distance * 2 / num_grid_y;
    const float gsz = max_neighbour_distance * 2 / num_grid_z;
    const size_t per_warp = (size_t)(num_c_out + 4 * num_total_grids) * sizeof(float);
    int warps = 4;
    while (warps > 1 && warps * per_warp > 96 * 1024) warps >>= 1;
    const size_t smem = warps * per_warp;
    if (smem > 200 * 1024) return fail(PDM_ERR_UNSUPPORTED, "stack_vector_pool: output row of %zu bytes does not fit on chip", per_warp);
    if (int rc = ensure_dynamic_smem((const void *)stack_vector_pool_kernel, smem)) return rc;
    stack_vector_pool_kernel<<<(m + warps - 1) / warps, warps * 32, smem, st>>>(
        b, m, support_xyz, support_features, xyz_batch_cnt, new_xyz, new_features, new_local_xyz, new_xyz_batch_cnt, num_grid_x,
        num_grid_y, num_grid_z, max_neighbour_distance, num_c_in, num_c_out, ceg, num_total_grids, point_cnt_of_grid, grouped_idxs,
        use_xyz, gsx, gsy, gsz, cum_sum_device, num_max_sum_points, nsample, neighbor_type, pooling_type);
    count_launch();
    PDM_CHECK_LAUNCH("stack_vector_pool");
    return PDM_OK;
}

extern "C" int pdm_stack_vector_pool_grad(int m, int num_c_out, int n, int num_c_in, int num_total_grids, int num_entries,
                                          const float *grad_new_features, const int *point_cnt_of_grid,
                                          const int *grouped_idxs, float *grad_support_features, void *stream) {
    PDM_NEG(m < 0 || num_c_out < 0 || n < 0 || num_c_in < 0 || num_total_grids <= 0 || num_entries < 0, "stack_vector_pool_grad");
    if (num_entries == 0 || num_c_in == 0) return PDM_OK;
    if (!grad_new_features || !point_cnt_of_grid || !grouped_idxs || !grad_support_features)
        return fail(PDM_ERR_INVALID_ARG, "stack_vector_pool_grad: null pointer");
    const int ceg = num_c_out / num_total_grids;
    if (ceg == 0) return fail(PDM_ERR_INVALID_ARG, "stack_vector_pool_grad: c_out < grids");
    const long long work = (long long)num_entries * num_c_in;
    stack_vector_pool_grad_kernel<<<(unsigned)((work + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        grad_new_features, point_cnt_of_grid, grouped_idxs, grad_support_features, num_c_out, num_c_in, ceg, num_total_grids, num_entries);
    count_launch();
    PDM_CHECK_LAUNCH("stack_vector_pool_grad");
    return PDM_OK;
}
