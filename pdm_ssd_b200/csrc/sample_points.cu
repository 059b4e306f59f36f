// sample_points.cu -- input staging on the GPU: DataProcessor.sample_points + the `points` part of collate_batch.
//
// Reference (host, numpy, per frame): pcdet/datasets/processor/data_processor.py:182-212 --
//   more points than NUM_POINTS: keep every far point (depth >= 40 m) and a random subset of the near ones without
//   replacement (or, when the far points alone exceed NUM_POINTS, a random subset of everything); fewer: keep all and
//   pad with a random choice without replacement; then shuffle --
// followed by DatasetTemplate.collate_batch (pcdet/datasets/dataset.py:237-244): frames concatenated with the batch
// index prepended as column 0.  At > 5 k frames/s the numpy choice / shuffle / pad of 16 frames per batch and the
// pageable copy behind it bound the pipeline; here the RAW ragged frames are uploaded once and
// sampled on the device.
//
// numpy's Mersenne-Twister stream cannot be reproduced, so the random choices are DEFINED by a counter-based hash
// (parity unpinned against the reference's RNG; oracle/sample_points_oracle.py restates this definition in numpy and the
// tests hold the kernel to it bit for bit, plus the reference's invariants):
//   k(stream, frame, i) = fmix32(seed ^ (frame+1)*0x9E3779B9 ^ (i+1)*0x85EBCA6B ^ stream*0xC2B2AE35)       (murmur3 finaliser)
//   d2 = (x*x + y*y) + z*z in fp32 (no fma);  far = d2 >= 1600
//   n > N and N > n_far : order the points by (far ? 0 : 1, k(1, frame, i)), stable; take the first N
//   n > N and N <= n_far: order by k(1, frame, i), stable; take the first N
//   n <= N              : all n points in their order, then the first N - n of the k(1)-ordered list (cyclically if N - n > n)
//   shuffle             : order the N selected entries by k(2, frame, j), j = position in the list above, stable
// Both orderings are stable LSB radix sorts over all frames at once (cub::DeviceSegmentedSort, library plumbing for a
// staging op); everything else is three small kernels.  No host synchronisation.
#include <cub/device/device_segmented_sort.cuh>

#include "common.cuh"

namespace pdm {

__host__ __device__ __forceinline__ unsigned sp_fmix32(unsigned h) {
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}
__host__ __device__ __forceinline__ unsigned sp_key(unsigned seed, unsigned stream, unsigned frame, unsigned i) {
    return sp_fmix32(seed ^ ((frame + 1u) * 0x9E3779B9u) ^ ((i + 1u) * 0x85EBCA6Bu) ^ (stream * 0xC2B2AE35u));
}

// offsets[b] = sum(counts[:b]) (b <= 65535 frames: one thread), n_far zeroed
__global__ void sp_offsets_kernel(int b, const int *__restrict__ counts, int *__restrict__ offsets, int *__restrict__ nfar) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int acc = 0;
        for (int i = 0; i < b; ++i) { offsets[i] = acc; acc += max(counts[i], 0); nfar[i] = 0; }
        offsets[b] = acc;
    }
}

// frame of raw point e (binary search in offsets), far flag, n_far per frame
__global__ void __launch_bounds__(256)
sp_far_kernel(int b, int c, const float *__restrict__ pts, const int *__restrict__ offsets, unsigned char *__restrict__ far,
              int *__restrict__ frame_of, int *__restrict__ nfar) {
    const int total = __ldg(offsets + b);
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    int lo = 0, hi = b;                     // largest f with offsets[f] <= e
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid) <= e) lo = mid; else hi = mid;
    }
    const float *p = pts + (size_t)e * c;
    const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    const bool f = d2 >= 1600.0f;
    far[e] = f ? 1 : 0;
    frame_of[e] = lo;
    if (f) atomicAdd(nfar + lo, 1);         // integer count: order-independent
}

__global__ void __launch_bounds__(256)
sp_keys1_kernel(int b, int n_out, unsigned seed, const int *__restrict__ offsets, const unsigned char *__restrict__ far,
                const int *__restrict__ frame_of, const int *__restrict__ nfar, unsigned long long *__restrict__ keys,
                int *__restrict__ vals) {
    const int total = __ldg(offsets + b);
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int f = __ldg(frame_of + e);
    const int i = e - __ldg(offsets + f);
    const int n = __ldg(offsets + f + 1) - __ldg(offsets + f);
    const bool keep_far_first = n > n_out && n_out > __ldg(nfar + f);
    const unsigned long long cls = (keep_far_first && !far[e]) ? 1ull : 0ull;
    keys[e] = (cls << 32) | sp_key(seed, 1u, (unsigned)f, (unsigned)i);
    vals[e] = i;
}

__global__ void sp_segments_kernel(int b, int n, int *__restrict__ o) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= b) o[i] = i * n;
}

// selection list of frame f (N entries) + shuffle keys
__global__ void __launch_bounds__(256)
sp_select_kernel(int b, int n_out, unsigned seed, const int *__restrict__ offsets, const int *__restrict__ sorted_idx,
                 unsigned long long *__restrict__ keys2, int *__restrict__ sel) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)b * n_out) return;
    const int f = (int)(t / n_out), j = (int)(t - (long long)f * n_out);
    const int off = __ldg(offsets + f), n = __ldg(offsets + f + 1) - off;
    int s = -1;                                                  // empty frame: no source point
    if (n > n_out) s = __ldg(sorted_idx + off + j);
    else if (n > 0) s = j < n ? j : __ldg(sorted_idx + off + (j - n) % n);
    sel[t] = s;
    keys2[t] = sp_key(seed, 2u, (unsigned)f, (unsigned)j);
}

__global__ void __launch_bounds__(256)
sp_gather_kernel(int b, int n_out, int c, const float *__restrict__ pts, const int *__restrict__ offsets,
                 const int *__restrict__ shuffled, float *__restrict__ out, int *__restrict__ choice) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)b * n_out) return;
    const int f = (int)(t / n_out);
    const int s = __ldg(shuffled + t);
    float *o = out + (size_t)t * (c + 1);
    o[0] = (float)f;                                             // collate_batch: batch index column (dataset.py:240-243)
    if (s >= 0) {
        const float *p = pts + ((size_t)__ldg(offsets + f) + s) * c;
        for (int k = 0; k < c; ++k) o[1 + k] = __ldg(p + k);
    } else {
        for (int k = 0; k < c; ++k) o[1 + k] = 0.f;
    }
    if (choice) choice[t] = s;
}

}  // namespace pdm

extern "C" int pdm_sample_points(int b, int total_points, int c, int num_points, unsigned seed, const float *points,
                                 const int *counts, float *out, int *choice, void *stream) {
    using namespace pdm;
    if (b < 0 || total_points < 0 || c < 3 || num_points <= 0) return fail(PDM_ERR_INVALID_ARG, "sample_points: bad size (need >= 3 channels)");
    if (b == 0) return PDM_OK;
    if (b > 65535) return fail(PDM_ERR_UNSUPPORTED, "sample_points: more than 65535 frames");
    if (!counts || !out || (total_points > 0 && !points)) return fail(PDM_ERR_INVALID_ARG, "sample_points: null pointer");
    const long long sel_total = (long long)b * num_points;
    if (sel_total >= 0x7fffffffLL) return fail(PDM_ERR_UNSUPPORTED, "sample_points: B * num_points must fit int32");
    cudaStream_t st = (cudaStream_t)stream;
    auto align = [](size_t v) { return (v + 255) / 256 * 256; };
    size_t t1 = 0, t2 = 0;
    cub::DeviceSegmentedSort::StableSortPairs(nullptr, t1, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                              (const int *)nullptr, (int *)nullptr, total_points, b, (const int *)nullptr, (const int *)nullptr, st);
    cub::DeviceSegmentedSort::StableSortPairs(nullptr, t2, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                              (const int *)nullptr, (int *)nullptr, (int)sel_total, b, (const int *)nullptr, (const int *)nullptr, st);
    const size_t nmax = (size_t)(total_points > sel_total ? total_points : sel_total) + 1;
    const size_t sz_k = align(nmax * 8), sz_i = align(nmax * 4), sz_off = align((size_t)(2 * b + 2) * 4), sz_far = align(nmax);
    const size_t temp = align(t1 > t2 ? t1 : t2);
    char *ws = static_cast<char *>(stream_scratch(st, 2 * sz_k + 4 * sz_i + 2 * sz_off + sz_far + temp));
    if (!ws) return PDM_ERR_INVALID_ARG;
    char *q = ws;
    auto take = [&](size_t bytes) { char *r = q; q += bytes; return r; };
    unsigned long long *k_in = (unsigned long long *)take(sz_k), *k_out = (unsigned long long *)take(sz_k);
    int *v_in = (int *)take(sz_i), *v_out = (int *)take(sz_i), *frame_of = (int *)take(sz_i), *sel = (int *)take(sz_i);
    int *offsets = (int *)take(sz_off), *seg2 = (int *)take(sz_off);
    unsigned char *far = (unsigned char *)take(sz_far);
    void *tmp = take(temp);
    int *nfar = offsets + b + 1;
    sp_offsets_kernel<<<1, 32, 0, st>>>(b, counts, offsets, nfar);
    count_launch();
    PDM_CHECK_LAUNCH("sample_points(offsets)");
    if (total_points > 0) {
        const unsigned g = (unsigned)((total_points + 255) / 256);
        sp_far_kernel<<<g, 256, 0, st>>>(b, c, points, offsets, far, frame_of, nfar);
        sp_keys1_kernel<<<g, 256, 0, st>>>(b, num_points, seed, offsets, far, frame_of, nfar, k_in, v_in);
        count_launch(2);
        PDM_CHECK_LAUNCH("sample_points(keys)");
        PDM_CHECK_CUDA(cub::DeviceSegmentedSort::StableSortPairs(tmp, t1, k_in, k_out, v_in, v_out, total_points, b, offsets, offsets + 1, st));
    }
    const unsigned g2 = (unsigned)((sel_total + 255) / 256);
    sp_select_kernel<<<g2, 256, 0, st>>>(b, num_points, seed, offsets, v_out, k_in, sel);
    count_launch();
    PDM_CHECK_LAUNCH("sample_points(select)");
    // second ordering: B equal segments of num_points
    sp_segments_kernel<<<(b + 256) / 256, 256, 0, st>>>(b, num_points, seg2);
    count_launch();
    PDM_CHECK_LAUNCH("sample_points(segments)");
    PDM_CHECK_CUDA(cub::DeviceSegmentedSort::StableSortPairs(tmp, t2, k_in, k_out, sel, v_in, (int)sel_total, b, seg2, seg2 + 1, st));
    sp_gather_kernel<<<g2, 256, 0, st>>>(b, num_points, c, points, offsets, v_in, out, choice);
    count_launch();
    PDM_CHECK_LAUNCH("sample_points(gather)");
    return PDM_OK;
}
