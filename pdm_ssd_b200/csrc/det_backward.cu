// det_backward.cu -- deterministic backward of gather / group / three_interpolate.
//
// The reference accumulates these gradients with atomicAdd (sampling_gpu.cu:53-70, group_points_gpu.cu:14-31,
// interpolate_gpu.cu:127-149): the fp32 summation order -- and so the last bits of the gradient -- changes from run
// to run.  Here every target element sums its contributions in ONE fixed order (ascending source position):
//   1. key[e] = batch * n + idx[e] for every source element e, value[e] = e;
//   2. a stable LSB radix sort by key (cub::DeviceRadixSort, library plumbing for a non-hot-path op) leaves the
//      sources of each target contiguous and in ascending e;
//   3. one thread per (target, channel chunk) finds its segment by binary search and adds it up serially.
// Same "+=" contract as the atomic kernels (the caller zero-fills, pointnet2_utils.py:68,148,192).
// Chosen by the Python Functions when torch.are_deterministic_algorithms_enabled().
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace pdm {

__global__ void __launch_bounds__(256)
det_keys_kernel(long long total, long long cols, int n, const int *__restrict__ idx, unsigned *__restrict__ keys,
                unsigned *__restrict__ vals) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long long bi = e / cols;
    keys[e] = (unsigned)(bi * n + __ldg(idx + e));
    vals[e] = (unsigned)e;
}

// grad_points[b, c, t] += sum over the sources e of target (b, t), ascending e, of grad_out[b, c, (e % cols) / div] * w[e]
template <int CH>
__global__ void __launch_bounds__(256)
det_segment_sum_kernel(long long total, long long cols, int div, int c, int n, long long targets,
                       const unsigned *__restrict__ keys, const unsigned *__restrict__ vals,
                       const float *__restrict__ grad_out, const float *__restrict__ weight, float *__restrict__ grad_points) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= targets) return;
    const int c0 = blockIdx.y * CH;
    long long lo = 0, hi = total;                  // first e with keys[e] >= t
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if ((long long)__ldg(keys + mid) < t) lo = mid + 1; else hi = mid;
    }
    const long long bi = t / n;
    const long long src_cols = cols / div;
    float acc[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) acc[k] = 0.f;
    bool any = false;
    for (long long e = lo; e < total && (long long)__ldg(keys + e) == t; ++e) {
        const long long src = __ldg(vals + e);
        const long long j = (src - bi * cols) / div;
        const float w = weight ? __ldg(weight + src) : 1.f;
        const float *g = grad_out + ((size_t)bi * c + c0) * src_cols + j;
#pragma unroll
        for (int k = 0; k < CH; ++k)
            if (c0 + k < c) acc[k] = __fadd_rn(acc[k], weight ? __fmul_rn(__ldg(g + (size_t)k * src_cols), w) : __ldg(g + (size_t)k * src_cols));
        any = true;
    }
    if (!any) return;
    float *dst = grad_points + ((size_t)bi * c + c0) * n + (t - bi * n);
#pragma unroll
    for (int k = 0; k < CH; ++k)
        if (c0 + k < c) dst[(size_t)k * n] = __fadd_rn(dst[(size_t)k * n], acc[k]);
}

static int scatter_add_det(int b, int c, int n, long long cols, int div, const float *grad_out, const int *idx,
                           const float *weight, float *grad_points, cudaStream_t st, const char *what) {
    if (b < 0 || c < 0 || n < 0 || cols < 0) return fail(PDM_ERR_INVALID_ARG, "%s: negative size", what);
    const long long total = (long long)b * cols, targets = (long long)b * n;
    if (total == 0 || c == 0 || n == 0) return PDM_OK;
    if (!grad_out || !idx || !grad_points) return fail(PDM_ERR_INVALID_ARG, "%s: null pointer", what);
    if (total >= 0xffffffffLL || targets >= 0xffffffffLL) return fail(PDM_ERR_UNSUPPORTED, "%s: more than 2^32 elements", what);
    int bits = 1;
    while ((1LL << bits) < targets) ++bits;
    size_t temp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp, (const unsigned *)nullptr, (unsigned *)nullptr, (const unsigned *)nullptr,
                                    (unsigned *)nullptr, (int)total, 0, bits, st);
    auto align = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t arr = align((size_t)total * 4);
    char *ws = static_cast<char *>(stream_scratch(st, 4 * arr + align(temp)));
    if (!ws) return PDM_ERR_INVALID_ARG;
    unsigned *k_in = (unsigned *)ws, *k_out = (unsigned *)(ws + arr), *v_in = (unsigned *)(ws + 2 * arr), *v_out = (unsigned *)(ws + 3 * arr);
    det_keys_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, cols, n, idx, k_in, v_in);
    count_launch();
    PDM_CHECK_LAUNCH(what);
    PDM_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(ws + 4 * arr, temp, k_in, k_out, v_in, v_out, (int)total, 0, bits, st));
    constexpr int CH = 8;
    const int chunks = (c + CH - 1) / CH;
    if (chunks > 65535) return fail(PDM_ERR_UNSUPPORTED, "%s: too many channels", what);
    dim3 grid((unsigned)((targets + 255) / 256), chunks);
    det_segment_sum_kernel<CH><<<grid, 256, 0, st>>>(total, cols, div, c, n, targets, k_out, v_out, grad_out, weight, grad_points);
    count_launch();
    PDM_CHECK_LAUNCH(what);
    return PDM_OK;
}

}  // namespace pdm

extern "C" {

int pdm_gather_points_grad_det(int b, int c, int n, int npoints, const float *grad_out, const int *idx, float *grad_points,
                               void *stream) {
    return pdm::scatter_add_det(b, c, n, npoints, 1, grad_out, idx, nullptr, grad_points, (cudaStream_t)stream, "gather_points_grad_det");
}

int pdm_group_points_grad_det(int b, int c, int n, int npoints, int nsample, const float *grad_out, const int *idx,
                              float *grad_points, void *stream) {
    return pdm::scatter_add_det(b, c, n, (long long)npoints * nsample, 1, grad_out, idx, nullptr, grad_points, (cudaStream_t)stream,
                                "group_points_grad_det");
}

int pdm_three_interpolate_grad_det(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight,
                                   float *grad_points, void *stream) {
    if (!weight && b > 0 && c > 0 && n > 0) return pdm::fail(PDM_ERR_INVALID_ARG, "three_interpolate_grad_det: null pointer");
    // source e = 3 j + k contributes grad_out[b, c, j] * weight[b, j, k] to target idx[b, j, k] of the m known points
    return pdm::scatter_add_det(b, c, m, (long long)n * 3, 3, grad_out, idx, weight, grad_points, (cudaStream_t)stream,
                                "three_interpolate_grad_det");
}

}  // extern "C"
