// ball_query.cu -- fixed-radius neighbour query of the set-abstraction path.
//
// Semantics (ball_query_gpu.cu:15-51): for each centre scan the points in ascending
// index k, keep the first `nsample` with d2 < radius^2 (d2 rounded as sqdist_ref,
// radius^2 = rn(radius*radius) in fp32), pad the row with the first hit, and leave a
// row with no hit untouched.
//
// v0 kernel: one thread per centre like the reference, but the points stream through
// shared memory in 12 KB tiles (one coalesced load per CTA instead of one broadcast
// global load per thread and point) and a CTA stops as soon as all of its centres are
// full.  Results are identical by construction: same scan order, same arithmetic.
#include "common.cuh"

namespace pdm {

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

}  // namespace pdm

extern "C" int pdm_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                              const float *xyz, int *idx, void *stream) {
    using namespace pdm;
    if (b < 0 || n < 0 || m < 0 || nsample < 0) return fail(PDM_ERR_INVALID_ARG, "ball_query: negative size");
    if (b == 0 || m == 0 || nsample == 0 || n == 0) return PDM_OK;
    if (!new_xyz || !xyz || !idx) return fail(PDM_ERR_INVALID_ARG, "ball_query: null pointer");
    if (b > 65535) return fail(PDM_ERR_UNSUPPORTED, "ball_query: batch %d > 65535", b);
    const float radius2 = radius * radius;  // fp32, as ball_query_gpu.cu:29
    dim3 grid((m + 127) / 128, b);
    ball_query_tiled_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(n, m, radius2, nsample, new_xyz, xyz, idx);
    count_launch();
    PDM_CHECK_LAUNCH("ball_query");
    return PDM_OK;
}
