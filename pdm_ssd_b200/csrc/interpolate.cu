// interpolate.cu -- feature-propagation ops: three_nn / three_interpolate (+grad).
//
// three_nn (interpolate_gpu.cu:16-59): for every "unknown" point the 3 nearest "known"
// points, strict `<` so the earliest index wins ties.  The reference keeps its running
// bests in double initialised to 1e40 and narrows to float on store; because every
// candidate is a float, that is equivalent to float bests initialised to +inf
// (1e40 narrows to +inf, and `inf < 1e40` is false exactly like `inf < inf`), which is
// what this kernel keeps in registers.  The known points are staged through shared
// memory in tiles (coalesced float4 loads, broadcast LDS reads) instead of every thread
// streaming them from global memory.
#include "common.cuh"

namespace pdm {

constexpr int kNNTile = 1024;  // known points per shared-memory tile (12 KB)

__global__ void __launch_bounds__(256)
three_nn_kernel(int n, int m, const float *__restrict__ unknown, const float *__restrict__ known,
                float *__restrict__ dist2, int *__restrict__ idx) {
    __shared__ float tile[kNNTile * 3];
    const int bi = blockIdx.y;
    const int pi = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = pi < n;
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (live) {
        const float *u = unknown + ((size_t)bi * n + pi) * 3;
        ux = __ldg(u);
        uy = __ldg(u + 1);
        uz = __ldg(u + 2);
    }
    const float *kn = known + (size_t)bi * m * 3;
    float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
    int i1 = 0, i2 = 0, i3 = 0;
    for (int base = 0; base < m; base += kNNTile) {
        const int cnt = min(kNNTile, m - base);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt * 3; t += blockDim.x) tile[t] = __ldg(kn + (size_t)base * 3 + t);
        __syncthreads();
        if (live) {
#pragma unroll 4
            for (int k = 0; k < cnt; ++k) {
                const float d = sqdist_ref(__fsub_rn(ux, tile[k * 3 + 0]), __fsub_rn(uy, tile[k * 3 + 1]),
                                           __fsub_rn(uz, tile[k * 3 + 2]));
                if (d < b3) {  // fast reject: not among the best three
                    const int kk = base + k;
                    if (d < b1) {
                        b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = kk;
                    } else if (d < b2) {
                        b3 = b2; i3 = i2; b2 = d; i2 = kk;
                    } else {
                        b3 = d; i3 = kk;
                    }
                }
            }
        }
    }
    if (live) {
        float *od = dist2 + ((size_t)bi * n + pi) * 3;
        int *oi = idx + ((size_t)bi * n + pi) * 3;
        od[0] = b1; od[1] = b2; od[2] = b3;
        oi[0] = i1; oi[1] = i2; oi[2] = i3;
    }
}

// out[b,c,j] = fma(w2,p2, fma(w0,p0, rn(w1*p1)))  -- the contraction nvcc applies to
// interpolate_gpu.cu:103.  One thread per output column j, looping over a chunk of
// channels so idx/weight are read once per CH channels (reference: once per channel).
template <int CH>
__global__ void __launch_bounds__(256)
three_interpolate_kernel(int c, int m, int n, const float *__restrict__ points,
                         const int *__restrict__ idx, const float *__restrict__ weight,
                         float *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int bi = blockIdx.z;
    const int c0 = blockIdx.y * CH;
    const int *id = idx + ((size_t)bi * n + j) * 3;
    const float *w = weight + ((size_t)bi * n + j) * 3;
    const int a0 = __ldg(id), a1 = __ldg(id + 1), a2 = __ldg(id + 2);
    const float w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    const float *src = points + ((size_t)bi * c + c0) * m;
    float *dst = out + ((size_t)bi * c + c0) * n + j;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        if (c0 + k < c) {
            const float *s = src + (size_t)k * m;
            float t = __fmul_rn(w1, __ldg(s + a1));
            t = __fmaf_rn(w0, __ldg(s + a0), t);
            dst[(size_t)k * n] = __fmaf_rn(w2, __ldg(s + a2), t);
        }
    }
}

// The same from shared memory: a CTA stages CHS channel rows of its frame (m floats each, coalesced 16-byte loads) and then
// serves ALL n output columns from there, 4 consecutive columns per thread.  A 4-byte gather of 32 random known points is 32
// cache-line look-ups in L1 for one warp instruction (the bound of the kernel above: 0.103 ms for 16 x 64 x 16384 outputs, the
// reference's kernel 0.108); from shared memory it is a few bank conflicts, and the stores are 16 bytes wide.
template <int CHS>
__global__ void __launch_bounds__(512)
three_interpolate_smem_kernel(int c, int m, int n, const float *__restrict__ points, const int *__restrict__ idx,
                              const float *__restrict__ weight, float *__restrict__ out) {
    extern __shared__ __align__(16) float rows[];          // [CHS][m]
    const int bi = blockIdx.y, c0 = blockIdx.x * CHS;
    const int nch = min(CHS, c - c0);
    const float *src = points + ((size_t)bi * c + c0) * m;
    for (int t = threadIdx.x; t < nch * m; t += blockDim.x) rows[t] = __ldg(src + t);      // rows are contiguous
    __syncthreads();
    const int *id = idx + (size_t)bi * n * 3;
    const float *w = weight + (size_t)bi * n * 3;
    float *dst = out + ((size_t)bi * c + c0) * n;
    const bool vec = (n & 3) == 0 && (((uintptr_t)dst) & 15) == 0;
    for (int j0 = threadIdx.x * 4; j0 < n; j0 += blockDim.x * 4) {
        int a[4][3];
        float ww[4][3];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int j = min(j0 + q, n - 1);
                a[q][k] = __ldg(id + (size_t)j * 3 + k);
                ww[q][k] = __ldg(w + (size_t)j * 3 + k);
            }
#pragma unroll
        for (int ch = 0; ch < CHS; ++ch) {
            if (ch < nch) {
                const float *r = rows + ch * m;
                float v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    v[q] = __fmaf_rn(ww[q][2], r[a[q][2]], __fmaf_rn(ww[q][0], r[a[q][0]], __fmul_rn(ww[q][1], r[a[q][1]])));
                float *o = dst + (size_t)ch * n + j0;
                if (vec) {
                    st_cs_f4(o, make_float4(v[0], v[1], v[2], v[3]));
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (j0 + q < n) o[q] = v[q];
                }
            }
        }
    }
}

// interpolate_gpu.cu:127-149: three atomicAdds of rn(g*w_k) per (b,c,j).
__global__ void __launch_bounds__(256)
three_interpolate_grad_kernel(int c, int n, int m, const float *__restrict__ grad_out,
                              const int *__restrict__ idx, const float *__restrict__ weight,
                              float *__restrict__ grad_points) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int bi = blockIdx.z, ci = blockIdx.y;
    const float g = __ldg(grad_out + ((size_t)bi * c + ci) * n + j);
    const int *id = idx + ((size_t)bi * n + j) * 3;
    const float *w = weight + ((size_t)bi * n + j) * 3;
    float *gp = grad_points + ((size_t)bi * c + ci) * m;
    atomicAdd(gp + __ldg(id + 0), __fmul_rn(g, __ldg(w + 0)));
    atomicAdd(gp + __ldg(id + 1), __fmul_rn(g, __ldg(w + 1)));
    atomicAdd(gp + __ldg(id + 2), __fmul_rn(g, __ldg(w + 2)));
}

}  // namespace pdm

extern "C" {

int pdm_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                 int *idx, void *stream) {
    using namespace pdm;
    if (b < 0 || n < 0 || m < 0) return fail(PDM_ERR_INVALID_ARG, "three_nn: negative size");
    if (b == 0 || n == 0) return PDM_OK;
    if (!unknown || !dist2 || !idx || (m > 0 && !known))
        return fail(PDM_ERR_INVALID_ARG, "three_nn: null pointer");
    if (b > 65535) return fail(PDM_ERR_UNSUPPORTED, "three_nn: batch %d > 65535", b);
    dim3 grid((n + 255) / 256, b);
    three_nn_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, m, unknown, known, dist2, idx);
    count_launch();
    PDM_CHECK_LAUNCH("three_nn");
    return PDM_OK;
}

int pdm_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                          const float *weight, float *out, void *stream) {
    using namespace pdm;
    if (b < 0 || c < 0 || n < 0 || m < 0) return fail(PDM_ERR_INVALID_ARG, "three_interpolate: negative size");
    if (b == 0 || c == 0 || n == 0) return PDM_OK;
    if (!points || !idx || !weight || !out) return fail(PDM_ERR_INVALID_ARG, "three_interpolate: null pointer");
    constexpr int CH = 8;
    const int chunks = (c + CH - 1) / CH;
    if (b > 65535 || chunks > 65535) return fail(PDM_ERR_UNSUPPORTED, "three_interpolate: b/c too large");
    // channel rows staged in shared memory when 4 (or 2) of them fit 64 KB and there are enough columns to amortise the staging
    static const bool smem_on = [] { const char *e = getenv("PDM_INTERP_SMEM"); return !(e && e[0] == '0'); }();
    if (smem_on && m > 0 && n >= m && n >= 1024 && ((uintptr_t)points & 15) == 0) {
        const int chs = (size_t)4 * m * 4 <= 64 * 1024 ? 4 : ((size_t)2 * m * 4 <= 64 * 1024 ? 2 : 0);
        if (chs) {
            const size_t smem = (size_t)chs * m * 4;
            dim3 g2((c + chs - 1) / chs, b);
            if (chs == 4) {
                if (int rc = ensure_dynamic_smem((const void *)three_interpolate_smem_kernel<4>, smem)) return rc;
                three_interpolate_smem_kernel<4><<<g2, 512, smem, (cudaStream_t)stream>>>(c, m, n, points, idx, weight, out);
            } else {
                if (int rc = ensure_dynamic_smem((const void *)three_interpolate_smem_kernel<2>, smem)) return rc;
                three_interpolate_smem_kernel<2><<<g2, 512, smem, (cudaStream_t)stream>>>(c, m, n, points, idx, weight, out);
            }
            count_launch();
            PDM_CHECK_LAUNCH("three_interpolate(smem)");
            return PDM_OK;
        }
    }
    dim3 grid((n + 255) / 256, chunks, b);
    three_interpolate_kernel<CH><<<grid, 256, 0, (cudaStream_t)stream>>>(c, m, n, points, idx, weight, out);
    count_launch();
    PDM_CHECK_LAUNCH("three_interpolate");
    return PDM_OK;
}

int pdm_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                               const float *weight, float *grad_points, void *stream) {
    using namespace pdm;
    if (b < 0 || c < 0 || n < 0 || m < 0)
        return fail(PDM_ERR_INVALID_ARG, "three_interpolate_grad: negative size");
    if (b == 0 || c == 0 || n == 0) return PDM_OK;
    if (!grad_out || !idx || !weight || !grad_points)
        return fail(PDM_ERR_INVALID_ARG, "three_interpolate_grad: null pointer");
    if (b > 65535 || c > 65535) return fail(PDM_ERR_UNSUPPORTED, "three_interpolate_grad: b/c > 65535");
    dim3 grid((n + 255) / 256, c, b);
    three_interpolate_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c, n, m, grad_out, idx, weight,
                                                                          grad_points);
    count_launch();
    PDM_CHECK_LAUNCH("three_interpolate_grad");
    return PDM_OK;
}

}  // extern "C"
