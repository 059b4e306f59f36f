// gather_group.cu -- index gathers of the set-abstraction path and their gradients.
//
//   gather_points : out[b,c,m]   = points[b,c,idx[b,m]]        (sampling_gpu.cu:15-31)
//   group_points  : out[b,c,p,s] = points[b,c,idx[b,p,s]]      (group_points_gpu.cu:53-72)
//
// Both are pure copies, so results are bit-exact by construction.  They are bound by
// the HBM write of `out` (group: 4*C*M*S bytes per frame); the reads are 4-byte random
// gathers out of a (C,N) slab that is L2-resident.  Layout of the work: one thread owns
// VEC=4 consecutive output columns j (16-byte index load, 16-byte streaming stores) and
// walks a chunk of CH channels with all CH*VEC gathers issued before the first store,
// so each thread keeps up to 32 independent L2 requests in flight.  The reference
// re-reads idx once per channel (grid.y = C); here it is read once per CH channels.
#include <stdlib.h>

#include "common.cuh"

namespace pdm {

template <int CH>
__global__ void __launch_bounds__(256)
group_points_vec4_kernel(int c, int n, int cols4, const float *__restrict__ points,
                         const int *__restrict__ idx, float *__restrict__ out) {
    // grid: x = column quads, y = channel chunk, z = batch
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= cols4) return;
    const int bi = blockIdx.z;
    const int c0 = blockIdx.y * CH;
    const size_t cols = (size_t)cols4 * 4;
    const int4 id = __ldg(reinterpret_cast<const int4 *>(idx + (size_t)bi * cols) + q);
    const float *src = points + ((size_t)bi * c + c0) * n;
    float *dst = out + ((size_t)bi * c + c0) * cols + (size_t)q * 4;
    float4 v[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        if (c0 + k < c) {
            const float *s = src + (size_t)k * n;
            v[k].x = __ldg(s + id.x);
            v[k].y = __ldg(s + id.y);
            v[k].z = __ldg(s + id.z);
            v[k].w = __ldg(s + id.w);
        }
    }
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        if (c0 + k < c) st_cs_f4(dst + (size_t)k * cols, v[k]);
    }
}

// Scalar variant for column counts that are not a multiple of 4 (or unaligned bases).
template <int CH>
__global__ void __launch_bounds__(256)
group_points_scalar_kernel(int c, int n, size_t cols, const float *__restrict__ points,
                           const int *__restrict__ idx, float *__restrict__ out) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    const int bi = blockIdx.z;
    const int c0 = blockIdx.y * CH;
    const int id = __ldg(idx + (size_t)bi * cols + j);
    const float *src = points + ((size_t)bi * c + c0) * n;
    float *dst = out + ((size_t)bi * c + c0) * cols + j;
    float v[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k)
        if (c0 + k < c) v[k] = __ldg(src + (size_t)k * n + id);
#pragma unroll
    for (int k = 0; k < CH; ++k)
        if (c0 + k < c) st_cs_f1(dst + (size_t)k * cols, v[k]);
}

// grad: grad_points[b,c,idx[b,j]] += grad_out[b,c,j]   (atomicAdd, as the reference:
// group_points_gpu.cu:14-31, sampling_gpu.cu:53-70; summation order is unspecified there too)
__global__ void __launch_bounds__(256)
scatter_add_kernel(int c, int n, size_t cols, const float *__restrict__ grad_out,
                   const int *__restrict__ idx, float *__restrict__ grad_points) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    const int bi = blockIdx.z, ci = blockIdx.y;
    const int id = __ldg(idx + (size_t)bi * cols + j);
    atomicAdd(grad_points + ((size_t)bi * c + ci) * n + id,
              __ldg(grad_out + ((size_t)bi * c + ci) * cols + j));
}

// QueryAndGroup.forward after the ball query, in one pass (pointnet2_utils.py:250-257):
//   out[b, 0:3, p, s]   = xyz[b, idx[b,p,s], :] - new_xyz[b, p, :]
//   out[b, 3+c, p, s]   = features[b, c, idx[b,p,s]]
// The reference materialises xyz^T, groups it, subtracts the centres in place, groups the features
// and concatenates: five full passes over (B, C, npoint, nsample) tensors.  Here the concatenated
// tensor is written once.  grid.y = 0 handles the three xyz channels, grid.y >= 1 feature chunks.
template <int CH>
__global__ void __launch_bounds__(256)
query_group_kernel(int c, int n, int npoints, int nsample, int use_xyz, const float *__restrict__ xyz,
                   const float *__restrict__ new_xyz, const float *__restrict__ feats,
                   const int *__restrict__ idx, float *__restrict__ out) {
    const size_t cols = (size_t)npoints * nsample;
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    const int bi = blockIdx.z;
    const int ctot = (use_xyz ? 3 : 0) + c;
    const int id = __ldg(idx + (size_t)bi * cols + j);
    float *dst = out + (size_t)bi * ctot * cols + j;
    if (use_xyz && blockIdx.y == 0) {
        const int p = (int)(j / nsample);
        const float *pp = xyz + ((size_t)bi * n + id) * 3;
        const float *qq = new_xyz + ((size_t)bi * npoints + p) * 3;
#pragma unroll
        for (int a = 0; a < 3; ++a) st_cs_f1(dst + (size_t)a * cols, __fsub_rn(__ldg(pp + a), __ldg(qq + a)));
        return;
    }
    const int c0 = ((int)blockIdx.y - (use_xyz ? 1 : 0)) * CH;
    const float *src = feats + ((size_t)bi * c + c0) * n + id;
    dst += (size_t)((use_xyz ? 3 : 0) + c0) * cols;
    float v[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k)
        if (c0 + k < c) v[k] = __ldg(src + (size_t)k * n);
#pragma unroll
    for (int k = 0; k < CH; ++k)
        if (c0 + k < c) st_cs_f1(dst + (size_t)k * cols, v[k]);
}

// Shared-memory variant for frames whose channel rows fit on chip (n <= 4096 with CHS = 2: SA2).
// The gathers of the kernel above are 4-byte reads of 32 different cache lines per warp instruction,
// and L1 looks up one line per cycle: 33.5 M gathered elements of SA2's (16,64,1024,32) tensor cost
// ~115 us of L1 time for 25 us worth of HBM writes.  Here a CTA stages CHS channel rows of its frame
// (or the frame's xyz) in shared memory with coalesced loads and gathers from there (a few bank
// conflicts instead of 32 line look-ups), loops over all columns, 4 per thread (16-byte index loads
// and 16-byte streaming stores).  grid.x = feature chunks [+ 1 for xyz], grid.y = frame.
template <int CHS>
__global__ void __launch_bounds__(512)
query_group_smem_kernel(int c, int n, int npoints, int nsample, int use_xyz, const float *__restrict__ xyz,
                        const float *__restrict__ new_xyz, const float *__restrict__ feats,
                        const int *__restrict__ idx, float *__restrict__ out) {
    extern __shared__ __align__(16) float qg_rows[];   // CHS * n floats, or 3 * n for the xyz CTA
    const int bi = blockIdx.y, tid = threadIdx.x;
    const size_t cols = (size_t)npoints * nsample;
    const int cols4 = (int)(cols / 4);                 // host guarantees cols % 4 == 0, nsample % 4 == 0
    const int ctot = (use_xyz ? 3 : 0) + c;
    const int nchunks = (c + CHS - 1) / CHS;
    const int4 *idx4 = reinterpret_cast<const int4 *>(idx + (size_t)bi * cols);
    if ((int)blockIdx.x == nchunks) {                  // xyz - centre, three channels
        const float *src = xyz + (size_t)bi * n * 3;
        for (int t = tid; t < n * 3; t += blockDim.x) qg_rows[t] = __ldg(src + t);
        __syncthreads();
        float *dst = out + (size_t)bi * ctot * cols;
        for (int q = tid; q < cols4; q += blockDim.x) {
            const int4 id = __ldg(idx4 + q);
            const float *cc = new_xyz + ((size_t)bi * npoints + (size_t)(q * 4) / nsample) * 3;   // the 4 columns share a centre
            const float cx = __ldg(cc), cy = __ldg(cc + 1), cz = __ldg(cc + 2);
            const int ids[4] = {id.x, id.y, id.z, id.w};
            float v[3][4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[0][k] = __fsub_rn(qg_rows[ids[k] * 3 + 0], cx);
                v[1][k] = __fsub_rn(qg_rows[ids[k] * 3 + 1], cy);
                v[2][k] = __fsub_rn(qg_rows[ids[k] * 3 + 2], cz);
            }
#pragma unroll
            for (int a = 0; a < 3; ++a)
                st_cs_f4(dst + (size_t)a * cols + (size_t)q * 4, make_float4(v[a][0], v[a][1], v[a][2], v[a][3]));
        }
        return;
    }
    const int c0 = blockIdx.x * CHS;
    const int cn = min(CHS, c - c0);
    const float *src = feats + ((size_t)bi * c + c0) * n;
    for (int t = tid; t < cn * n; t += blockDim.x) qg_rows[t] = __ldg(src + t);
    __syncthreads();
    float *dst = out + ((size_t)bi * ctot + (use_xyz ? 3 : 0) + c0) * cols;
    for (int q = tid; q < cols4; q += blockDim.x) {
        const int4 id = __ldg(idx4 + q);
#pragma unroll
        for (int k = 0; k < CHS; ++k) {
            if (k < cn) {
                const float *r = qg_rows + k * n;
                st_cs_f4(dst + (size_t)k * cols + (size_t)q * 4, make_float4(r[id.x], r[id.y], r[id.z], r[id.w]));
            }
        }
    }
}

static int launch_group(int b, int c, int n, size_t cols, const float *points, const int *idx,
                        float *out, cudaStream_t st, const char *what) {
    if (b < 0 || c < 0 || n < 0) return fail(PDM_ERR_INVALID_ARG, "%s: negative size", what);
    if (b == 0 || c == 0 || cols == 0) return PDM_OK;
    if (!points || !idx || !out) return fail(PDM_ERR_INVALID_ARG, "%s: null pointer", what);
    if (b > 65535) return fail(PDM_ERR_UNSUPPORTED, "%s: batch %d > 65535", what, b);
    constexpr int CH = 8;
    const int chunks = (c + CH - 1) / CH;
    if (chunks > 65535) return fail(PDM_ERR_UNSUPPORTED, "%s: too many channels %d", what, c);
    const bool vec_ok = (cols % 4 == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    {
        // channel rows in shared memory when they fit (see query_group_smem_kernel)
        constexpr int CHS = 2;
        const size_t smem = (size_t)n * 4 * CHS;
        static const bool off = [] { const char *e = getenv("PDM_QG_SMEM"); return e && e[0] == 'o'; }();   // A/B knob
        if (!off && vec_ok && c >= 8 && cols >= 4096 && smem <= 64 * 1024 && cols < 0x7fffffffu) {
            auto kern = query_group_smem_kernel<CHS>;
            if (int rc = ensure_dynamic_smem((const void *)kern, smem)) return rc;
            dim3 grid((c + CHS - 1) / CHS, b);
            kern<<<grid, 512, smem, st>>>(c, n, (int)cols, 1, 0, nullptr, nullptr, points, idx, out);
            count_launch();
            PDM_CHECK_LAUNCH(what);
            return PDM_OK;
        }
    }
    if (vec_ok) {
        const int cols4 = (int)(cols / 4);
        dim3 grid((cols4 + 255) / 256, chunks, b);
        if (c >= CH) {
            prefer_max_smem((const void *)group_points_vec4_kernel<CH>);
            group_points_vec4_kernel<CH><<<grid, 256, 0, st>>>(c, n, cols4, points, idx, out);
        } else {
            grid.y = c;  // few channels: one per block row keeps registers low
            prefer_max_smem((const void *)group_points_vec4_kernel<1>);
            group_points_vec4_kernel<1><<<grid, 256, 0, st>>>(c, n, cols4, points, idx, out);
        }
    } else {
        dim3 grid((unsigned)((cols + 255) / 256), chunks, b);
        prefer_max_smem((const void *)group_points_scalar_kernel<CH>);
        group_points_scalar_kernel<CH><<<grid, 256, 0, st>>>(c, n, cols, points, idx, out);
    }
    count_launch();
    PDM_CHECK_LAUNCH(what);
    return PDM_OK;
}

static int launch_scatter_add(int b, int c, int n, size_t cols, const float *grad_out,
                              const int *idx, float *grad_points, cudaStream_t st,
                              const char *what) {
    if (b < 0 || c < 0 || n < 0) return fail(PDM_ERR_INVALID_ARG, "%s: negative size", what);
    if (b == 0 || c == 0 || cols == 0) return PDM_OK;
    if (!grad_out || !idx || !grad_points) return fail(PDM_ERR_INVALID_ARG, "%s: null pointer", what);
    if (b > 65535 || c > 65535) return fail(PDM_ERR_UNSUPPORTED, "%s: b/c > 65535", what);
    dim3 grid((unsigned)((cols + 255) / 256), c, b);
    scatter_add_kernel<<<grid, 256, 0, st>>>(c, n, cols, grad_out, idx, grad_points);
    count_launch();
    PDM_CHECK_LAUNCH(what);
    return PDM_OK;
}

}  // namespace pdm

extern "C" {

int pdm_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx,
                      float *out, void *stream) {
    if (npoints < 0) return pdm::fail(PDM_ERR_INVALID_ARG, "gather_points: negative size");
    return pdm::launch_group(b, c, n, (size_t)npoints, points, idx, out, (cudaStream_t)stream,
                             "gather_points");
}

int pdm_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                     const int *idx, float *out, void *stream) {
    if (npoints < 0 || nsample < 0) return pdm::fail(PDM_ERR_INVALID_ARG, "group_points: negative size");
    return pdm::launch_group(b, c, n, (size_t)npoints * nsample, points, idx, out,
                             (cudaStream_t)stream, "group_points");
}

int pdm_query_and_group(int b, int c, int n, int npoints, int nsample, int use_xyz, const float *xyz,
                        const float *new_xyz, const float *features, const int *idx, float *out,
                        void *stream) {
    using namespace pdm;
    if (b < 0 || c < 0 || n < 0 || npoints < 0 || nsample < 0)
        return fail(PDM_ERR_INVALID_ARG, "query_and_group: negative size");
    if (!use_xyz && c == 0) return fail(PDM_ERR_INVALID_ARG, "query_and_group: no xyz and no features");
    if (b == 0 || npoints == 0 || nsample == 0) return PDM_OK;
    if (!idx || !out || (use_xyz && (!xyz || !new_xyz)) || (c > 0 && !features))
        return fail(PDM_ERR_INVALID_ARG, "query_and_group: null pointer");
    constexpr int CH = 8;
    const int chunks = (c + CH - 1) / CH + (use_xyz ? 1 : 0);
    if (b > 65535 || chunks > 65535) return fail(PDM_ERR_UNSUPPORTED, "query_and_group: b/c too large");
    const size_t cols = (size_t)npoints * nsample;
    {
        // channel rows in shared memory when they fit and there is enough to gather (see the kernel)
        constexpr int CHS = 2;
        const size_t smem = (size_t)n * 4 * (use_xyz ? 3 : CHS);
        static const bool off = [] { const char *e = getenv("PDM_QG_SMEM"); return e && e[0] == 'o'; }();   // A/B knob
        if (!off && c >= 8 && smem <= 64 * 1024 && nsample % 4 == 0 && cols < 0x7fffffffu &&
            ((reinterpret_cast<uintptr_t>(idx) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
            auto kern = query_group_smem_kernel<CHS>;
            if (int rc = ensure_dynamic_smem((const void *)kern, smem)) return rc;
            dim3 grid((c + CHS - 1) / CHS + (use_xyz ? 1 : 0), b);
            kern<<<grid, 512, smem, (cudaStream_t)stream>>>(c, n, npoints, nsample, use_xyz ? 1 : 0, xyz, new_xyz,
                                                           features, idx, out);
            count_launch();
            PDM_CHECK_LAUNCH("query_and_group(smem)");
            return PDM_OK;
        }
    }
    dim3 grid((unsigned)((cols + 255) / 256), chunks, b);
    prefer_max_smem((const void *)query_group_kernel<CH>);
    query_group_kernel<CH><<<grid, 256, 0, (cudaStream_t)stream>>>(c, n, npoints, nsample, use_xyz ? 1 : 0, xyz,
                                                                  new_xyz, features, idx, out);
    count_launch();
    PDM_CHECK_LAUNCH("query_and_group");
    return PDM_OK;
}

int pdm_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int *idx,
                           float *grad_points, void *stream) {
    if (npoints < 0) return pdm::fail(PDM_ERR_INVALID_ARG, "gather_points_grad: negative size");
    return pdm::launch_scatter_add(b, c, n, (size_t)npoints, grad_out, idx, grad_points,
                                   (cudaStream_t)stream, "gather_points_grad");
}

int pdm_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                          const int *idx, float *grad_points, void *stream) {
    if (npoints < 0 || nsample < 0)
        return pdm::fail(PDM_ERR_INVALID_ARG, "group_points_grad: negative size");
    return pdm::launch_scatter_add(b, c, n, (size_t)npoints * nsample, grad_out, idx, grad_points,
                                   (cudaStream_t)stream, "group_points_grad");
}

}  // extern "C"
