// Standalone check of the tcgen05 building block used by the tensor-core SA kernel:
//   D[128 x N] = A[128 x K] * W[N x K]^T   with fp32 operands split into tf32 hi/lo parts
//   (3 MMAs per k-step: hi*hi + lo*hi + hi*lo), accumulators in TMEM, operands in shared memory in
//   the K-major no-swizzle core-matrix layout  smem[kchunk][row][4 floats].
// Compares with a double-precision CPU product.  All waits are bounded (no hang on a bad descriptor).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 bytes, 128 contiguous bytes.
// start address >> 4 in bits [0,14); LBO (byte offset between the two 16-byte K chunks of one MMA)
// >> 4 in bits [16,30); SBO (byte offset between 8-row groups) >> 4 in bits [32,46); version 1 at bit 46.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

constexpr int M = 128;

__global__ void __launch_bounds__(128, 1)
tc_gemm(int K, int N, const float *__restrict__ A, const float *__restrict__ W, float *__restrict__ D, int *err) {
    extern __shared__ __align__(128) unsigned char smem[];
    // layout: A_hi[K/4][128][4], A_lo, W_hi[K/4][N][4], W_lo
    float *a_hi = reinterpret_cast<float *>(smem);
    float *a_lo = a_hi + K * M;
    float *w_hi = a_lo + K * M;
    float *w_lo = w_hi + K * N;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    // operands -> shared, split into tf32-exact hi and remainder lo
    for (int i = tid; i < K * M; i += 128) {
        const int r = i / K, k = i % K;
        const float v = A[i];
        const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        const int off = ((k >> 2) * M + r) * 4 + (k & 3);
        a_hi[off] = h;
        a_lo[off] = v - h;
    }
    for (int i = tid; i < K * N; i += 128) {
        const int n = i / K, k = i % K;
        const float v = W[i];
        const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        const int off = ((k >> 2) * N + n) * 4 + (k & 3);
        w_hi[off] = h;
        w_lo[off] = v - h;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (tensor core)
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;

    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32(M, N);
        const uint32_t a_lbo = M * 16, w_lbo = N * 16, sbo = 128;
        for (int kk = 0; kk < K / 8; ++kk) {
            const uint32_t aoff = kk * 2 * M * 16, woff = kk * 2 * N * 16;
            const uint64_t ah = make_desc(smem_u32(a_hi) + aoff, a_lbo, sbo), al = make_desc(smem_u32(a_lo) + aoff, a_lbo, sbo);
            const uint64_t wh = make_desc(smem_u32(w_hi) + woff, w_lbo, sbo), wl = make_desc(smem_u32(w_lo) + woff, w_lbo, sbo);
            mma_tf32(tmem, ah, wh, idesc, kk > 0);
            mma_tf32(tmem, al, wh, idesc, 1);
            mma_tf32(tmem, ah, wl, idesc, 1);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    // bounded wait on the commit
    {
        uint32_t done = 0;
        for (int it = 0; it < (1 << 22) && !done; ++it) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        }
        if (!done && tid == 0) *err = 1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    // epilogue: warp w reads TMEM lanes 32w..32w+31 (its rows), 32 columns at a time
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                       "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                       "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128));
}

int main() {
    const int cases[][2] = {{64, 64}, {72, 64}, {64, 128}, {8, 16}, {128, 128}};
    for (auto &c : cases) {
        const int K = c[0], N = c[1];
        std::vector<float> A(M * K), W(N * K), D(M * N);
        srand(K * 1000 + N);
        for (auto &x : A) x = (rand() / (float)RAND_MAX - 0.5f) * 4.f;
        for (auto &x : W) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
        float *dA, *dW, *dD; int *dErr;
        cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dErr, 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(dErr, 0, 4); cudaMemset(dD, 0, D.size() * 4);
        const size_t smem = (size_t)(2 * K * M + 2 * K * N) * 4;
        cudaFuncSetAttribute(tc_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tc_gemm<<<1, 128, smem>>>(K, N, dA, dW, dD, dErr);
        cudaError_t e = cudaDeviceSynchronize();
        int herr = 0;
        cudaMemcpy(&herr, dErr, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int r = 0; r < M; ++r)
            for (int n = 0; n < N; ++n) {
                double s = 0;
                for (int k = 0; k < K; ++k) s += (double)A[r * K + k] * W[n * K + k];
                maxerr = fmax(maxerr, fabs(s - D[r * N + n]));
                maxref = fmax(maxref, fabs(s));
            }
        printf("K=%3d N=%3d  cuda=%s  timeout=%d  max|err|=%.3e  max|ref|=%.3f  rel=%.3e\n", K, N, cudaGetErrorString(e), herr, maxerr, maxref, maxerr / maxref);
        cudaFree(dA); cudaFree(dW); cudaFree(dD); cudaFree(dErr);
    }
    return 0;
}
