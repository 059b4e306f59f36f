// Raw issue / execution rate of tcgen05.mma.cta_group::1.kind::f16 (M = 128, K = 16, bf16) from shared-memory operands in
// the K-major no-swizzle layout, as a function of N, of the accumulator pattern and of the A-descriptor geometry used by
// conv_tc.cu (LBO = 288, SBO = 1152, start shifted by 16 bytes) -- to tell issue-rate limits from operand-bandwidth limits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

// mode bits: 1 = alternate between 4 accumulators, 2 = conv-style A geometry (LBO 288 / SBO 1152), 4 = A start shifted by 16 B,
//            8 = rotate through 9 different A start addresses (taps) and 3 B addresses
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, int mode, long long *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t dummy_full, dummy_empty;
    __shared__ uint32_t tmem_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&dummy_full)), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&dummy_empty)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u + i;   // finite bf16 pairs
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_s;
    if (warp == 1) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 96 * 1024;
        const uint32_t albo = (mode & 2) ? 288u : 2048u, asbo = (mode & 2) ? 1152u : 128u;
        const uint32_t blbo = (uint32_t)N * 16u, bsbo = 128u;
        const long long t0 = clock64();
        if (mode & 32) {
            // loop-invariant descriptors: the loop body is 6 x UTCHMMA and the loop counter -- the pure issue cost
            const uint64_t ad = desc(a0, albo, asbo), ad2 = desc(a0 + 20736u, albo, asbo);
            const uint64_t bd = desc(b0, blbo, bsbo), bd2 = desc(b0 + 8192u, blbo, bsbo);
            const uint32_t d = tmem;
            if (elect_one()) {
#pragma unroll 1
                for (int it = 0; it < iters; ++it) {
                    mma(d, ad, bd, idesc, 1);
                    mma(d, ad2, bd, idesc, 1);
                    mma(d, ad, bd2, idesc, 1);
                    mma(d, ad + 36, bd + 16, idesc, 1);
                    mma(d, ad2 + 36, bd + 16, idesc, 1);
                    mma(d, ad + 36, bd2 + 16, idesc, 1);
                }
            }
            __syncwarp();
        } else
        for (int it = 0; it < iters; ++it) {
            const uint32_t shift = (mode & 4) ? 16u : 0u;
            const uint32_t tap = (mode & 8) ? (uint32_t)(it % 9) : 0u;
            const uint32_t aaddr = a0 + shift + (tap / 3) * asbo + (tap % 3) * 16u;
            const uint64_t ad = desc(aaddr, albo, asbo), ad2 = desc(aaddr + 20736u, albo, asbo);
            const uint64_t bd = desc(b0 + ((mode & 8) ? (uint32_t)(it % 3) * 16384u : 0u), blbo, bsbo), bd2 = desc(b0 + 8192u, blbo, bsbo);
            const uint32_t d = tmem + ((mode & 1) ? (uint32_t)((it & 3) * 128) : 0u);
            if (elect_one()) {
                mma(d, ad, bd, idesc, 1);
                mma(d, ad2, bd, idesc, 1);
                mma(d, ad, bd2, idesc, 1);
                mma(d, ad + 36, bd + (blbo >> 3), idesc, 1);
                mma(d, ad2 + 36, bd + (blbo >> 3), idesc, 1);
                mma(d, ad + 36, bd2 + (blbo >> 3), idesc, 1);
            }
            __syncwarp();
            // mode bit 16: the per-block bookkeeping of conv_tc_kernel every 12 MMAs: a wait on an already-complete barrier
            // (parity 1 of a fresh barrier), a fence, and a commit to a barrier nobody waits on
            if ((mode & 16) && (it & 1)) {
                uint32_t ok = 0;
                while (!ok)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                                 : "=r"(ok) : "r"(smem_u32(&dummy_full)), "r"(1) : "memory");
                asm volatile("tcgen05.fence::after_thread_sync;");
                if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&dummy_empty)) : "memory");
                __syncwarp();
            }
        }
        if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
        __syncwarp();
        uint32_t done = 0;
        for (long long spin = 0; spin < (1ll << 26) && !done; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        const long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) out[blockIdx.x] = done ? (t1 - t0) : -1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

int main() {
    long long *d_out, h_out[148];
    cudaMalloc(&d_out, sizeof(h_out));
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2000;
    for (int N : {16, 64, 128, 256}) {
        for (int mode : {0, 32}) {
            if (N == 256 && (mode & 1)) continue;
            for (int grid : {148}) {
                rate_kernel<<<grid, 128, smem>>>(N, iters, mode, d_out);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(h_out, d_out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
                long long mx = 0;
                for (int i = 0; i < grid; ++i) mx = h_out[i] > mx ? h_out[i] : mx;
                printf("N=%3d mode=%2d grid=%3d  %s  %.1f cycles per MMA\n", N, mode, grid, cudaGetErrorString(e), (double)mx / (6.0 * iters));
            }
        }
    }
    return 0;
}
