// common.cuh -- shared helpers for libpdmops (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/pdm_ops.h"

namespace pdm {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

int fail(int code, const char *fmt, ...);
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Call right after a launch: returns non-zero (and records the message) on a launch error.
#define PDM_CHECK_LAUNCH(what)                                                         \
    do {                                                                               \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess)                                                        \
            return pdm::fail((int)e__, "%s: %s", what, cudaGetErrorString(e__));       \
    } while (0)

#define PDM_CHECK_CUDA(expr)                                                           \
    do {                                                                               \
        cudaError_t e__ = (expr);                                                      \
        if (e__ != cudaSuccess)                                                        \
            return pdm::fail((int)e__, "%s: %s", #expr, cudaGetErrorString(e__));      \
    } while (0)

// Squared distance with the reference's exact rounding sequence.  nvcc 12.9 contracts
//   (ax-bx)*(ax-bx) + (ay-by)*(ay-by) + (az-bz)*(az-bz)
// in sampling_gpu.cu:139, ball_query_gpu.cu:38 and interpolate_gpu.cu:41 into
//   FMUL t = dy*dy ; FFMA t = dx*dx + t ; FFMA d = dz*dz + t
// (read from the reference SASS).  Spelled with intrinsics so no optimisation
// level or future compiler can re-associate it.
__device__ __forceinline__ float sqdist_ref(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Two fp32 multiply-adds in one instruction (fma.rn.f32x2 -> FFMA2 on sm_100a): c0 = a * b0 + c0, c1 = a * b1 + c1, each half
// an IEEE fma (bit-identical to fmaf).  The CUDA-core MLP kernels (sa_rows.cu, point_head.cu) are bound by instruction ISSUE, not by the fp32 pipe: an FFMA2 takes one
// issue slot for two pipe cycles, which frees the other slot for the LDS / index arithmetic around it.
__device__ __forceinline__ void ffma2(float &c0, float &c1, float a, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %2};\n\tmov.b64 rb, {%3, %4};\n\tmov.b64 rc, {%0, %1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "+f"(c0), "+f"(c1) : "f"(a), "f"(b0), "f"(b1));
}


// fp32 carried as two bf16 values for the tensor-core layers (conv_tc.cu): hi = bf16(v), lo = bf16(v - hi);
// hi + lo equals v to 2^-17 relative.
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(__fsub_rn(v, __bfloat162float(hi)));
}
__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
// eight fp32 values -> the two 16-byte rows (hi, lo) of one pixel's 8-channel chunk in the split NHWC8 layout
__device__ __forceinline__ void split8_bf16(const float *f, uint4 &hi, uint4 &lo) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        __nv_bfloat16 h0, l0, h1, l1;
        split_bf16(f[2 * q], h0, l0);
        split_bf16(f[2 * q + 1], h1, l1);
        hw[q] = pack_bf16x2(h0, h1);
        lw[q] = pack_bf16x2(l0, l1);
    }
    hi = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    lo = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}
// the inverse for one 16-byte row pair: value = hi + lo
__device__ __forceinline__ void join8_bf16(const uint4 &hi, const uint4 &lo, float *f) {
    const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        f[2 * q] = __fadd_rn(__uint_as_float(hw[q] << 16), __uint_as_float(lw[q] << 16));
        f[2 * q + 1] = __fadd_rn(__uint_as_float(hw[q] & 0xffff0000u), __uint_as_float(lw[q] & 0xffff0000u));
    }
}

// Streaming (evict-first) 128-bit store for write-once outputs.
__device__ __forceinline__ void st_cs_f4(float *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_cs_f1(float *p, float v) {
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// cudaFuncAttributeMaxDynamicSharedMemorySize, set once per (kernel, device) and only raised:
// keeps driver calls out of the steady state (and out of CUDA-graph capture).
int ensure_dynamic_smem(const void *func, size_t bytes);

// cudaFuncAttributePreferredSharedMemoryCarveout = max shared, set once per (kernel, device).
// The on-chip FPS kernel needs ~197 KB of shared memory per CTA; an SM can only take such a CTA in
// the max-shared L1/shared split, and it has to drain before the split changes.  When the small
// kernels of the chain (ball query, grouping, gather) run with their default (L1-heavy) preference
// between FPS launches of other streams, the SMs keep flipping between the two splits and FPS CTAs
// wait for drained SMs.  Asking for the same split everywhere removes the flips.
// PDM_CARVEOUT=off disables it (A/B measurements).
void prefer_max_smem(const void *func);

// Library scratch (ball-query grid, NMS masks, neck work lists, FPS throughput mode): one buffer per (device,
// stream) for eager calls, one PRIVATE buffer per stream capture for calls recorded into a CUDA graph (see
// capi.cu for the ownership rules).  Stream-ordered cudaMallocAsync inside a captured step turned every graph
// launch into ~1 ms of driver work once four processes drove four GPUs; a plain pointer costs nothing.
// Returns nullptr and records an error if a captured call needs more than the warm-up run prepared.
void *stream_scratch(cudaStream_t st, size_t bytes);

// reference host helper cuda_utils.h:10-14 (block size the reference FPS would use);
// decides the tie-break order our FPS has to reproduce.
int ref_fps_block_size(int n);

}  // namespace pdm
