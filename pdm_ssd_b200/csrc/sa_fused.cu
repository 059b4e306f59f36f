// sa_fused.cu -- one set-abstraction scale in a single kernel:
//   grouping (xyz - centre, features) -> shared MLP (1x1 conv + folded eval-mode BN + ReLU) x L
//   -> max over the nsample neighbours.
//
// Replaces, for inference, the chain QueryAndGroup.forward (pointnet2_utils.py:241-264: two
// group_points launches, a subtraction, a cat) + `self.mlps[i]` (Conv2d 1x1 / BatchNorm2d / ReLU
// per layer, pointnet2_modules.py:40,90-97) + F.max_pool2d (pointnet2_modules.py:41-52).  In the
// reference every one of those steps round-trips a (B, C, npoint, nsample) tensor through HBM
// (~3.3 GB per SA layer at batch 16); here a CTA keeps the activations of
// 128 rows (= 128/nsample centres x nsample neighbours) in shared memory from the gather to the
// max-pool, and only the (B, C_out, npoint) result is written.
//
// Arithmetic: fp32 FMA on the CUDA cores (exact to fp32 rounding; the reference runs the same
// GEMMs in fp32 through cuDNN/cuBLAS with a different summation order, so outputs agree to
// ~1e-6 relative, far inside the 1e-3 budget).  This file holds the GENERIC kernels of the op --
// sa_fused_kernel (CUDA cores, any widths <= 128, nsample 4..128) and round 1's tcgen05 kernel
// (tf32 hi/lo split, one tile per CTA) -- and the dispatch of pdm_sa_fused_forward[_v2]: narrow stacks
// go to the thread-per-row kernel of sa_rows.cu, GEMM-sized ones to the persistent warp-specialised
// tcgen05 kernel of sa_tc.cu; what those do not take falls through to the kernels below.
//
// Thread tile: 8 rows x 4 output channels, activations stored channel-major A[c][row] so that a
// thread's 8 rows are two 16-byte shared loads and its 4 weights one 16-byte (broadcast) load per
// k: 32 FMAs per 3 LDS.128.
#include <stdlib.h>

#include "common.cuh"

namespace pdm {

constexpr int kSARows = 128;      // rows (centre, neighbour) per CTA
constexpr int kSAThreads = 256;
constexpr int kSAMaxLayers = 4;
constexpr int kSAMaxC = 128;      // widest layer kept in shared memory

struct SAFusedParams {
    int n, m, c_feat, nsample, use_xyz, n_layers;
    int width[kSAMaxLayers + 1];   // width[0] = 3*use_xyz + c_feat
    int wpad[kSAMaxLayers + 1];    // width rounded up to a multiple of 4
    int woff[kSAMaxLayers];        // offset (floats) of layer l's transposed weights Wt[k][wpad] in `packed`
    int boff[kSAMaxLayers];        // offset of its bias (wpad floats)
    int act_floats;                // size of one activation buffer: max padded width * kSARows
    int w_floats;                  // size of the weight buffer: max_l width[l] * wpad[l+1]
};

__global__ void __launch_bounds__(kSAThreads)
sa_fused_kernel(SAFusedParams P, const float *__restrict__ xyz, const float *__restrict__ feats,
                const float *__restrict__ new_xyz, const int *__restrict__ idx,
                const float *__restrict__ packed, float *__restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    // sized by the host for THIS scale's widths: narrow MLPs (SA1: 4-16-16-32) take ~37 KB and
    // several CTAs share an SM, wide ones (128 channels) take the full 193 KB
    float *bufA = smem;                           // [max width][kSARows]
    float *bufB = bufA + P.act_floats;            // [max width][kSARows]
    float *wsm = bufB + P.act_floats;             // [k][wpad] of the current layer
    float *bsm = wsm + P.w_floats;                // [wpad]
    const int tid = threadIdx.x;
    const int bi = blockIdx.y;
    const int S = P.nsample;
    const int cpb = kSARows / S;                  // centres per CTA
    const int m0 = blockIdx.x * cpb;

    // ---- gather: A[c][r], r = i*S + s ------------------------------------------------------
    {
        const int r = tid & (kSARows - 1);
        const int half = tid >> 7;                // two threads per row split the channels
        const int i = r / S, s = r - i * S;
        const int mc = min(m0 + i, P.m - 1);      // rows of centres past the end are computed and dropped
        const int id = __ldg(idx + ((size_t)bi * P.m + mc) * S + s);
        int c0 = 0;
        if (P.use_xyz) {
            if (half == 0) {
                const float *pp = xyz + ((size_t)bi * P.n + id) * 3;
                const float *qq = new_xyz + ((size_t)bi * P.m + mc) * 3;
#pragma unroll
                for (int a = 0; a < 3; ++a) bufA[a * kSARows + r] = __fsub_rn(__ldg(pp + a), __ldg(qq + a));
            }
            c0 = 3;
        }
        const float *f = feats + (size_t)bi * P.c_feat * P.n + id;
        for (int c = half; c < P.c_feat; c += 2) bufA[(c0 + c) * kSARows + r] = __ldg(f + (size_t)c * P.n);
    }

    float *cur = bufA, *nxt = bufB;
    // thread tile: rows {rt*4..rt*4+3} and {64+rt*4..64+rt*4+3} (two conflict-free 16-byte loads
    // per k: consecutive lanes read consecutive 16-byte chunks) x 4 output channels
    const int rt = tid & 15;
    const int half = (tid >> 4) & 1;              // which of the warp's two column tiles
    const int ctw = (tid >> 5) * 2;               // first column tile of the warp
    for (int l = 0; l < P.n_layers; ++l) {
        const int cin = P.width[l], cout = P.width[l + 1], cpad = P.wpad[l + 1];
        __syncthreads();                          // `cur` complete; previous layer's weights dead
        for (int t = tid; t < cin * cpad; t += kSAThreads) wsm[t] = __ldg(packed + P.woff[l] + t);
        for (int t = tid; t < cpad; t += kSAThreads) bsm[t] = __ldg(packed + P.boff[l] + t);
        __syncthreads();
        const bool last = l + 1 == P.n_layers;
        for (int ctb = ctw; ctb * 4 < cpad; ctb += 16) {  // warp-uniform trip count
            const int ct = ctb + half;
            const bool active = ct * 4 < cpad;    // odd number of column tiles: the upper half-warp idles
            const int ctc = active ? ct : ctb;    // (it recomputes the lower tile; results dropped)
            float acc[8][4];
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[a][q] = 0.f;
            const float *ap = cur + rt * 4;
            const float *wp = wsm + ctc * 4;
#pragma unroll 4
            for (int k = 0; k < cin; ++k) {
                const float4 a0 = *reinterpret_cast<const float4 *>(ap + k * kSARows);
                const float4 a1 = *reinterpret_cast<const float4 *>(ap + k * kSARows + 64);
                const float4 w4 = *reinterpret_cast<const float4 *>(wp + k * cpad);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int a = 0; a < 8; ++a)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[a][q] = fmaf(av[a], wv[q], acc[a][q]);
            }
            const float4 b4 = *reinterpret_cast<const float4 *>(bsm + ctc * 4);
            const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
            if (!last) {
                if (active) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float o[8];
#pragma unroll
                        for (int a = 0; a < 8; ++a) o[a] = fmaxf(acc[a][q] + bv[q], 0.f);
                        float *dst = nxt + (ct * 4 + q) * kSARows + rt * 4;
                        *reinterpret_cast<float4 *>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<float4 *>(dst + 64) = make_float4(o[4], o[5], o[6], o[7]);
                    }
                }
            } else {
                // max over the nsample rows of a centre.  Rows 0..63 and 64..127 are two groups of
                // 16 four-row tiles; a centre spans S/4 adjacent tiles of one group (S <= 64) or
                // both groups (S = 128).
                const int tpc = min(S >> 2, 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float lo = fmaxf(acc[0][q] + bv[q], 0.f), hi = fmaxf(acc[4][q] + bv[q], 0.f);
#pragma unroll
                    for (int a = 1; a < 4; ++a) {
                        lo = fmaxf(lo, fmaxf(acc[a][q] + bv[q], 0.f));
                        hi = fmaxf(hi, fmaxf(acc[4 + a][q] + bv[q], 0.f));
                    }
                    for (int o = 1; o < tpc; o <<= 1) {
                        lo = fmaxf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
                    }
                    const int col = ct * 4 + q;
                    if (active && col < cout && (rt % tpc) == 0) {
                        if (S == kSARows) {
                            if (m0 < P.m) out[((size_t)bi * cout + col) * P.m + m0] = fmaxf(lo, hi);
                        } else {
                            const int c_lo = m0 + (rt * 4) / S, c_hi = m0 + (64 + rt * 4) / S;
                            if (c_lo < P.m) out[((size_t)bi * cout + col) * P.m + c_lo] = lo;
                            if (c_hi < P.m) out[((size_t)bi * cout + col) * P.m + c_hi] = hi;
                        }
                    }
                }
            }
        }
        float *tmp = cur; cur = nxt; nxt = tmp;
    }
}


// ---------------------------------------------------------------------------------------------
// Tensor-core variant (tcgen05 + TMEM).  Same data flow, but every layer
//     D[128 rows x N] = A[128 x K] * W[N x K]^T
// is issued by ONE thread as tcgen05.mma.kind::tf32 instructions (M = 128, K = 8 per instruction)
// with the accumulator in tensor memory.  To stay at fp32 accuracy (budget 1e-3; plain TF32 would
// sit right at it after three layers) each operand is split into a tf32-exact high part and a
// remainder, and the product is formed as  hi*hi + lo*hi + hi*lo  (3 MMAs per k-step; measured
// 1.1e-6 relative against fp64 in tools/micro/tc_gemm_test.cu).
// Operands live in shared memory in the K-major, no-swizzle core-matrix layout
//     smem[k/4][row][4 floats]      (8 rows x 16 bytes = one 128-byte core matrix;
//                                    LBO = rows*16 between the two 16-byte K chunks of an MMA,
//                                    SBO = 128 between 8-row groups)
// which is exactly what a thread-per-row epilogue writes with conflict-free 16-byte stores, so the
// activations never leave the SM between layers: TMEM -> registers (tcgen05.ld, one row per
// thread) -> bias + ReLU -> hi/lo split -> shared memory -> next layer's MMA.
// The max-pool uses redux.sync on the bit patterns (values are >= 0 after ReLU).
// Every mbarrier wait is bounded: a wrong descriptor can produce wrong numbers, never a hang.
constexpr int kTCThreads = 256;
constexpr int kTCSub = 64;         // output columns staged (and multiplied) at a time

// Host-prepared operand buffer `packed_tc` (see pointnet2_modules._pack_folded_tc):
//   bias[n_layers][128], then for every layer, for every block of <= 64 output columns:
//   W_hi[kpad/4][ns][4], W_lo[kpad/4][ns][4]   -- already split and already in the MMA layout,
// so staging a block is a straight 16-byte copy.
struct SATCParams {
    int n, m, c_feat, nsample, use_xyz, n_layers;
    int width[kSAMaxLayers + 1];
    int kpad[kSAMaxLayers];        // input width of layer l rounded up to 8
    int npad[kSAMaxLayers];        // output width rounded up to 16
    int woff[kSAMaxLayers];        // offset (floats) of layer l's first block in packed_tc
    int a_floats;                  // floats of ONE of A_hi / A_lo
    int w_floats;                  // floats of ONE of W_hi / W_lo (largest block)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void split_tf32(float v, float &hi, float &lo) {
    hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);   // what the tensor core reads of v
    lo = v - hi;                                              // exact in fp32
}

__global__ void __launch_bounds__(kTCThreads, 3)
sa_fused_tc_kernel(SATCParams P, const float *__restrict__ xyz, const float *__restrict__ feats,
                   const float *__restrict__ new_xyz, const int *__restrict__ idx,
                   const float *__restrict__ packed_tc, float *__restrict__ out, int *__restrict__ err) {
    extern __shared__ __align__(128) float smem[];
    float *a_hi = smem;
    float *a_lo = a_hi + P.a_floats;
    float *w_hi = a_lo + P.a_floats;
    float *w_lo = w_hi + P.w_floats;
    unsigned *cmax = reinterpret_cast<unsigned *>(w_lo + P.w_floats);   // [centres per CTA][128] max-pool accumulators
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bi = blockIdx.y;
    const int S = P.nsample;
    const int cpb = kSARows / S;
    const int m0 = blockIdx.x * cpb;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    for (int t = tid; t < cpb * kSAMaxC; t += kTCThreads) cmax[t] = 0u;

    // ---- gather the (xyz - centre, features) rows: 4 input channels -> one 16-byte store --------
    {
        const int r = tid & (kSARows - 1);
        const int i = r / S, sidx = r - i * S;
        const int mc = min(m0 + i, P.m - 1);
        const int id = __ldg(idx + ((size_t)bi * P.m + mc) * S + sidx);
        const float *pp = xyz + ((size_t)bi * P.n + id) * 3;
        const float *qq = new_xyz + ((size_t)bi * P.m + mc) * 3;
        const float *f = feats + (size_t)bi * P.c_feat * P.n + id;
        const int c0 = P.use_xyz ? 3 : 0;
        for (int k4 = (tid >> 7); k4 * 4 < P.kpad[0]; k4 += 2) {
            float h[4], lo4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k = k4 * 4 + e;
                float v = 0.f;                                   // K padding
                if (k < c0) v = __fsub_rn(__ldg(pp + k), __ldg(qq + k));
                else if (k < P.width[0]) v = __ldg(f + (size_t)(k - c0) * P.n);
                split_tf32(v, h[e], lo4[e]);
            }
            const int off = (k4 * kSARows + r) * 4;
            *reinterpret_cast<float4 *>(a_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4 *>(a_lo + off) = make_float4(lo4[0], lo4[1], lo4[2], lo4[3]);
        }
    }

    uint32_t phase = 0;
    for (int l = 0; l < P.n_layers; ++l) {
        const int cout = P.width[l + 1], K = P.kpad[l], N = P.npad[l];
        const float *bias = packed_tc + l * kSAMaxC;
        const float *wsrc = packed_tc + P.woff[l];
        for (int n0 = 0; n0 < N; n0 += kTCSub) {
            const int ns = min(kTCSub, N - n0);
            // stage this block of weights (pre-split, pre-laid-out): 2 * K * ns floats, 16 bytes at a time
            const int blk4 = K * ns / 4;
            const float4 *src_hi = reinterpret_cast<const float4 *>(wsrc);
            const float4 *src_lo = src_hi + blk4;
            for (int t = tid; t < blk4; t += kTCThreads) {
                reinterpret_cast<float4 *>(w_hi)[t] = __ldg(src_hi + t);
                reinterpret_cast<float4 *>(w_lo)[t] = __ldg(src_lo + t);
            }
            wsrc += 2 * K * ns;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncthreads();
            asm volatile("tcgen05.fence::after_thread_sync;");
            if (warp == 0) {
                if (lane == 0) {
                    const uint32_t tmem = tmem_base_s + n0;
                    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ns >> 3) << 17) | ((uint32_t)(kSARows >> 4) << 24);
                    const uint32_t a_lbo = kSARows * 16, w_lbo = ns * 16, sbo = 128;
                    for (int kk = 0; kk < K / 8; ++kk) {
                        const uint32_t aoff = kk * 2 * a_lbo, woff = kk * 2 * w_lbo;
                        const uint64_t ah = umma_desc_kmajor(smem_u32(a_hi) + aoff, a_lbo, sbo);
                        const uint64_t al = umma_desc_kmajor(smem_u32(a_lo) + aoff, a_lbo, sbo);
                        const uint64_t wh = umma_desc_kmajor(smem_u32(w_hi) + woff, w_lbo, sbo);
                        const uint64_t wl = umma_desc_kmajor(smem_u32(w_lo) + woff, w_lbo, sbo);
                        umma_tf32(tmem, ah, wh, idesc, kk > 0);
                        umma_tf32(tmem, al, wh, idesc, 1);
                        umma_tf32(tmem, ah, wl, idesc, 1);
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
                }
                __syncwarp();
                // bounded wait (warp 0 only; the others sleep in the barrier below): a wrong descriptor
                // may give wrong numbers but never a hang
                uint32_t done = 0;
                for (int it = 0; it < (1 << 20) && !done; ++it)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
                if (!done) {              // never continue past an unsatisfied barrier: the launch fails loudly
                    if (lane == 0 && err) atomicExch(err, 1);
                    __threadfence_system();
                    __trap();
                }
            }
            phase ^= 1u;
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncthreads();                                 // MMAs of this block complete: W may be overwritten, D is readable
            asm volatile("tcgen05.fence::after_thread_sync;");
        }
        // ---- epilogue: thread = one row (TMEM lane); warps w and w+4 split the 32-column chunks ----
        const bool last = l + 1 == P.n_layers;
        const int row = (warp & 3) * 32 + lane;
        const int nchunks = (N + 31) / 32;
        const int Knext = last ? 0 : P.kpad[l + 1];
        const uint32_t tmem = tmem_base_s;
        for (int ch = (warp >> 2); ch < nchunks; ch += 2) {
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + ch * 32;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                         "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (!last) {
                // bias + ReLU, split, store as next layer's A (columns >= cout are exact zeros: zero weights, zero bias)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int n0 = ch * 32 + q * 4;
                    if (n0 < Knext) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias + n0));
                        const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
                        float h[4], lo4[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            split_tf32(fmaxf(__uint_as_float(v[q * 4 + e]) + bv[e], 0.f), h[e], lo4[e]);
                        const int off = ((n0 >> 2) * kSARows + row) * 4;
                        *reinterpret_cast<float4 *>(a_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<float4 *>(a_lo + off) = make_float4(lo4[0], lo4[1], lo4[2], lo4[3]);
                    }
                }
            } else {
                // max over the nsample rows of a centre (values >= 0: bit patterns order like floats)
                const unsigned gmask = S >= 32 ? 0xffffffffu : (((1u << S) - 1u) << ((lane / S) * S));
                const int centre_local = row / S;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int col = ch * 32 + j;
                    const unsigned bits = __float_as_uint(fmaxf(__uint_as_float(v[j]) + __ldg(bias + col), 0.f));
                    const unsigned mx = __reduce_max_sync(gmask, bits);
                    if (col < cout && (lane % (S >= 32 ? 32 : S)) == 0) atomicMax(&cmax[centre_local * kSAMaxC + col], mx);
                }
            }
        }
        // (the next block's staging starts with fence + __syncthreads, which also orders these TMEM reads
        //  and A stores before the next MMAs)
        asm volatile("tcgen05.fence::before_thread_sync;");
    }
    __syncthreads();
    const int cout = P.width[P.n_layers];
    for (int t = tid; t < cpb * cout; t += kTCThreads) {
        const int ci = t / cout, col = t - ci * cout;
        if (m0 + ci < P.m) out[((size_t)bi * cout + col) * P.m + m0 + ci] = __uint_as_float(cmax[ci * kSAMaxC + col]);
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base_s), "r"(128));
}

}  // namespace pdm

namespace pdm {
int sa_tc3_try(int b, int n, int m, int c_feat, int nsample, int use_xyz, const float *xyz, const float *feats,
               const float *feats_pm, const float *new_xyz, const int *idx, int n_layers, const int *widths,
               const void *wpacked, const float *bias, float *out, float *out_pm, cudaStream_t st);   // sa_tc.cu
int sa_rows_try(int b, int n, int m, int c_feat, int nsample, int use_xyz, const float *xyz, const float *feats,
                const float *new_xyz, const int *idx, int n_layers, const int *widths, const float *packed, float *out,
                float *out_pm, cudaStream_t st);   // sa_rows.cu
}

// packed: for each layer l, Wt[k][wpad(l+1)] (k < width[l]; transposed, BN folded, zero padded
// columns) followed by bias[wpad(l+1)]; offsets are derived here from `widths`.
// packed_tc (optional): operands of the tensor-core kernel, see SATCParams.
extern "C" int pdm_sa_fused_forward_v2(int b, int n, int m, int c_feat, int nsample, int use_xyz,
                                       const float *xyz, const float *features, const float *features_pm,
                                       const float *new_xyz, const int *idx, int n_layers, const int *widths,
                                       const float *packed, const float *packed_tc, const void *packed_tc3,
                                       const float *bias_tc3, float *out, float *out_pm, void *stream);

extern "C" int pdm_sa_fused_forward(int b, int n, int m, int c_feat, int nsample, int use_xyz,
                                    const float *xyz, const float *features, const float *new_xyz,
                                    const int *idx, int n_layers, const int *widths,
                                    const float *packed, const float *packed_tc, float *out, void *stream) {
    return pdm_sa_fused_forward_v2(b, n, m, c_feat, nsample, use_xyz, xyz, features, nullptr, new_xyz, idx, n_layers, widths,
                                   packed, packed_tc, nullptr, nullptr, out, nullptr, stream);
}

namespace pdm {
// (B, C, M) -> (B, M, C) for the kernels that do not write the point-major copy themselves
__global__ void __launch_bounds__(256)
to_point_major_kernel(int bm_total, int m, int c, const float *__restrict__ in, float *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)bm_total * c) return;
    const int ch = (int)(i % c);
    const long long cm = i / c;
    const int bi = (int)(cm / m), mi = (int)(cm - (long long)bi * m);
    out[i] = __ldg(in + ((size_t)bi * c + ch) * m + mi);
}
static int finish_point_major(int b, int m, int c, const float *out, float *out_pm, cudaStream_t st) {
    if (!out_pm) return PDM_OK;
    const long long total = (long long)b * m * c;
    to_point_major_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(b * m, m, c, out, out_pm);
    count_launch();
    PDM_CHECK_LAUNCH("sa_fused_forward(point-major copy)");
    return PDM_OK;
}
}  // namespace pdm

extern "C" int pdm_sa_fused_forward_v2(int b, int n, int m, int c_feat, int nsample, int use_xyz,
                                       const float *xyz, const float *features, const float *features_pm,
                                       const float *new_xyz, const int *idx, int n_layers, const int *widths,
                                       const float *packed, const float *packed_tc, const void *packed_tc3,
                                       const float *bias_tc3, float *out, float *out_pm, void *stream) {
    using namespace pdm;
    if (b < 0 || n < 0 || m < 0 || c_feat < 0 || nsample <= 0)
        return fail(PDM_ERR_INVALID_ARG, "sa_fused_forward: bad size");
    if (n_layers < 1 || n_layers > kSAMaxLayers) return fail(PDM_ERR_UNSUPPORTED, "sa_fused_forward: %d layers", n_layers);
    if (nsample < 4 || kSARows % nsample != 0 || (nsample & (nsample - 1)) != 0)
        return fail(PDM_ERR_UNSUPPORTED, "sa_fused_forward: nsample %d (need a power of two in 4..128)", nsample);
    if (!widths || widths[0] != (use_xyz ? 3 : 0) + c_feat)
        return fail(PDM_ERR_INVALID_ARG, "sa_fused_forward: widths[0] must be 3*use_xyz + c_feat");
    SAFusedParams P;
    P.n = n; P.m = m; P.c_feat = c_feat; P.nsample = nsample; P.use_xyz = use_xyz ? 1 : 0; P.n_layers = n_layers;
    int off = 0;
    for (int l = 0; l <= n_layers; ++l) {
        if (widths[l] < 1 || widths[l] > kSAMaxC) return fail(PDM_ERR_UNSUPPORTED, "sa_fused_forward: width %d", widths[l]);
        P.width[l] = widths[l];
        P.wpad[l] = (widths[l] + 3) / 4 * 4;
    }
    int maxw = 0, maxwt = 0;
    for (int l = 0; l <= n_layers; ++l) maxw = P.wpad[l] > maxw ? P.wpad[l] : maxw;
    for (int l = 0; l < n_layers; ++l) {
        P.woff[l] = off; off += P.width[l] * P.wpad[l + 1];
        P.boff[l] = off; off += P.wpad[l + 1];
        maxwt = P.width[l] * P.wpad[l + 1] > maxwt ? P.width[l] * P.wpad[l + 1] : maxwt;
    }
    P.act_floats = maxw * kSARows;
    P.w_floats = (maxwt + 3) / 4 * 4;
    if (b == 0 || m == 0) return PDM_OK;
    if (!xyz || !new_xyz || !idx || !packed || !out || (c_feat > 0 && !features))
        return fail(PDM_ERR_INVALID_ARG, "sa_fused_forward: null pointer");
    if (b > 65535) return fail(PDM_ERR_UNSUPPORTED, "sa_fused_forward: batch > 65535");
    // narrow MLPs (first SA layer): thread-per-row kernel, everything in registers (sa_rows.cu); PDM_SA_ROWS=0 disables
    {
        const char *er = getenv("PDM_SA_ROWS");
        const bool rows_off = er && er[0] == '0';
        if (!rows_off) {
            const int rc = sa_rows_try(b, n, m, c_feat, nsample, use_xyz, xyz, features, new_xyz, idx, n_layers, widths, packed, out,
                                       out_pm, (cudaStream_t)stream);
            if (rc >= 0 || rc < -1) return rc;
        }
    }
    // genuine GEMMs (second SA layer): persistent warp-specialised tcgen05 kernel (sa_tc.cu); PDM_SA_TC3=0 disables
    {
        const char *e3 = getenv("PDM_SA_TC3");      // read per call: tools toggle it within a process
        const bool tc3_off = e3 && e3[0] == '0';
        const char *env = getenv("PDM_SA_TC");
        if (!tc3_off && packed_tc3 && bias_tc3 && !(env && env[0] == '0')) {
            const int rc = sa_tc3_try(b, n, m, c_feat, nsample, use_xyz, xyz, features, features_pm, new_xyz, idx, n_layers, widths,
                                      packed_tc3, bias_tc3, out, out_pm, (cudaStream_t)stream);
            if (rc >= 0 || rc < -1) return rc;
        }
    }
    // tensor-core path (tcgen05) whenever its operands were supplied and the scale fits;
    // PDM_SA_TC=0 forces the CUDA-core kernel
    {
        const char *env = getenv("PDM_SA_TC");
        const bool want_tc = packed_tc != nullptr && !(env && env[0] == '0');
        SATCParams T;
        T.n = n; T.m = m; T.c_feat = c_feat; T.nsample = nsample; T.use_xyz = use_xyz ? 1 : 0; T.n_layers = n_layers;
        int amax = 0, wmax = 0, woff = n_layers * kSAMaxC;
        for (int l = 0; l <= n_layers; ++l) T.width[l] = P.width[l];
        for (int l = 0; l < n_layers; ++l) {
            T.kpad[l] = (P.width[l] + 7) / 8 * 8;
            T.npad[l] = (P.width[l + 1] + 15) / 16 * 16;
            T.woff[l] = woff;
            woff += 2 * T.kpad[l] * T.npad[l];
            amax = T.kpad[l] > amax ? T.kpad[l] : amax;
            const int ns = T.npad[l] < kTCSub ? T.npad[l] : kTCSub;
            wmax = T.kpad[l] * ns > wmax ? T.kpad[l] * ns : wmax;
        }
        // Tensor cores only where the MLP is a genuine GEMM: below ~4k multiply-adds per row (SA1's
        // 4-16-16-32 stack has 832) the tile is bound by the gather and the per-layer barriers and the
        // CUDA-core kernel is faster (0.46 vs 0.70 ms at batch 16); SA2's 67-64-64-128 stack (16.6k)
        // runs 1.6x faster on tcgen05 (0.56 vs 0.89 ms).  PDM_SA_TC=1 forces the tensor-core kernel.
        long long macs_per_row = 0;
        for (int l = 0; l < n_layers; ++l) macs_per_row += (long long)P.width[l] * P.width[l + 1];
        const bool forced = env && env[0] == '1';
        // layer l+1 reads what layer l's epilogue wrote: its K padding must be covered by layer l's N padding
        bool ok = want_tc && nsample >= 8 && (forced || macs_per_row >= 4096);
        for (int l = 0; l + 1 < n_layers; ++l) ok = ok && T.kpad[l + 1] <= T.npad[l];
        T.a_floats = amax * kSARows;
        T.w_floats = wmax;
        const int cpb = kSARows / nsample;
        const size_t smem_tc = sizeof(float) * ((size_t)2 * T.a_floats + 2 * T.w_floats + (size_t)cpb * kSAMaxC);
        if (ok && smem_tc <= 220 * 1024) {
            if (int rc = ensure_dynamic_smem((const void *)sa_fused_tc_kernel, smem_tc)) return rc;
            dim3 grid((m + cpb - 1) / cpb, b);
            sa_fused_tc_kernel<<<grid, kTCThreads, smem_tc, (cudaStream_t)stream>>>(T, xyz, features, new_xyz, idx, packed_tc, out, nullptr);
            count_launch();
            PDM_CHECK_LAUNCH("sa_fused_forward(tcgen05)");
            return finish_point_major(b, m, widths[n_layers], out, out_pm, (cudaStream_t)stream);
        }
    }
    const size_t smem = sizeof(float) * ((size_t)2 * P.act_floats + P.w_floats + kSAMaxC);
    if (int rc = ensure_dynamic_smem((const void *)sa_fused_kernel, smem)) return rc;
    const int cpb = kSARows / nsample;
    dim3 grid((m + cpb - 1) / cpb, b);
    sa_fused_kernel<<<grid, kSAThreads, smem, (cudaStream_t)stream>>>(P, xyz, features, new_xyz, idx, packed, out);
    count_launch();
    PDM_CHECK_LAUNCH("sa_fused_forward");
    return finish_point_major(b, m, widths[n_layers], out, out_pm, (cudaStream_t)stream);
}
