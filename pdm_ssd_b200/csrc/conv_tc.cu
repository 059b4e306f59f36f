// conv_tc.cu -- 3x3 (or 1x1) convolution + folded eval-mode BatchNorm + activation on the BEV map as an
// implicit GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in tensor memory),
// operands staged by TMA.  Replaces, for inference, the Conv2d/BatchNorm2d/ReLU stacks of the
// "context learning" block (pcdet/models/backbones_2d/base_bev_backbone.py:27-47) and of the heatmap
// branch of the hybrid head (pcdet/models/dense_heads/center_head.py:12-46), which the reference runs
// as cuDNN fp32 kernels.
//
// Accuracy: the budget is 1e-3 relative in fp32 (BASELINE.json north_star).  Every fp32 operand is
// carried as TWO bf16 values, hi = bf16(v) and lo = bf16(v - hi), and a product is formed as
// hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM (3 tcgen05.mma.kind::f16 per k-step): the
// dropped lo*lo term is 2^-18 relative, measured ~1e-5 against torch fp32 after a 5-conv stack.
//
// Activation layout between the layers ("split NHWC8", written by the producing kernel's epilogue and
// by the neck's dense writer, read by TMA):   bf16 [plane: hi|lo][B][Y][C/8][X][8]
// i.e. per image row the channels come in chunks of 8 and each chunk holds the whole row.  A TMA box
// {(H)x8 elements, 4 chunks, H rows} of that tensor lands in shared memory as  [row][chunk][x][8]  which
// IS the K-major, no-swizzle core-matrix layout of tcgen05 (core matrix = 8 consecutive pixels x 16
// bytes, contiguous; LBO = one chunk row = H*16 bytes; SBO = one image row = 4*H*16 bytes).  So ONE halo
// tile (18 x 18 pixels for a 16 x 16 output unit) serves all nine taps: tap (ky,kx) is the same buffer
// with the descriptor start address moved by ky rows and kx pixels -- no im2col copy, the input is read
// from L2/HBM 1.27x instead of 9x, and out-of-image halo cells are zero-filled by TMA (= conv padding).
//
// CTA = persistent, warp-specialised:  warp 0 lane 0 streams halo tiles (TMA tensor loads, 2-stage ring
// of 32-channel chunks), warp 1 lane 0 streams the pre-packed weight blocks (1-D bulk copies, 3-stage
// ring, one (32-channel chunk, tap) block per stage), warp 2 lane 0 issues the MMAs, warp 3 owns the
// TMEM allocation, warps 4-7 drain the accumulators (tcgen05.ld -> bias + activation -> hi/lo split ->
// global).  A CTA works on TWO 16x16 units at a time (4 accumulators of 128 rows = 4 x npad TMEM
// columns) so that every weight block fetched from L2 feeds 512 output pixels.
// Every mbarrier wait is bounded: on a time-out the kernel records a code in the error word and traps
// (the launch fails loudly; it never hangs and never continues past an unsatisfied barrier).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include <mutex>

#include "tc_common.cuh"

namespace pdm {

constexpr int kCvThreads = 256;
constexpr int kCvMaxAStages = 4;
constexpr int kCvMaxWStages = 36;     // resident mode keeps every (chunk, tap) block: up to 4 chunks x 9 taps
constexpr int kCvMaxN = 128;
constexpr int kCvSmemBudget = 227 * 1024 - 1536;   // dynamic part: 227 KB per CTA minus static data and alignment slack

struct ConvTCParams {
    int B, Y, X, cin, cout, npad, ksize;
    int halo;            // 16 + ksize - 1 pixels per side of a unit's input tile
    int units_x, units_y, total_units, units_per_cta;
    int nu;              // units per tile: 2 (4 accumulators) when two sets of them fit in TMEM, else 1
    int a_unit_bytes;    // one (unit, plane) box: halo rows x 4 chunks x halo pixels x 16 bytes
    int w_block_bytes;   // one (chunk, tap) weight block: 2 planes x 4 chunks x npad rows x 16 bytes
    int w_stage_bytes;   // one ring stage = the ksize taps of one (chunk, ky) filter row: ksize blocks, contiguous in w_packed
    int a_stages, w_stages;
    int w_resident;      // every weight block has its own slot: loaded once per CTA, never released
    int act;             // 0 none, 1 relu, 2 sigmoid
    int tmem_cols;
    int ncat;            // N-concatenated weights (npad <= 64): blocks packed [chunk][hi|lo][n][8], so ONE MMA of N = 2 npad forms
                         // A_hi x [W_hi; W_lo] in two column ranges of the accumulator (summed in the epilogue) and a second one of
                         // N = npad adds A_lo x W_hi: 64 + 48 cycles per k-step instead of 3 x 48 (an N = 64 MMA is bound by the
                         // 128 B/clk of shared-memory operand bandwidth, not by the tensor pipe)
    int accw;            // TMEM columns per accumulator: npad, or 2 npad with ncat
    int debug;           // PDM_CONV_DEBUG (measurement only, results wrong): 1 halo loads only while the ring fills,
                         // 2 same for weight blocks, 4 no global stores in the epilogue
};

struct __align__(8) ConvBarriers {
    uint64_t a_full[kCvMaxAStages], a_empty[kCvMaxAStages];
    uint64_t w_full[kCvMaxWStages], w_empty[kCvMaxWStages];
    uint64_t acc_full[2], acc_empty[2];
};

// 16 accumulator values of one pixel -> bias + activation, then the two 16-byte hi/lo rows per 8-channel chunk
__device__ __forceinline__ float cv_act(float a, int act) {
    if (act == 1) return fmaxf(a, 0.f);
    if (act == 2) return 1.f / (1.f + __expf(-a));
    return a;
}
__device__ __forceinline__ void cv_split2(float a, float b, uint32_t &hi, uint32_t &lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);                 // one cvt for the pair
    hi = *reinterpret_cast<const uint32_t *>(&h);
    const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}

__global__ void __launch_bounds__(kCvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_in, const ConvTCParams P,
               const unsigned char *__restrict__ w_packed, const float *__restrict__ bias,
               __nv_bfloat16 *__restrict__ out_split, float *__restrict__ out_nchw, int *__restrict__ err) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ ConvBarriers bars;
    __shared__ uint32_t tmem_base_s;
    __shared__ float bias_s[kCvMaxN];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem0 = (cv_smem_u32(smem_raw) + 127u) & ~127u;
    const int NU = P.nu;
    const uint32_t a_stage_bytes = (uint32_t)(2 * NU) * (uint32_t)P.a_unit_bytes;
    const uint32_t a_base = smem0;
    const uint32_t w_base = smem0 + (uint32_t)P.a_stages * a_stage_bytes;

    const int u_first = blockIdx.x * P.units_per_cta;
    const int u_last = min(u_first + P.units_per_cta, P.total_units);
    const int n_tiles = (u_last - u_first + NU - 1) / NU;
    const int n_kc = P.cin / 32, pad = (P.ksize - 1) / 2;

    if (tid == 0) {
        for (int s = 0; s < P.a_stages; ++s) { mbar_init(cv_smem_u32(&bars.a_full[s]), 1); mbar_init(cv_smem_u32(&bars.a_empty[s]), 1); }
        for (int s = 0; s < P.w_stages; ++s) { mbar_init(cv_smem_u32(&bars.w_full[s]), 1); mbar_init(cv_smem_u32(&bars.w_empty[s]), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(cv_smem_u32(&bars.acc_full[s]), 1); mbar_init(cv_smem_u32(&bars.acc_empty[s]), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(cv_smem_u32(&tmem_base_s)), "r"(P.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid < kCvMaxN) bias_s[tid] = tid < P.cout ? __ldg(bias + tid) : 0.f;
    if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_in) : "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===== halo-tile producer (TMA tensor loads) =====
        {
            int stage = 0; uint32_t phase = 0;
            for (int t = 0; t < n_tiles; ++t) {
                const int u0 = u_first + NU * t;
                const int nun = min(NU, u_last - u0);
                for (int kc = 0; kc < n_kc; ++kc) {
                    mbar_wait(cv_smem_u32(&bars.a_empty[stage]), phase ^ 1u, err, 101);
                    const uint32_t full = cv_smem_u32(&bars.a_full[stage]);
                    if ((P.debug & 1) && (t * n_kc + kc) >= P.a_stages) {
                        if (elect_one()) mbar_arrive(full);
                        __syncwarp();
                        if (++stage == P.a_stages) { stage = 0; phase ^= 1u; }
                        continue;
                    }
                    if (elect_one()) mbar_arrive_expect_tx(full, (uint32_t)(nun * 2 * P.a_unit_bytes));
                    __syncwarp();
                    for (int un = 0; un < nun; ++un) {
                        const int u = u0 + un;
                        const int ux = u % P.units_x, r = u / P.units_x;
                        const int uy = r % P.units_y, b = r / P.units_y;
                        const uint32_t dst = a_base + stage * a_stage_bytes + (uint32_t)(un * 2 * P.a_unit_bytes);
                        if (elect_one()) {
                            tma_load_5d(dst, &tmap_in, full, (ux * 16 - pad) * 8, kc * 4, uy * 16 - pad, b, 0);
                            tma_load_5d(dst + (uint32_t)P.a_unit_bytes, &tmap_in, full, (ux * 16 - pad) * 8, kc * 4, uy * 16 - pad, b, 1);
                        }
                        __syncwarp();
                    }
                    if (++stage == P.a_stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== weight-block producer (1-D bulk copies; the blocks are pre-packed in consumption order) =====
        {
            int stage = 0; uint32_t phase = 0;
            const int rounds = P.w_resident ? min(n_tiles, 1) : n_tiles;
            for (int t = 0; t < rounds; ++t) {
                const unsigned char *src = w_packed;
                for (int rb = 0; rb < n_kc * P.ksize; ++rb) {
                    mbar_wait(cv_smem_u32(&bars.w_empty[stage]), phase ^ 1u, err, 102);
                    const uint32_t full = cv_smem_u32(&bars.w_full[stage]);
                    if ((P.debug & 2) && (t * n_kc * P.ksize + rb) >= P.w_stages) {
                        if (elect_one()) mbar_arrive(full);
                        __syncwarp();
                        src += P.w_stage_bytes;
                        if (++stage == P.w_stages) { stage = 0; phase ^= 1u; }
                        continue;
                    }
                    if (elect_one()) {
                        mbar_arrive_expect_tx(full, (uint32_t)P.w_stage_bytes);
                        bulk_load_1d(w_base + stage * (uint32_t)P.w_stage_bytes, src, (uint32_t)P.w_stage_bytes, full);
                    }
                    __syncwarp();
                    src += P.w_stage_bytes;
                    if (++stage == P.w_stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 2) {
        // ===== MMA issuer =====
        // The issue rate of a single thread is what bounds MMAs with N <= 128 (tools/micro/umma_rate.cu: 64 cycles per
        // 128x128x16 MMA with loop-invariant descriptors, 80+ as soon as a handful of address instructions sit between
        // two MMAs).  So: one barrier wait / commit per filter ROW (ksize taps = 36 or 72 MMAs), all of them issued from
        // one elect region, every descriptor = a per-row base low word + a loop-invariant constant (32-bit adds; the high
        // words never change).
        {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P.npad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t idesc_cat = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * P.npad) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t H16 = (uint32_t)P.halo * 16u;          // one chunk row of the halo tile
            const uint32_t a_lbo = H16, a_sbo = 4u * H16;
            const uint32_t w_lbo = (uint32_t)(P.ncat ? 2 * P.npad : P.npad) * 16u;     // one 8-channel chunk of a weight block
            const uint32_t a_hi_word = cv_desc_hi(a_sbo), w_hi_word = cv_desc_hi(128u);
            const uint32_t a_lbo_f = ((a_lbo >> 4) & 0x3fffu) << 16, w_lbo_f = ((w_lbo >> 4) & 0x3fffu) << 16;
            // descriptor low-word increments (units of 16 bytes)
            const uint32_t A_PLANE = (uint32_t)P.a_unit_bytes >> 4, A_UNIT = 2u * A_PLANE, A_KS = (2u * a_lbo) >> 4;
            const uint32_t W_PLANE = (4u * w_lbo) >> 4, W_KS = (2u * w_lbo) >> 4, W_TAP = (uint32_t)P.w_block_bytes >> 4;   // (W_PLANE: !ncat)
            const uint32_t ACCW = (uint32_t)P.accw;
            int as = 0, ws = 0; uint32_t aph = 0, wph = 0;
            for (int t = 0; t < n_tiles; ++t) {
                const int nun = min(NU, u_last - (u_first + NU * t));
                const int set = t & 1;
                const uint32_t use = (uint32_t)(t >> 1);
                mbar_wait(cv_smem_u32(&bars.acc_empty[set]), (use & 1u) ^ 1u, err, 103);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t acc0 = tmem_base + (uint32_t)(set * 2 * NU) * ACCW;
                for (int kc = 0; kc < n_kc; ++kc) {
                    mbar_wait(cv_smem_u32(&bars.a_full[as]), aph, err, 104);
                    const uint32_t a_st = a_base + as * a_stage_bytes;
                    for (int ky = 0; ky < P.ksize; ++ky) {
                        if (!P.w_resident || t == 0) mbar_wait(cv_smem_u32(&bars.w_full[ws]), wph, err, 105);
                        asm volatile("tcgen05.fence::after_thread_sync;");
                        const uint32_t a_row = (((a_st + (uint32_t)ky * a_sbo) >> 4) & 0x3fffu) | a_lbo_f;      // (kx, un, mh, ks, plane) = 0
                        const uint32_t w_row = (((w_base + ws * (uint32_t)P.w_stage_bytes) >> 4) & 0x3fffu) | w_lbo_f;
                        const uint32_t first = (uint32_t)((kc | ky) != 0);
                        if (elect_one()) {
                            if (P.ncat) {
#pragma unroll 1
                                for (int un = 0; un < nun; ++un) {
                                    const uint32_t a_un = a_row + (uint32_t)un * A_UNIT;
                                    const uint32_t d_un = acc0 + (uint32_t)(un * 2) * ACCW;
#pragma unroll
                                    for (int kx = 0; kx < 3; ++kx) {
                                        if (kx < P.ksize) {
#pragma unroll
                                            for (int mh = 0; mh < 2; ++mh) {
                                                const uint32_t d = d_un + (uint32_t)mh * ACCW;
#pragma unroll
                                                for (int ks = 0; ks < 2; ++ks) {
                                                    const uint32_t ah = a_un + (uint32_t)(kx + 8 * mh) + (uint32_t)ks * A_KS;
                                                    const uint32_t wh = w_row + (uint32_t)kx * W_TAP + (uint32_t)ks * W_KS;
                                                    umma_bf16_lohi(d, ah, a_hi_word, wh, w_hi_word, idesc_cat, (kx | ks) ? 1u : first);   // A_hi x [W_hi; W_lo]
                                                    umma_bf16_lohi(d, ah + A_PLANE, a_hi_word, wh, w_hi_word, idesc, 1u);                 // A_lo x W_hi
                                                }
                                            }
                                        }
                                    }
                                }
                            } else {
#pragma unroll 1
                            for (int un = 0; un < nun; ++un) {
                                const uint32_t a_un = a_row + (uint32_t)un * A_UNIT;
                                const uint32_t d_un = acc0 + (uint32_t)(un * 2) * ACCW;
#pragma unroll
                                for (int kx = 0; kx < 3; ++kx) {
                                    if (kx < P.ksize) {
#pragma unroll
                                        for (int mh = 0; mh < 2; ++mh) {
                                            const uint32_t d = d_un + (uint32_t)mh * ACCW;
#pragma unroll
                                            for (int ks = 0; ks < 2; ++ks) {
                                                const uint32_t ah = a_un + (uint32_t)(kx + 8 * mh) + (uint32_t)ks * A_KS;
                                                const uint32_t wh = w_row + (uint32_t)kx * W_TAP + (uint32_t)ks * W_KS;
                                                umma_bf16_lohi(d, ah, a_hi_word, wh, w_hi_word, idesc, (kx | ks) ? 1u : first);
                                                umma_bf16_lohi(d, ah + A_PLANE, a_hi_word, wh, w_hi_word, idesc, 1u);
                                                umma_bf16_lohi(d, ah, a_hi_word, wh + W_PLANE, w_hi_word, idesc, 1u);
                                            }
                                        }
                                    }
                                }
                            }
                            }
                            if (!P.w_resident) umma_commit(cv_smem_u32(&bars.w_empty[ws]));   // filter row free once these MMAs retire
                            if (ky + 1 == P.ksize) umma_commit(cv_smem_u32(&bars.a_empty[as]));   // halo chunk free
                            if (ky + 1 == P.ksize && kc + 1 == n_kc) umma_commit(cv_smem_u32(&bars.acc_full[set]));   // tile complete
                        }
                        __syncwarp();
                        if (++ws == P.w_stages) { ws = 0; wph ^= 1u; }
                    }
                    if (++as == P.a_stages) { as = 0; aph ^= 1u; }
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM lane = GEMM row = pixel (row r: image row r/8 of the unit, pixel r%8 of the half) =====
        const int wq = warp & 3;
        const int C8 = P.cout >> 3;
        const size_t plane = (size_t)P.B * P.Y * C8 * P.X;
        for (int t = 0; t < n_tiles; ++t) {
            const int u0 = u_first + NU * t;
            const int nun = min(NU, u_last - u0);
            const int set = t & 1;
            const uint32_t use = (uint32_t)(t >> 1);
            mbar_wait(cv_smem_u32(&bars.acc_full[set]), use & 1u, err, 106);
            asm volatile("tcgen05.fence::after_thread_sync;");
            for (int un = 0; un < nun; ++un) {
                const int u = u0 + un;
                const int ux = u % P.units_x, r = u / P.units_x;
                const int uy = r % P.units_y, b = r / P.units_y;
                const int y = uy * 16 + wq * 4 + (lane >> 3);
#pragma unroll 1
                for (int mh = 0; mh < 2; ++mh) {
                    const int x = ux * 16 + mh * 8 + (lane & 7);
                    const bool valid = y < P.Y && x < P.X;
                    const uint32_t col0 = (uint32_t)((set * 2 * NU + un * 2 + mh) * P.accw);
                    const size_t row0 = (((size_t)b * P.Y + y) * C8) * P.X + x;
#pragma unroll 1
                    for (int ch = 0; ch * 32 < P.npad; ++ch) {
                        uint32_t v[32];
                        const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + col0 + (uint32_t)(ch * 32);
                        if (P.npad >= 32) {
                            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                                         "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                                         : "r"(taddr));
                        } else {
                            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                                         : "r"(taddr));
#pragma unroll
                            for (int j = 16; j < 32; ++j) v[j] = 0u;
                        }
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (P.ncat) {      // the A_hi x W_lo part sits npad columns further: add it
                            uint32_t w2[32];
                            if (P.npad >= 32) {
                                tmem_ld_x32(taddr + (uint32_t)P.npad, w2);
                            } else {
                                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                                             : "=r"(w2[0]), "=r"(w2[1]), "=r"(w2[2]), "=r"(w2[3]), "=r"(w2[4]), "=r"(w2[5]), "=r"(w2[6]), "=r"(w2[7]),
                                               "=r"(w2[8]), "=r"(w2[9]), "=r"(w2[10]), "=r"(w2[11]), "=r"(w2[12]), "=r"(w2[13]), "=r"(w2[14]), "=r"(w2[15])
                                             : "r"(taddr + (uint32_t)P.npad));
                                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                                for (int j = 16; j < 32; ++j) w2[j] = 0u;
                            }
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__fadd_rn(__uint_as_float(v[j]), __uint_as_float(w2[j])));
                        }
                        if (valid && !(P.debug & 4)) {
#pragma unroll
                            for (int h = 0; h < 4; ++h) {
                                const int c8 = ch * 4 + h;
                                if (out_split && c8 < C8) {
                                    uint32_t hw[4], lw[4];
#pragma unroll
                                    for (int q = 0; q < 4; ++q) {
                                        const int j = h * 8 + 2 * q;
                                        cv_split2(cv_act(__uint_as_float(v[j]) + bias_s[ch * 32 + j], P.act),
                                                  cv_act(__uint_as_float(v[j + 1]) + bias_s[ch * 32 + j + 1], P.act), hw[q], lw[q]);
                                    }
                                    const size_t row = row0 + (size_t)c8 * P.X;
                                    *reinterpret_cast<uint4 *>(out_split + row * 8) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                                    *reinterpret_cast<uint4 *>(out_split + (plane + row) * 8) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                                }
                            }
                            if (out_nchw) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    const int c = ch * 32 + j;
                                    if (c < P.cout)
                                        out_nchw[(((size_t)b * P.cout + c) * P.Y + y) * P.X + x] = cv_act(__uint_as_float(v[j]) + bias_s[c], P.act);
                                }
                            }
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) mbar_arrive(cv_smem_u32(&bars.acc_empty[set]));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (warp == 3) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P.tmem_cols));
}

// ---- layout conversions (API edges and tests; inside the detector the producers write the split form) ----
// fp32 (B,C,Y,X) -> split NHWC8.  Thread = (b, y, chunk, x): 8 strided reads (coalesced across x), two 16-byte stores.
__global__ void __launch_bounds__(256)
split_from_nchw_kernel(int B, int C, int Y, int X, const float *__restrict__ in, __nv_bfloat16 *__restrict__ out) {
    const int C8 = C >> 3;
    const size_t total = (size_t)B * Y * C8 * X;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % X);
        size_t r = i / X;
        const int c8 = (int)(r % C8); r /= C8;
        const int y = (int)(r % Y);
        const int b = (int)(r / Y);
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __ldg(in + (((size_t)b * C + c8 * 8 + e) * Y + y) * X + x);
        uint4 hq, lq;
        split8_bf16(f, hq, lq);
        *reinterpret_cast<uint4 *>(out + i * 8) = hq;
        *reinterpret_cast<uint4 *>(out + (total + i) * 8) = lq;
    }
}

// split NHWC8 -> fp32 (B,C,Y,X): value = hi + lo
__global__ void __launch_bounds__(256)
split_to_nchw_kernel(int B, int C, int Y, int X, const __nv_bfloat16 *__restrict__ in, float *__restrict__ out) {
    const int C8 = C >> 3;
    const size_t total = (size_t)B * Y * C8 * X;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % X);
        size_t r = i / X;
        const int c8 = (int)(r % C8); r /= C8;
        const int y = (int)(r % Y);
        const int b = (int)(r / Y);
        const uint4 h = *reinterpret_cast<const uint4 *>(in + i * 8);
        const uint4 l = *reinterpret_cast<const uint4 *>(in + (total + i) * 8);
        float f[8];
        join8_bf16(h, l, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) out[(((size_t)b * C + c8 * 8 + e) * Y + y) * X + x] = f[e];
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point: no link-time dependency on libcuda
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            (void)cudaGetLastError();
    });
    return fn;
}

// one device error word per process and device, checked lazily (a time-out traps anyway)
static int *conv_err_word() {
    static std::mutex mu;
    static int *words[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!words[dev]) {
        if (cudaMalloc(&words[dev], sizeof(int)) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
        cudaMemset(words[dev], 0, sizeof(int));
    }
    return words[dev];
}

}  // namespace pdm

extern "C" int pdm_act_split_bytes(int b, int c, int y, int x, long long *bytes) {
    if (b < 0 || c < 0 || y < 0 || x < 0 || (c & 7) || !bytes) return pdm::fail(PDM_ERR_INVALID_ARG, "act_split_bytes: bad size");
    *bytes = (long long)2 * b * y * c * x * 2;
    return PDM_OK;
}

extern "C" int pdm_act_split_from_nchw(int b, int c, int y, int x, const float *in, void *out_split, void *stream) {
    using namespace pdm;
    if (b < 0 || c < 0 || y < 0 || x < 0 || (c & 7)) return fail(PDM_ERR_INVALID_ARG, "act_split_from_nchw: bad size (channels must be a multiple of 8)");
    const size_t total = (size_t)b * y * (c >> 3) * x;
    if (total == 0) return PDM_OK;
    if (!in || !out_split) return fail(PDM_ERR_INVALID_ARG, "act_split_from_nchw: null pointer");
    const int grid = (int)((total + 255) / 256 < (size_t)kNumSMs * 16 ? (total + 255) / 256 : (size_t)kNumSMs * 16);
    split_from_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(b, c, y, x, in, (__nv_bfloat16 *)out_split);
    count_launch();
    PDM_CHECK_LAUNCH("act_split_from_nchw");
    return PDM_OK;
}

extern "C" int pdm_act_split_to_nchw(int b, int c, int y, int x, const void *in_split, float *out, void *stream) {
    using namespace pdm;
    if (b < 0 || c < 0 || y < 0 || x < 0 || (c & 7)) return fail(PDM_ERR_INVALID_ARG, "act_split_to_nchw: bad size (channels must be a multiple of 8)");
    const size_t total = (size_t)b * y * (c >> 3) * x;
    if (total == 0) return PDM_OK;
    if (!in_split || !out) return fail(PDM_ERR_INVALID_ARG, "act_split_to_nchw: null pointer");
    const int grid = (int)((total + 255) / 256 < (size_t)kNumSMs * 16 ? (total + 255) / 256 : (size_t)kNumSMs * 16);
    split_to_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(b, c, y, x, (const __nv_bfloat16 *)in_split, out);
    count_launch();
    PDM_CHECK_LAUNCH("act_split_to_nchw");
    return PDM_OK;
}

extern "C" int pdm_conv_tc_forward(int b, int y, int x, int cin, int cout, int ksize, const void *in_split,
                                   const void *w_packed, const float *bias, int act, void *out_split, float *out_nchw,
                                   void *stream) {
    using namespace pdm;
    if (b < 0 || y < 0 || x < 0 || cin <= 0 || cout <= 0) return fail(PDM_ERR_INVALID_ARG, "conv_tc_forward: bad size");
    if (ksize != 1 && ksize != 3) return fail(PDM_ERR_UNSUPPORTED, "conv_tc_forward: kernel size %d (1 or 3)", ksize);
    if (cin % 32 != 0 || cin > 512) return fail(PDM_ERR_UNSUPPORTED, "conv_tc_forward: cin %d (multiple of 32, <= 512)", cin);
    if (cout > kCvMaxN) return fail(PDM_ERR_UNSUPPORTED, "conv_tc_forward: cout %d > %d", cout, kCvMaxN);
    if (act < 0 || act > 2) return fail(PDM_ERR_INVALID_ARG, "conv_tc_forward: act %d", act);
    if (out_split && (cout & 7)) return fail(PDM_ERR_UNSUPPORTED, "conv_tc_forward: split output needs cout %% 8 == 0");
    if (b == 0 || y == 0 || x == 0) return PDM_OK;
    if (!in_split || !w_packed || !bias || (!out_split && !out_nchw)) return fail(PDM_ERR_INVALID_ARG, "conv_tc_forward: null pointer");
    if ((x * 16) % 16 != 0 || ((uintptr_t)in_split & 15) || ((uintptr_t)w_packed & 15))
        return fail(PDM_ERR_INVALID_ARG, "conv_tc_forward: operands must be 16-byte aligned");
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(PDM_ERR_UNSUPPORTED, "conv_tc_forward: cuTensorMapEncodeTiled is not available from this driver");

    ConvTCParams P;
    P.B = b; P.Y = y; P.X = x; P.cin = cin; P.cout = cout; P.ksize = ksize; P.act = act;
    P.npad = cout <= 16 ? 16 : cout <= 32 ? 32 : cout <= 64 ? 64 : 128;
    P.halo = 16 + ksize - 1;
    P.units_x = (x + 15) / 16; P.units_y = (y + 15) / 16;
    P.total_units = b * P.units_x * P.units_y;
    // two accumulator sets always (the epilogue of a tile overlaps the MMAs of the next): a tile is two 16x16 units
    // (4 accumulators) when 8 * npad TMEM columns fit, one unit otherwise
    static const bool ncat_on = [] { const char *e = getenv("PDM_CONV_NCAT"); return !(e && e[0] == '0'); }();
    P.ncat = (P.npad <= 64 && ncat_on) ? 1 : 0;      // the packing (conv_tc.py pack_conv_weight) follows the same rule
    P.accw = P.ncat ? 2 * P.npad : P.npad;
    P.nu = 8 * P.accw <= 512 ? 2 : 1;
    int per = (P.total_units + P.nu * kNumSMs - 1) / (P.nu * kNumSMs) * P.nu;
    if (per < P.nu) per = P.nu;
    P.units_per_cta = per;
    const int grid = (P.total_units + per - 1) / per;
    P.a_unit_bytes = P.halo * 4 * P.halo * 16;
    P.w_block_bytes = 2 * 4 * P.npad * 16;
    P.w_stage_bytes = ksize * P.w_block_bytes;
    int cols = 4 * P.nu * P.accw, pw = 32;
    while (pw < cols) pw <<= 1;
    P.tmem_cols = pw;
    // shared memory: halo ring (2 stages when a tile has two units, else 3), the rest for filter rows (ksize weight
    // blocks each); if every row of the layer fits they stay resident
    const int a_stage = 2 * P.nu * P.a_unit_bytes;
    const int n_rows = (cin / 32) * ksize;
    P.a_stages = P.nu == 2 ? 2 : 3;
    int w_fit = (kCvSmemBudget - P.a_stages * a_stage) / P.w_stage_bytes;
    if (w_fit < 2 && P.a_stages > 2) { P.a_stages = 2; w_fit = (kCvSmemBudget - P.a_stages * a_stage) / P.w_stage_bytes; }
    if (w_fit < 2) return fail(PDM_ERR_UNSUPPORTED, "conv_tc_forward: operands do not fit in shared memory");
    P.w_resident = (w_fit >= n_rows && n_rows <= kCvMaxWStages) ? 1 : 0;
    P.w_stages = P.w_resident ? n_rows : (w_fit < 4 ? w_fit : 4);
    static const int dbg = [] { const char *e = getenv("PDM_CONV_DEBUG"); return e ? atoi(e) : 0; }();
    P.debug = dbg;

    CUtensorMap tmap;
    const cuuint64_t gdim[5] = {(cuuint64_t)x * 8, (cuuint64_t)(cin / 8), (cuuint64_t)y, (cuuint64_t)b, 2};
    const cuuint64_t gstr[4] = {(cuuint64_t)x * 16, (cuuint64_t)x * 16 * (cin / 8), (cuuint64_t)x * 16 * (cin / 8) * y,
                                (cuuint64_t)x * 16 * (cin / 8) * y * b};
    const cuuint32_t box[5] = {(cuuint32_t)P.halo * 8, 4, (cuuint32_t)P.halo, 1, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void *>(in_split), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(PDM_ERR_INVALID_ARG, "conv_tc_forward: cuTensorMapEncodeTiled failed (%d)", (int)cr);

    const size_t smem = (size_t)P.a_stages * a_stage + (size_t)P.w_stages * P.w_stage_bytes + 128;
    if (int rc = ensure_dynamic_smem((const void *)conv_tc_kernel, smem)) return rc;
    conv_tc_kernel<<<grid, kCvThreads, smem, (cudaStream_t)stream>>>(tmap, P, (const unsigned char *)w_packed, bias,
                                                                    (__nv_bfloat16 *)out_split, out_nchw, conv_err_word());
    count_launch();
    PDM_CHECK_LAUNCH("conv_tc_forward");
    return PDM_OK;
}
