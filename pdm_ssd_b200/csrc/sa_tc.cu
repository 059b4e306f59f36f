// sa_tc.cu -- fused set-abstraction scale on the tensor cores, persistent and warp-specialised.
//
// Same contract as sa_fused_kernel (sa_fused.cu): grouping (xyz - centre, features) -> [Conv2d 1x1 + folded eval-mode
// BatchNorm + ReLU] x L -> max over the nsample neighbours (pointnet2_utils.py:241-264, pointnet2_modules.py:40-52,
// 90-97), for scales whose layers are genuine GEMMs (KITTI SA2: 35-64-64-128, 14.5 k multiply-adds per row).
// Round 1's tcgen05 kernel ran one 128-row tile per CTA with every stage serialised behind __syncthreads and re-staged
// ~100 KB of weights per CTA (4096 CTAs): tensor pipe 8.5 % active, 0.56 ms per batch of 16.  This one:
//   * a CTA per SM walks over tiles of 128 rows (4 centres x 32 neighbours); the weights of ALL layers are brought in
//     ONCE per CTA by bulk copies (cp.async.bulk -> UBLKCP) and stay in shared memory;
//   * warps 12-15 gather the next tile (one row per thread) into a double-buffered A operand while
//     warp 1 issues the MMAs of the two current tiles (tcgen05.mma.kind::f16, accumulators in TMEM) and
//     warps 4-11 run the epilogues (tcgen05.ld -> bias + ReLU -> operand of the next layer, or the max-pool);
//     hand-offs are mbarriers, no CTA-wide barrier in the steady state;
//   * TWO tiles are in flight per CTA (slots A / B with their own accumulators in TMEM and their own inter-layer operand):
//     the MMA warp issues A.L1 B.L1 A.L2 B.L2 ..., the epilogue warps run A.E1 B.E1 A.E2 B.E2 ..., so the tensor pipe works
//     on one tile while the epilogue of the other drains -- the three layers of ONE tile are strictly serial (six dependent
//     hand-offs), which left every unit idle most of the time when a CTA had a single tile in flight (tensor pipe 20 %);
//   * fp32-grade accuracy without fp32 tensor math: every operand is carried as TWO bf16 values hi + lo (16 mantissa
//     bits) and a product is formed as hh + hl + lh (dropped lo*lo <= 2^-18), the scheme of the convolutions (conv_tc.cu):
//     ~1e-5 against torch fp32 after three layers (budget 1e-3).  (The first version carried three planes hi/mid/lo, six
//     MMAs per product, 1.5e-6: its 264 KB for two tiles in flight do not fit an SM, 202 KB with two planes do.)
// Operand layout: K-major, no swizzle: plane[k/8][row][8 bf16] (core matrix = 8 rows x 16 bytes, LBO = rows*16,
// SBO = 128), exactly what a row-per-thread epilogue writes with conflict-free 16-byte stores.
// Every mbarrier wait is bounded; a time-out records a code in the error word and traps (never continues).
#include <stdlib.h>

#include <mutex>

#include "tc_common.cuh"

namespace pdm {

constexpr int kStRows = 128;
constexpr int kStThreads = 512;          // warp 0: weights, 1: MMA, 2: TMEM, 3: idle, 4-11: epilogue, 12-15: gather
constexpr int kStMaxLayers = 3;
constexpr int kStMaxHidden = 64;         // widest intermediate layer (operand buffer of the next layer)
constexpr int kStMaxOut = 128;

struct SATc3Params {
    int b, n, m, c_feat, nsample, use_xyz, n_layers;
    int width[kStMaxLayers + 1];
    int kpad[kStMaxLayers];      // input width of layer l rounded up to 16
    int npad[kStMaxLayers];      // output width rounded up to 16
    int woff[kStMaxLayers];      // byte offset of layer l's weight planes in the packed buffer / in shared memory
    int wbytes;                  // all weight planes
    int a1_bytes;                // one gather buffer: 2 planes x (kpad[0]/8) x 128 x 16
    int a23_bytes;               // inter-layer operand of one slot: 2 planes x (kStMaxHidden/8) x 128 x 16
    int tiles;                   // ceil(b * m * nsample / 128)
    int dcol[kStMaxLayers];      // TMEM column of layer l's accumulator inside a slot
    int slot_cols;               // TMEM columns of one slot (sum of npad)
    int tmem_cols;
    int debug;                   // PDM_SA_TC3_DEBUG (measurement only, results wrong): 1 no gather, 2 no hidden epilogue work,
                                 // 4 no max-pool work, 8 no MMAs
};

struct __align__(8) SATcBarriers {
    uint64_t w_full;
    uint64_t a1_full[2], a1_empty[2];
    uint64_t d_full[2][kStMaxLayers];       // [slot] layer l's accumulator complete (tcgen05.commit)
    uint64_t a_next_full[2][kStMaxLayers];  // [slot] operand of layer l+1 written by the epilogue of layer l
};

// v -> two bf16 with hi + lo == v to 2^-18; pairs packed for 16-byte stores
__device__ __forceinline__ void split2_pair(float a, float b, uint32_t &hi, uint32_t &lo) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t *>(&t);
    const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xffff0000u);
    t = __floats2bfloat162_rn(ra, rb);
    lo = *reinterpret_cast<const uint32_t *>(&t);
}

// eight consecutive K values of one row -> the row's 16-byte slot in each of the two planes of an operand buffer
__device__ __forceinline__ void store_split2(unsigned char *buf, int plane_bytes, int chunk, int row, const float *v) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) split2_pair(v[2 * q], v[2 * q + 1], h[q], l[q]);
    unsigned char *p = buf + ((size_t)chunk * kStRows + row) * 16;
    *reinterpret_cast<uint4 *>(p) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4 *>(p + plane_bytes) = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(kStThreads, 1)
sa_tc3_kernel(const SATc3Params P, const float *__restrict__ xyz, const float *__restrict__ feats,
              const float *__restrict__ feats_pm, const float *__restrict__ new_xyz, const int *__restrict__ idx,
              const unsigned char *__restrict__ wpacked, const float *__restrict__ bias, float *__restrict__ out,
              float *__restrict__ out_pm, int *__restrict__ err) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ SATcBarriers bars;
    __shared__ uint32_t tmem_base_s;
    __shared__ float bias_s[kStMaxLayers][kStMaxOut];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char *smem = smem_raw + ((128u - (cv_smem_u32(smem_raw) & 127u)) & 127u);
    unsigned char *w_s = smem;                                  // all layers' weight planes
    unsigned char *a1_s = w_s + P.wbytes;                       // [2] gather buffers
    unsigned char *a23_s = a1_s + 2 * P.a1_bytes;               // [2] operand of layers 2..L, one per slot
    const int L = P.n_layers;
    const int my_tiles = (P.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (tid == 0) {
        mbar_init(cv_smem_u32(&bars.w_full), 1);
        for (int s = 0; s < 2; ++s) { mbar_init(cv_smem_u32(&bars.a1_full[s]), 4); mbar_init(cv_smem_u32(&bars.a1_empty[s]), 1); }
        for (int sl = 0; sl < 2; ++sl)
            for (int l = 0; l < kStMaxLayers; ++l) { mbar_init(cv_smem_u32(&bars.d_full[sl][l]), 1); mbar_init(cv_smem_u32(&bars.a_next_full[sl][l]), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(cv_smem_u32(&tmem_base_s)), "r"(P.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int t = tid; t < kStMaxLayers * kStMaxOut; t += kStThreads) {
        const int l = t / kStMaxOut, c = t - l * kStMaxOut;
        bias_s[l][c] = (l < L && c < P.width[l + 1]) ? __ldg(bias + l * kStMaxOut + c) : 0.f;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===== weights: once per CTA =====
        if (elect_one()) {
            const uint32_t full = cv_smem_u32(&bars.w_full);
            mbar_arrive_expect_tx(full, (uint32_t)P.wbytes);
            for (uint32_t off = 0; off < (uint32_t)P.wbytes; off += 32768u) {
                const uint32_t nb = min(32768u, (uint32_t)P.wbytes - off);
                bulk_load_1d(cv_smem_u32(w_s) + off, wpacked + off, nb, full);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t w_hi_word = cv_desc_hi(128u), a_hi_word = cv_desc_hi(128u);
        const uint32_t a_lbo_f = (((uint32_t)kStRows * 16u) >> 4) << 16;
        mbar_wait(cv_smem_u32(&bars.w_full), 0, err, 201);
        for (int t0 = 0; t0 < my_tiles; t0 += 2) {
            for (int l = 0; l < L; ++l) {
                for (int st = 0; st < 2 && t0 + st < my_tiles; ++st) {       // slot = tile parity
                    const uint32_t ph = (uint32_t)((t0 + st) >> 1) & 1u;
                    if (l == 0) mbar_wait(cv_smem_u32(&bars.a1_full[st]), ph, err, 202);
                    else mbar_wait(cv_smem_u32(&bars.a_next_full[st][l - 1]), ph, err, 203);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const int K = P.kpad[l], N = P.npad[l];
                    const uint32_t a_base = l == 0 ? cv_smem_u32(a1_s) + (uint32_t)(st * P.a1_bytes) : cv_smem_u32(a23_s) + (uint32_t)(st * P.a23_bytes);
                    const uint32_t w_base = cv_smem_u32(w_s) + (uint32_t)P.woff[l];
                    // descriptor low words (start address >> 4 | LBO) + constants in units of 16 bytes: below N = 256 the issuing
                    // thread bounds the MMA rate (tools/micro/umma_rate.cu), so nothing but adds sits between two MMAs
                    const uint32_t a_lo0 = ((a_base >> 4) & 0x3fffu) | a_lbo_f;
                    const uint32_t w_lo0 = ((w_base >> 4) & 0x3fffu) | (((((uint32_t)N * 16u) >> 4) & 0x3fffu) << 16);
                    const uint32_t A_PLANE = (uint32_t)(K / 8) * kStRows, W_PLANE = (uint32_t)(K / 8) * (uint32_t)N;   // (bytes >> 4)
                    const uint32_t A_KS = 2u * kStRows, W_KS = 2u * (uint32_t)N;
                    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kStRows >> 4) << 24);
                    const uint32_t d = tmem_base + (uint32_t)(st * P.slot_cols + P.dcol[l]);
                    if (!(P.debug & 8) && elect_one()) {
#pragma unroll 1
                        for (int ks = 0; ks < K / 16; ++ks) {
                            const uint32_t a0 = a_lo0 + (uint32_t)ks * A_KS, w0 = w_lo0 + (uint32_t)ks * W_KS;
                            umma_bf16_lohi(d, a0, a_hi_word, w0, w_hi_word, idesc, (uint32_t)(ks != 0));     // hi*hi
                            umma_bf16_lohi(d, a0, a_hi_word, w0 + W_PLANE, w_hi_word, idesc, 1u);            // hi*lo
                            umma_bf16_lohi(d, a0 + A_PLANE, a_hi_word, w0, w_hi_word, idesc, 1u);            // lo*hi
                        }
                    }
                    __syncwarp();
                    if (elect_one()) {
                        if (l == 0) umma_commit(cv_smem_u32(&bars.a1_empty[st]));      // gather buffer free once layer 1 retires
                        umma_commit(cv_smem_u32(&bars.d_full[st][l]));
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 4 && warp < 12) {
        // ===== epilogues: thread = row = TMEM lane; the two warps of a lane quadrant take alternate 32-column chunks =====
        const int wq = warp & 3, row = wq * 32 + lane, half = (warp - 4) >> 2;
        const int cout = P.width[L];
        for (int t0 = 0; t0 < my_tiles; t0 += 2) {
            for (int l = 0; l < L; ++l) {
                for (int st = 0; st < 2 && t0 + st < my_tiles; ++st) {
                    const uint32_t ph = (uint32_t)((t0 + st) >> 1) & 1u;
                    const long long tile = (long long)blockIdx.x + (long long)(t0 + st) * gridDim.x;
                    unsigned char *a23 = a23_s + st * P.a23_bytes;
                    mbar_wait(cv_smem_u32(&bars.d_full[st][l]), ph, err, 204);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const int N = P.npad[l];
                    const bool last = l + 1 == L;
                    for (int ch = half; ch * 32 < N; ch += 2) {
                        uint32_t v[32];
                        tmem_ld_x32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(st * P.slot_cols + P.dcol[l] + ch * 32), v);
                        if (!last) {
                            if (P.debug & 2) continue;
                            // bias + ReLU -> next layer's operand (columns >= width are exact zeros: zero weights, zero bias)
                            const int knext = P.kpad[l + 1];
                            const int a23_plane = (knext / 8) * kStRows * 16;      // plane stride of the operand the next MMAs read
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int c0 = ch * 32 + q * 8;
                                if (c0 < knext) {
                                    float f[8];
#pragma unroll
                                    for (int e = 0; e < 8; ++e) f[e] = fmaxf(__uint_as_float(v[q * 8 + e]) + bias_s[l][c0 + e], 0.f);
                                    store_split2(a23, a23_plane, c0 >> 3, row, f);
                                }
                            }
                        } else {
                            // max over the 32 neighbours of the warp's centre (values >= 0: bit patterns order like floats);
                            // lane j keeps column j of the chunk, so the stores below are one value per lane
                            unsigned keep = 0u;
                            if (P.debug & 4) continue;
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const unsigned bits = __float_as_uint(fmaxf(__uint_as_float(v[j]) + bias_s[l][ch * 32 + j], 0.f));
                                const unsigned mx = __reduce_max_sync(0xffffffffu, bits);
                                if (lane == j) keep = mx;
                            }
                            const long long centre = tile * (kStRows / 32) + wq;
                            const int c = ch * 32 + lane;
                            if (c < cout && centre < (long long)P.b * P.m) {
                                const int bi = (int)(centre / P.m), mi = (int)(centre - (long long)bi * P.m);
                                out[((size_t)bi * cout + c) * P.m + mi] = __uint_as_float(keep);
                                if (out_pm) out_pm[(size_t)centre * cout + c] = __uint_as_float(keep);
                            }
                        }
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;");
                    if (!last) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core
                        __syncwarp();
                        if (lane == 0) mbar_arrive(cv_smem_u32(&bars.a_next_full[st][l]));
                    }
                }
            }
        }
    } else if (warp >= 12) {
        // ===== gather: one row per thread into the double-buffered layer-1 operand =====
        const int row = (warp - 12) * 32 + lane;
        const int S = P.nsample;
        const int a1_plane = (P.kpad[0] / 8) * kStRows * 16;
        const long long total_rows = (long long)P.b * P.m * S;
        const int c0 = P.use_xyz ? 3 : 0;
        for (int t = 0; t < my_tiles; ++t) {
            const int st = t & 1;
            const uint32_t ph_ring = (uint32_t)(t >> 1) & 1u;
            const long long tile = (long long)blockIdx.x + (long long)t * gridDim.x;
            long long r = tile * kStRows + row;
            if (r >= total_rows) r = total_rows - 1;                 // rows past the end are computed and dropped
            const long long centre = r / S;
            const int bi = (int)(centre / P.m);
            const int id = __ldg(idx + r);
            mbar_wait(cv_smem_u32(&bars.a1_empty[st]), ph_ring ^ 1u, err, 205);
            unsigned char *buf = a1_s + st * P.a1_bytes;
            const float *pp = xyz + ((size_t)bi * P.n + id) * 3;
            const float *qq = new_xyz + (size_t)centre * 3;
            const float *fc = feats + (size_t)bi * P.c_feat * P.n + id;                     // channel-major (B, C, N)
            const float *fp = feats_pm ? feats_pm + ((size_t)bi * P.n + id) * P.c_feat : nullptr;   // point-major (B, N, C)
            for (int k8 = 0; k8 * 8 < P.kpad[0] && !(P.debug & 1); ++k8) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int k = k8 * 8 + e;
                    float val = 0.f;                                                        // K padding
                    if (k < c0) val = __fsub_rn(__ldg(pp + k), __ldg(qq + k));
                    else if (k < P.width[0]) val = fp ? __ldg(fp + (k - c0)) : __ldg(fc + (size_t)(k - c0) * P.n);
                    f[e] = val;
                }
                store_split2(buf, a1_plane, k8, row, f);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(cv_smem_u32(&bars.a1_full[st]));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P.tmem_cols));
}

// one device error word per device (a time-out traps anyway; the word says which barrier)
static int *sa_tc_err_word() {
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

// Returns -1 when the scale does not fit this kernel (the caller falls back), else a PDM code.
// wpacked: bf16 planes per layer l: [hi|lo][kpad/8][npad][8] of the BN-folded W'[n][k]; bias: fp32 [L][128].
int sa_tc3_try(int b, int n, int m, int c_feat, int nsample, int use_xyz, const float *xyz, const float *feats,
               const float *feats_pm, const float *new_xyz, const int *idx, int n_layers, const int *widths,
               const void *wpacked, const float *bias, float *out, float *out_pm, cudaStream_t st) {
    if (!wpacked || !bias || nsample != 32 || n_layers < 2 || n_layers > kStMaxLayers) return -1;
    SATc3Params P;
    P.b = b; P.n = n; P.m = m; P.c_feat = c_feat; P.nsample = nsample; P.use_xyz = use_xyz ? 1 : 0; P.n_layers = n_layers;
    int off = 0, col = 0;
    for (int l = 0; l <= n_layers; ++l) P.width[l] = widths[l];
    for (int l = 0; l < n_layers; ++l) {
        P.kpad[l] = (widths[l] + 15) / 16 * 16;
        P.npad[l] = (widths[l + 1] + 15) / 16 * 16;
        if (l + 1 < n_layers && (P.npad[l] > kStMaxHidden)) return -1;
        if (P.npad[l] > kStMaxOut || P.kpad[l] > 96) return -1;
        P.woff[l] = off;
        off += 2 * (P.kpad[l] / 8) * P.npad[l] * 16;
        P.dcol[l] = col;
        col += P.npad[l];
    }
    for (int l = 0; l + 1 < n_layers; ++l)
        if (P.kpad[l + 1] > P.npad[l] || P.kpad[l + 1] > kStMaxHidden) return -1;     // layer l+1 reads what layer l's epilogue wrote
    P.wbytes = off;
    P.a1_bytes = 2 * (P.kpad[0] / 8) * kStRows * 16;
    P.a23_bytes = 2 * (kStMaxHidden / 8) * kStRows * 16;
    P.slot_cols = col;
    int pw = 32;
    while (pw < 2 * col) pw <<= 1;
    if (pw > 512) return -1;
    P.tmem_cols = pw;
    const long long rows = (long long)b * m * nsample;
    P.tiles = (int)((rows + kStRows - 1) / kStRows);
    const char *dbg = getenv("PDM_SA_TC3_DEBUG");
    P.debug = dbg ? atoi(dbg) : 0;
    const size_t smem = (size_t)P.wbytes + 2 * (size_t)P.a1_bytes + 2 * (size_t)P.a23_bytes + 128;
    if (smem > 227 * 1024 - 2560) return -1;
    if (((uintptr_t)wpacked & 15) != 0) return fail(PDM_ERR_INVALID_ARG, "sa_fused_forward: tensor-core weights must be 16-byte aligned");
    if (int rc = ensure_dynamic_smem((const void *)sa_tc3_kernel, smem)) return rc;
    const int grid = P.tiles < kNumSMs ? P.tiles : kNumSMs;
    sa_tc3_kernel<<<grid, kStThreads, smem, st>>>(P, xyz, feats, feats_pm, new_xyz, idx, (const unsigned char *)wpacked, bias, out, out_pm,
                                                 sa_tc_err_word());
    count_launch();
    PDM_CHECK_LAUNCH("sa_fused_forward(tcgen05 x3)");
    return PDM_OK;
}

}  // namespace pdm
