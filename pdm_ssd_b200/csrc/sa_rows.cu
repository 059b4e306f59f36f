// sa_rows.cu -- fused set-abstraction scale for NARROW shared MLPs (first SA layer: 3 + C_feat <= 8 input channels,
// hidden widths 16/32, output width 32/64): one THREAD per (centre, neighbour) row, every activation in registers.
//
// Same contract as sa_fused_kernel (sa_fused.cu): grouping (xyz - centre, features) -> Conv2d 1x1 + folded eval-mode
// BatchNorm + ReLU per layer -> max over the nsample neighbours (pointnet2_utils.py:241-264, pointnet2_modules.py:
// 40-52,90-97).  At these widths the layers are not GEMMs worth staging: the KITTI SA1 stack 4-16-16-32 is 832
// multiply-adds per row, while the 128-row shared-memory tiling of sa_fused_kernel spends its time in the gather,
// two barriers per layer and shared-memory round trips of the activations (0.46 ms per batch of 16).  Here a warp owns
// the 32 neighbours of one centre (nsample 32; two centres per warp for nsample 16): the gather is one index load and
// a handful of coordinate / feature loads per thread, the whole MLP runs out of registers with the weights broadcast
// from shared memory (one LDS.128 per four multiply-adds, conflict-free), and the max-pool is a `redux.sync` per
// output channel on the bit patterns (values are >= 0 after ReLU).  Outputs are staged per CTA so that every global
// store fills whole 32-byte sectors of the (B, C_out, M) tensor; optionally the same features are also written
// point-major (B, M, C_out), the layout the next layer's gather and the detector's point_features want.
#include "common.cuh"

namespace pdm {

constexpr int kRowsThreads = 256;

struct SARowsParams {
    int n, m, c_feat, nsample, use_xyz;
    int width0;                 // real input width (<= K0)
    int woff[3], boff[3];       // offsets (floats) into `packed` (layout of _pack_folded: Wt[k][pad4(out)], bias[pad4(out)])
    int groups;                 // ceil(b * m / centres per CTA)
    int b;
};

// acc[j] = bias[j] + sum_k in[k] * W[k][j], weights broadcast from shared memory, then ReLU
template <int KIN, int COUT>
__device__ __forceinline__ void rows_layer(const float (&in)[KIN], float (&acc)[COUT], const float *__restrict__ w, const float *__restrict__ bias) {
#pragma unroll
    for (int j = 0; j < COUT; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4 *>(bias + j);
        acc[j] = b4.x; acc[j + 1] = b4.y; acc[j + 2] = b4.z; acc[j + 3] = b4.w;
    }
#pragma unroll
    for (int k = 0; k < KIN; ++k) {
        const float a = in[k];
#pragma unroll
        for (int j = 0; j < COUT; j += 4) {
            const float4 w4 = *reinterpret_cast<const float4 *>(w + k * COUT + j);
            ffma2(acc[j], acc[j + 1], a, w4.x, w4.y);
            ffma2(acc[j + 2], acc[j + 3], a, w4.z, w4.w);
        }
    }
#pragma unroll
    for (int j = 0; j < COUT; ++j) acc[j] = fmaxf(acc[j], 0.f);
}

// the same for the TWO rows of a thread: every weight vector read from shared memory feeds both (ncu after the FFMA2 change:
// short_scoreboard + mio_throttle on top -- the broadcast LDS.128 per two FFMA2 was the bottleneck)
template <int KIN, int COUT>
__device__ __forceinline__ void rows_layer2(const float (&in)[2][KIN], float (&acc)[2][COUT], const float *__restrict__ w,
                                            const float *__restrict__ bias) {
#pragma unroll
    for (int j = 0; j < COUT; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4 *>(bias + j);
#pragma unroll
        for (int r = 0; r < 2; ++r) { acc[r][j] = b4.x; acc[r][j + 1] = b4.y; acc[r][j + 2] = b4.z; acc[r][j + 3] = b4.w; }
    }
#pragma unroll
    for (int k = 0; k < KIN; ++k) {
#pragma unroll
        for (int j = 0; j < COUT; j += 4) {
            const float4 w4 = *reinterpret_cast<const float4 *>(w + k * COUT + j);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                ffma2(acc[r][j], acc[r][j + 1], in[r][k], w4.x, w4.y);
                ffma2(acc[r][j + 2], acc[r][j + 3], in[r][k], w4.z, w4.w);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int j = 0; j < COUT; ++j) acc[r][j] = fmaxf(acc[r][j], 0.f);
}

template <int K0, int C1, int C2, int C3>     // C3 == 0: two layers, C2 is the output width
__global__ void __launch_bounds__(kRowsThreads)
sa_rows_kernel(SARowsParams P, const float *__restrict__ xyz, const float *__restrict__ feats, const float *__restrict__ new_xyz,
               const int *__restrict__ idx, const float *__restrict__ packed, float *__restrict__ out, float *__restrict__ out_pm) {
    constexpr int COUT = C3 > 0 ? C3 : C2;
    constexpr int CLAST_IN = C3 > 0 ? C2 : C1;
    __shared__ __align__(16) float w1[K0 * C1], b1[C1], w2[C1 * C2], b2[C2];
    __shared__ __align__(16) float w3[C3 > 0 ? C2 * C3 : 4], b3[C3 > 0 ? C3 : 4];
    __shared__ unsigned tile[COUT][2 * kRowsThreads / 16 + 1];    // [channel][centre of this CTA iteration], bit patterns
    const int tid = threadIdx.x, lane = tid & 31;
    // ---- weights -> shared memory (zero rows for the padded input channels) ---------------------------------------
    for (int t = tid; t < K0 * C1; t += kRowsThreads) w1[t] = (t / C1) < P.width0 ? __ldg(packed + P.woff[0] + t) : 0.f;
    for (int t = tid; t < C1; t += kRowsThreads) b1[t] = __ldg(packed + P.boff[0] + t);
    for (int t = tid; t < C1 * C2; t += kRowsThreads) w2[t] = __ldg(packed + P.woff[1] + t);
    for (int t = tid; t < C2; t += kRowsThreads) b2[t] = __ldg(packed + P.boff[1] + t);
    if (C3 > 0) {
        for (int t = tid; t < C2 * C3; t += kRowsThreads) w3[t] = __ldg(packed + P.woff[2] + t);
        for (int t = tid; t < C3; t += kRowsThreads) b3[t] = __ldg(packed + P.boff[2] + t);
    }
    __syncthreads();
    const int S = P.nsample;                      // 16 or 32
    const int cph = kRowsThreads / S;             // centres per half of a CTA iteration
    const int cpb = 2 * cph;                      // centres per CTA iteration: a thread owns one row of TWO centres
    const int total = P.b * P.m;
    const unsigned gmask = S == 32 ? 0xffffffffu : (0xffffu << (lane & 16));
    for (int g = blockIdx.x; g < P.groups; g += gridDim.x) {
        const int cl = tid / S, s = tid - cl * S;
        float in[2][K0];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int ci = min(g * cpb + r * cph + cl, total - 1);          // centres past the end are computed and dropped
            const int bi = ci / P.m;
            const int id = __ldg(idx + (size_t)ci * S + s);
#pragma unroll
            for (int k = 0; k < K0; ++k) in[r][k] = 0.f;
            int c0 = 0;
            if (P.use_xyz) {
                const float *pp = xyz + ((size_t)bi * P.n + id) * 3;
                const float *qq = new_xyz + (size_t)ci * 3;
#pragma unroll
                for (int a = 0; a < 3; ++a) in[r][a] = __fsub_rn(__ldg(pp + a), __ldg(qq + a));
                c0 = 3;
            }
            const float *f = feats + (size_t)bi * P.c_feat * P.n + id;
#pragma unroll
            for (int k = 0; k < K0; ++k)
                if (k >= c0 && k - c0 < P.c_feat) in[r][k] = __ldg(f + (size_t)(k - c0) * P.n);
        }
        float a1[2][C1], a2[2][C2];
        rows_layer2<K0, C1>(in, a1, w1, b1);
        rows_layer2<C1, C2>(a1, a2, w2, b2);
        // last layer in chunks of 8 output channels straight into the max-pool
        if (C3 > 0) {
#pragma unroll
            for (int j0 = 0; j0 < COUT; j0 += 8) {
                float acc[2][8];
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[r][j] = b3[j0 + j];
#pragma unroll
                for (int k = 0; k < CLAST_IN; ++k) {
                    const float4 wa = *reinterpret_cast<const float4 *>(w3 + k * COUT + j0);
                    const float4 wb = *reinterpret_cast<const float4 *>(w3 + k * COUT + j0 + 4);
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const float a = a2[r][k];
                        ffma2(acc[r][0], acc[r][1], a, wa.x, wa.y); ffma2(acc[r][2], acc[r][3], a, wa.z, wa.w);
                        ffma2(acc[r][4], acc[r][5], a, wb.x, wb.y); ffma2(acc[r][6], acc[r][7], a, wb.z, wb.w);
                    }
                }
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const unsigned mx = __reduce_max_sync(gmask, __float_as_uint(fmaxf(acc[r][j], 0.f)));
                        if (s == 0) tile[j0 + j][r * cph + cl] = mx;
                    }
            }
        } else {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int j = 0; j < COUT; ++j) {
                    const unsigned mx = __reduce_max_sync(gmask, __float_as_uint(a2[r][j]));
                    if (s == 0) tile[j][r * cph + cl] = mx;
                }
        }
        __syncthreads();
        // ---- stores: (B, C_out, M) in runs of `cpb` consecutive centres per channel; optional point-major copy -----
        for (int t = tid; t < COUT * cpb; t += kRowsThreads) {
            const int c = t / cpb, q = t - c * cpb;
            const int cj = g * cpb + q;
            if (cj < total) {
                const int bj = cj / P.m, mj = cj - bj * P.m;
                out[((size_t)bj * COUT + c) * P.m + mj] = __uint_as_float(tile[c][q]);
            }
        }
        if (out_pm) {
            for (int t = tid; t < COUT * cpb; t += kRowsThreads) {
                const int q = t / COUT, c = t - q * COUT;
                const int cj = g * cpb + q;
                if (cj < total) out_pm[(size_t)cj * COUT + c] = __uint_as_float(tile[c][q]);
            }
        }
        __syncthreads();
    }
}

template <int K0, int C1, int C2, int C3>
static int launch_rows(const SARowsParams &P, const float *xyz, const float *feats, const float *new_xyz, const int *idx,
                       const float *packed, float *out, float *out_pm, cudaStream_t st) {
    const int grid = P.groups < kNumSMs * 6 ? P.groups : kNumSMs * 6;
    sa_rows_kernel<K0, C1, C2, C3><<<grid, kRowsThreads, 0, st>>>(P, xyz, feats, new_xyz, idx, packed, out, out_pm);
    count_launch();
    PDM_CHECK_LAUNCH("sa_fused_forward(rows)");
    return PDM_OK;
}

// Returns -1 when the scale does not fit this kernel (the caller falls back to sa_fused_kernel), else a PDM code.
int sa_rows_try(int b, int n, int m, int c_feat, int nsample, int use_xyz, const float *xyz, const float *feats,
                const float *new_xyz, const int *idx, int n_layers, const int *widths, const float *packed, float *out,
                float *out_pm, cudaStream_t st) {
    if (n_layers < 2 || n_layers > 3 || (nsample != 16 && nsample != 32) || widths[0] > 8) return -1;
    for (int l = 1; l <= n_layers; ++l)
        if (widths[l] % 4 != 0) return -1;
    SARowsParams P;
    P.n = n; P.m = m; P.c_feat = c_feat; P.nsample = nsample; P.use_xyz = use_xyz ? 1 : 0; P.width0 = widths[0]; P.b = b;
    int off = 0;
    for (int l = 0; l < n_layers; ++l) {
        P.woff[l] = off; off += widths[l] * widths[l + 1];       // widths are multiples of 4: pad4(w) == w
        P.boff[l] = off; off += widths[l + 1];
    }
    const int cpb = 2 * kRowsThreads / nsample;      // a thread owns one row of two centres
    P.groups = (int)(((long long)b * m + cpb - 1) / cpb);
    const int k0 = widths[0] <= 4 ? 4 : 8, c1 = widths[1], c2 = widths[2], c3 = n_layers == 3 ? widths[3] : 0;
#define PDM_ROWS(K0, C1, C2, C3) \
    if (k0 == K0 && c1 == C1 && c2 == C2 && c3 == C3) return launch_rows<K0, C1, C2, C3>(P, xyz, feats, new_xyz, idx, packed, out, out_pm, st)
    PDM_ROWS(4, 16, 16, 32); PDM_ROWS(8, 16, 16, 32);
    PDM_ROWS(4, 16, 32, 0);  PDM_ROWS(8, 16, 32, 0);
    PDM_ROWS(4, 32, 32, 64); PDM_ROWS(8, 32, 32, 64);
    PDM_ROWS(4, 16, 32, 64); PDM_ROWS(8, 16, 32, 64);
    PDM_ROWS(4, 32, 64, 0);  PDM_ROWS(8, 32, 64, 0);
#undef PDM_ROWS
    return -1;
}

}  // namespace pdm
