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
// ~1e-6 relative, far inside the 1e-3 budget).  The layers are genuine GEMMs of shape
// [128 rows x C_in] x [C_in x C_out] per CTA; a tcgen05 (TF32 / bf16x3) version of the inner
// product is the planned follow-up -- the data flow stays as is.
//
// Thread tile: 8 rows x 4 output channels, activations stored channel-major A[c][row] so that a
// thread's 8 rows are two 16-byte shared loads and its 4 weights one 16-byte (broadcast) load per
// k: 32 FMAs per 3 LDS.128.
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

}  // namespace pdm

// packed: for each layer l, Wt[k][wpad(l+1)] (k < width[l]; transposed, BN folded, zero padded
// columns) followed by bias[wpad(l+1)]; offsets are derived here from `widths`.
extern "C" int pdm_sa_fused_forward(int b, int n, int m, int c_feat, int nsample, int use_xyz,
                                    const float *xyz, const float *features, const float *new_xyz,
                                    const int *idx, int n_layers, const int *widths,
                                    const float *packed, float *out, void *stream) {
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
    const size_t smem = sizeof(float) * ((size_t)2 * P.act_floats + P.w_floats + kSAMaxC);
    if (int rc = ensure_dynamic_smem((const void *)sa_fused_kernel, smem)) return rc;
    const int cpb = kSARows / nsample;
    dim3 grid((m + cpb - 1) / cpb, b);
    sa_fused_kernel<<<grid, kSAThreads, smem, (cudaStream_t)stream>>>(P, xyz, features, new_xyz, idx, packed, out);
    count_launch();
    PDM_CHECK_LAUNCH("sa_fused_forward");
    return PDM_OK;
}
