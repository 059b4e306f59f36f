// point_head.cu -- the per-point half of the hybrid head in one kernel, and the small per-point Linear of
// the neck.
//
// pdm_point_head_forward fuses, for every sampled centre (row):
//   * feature fusion: [point feature | BEV context feature under the point] -- the BEV feature is read from
//     the split NHWC8 map the heatmap branch's shared conv wrote (conv_tc.cu), the scene-heatmap value from
//     the fp32 (B, n_class, Y, X) map; pillar of a point = floor((p - range_min) / voxel), the neck's and
//     pcdet's convention (dynamic_voxel_vfe.py:60-71), clamped to the map;
//   * the two FC stacks of PointHeadTemplate.make_fc_layers (point_head_template.py:36-47:
//     Linear(bias=False) + BatchNorm1d + ReLU, then Linear(bias=True)) for classes and box residuals, with the
//     eval-mode BatchNorm folded into the first Linear; both hidden layers are computed as ONE
//     [rows x C_in] x [C_in x (Hc + Hb)] product;
//   * score calibration: sigmoid(cls) * sqrt(heatmap at the point), best class;
//   * PointResidualCoder.decode_torch with mean sizes (box_coder_utils.py:189-222).
// CUDA cores, fp32 FMA: 0.85 GMAC for 16 384 points -- a GEMM of this size does not amortise a tensor-core
// pipeline (BASELINE north_star: "tcgen05 only where they are genuine dense GEMMs").  A CTA owns 64 rows;
// activations sit channel-major in shared memory ([c][row]); thread tile = 4 rows x 4 hidden units.
#include "common.cuh"

namespace pdm {

constexpr int kPHRows = 64;
constexpr int kPHThreads = 256;
constexpr int kPHColBlock = 64;     // hidden units per weight block staged in shared memory
constexpr int kPHMaxHidden = 256;
constexpr int kPHMaxClass = 8;

struct PointHeadParams {
    int P, B, Cp, Cs, Y, X, ncls, hc, hb;   // hc / hb: hidden widths of the class / box stacks
    float xmin, ymin, vx, vy;
};

__global__ void __launch_bounds__(kPHThreads)
point_head_kernel(PointHeadParams Q, const float *__restrict__ coords, const float *__restrict__ pfeat,
                  const __nv_bfloat16 *__restrict__ bev_split, const float *__restrict__ hm,
                  const float *__restrict__ w1t /*[Cin][hc+hb]*/, const float *__restrict__ b1 /*[hc+hb]*/,
                  const float *__restrict__ w2c /*[ncls][hc]*/, const float *__restrict__ b2c,
                  const float *__restrict__ w2b /*[8][hb]*/, const float *__restrict__ b2b,
                  const float *__restrict__ mean_size /*[ncls][3]*/,
                  float *__restrict__ scores /*(P,ncls)*/, float *__restrict__ boxes /*(P,7)*/,
                  float *__restrict__ best /*(P)*/, int *__restrict__ label /*(P)*/,
                  float *__restrict__ cls_raw /*(P,ncls) optional*/, float *__restrict__ box_raw /*(P,8) optional*/) {
    extern __shared__ __align__(16) float sm[];
    const int Cin = Q.Cp + Q.Cs, H = Q.hc + Q.hb;
    float *A = sm;                                  // [Cin][64]
    float *Hd = A + Cin * kPHRows;                  // [H][64]
    float *Wb = Hd + H * kPHRows;                   // [Cin][64] current weight block
    float *out2 = Wb + Cin * kPHColBlock;           // [ncls + 8][64]
    float *hm_s = out2 + (kPHMaxClass + 8) * kPHRows;   // [ncls][64]
    const int tid = threadIdx.x;
    const int p0 = blockIdx.x * kPHRows;

    // ---- gather: lanes over rows, so every shared store is conflict-free -----------------------------------
    {
        const int r = tid & (kPHRows - 1), part = tid >> 6;          // 4 threads per row split the channels
        const int p = min(p0 + r, Q.P - 1);
        const float *pc = coords + (size_t)p * 4;
        const int b = min(max((int)__ldg(pc), 0), Q.B - 1);
        const int cx = min(max((int)floorf(__fdiv_rn(__fsub_rn(__ldg(pc + 1), Q.xmin), Q.vx)), 0), Q.X - 1);
        const int cy = min(max((int)floorf(__fdiv_rn(__fsub_rn(__ldg(pc + 2), Q.ymin), Q.vy)), 0), Q.Y - 1);
        const float *f = pfeat + (size_t)p * Q.Cp;
        for (int c = part * 4; c < Q.Cp; c += 16) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(f + c));
            A[(c + 0) * kPHRows + r] = v.x; A[(c + 1) * kPHRows + r] = v.y;
            A[(c + 2) * kPHRows + r] = v.z; A[(c + 3) * kPHRows + r] = v.w;
        }
        const int C8 = Q.Cs >> 3;
        const size_t lo_off = (size_t)Q.B * Q.Y * C8 * Q.X * 8;
        for (int c8 = part; c8 < C8; c8 += 4) {
            const __nv_bfloat16 *src = bev_split + ((((size_t)b * Q.Y + cy) * C8 + c8) * Q.X + cx) * 8;
            const uint4 hq = __ldg(reinterpret_cast<const uint4 *>(src));
            const uint4 lq = __ldg(reinterpret_cast<const uint4 *>(src + lo_off));
            float v[8];
            join8_bf16(hq, lq, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) A[(Q.Cp + c8 * 8 + e) * kPHRows + r] = v[e];
        }
        for (int k = part; k < Q.ncls; k += 4)
            hm_s[k * kPHRows + r] = __ldg(hm + (((size_t)b * Q.ncls + k) * Q.Y + cy) * Q.X + cx);
    }

    // ---- hidden layers of both stacks: Hd[h][r] = relu(sum_c A[c][r] * W1t[c][h] + b1[h]) ---------------------
    const int rt = tid & 15, ct = tid >> 4;      // 16 row tiles x 16 column tiles of 4
    for (int h0 = 0; h0 < H; h0 += kPHColBlock) {
        const int hn = min(kPHColBlock, H - h0);
        __syncthreads();                          // A complete / previous block's weights dead
        for (int t = tid; t < Cin * (kPHColBlock / 4); t += kPHThreads) {
            const int c = t / (kPHColBlock / 4), q = t - c * (kPHColBlock / 4);
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q * 4 < hn) w = __ldg(reinterpret_cast<const float4 *>(w1t + (size_t)c * H + h0 + q * 4));
            *reinterpret_cast<float4 *>(Wb + c * kPHColBlock + q * 4) = w;
        }
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[a][q] = 0.f;
#pragma unroll 4
        for (int c = 0; c < Cin; ++c) {
            const float4 a4 = *reinterpret_cast<const float4 *>(A + c * kPHRows + rt * 4);
            const float4 w4 = *reinterpret_cast<const float4 *>(Wb + c * kPHColBlock + ct * 4);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                ffma2(acc[a][0], acc[a][1], av[a], wv[0], wv[1]);
                ffma2(acc[a][2], acc[a][3], av[a], wv[2], wv[3]);
            }
        }
        if (ct * 4 < hn) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int h = h0 + ct * 4 + q;
                const float bb = __ldg(b1 + h);
                *reinterpret_cast<float4 *>(Hd + h * kPHRows + rt * 4) =
                    make_float4(fmaxf(acc[0][q] + bb, 0.f), fmaxf(acc[1][q] + bb, 0.f), fmaxf(acc[2][q] + bb, 0.f), fmaxf(acc[3][q] + bb, 0.f));
            }
        }
    }
    __syncthreads();

    // ---- output layers: ncls class logits from Hd[0:hc], 8 box residuals from Hd[hc:hc+hb] -------------------
    {
        const int r = tid & (kPHRows - 1), part = tid >> 6;
        const int nout = Q.ncls + 8;
        for (int o = part; o < nout; o += 4) {
            const bool is_cls = o < Q.ncls;
            const float *w = is_cls ? w2c + (size_t)o * Q.hc : w2b + (size_t)(o - Q.ncls) * Q.hb;
            const float *hsrc = is_cls ? Hd : Hd + Q.hc * kPHRows;
            const int n = is_cls ? Q.hc : Q.hb;
            float s = 0.f;
            for (int k = 0; k < n; ++k) s = fmaf(hsrc[k * kPHRows + r], __ldg(w + k), s);
            out2[o * kPHRows + r] = s + (is_cls ? __ldg(b2c + o) : __ldg(b2b + o - Q.ncls));
        }
    }
    __syncthreads();

    // ---- scores, best class, box decode --------------------------------------------------------------------
    if (tid < kPHRows && p0 + tid < Q.P) {
        const int r = tid, p = p0 + r;
        float bs = -1.f;
        int bl = 0;
        for (int k = 0; k < Q.ncls; ++k) {
            const float logit = out2[k * kPHRows + r];
            const float s = (1.f / (1.f + expf(-logit))) * sqrtf(hm_s[k * kPHRows + r]);
            scores[(size_t)p * Q.ncls + k] = s;
            if (cls_raw) cls_raw[(size_t)p * Q.ncls + k] = logit;
            if (s > bs) { bs = s; bl = k; }           // first maximum, like torch.max
        }
        best[p] = bs;
        label[p] = bl;
        float e[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            e[j] = out2[(Q.ncls + j) * kPHRows + r];
            if (box_raw) box_raw[(size_t)p * 8 + j] = e[j];
        }
        const float dxa = __ldg(mean_size + bl * 3), dya = __ldg(mean_size + bl * 3 + 1), dza = __ldg(mean_size + bl * 3 + 2);
        const float diag = sqrtf(__fadd_rn(__fmul_rn(dxa, dxa), __fmul_rn(dya, dya)));
        const float *pc = coords + (size_t)p * 4;
        float *o = boxes + (size_t)p * 7;
        o[0] = __fadd_rn(__fmul_rn(e[0], diag), __ldg(pc + 1));
        o[1] = __fadd_rn(__fmul_rn(e[1], diag), __ldg(pc + 2));
        o[2] = __fadd_rn(__fmul_rn(e[2], dza), __ldg(pc + 3));
        o[3] = __fmul_rn(expf(e[3]), dxa);
        o[4] = __fmul_rn(expf(e[4]), dya);
        o[5] = __fmul_rn(expf(e[5]), dza);
        o[6] = atan2f(e[7], e[6]);                    // encoding order [.., cos, sin] (box_coder_utils.py:197)
    }
}

// out (P, nout) = x (P, C) W^T + b: one warp per row, lanes over channels (coalesced row reads), shuffle reduce.
// The neck's per-centre SH coefficients (nn.Linear(C, (L+1)^2), SPEC_PDM.md): 16 384 x 128 x 9.
__global__ void __launch_bounds__(256)
linear_rows_kernel(int P, int C, int nout, const float *__restrict__ x, const float *__restrict__ w /*[nout][C]*/,
                   const float *__restrict__ b, float *__restrict__ out) {
    extern __shared__ float wsm[];                 // [nout][C]
    for (int t = threadIdx.x; t < nout * C; t += blockDim.x) wsm[t] = __ldg(w + t);
    __syncthreads();
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int p = blockIdx.x * wpb + (threadIdx.x >> 5); p < P; p += gridDim.x * wpb) {
        float acc[16];
#pragma unroll
        for (int o = 0; o < 16; ++o) acc[o] = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float v = __ldg(x + (size_t)p * C + c);
#pragma unroll
            for (int o = 0; o < 16; ++o)
                if (o < nout) acc[o] = fmaf(v, wsm[o * C + c], acc[o]);
        }
#pragma unroll
        for (int o = 0; o < 16; ++o) {
            if (o < nout) {
                float s = acc[o];
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
                if (lane == 0) out[(size_t)p * nout + o] = s + (b ? __ldg(b + o) : 0.f);
            }
        }
    }
}

}  // namespace pdm

extern "C" int pdm_linear_rows(int p, int c, int nout, const float *x, const float *w, const float *b, float *out,
                               void *stream) {
    using namespace pdm;
    if (p < 0 || c <= 0 || nout <= 0) return fail(PDM_ERR_INVALID_ARG, "linear_rows: bad size");
    if (nout > 16 || (size_t)nout * c * 4 > 96 * 1024) return fail(PDM_ERR_UNSUPPORTED, "linear_rows: nout %d (<= 16), C %d", nout, c);
    if (p == 0) return PDM_OK;
    if (!x || !w || !out) return fail(PDM_ERR_INVALID_ARG, "linear_rows: null pointer");
    const size_t smem = (size_t)nout * c * 4;
    if (int rc = ensure_dynamic_smem((const void *)linear_rows_kernel, smem)) return rc;
    const int grid = (p + 7) / 8 < kNumSMs * 8 ? (p + 7) / 8 : kNumSMs * 8;
    linear_rows_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(p, c, nout, x, w, b, out);
    count_launch();
    PDM_CHECK_LAUNCH("linear_rows");
    return PDM_OK;
}

extern "C" int pdm_point_head_forward(int p, int batch, int c_point, int c_bev, int y, int x, int n_class, int hidden_cls,
                                      int hidden_box, const float *range_min_xy, const float *voxel_xy,
                                      const float *point_coords, const float *point_features, const void *bev_split,
                                      const float *heatmap, const float *w1t, const float *b1, const float *w2_cls,
                                      const float *b2_cls, const float *w2_box, const float *b2_box, const float *mean_size,
                                      float *scores, float *boxes, float *best_score, int *best_label, float *cls_raw,
                                      float *box_raw, void *stream) {
    using namespace pdm;
    if (p < 0 || batch <= 0 || c_point <= 0 || c_bev < 0 || y <= 0 || x <= 0) return fail(PDM_ERR_INVALID_ARG, "point_head_forward: bad size");
    if (n_class < 1 || n_class > kPHMaxClass) return fail(PDM_ERR_UNSUPPORTED, "point_head_forward: %d classes (<= %d)", n_class, kPHMaxClass);
    if ((c_point & 3) || (c_bev & 7)) return fail(PDM_ERR_UNSUPPORTED, "point_head_forward: point channels %% 4 and BEV channels %% 8 must be 0");
    if (hidden_cls <= 0 || hidden_box <= 0 || (hidden_cls & 3) || (hidden_box & 3) || hidden_cls + hidden_box > kPHMaxHidden)
        return fail(PDM_ERR_UNSUPPORTED, "point_head_forward: hidden widths %d + %d (multiples of 4, sum <= %d)", hidden_cls, hidden_box, kPHMaxHidden);
    if (p == 0) return PDM_OK;
    if (!range_min_xy || !voxel_xy || !point_coords || !point_features || (c_bev > 0 && !bev_split) || !heatmap || !w1t || !b1 ||
        !w2_cls || !b2_cls || !w2_box || !b2_box || !mean_size || !scores || !boxes || !best_score || !best_label)
        return fail(PDM_ERR_INVALID_ARG, "point_head_forward: null pointer");
    PointHeadParams Q;
    Q.P = p; Q.B = batch; Q.Cp = c_point; Q.Cs = c_bev; Q.Y = y; Q.X = x; Q.ncls = n_class; Q.hc = hidden_cls; Q.hb = hidden_box;
    Q.xmin = range_min_xy[0]; Q.ymin = range_min_xy[1]; Q.vx = voxel_xy[0]; Q.vy = voxel_xy[1];
    const int cin = c_point + c_bev, H = hidden_cls + hidden_box;
    const size_t smem = sizeof(float) * ((size_t)cin * kPHRows + (size_t)H * kPHRows + (size_t)cin * kPHColBlock +
                                         (size_t)(kPHMaxClass + 8) * kPHRows + (size_t)kPHMaxClass * kPHRows);
    if (smem > 220 * 1024) return fail(PDM_ERR_UNSUPPORTED, "point_head_forward: %d input channels do not fit in shared memory", cin);
    if (int rc = ensure_dynamic_smem((const void *)point_head_kernel, smem)) return rc;
    point_head_kernel<<<(p + kPHRows - 1) / kPHRows, kPHThreads, smem, (cudaStream_t)stream>>>(
        Q, point_coords, point_features, (const __nv_bfloat16 *)bev_split, heatmap, w1t, b1, w2_cls, b2_cls, w2_box, b2_box,
        mean_size, scores, boxes, best_score, best_label, cls_raw, box_raw);
    count_launch();
    PDM_CHECK_LAUNCH("point_head_forward");
    return PDM_OK;
}
