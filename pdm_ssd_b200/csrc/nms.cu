// nms.cu -- rotated bird's-eye-view IoU and greedy NMS, batched over frames, without host round trips.
//
// Reference: pcdet/ops/iou3d_nms (iou3d_nms_kernel.cu:99-234 box_overlap / iou_bev, :295-341
// nms_kernel, iou3d_nms.cpp:137-183 nms_gpu).  There the suppression bit-mask is computed on the GPU,
// copied to the host (cudaMalloc + blocking cudaMemcpy per call) and the greedy pass runs on the CPU,
// once per frame (detector3d_template.py:199) -- a device synchronisation per frame and class.
// Here one launch builds the masks of ALL frames (upper-triangular 64x64 tiles only: the greedy pass
// never looks at a lower-indexed box again) and a second launch runs the greedy pass on the GPU, one
// warp per frame with the "removed" bit set spread over the lanes and the mask rows prefetched.
// The keep order is the reference's: ascending position in the score-sorted list.
//
// The overlap follows the reference's construction, margins included, so that keep decisions agree:
// intersection points of the 4x4 edge pairs (bounding-box rejection, straddle test, EPS = 1e-8 branch
// in the line intersection), corners of one box inside the other with a 1e-2 margin, the points
// ordered by atan2 around their centroid (bubble sort), shoelace area; IoU = s / max(sa + sb - s, EPS).
// This file is compiled with nvcc's default fused-multiply-add contraction (like the reference), not
// with --fmad=false as the bit-exact sampling kernels are.
#include "common.cuh"

namespace pdm {

constexpr float kNmsEps = 1e-8f;
constexpr float kInBoxMargin = 1e-2f;
constexpr int kTile = 64;  // boxes per mask tile = bits per mask word

struct V2 {
    float x, y;
};
__device__ __forceinline__ V2 mk(float x, float y) { V2 v; v.x = x; v.y = y; return v; }
__device__ __forceinline__ float cross_o(const V2 &a, const V2 &b, const V2 &o) {
    return (a.x - o.x) * (b.y - o.y) - (b.x - o.x) * (a.y - o.y);
}

// corners of a box [x, y, z, dx, dy, dz, heading] in the order (-,-), (+,-), (+,+), (-,+), rotated by heading
__device__ __forceinline__ void box_corners(const float *box, V2 *c) {
    const float hx = box[3] / 2, hy = box[4] / 2;
    const float cs = cos(box[6]), sn = sin(box[6]);
    const float px[4] = {box[0] - hx, box[0] + hx, box[0] + hx, box[0] - hx};
    const float py[4] = {box[1] - hy, box[1] - hy, box[1] + hy, box[1] + hy};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        c[k].x = (px[k] - box[0]) * cs + (py[k] - box[1]) * (-sn) + box[0];
        c[k].y = (px[k] - box[0]) * sn + (py[k] - box[1]) * cs + box[1];
    }
    c[4] = c[0];
}

__device__ __forceinline__ bool inside_with_margin(const float *box, const V2 &p) {
    const float cs = cos(-box[6]), sn = sin(-box[6]);  // rotate the point into the box frame
    const float rx = (p.x - box[0]) * cs + (p.y - box[1]) * (-sn);
    const float ry = (p.x - box[0]) * sn + (p.y - box[1]) * cs;
    return fabs(rx) < box[3] / 2 + kInBoxMargin && fabs(ry) < box[4] / 2 + kInBoxMargin;
}

// segment p0->p1 against q0->q1 (iou3d_nms_kernel.cu:62-91)
__device__ __forceinline__ bool seg_intersection(const V2 &p1, const V2 &p0, const V2 &q1, const V2 &q0, V2 &out) {
    const bool boxes_touch = fminf(p0.x, p1.x) <= fmaxf(q0.x, q1.x) && fminf(q0.x, q1.x) <= fmaxf(p0.x, p1.x) &&
                             fminf(p0.y, p1.y) <= fmaxf(q0.y, q1.y) && fminf(q0.y, q1.y) <= fmaxf(p0.y, p1.y);
    if (!boxes_touch) return false;
    const float s1 = cross_o(q0, p1, p0);
    const float s2 = cross_o(p1, q1, p0);
    const float s3 = cross_o(p0, q1, q0);
    const float s4 = cross_o(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return false;
    const float s5 = cross_o(q1, p1, p0);
    if (fabs(s5 - s1) > kNmsEps) {
        out.x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        out.y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        const float D = a0 * b1 - a1 * b0;
        out.x = (b0 * c1 - b1 * c0) / D;
        out.y = (a1 * c0 - a0 * c1) / D;
    }
    return true;
}

__device__ float bev_overlap(const float *a, const float *b) {
    V2 ca[5], cb[5];
    box_corners(a, ca);
    box_corners(b, cb);
    V2 pts[16];
    int cnt = 0;
    float sx = 0.f, sy = 0.f;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (cnt < 16 && seg_intersection(ca[i + 1], ca[i], cb[j + 1], cb[j], pts[cnt])) {
                sx = sx + pts[cnt].x;
                sy = sy + pts[cnt].y;
                ++cnt;
            }
    for (int k = 0; k < 4; ++k) {
        // (the reference's 16-slot array can overflow when margin corners add to 8 edge crossings,
        //  iou3d_nms_kernel.cu:150-172; such points are dropped here instead of written past the end)
        if (cnt < 16 && inside_with_margin(a, cb[k])) {
            sx = sx + cb[k].x; sy = sy + cb[k].y;
            pts[cnt++] = cb[k];
        }
        if (cnt < 16 && inside_with_margin(b, ca[k])) {
            sx = sx + ca[k].x; sy = sy + ca[k].y;
            pts[cnt++] = ca[k];
        }
    }
    const V2 ctr = mk(sx / cnt, sy / cnt);
    // ascending polar angle around the centroid: the reference bubble-sorts with atan2 evaluated inside
    // every comparison (:189-199); the angles are pure functions of the points, so computing them once
    // and carrying them through the same swaps gives the same order
    float ang[16];
    for (int k = 0; k < cnt; ++k) ang[k] = atan2(pts[k].y - ctr.y, pts[k].x - ctr.x);
    for (int j = 0; j < cnt - 1; ++j)
        for (int i = 0; i < cnt - j - 1; ++i)
            if (ang[i] > ang[i + 1]) {
                const V2 t = pts[i];
                pts[i] = pts[i + 1];
                pts[i + 1] = t;
                const float ta = ang[i];
                ang[i] = ang[i + 1];
                ang[i + 1] = ta;
            }
    float area = 0.f;
    for (int k = 0; k < cnt - 1; ++k) {
        const V2 u = mk(pts[k].x - pts[0].x, pts[k].y - pts[0].y), v = mk(pts[k + 1].x - pts[0].x, pts[k + 1].y - pts[0].y);
        area += u.x * v.y - u.y * v.x;
    }
    return fabs(area) / 2.0f;
}

__device__ __forceinline__ float bev_iou(const float *a, const float *b) {
    const float sa = a[3] * a[4], sb = b[3] * b[4];
    const float s = bev_overlap(a, b);
    return s / fmaxf(sa + sb - s, kNmsEps);
}

// axis-aligned BEV IoU, heading ignored (iou_normal, iou3d_nms_kernel.cu:341-352)
__device__ __forceinline__ float bev_iou_normal(const float *a, const float *b) {
    const float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
    const float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
    const float width = fmaxf(right - left, 0.f), height = fmaxf(bottom - top, 0.f);
    const float inter = width * height;
    return inter / fmaxf(a[3] * a[4] + b[3] * b[4] - inter, kNmsEps);
}

// overlap of box i of `a` with box i of `b` (paired_boxes_overlap_kernel / boxes_aligned_overlap_kernel,
// iou3d_nms_kernel.cu:251-277)
__global__ void __launch_bounds__(256)
paired_bev_kernel(int n, const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ ans) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ans[i] = bev_overlap(a + (size_t)i * 7, b + (size_t)i * 7);
}

// pairwise IoU (iou3d_nms_kernel.cu:279-293) or overlap area (:236-249): ans[i, j] = f(a_i, b_j)
template <bool IOU>
__global__ void __launch_bounds__(256)
pair_bev_kernel(int na, const float *__restrict__ a, int nb, const float *__restrict__ b, float *__restrict__ ans) {
    const int j = blockIdx.x * 16 + (threadIdx.x & 15), i = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (i >= na || j >= nb) return;
    const float *pa = a + (size_t)i * 7, *pb = b + (size_t)j * 7;
    ans[(size_t)i * nb + j] = IOU ? bev_iou(pa, pb) : bev_overlap(pa, pb);
}

// mask[f, i, w] bit t  <=>  IoU(box i, box 64w + t) > thresh, for 64w + t > i.  Upper-triangular 64x64 tiles.
// NORMAL: axis-aligned IoU (nms_normal_kernel, iou3d_nms_kernel.cu:355-398) instead of the rotated one.
//
// The rotated overlap costs a few thousand instructions and only a few percent of the pairs of a tile can overlap
// at all, so a thread-per-row loop over 64 columns leaves most lanes of a warp idle behind the one lane that found
// a candidate (round 1: 0.62 ms for 16 x 1024 boxes, the third largest kernel of the detector).  Here a tile is
// processed in two dense phases by 128 threads:
//   1. every (row, 32 columns) pair of the tile gets a candidate bit-mask from two exact rejection tests -- disjoint
//      circumscribed circles, disjoint axis-aligned bounding boxes, both grown by far more than the reference's 1e-2
//      corner margin and any rounding: such boxes have no edge crossing and no corner inside the other, the
//      reference's overlap is exactly 0 and `0 > thresh` is false for thresh >= 0;
//   2. the candidates are numbered by a prefix sum over the 128 masks and handed out round-robin, one rotated IoU
//      per thread per step, results OR-ed into shared memory: all lanes stay busy until the list is empty.
constexpr int kMaskThreads = 128;

template <bool NORMAL>
__global__ void __launch_bounds__(kMaskThreads)
nms_mask_kernel(int k, int words, const int *__restrict__ counts, float thresh, const float *__restrict__ boxes,
                unsigned long long *__restrict__ mask) {
    const int f = blockIdx.z, rt = blockIdx.y, ct = blockIdx.x;
    const int n = counts ? min(__ldg(counts + f), k) : k;
    const int tid = threadIdx.x;
    __shared__ float rbox[kTile * 7], cbox[kTile * 7];
    __shared__ float rext[kTile * 3], cext[kTile * 3];        // per box: radius, half extent x, half extent y of the AABB
    __shared__ unsigned cand[kMaskThreads], res[kMaskThreads]; // entry e = half * 64 + row: columns half*32 .. half*32+31
    __shared__ int pref[kMaskThreads + 1];
    const float *fb = boxes + (size_t)f * k * 7;
    const bool live = ct >= rt && rt * kTile < n && ct * kTile < n;
    if (!live) {                                               // the sweep ORs whole rows: lower-triangular words must be 0
        if (tid < kTile && rt * kTile + tid < k) mask[((size_t)f * k + rt * kTile + tid) * words + ct] = 0ull;
        return;
    }
    const int rn = min(kTile, n - rt * kTile), cn = min(kTile, n - ct * kTile);
    for (int t = tid; t < rn * 7; t += kMaskThreads) rbox[t] = __ldg(fb + (size_t)rt * kTile * 7 + t);
    for (int t = tid; t < cn * 7; t += kMaskThreads) cbox[t] = __ldg(fb + (size_t)ct * kTile * 7 + t);
    __syncthreads();
    {
        const bool isrow = tid < kTile;
        const int j = tid & (kTile - 1);
        if (j < (isrow ? rn : cn)) {
            const float *bx = (isrow ? rbox : cbox) + j * 7;
            float *ex = (isrow ? rext : cext) + j * 3;
            const float cs = fabsf(cosf(bx[6])), sn = fabsf(sinf(bx[6]));
            ex[0] = 0.5f * sqrtf(bx[3] * bx[3] + bx[4] * bx[4]) * 1.001f + 0.1f;
            ex[1] = 0.5f * (cs * fabsf(bx[3]) + sn * fabsf(bx[4])) * 1.001f + 0.1f;
            ex[2] = 0.5f * (sn * fabsf(bx[3]) + cs * fabsf(bx[4])) * 1.001f + 0.1f;
        }
    }
    __syncthreads();
    // ---- phase 1: candidate masks -----------------------------------------------------------------------------------
    {
        const int row = tid & (kTile - 1), half = tid >> 6;
        unsigned bits = 0u;
        if (row < rn) {
            const float *me = rbox + row * 7, *mx = rext + row * 3;
            for (int q = 0; q < 32; ++q) {
                const int t = half * 32 + q;
                if (t >= cn || (ct == rt && t <= row)) continue;
                const float *ob = cbox + t * 7, *ox = cext + t * 3;
                const float ddx = ob[0] - me[0], ddy = ob[1] - me[1];
                const float rr = mx[0] + ox[0];
                const bool far = thresh >= 0.f && !NORMAL &&
                                 (ddx * ddx + ddy * ddy > rr * rr || fabsf(ddx) > mx[1] + ox[1] || fabsf(ddy) > mx[2] + ox[2]);
                if (!far) bits |= 1u << q;
            }
        }
        cand[tid] = bits;
        res[tid] = 0u;
    }
    __syncthreads();
    // ---- prefix sum of the candidate counts (128 entries: one warp scan per 32, then 4 partials) -------------------------
    {
        const int lane = tid & 31, w = tid >> 5;
        int v = __popc(cand[tid]), incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        __shared__ int wsum[4];
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        int base = 0;
        for (int q = 0; q < w; ++q) base += wsum[q];
        pref[tid + 1] = base + incl;
        if (tid == 0) pref[0] = 0;
    }
    __syncthreads();
    // ---- phase 2: one IoU per thread per step over the dense candidate list ------------------------------------------------
    const int total = pref[kMaskThreads];
    for (int c = tid; c < total; c += kMaskThreads) {
        int lo = 0, hi = kMaskThreads;                         // largest e with pref[e] <= c
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (pref[mid] <= c) lo = mid; else hi = mid;
        }
        const int e = lo, q = __fns(cand[e], 0, c - pref[e] + 1);
        const int row = e & (kTile - 1), t = (e >> 6) * 32 + q;
        const float iou = NORMAL ? bev_iou_normal(rbox + row * 7, cbox + t * 7) : bev_iou(rbox + row * 7, cbox + t * 7);
        if (iou > thresh) atomicOr(&res[e], 1u << q);
    }
    __syncthreads();
    if (tid < kTile && rt * kTile + tid < k)
        mask[((size_t)f * k + rt * kTile + tid) * words + ct] =
            tid < rn ? ((unsigned long long)res[tid] | ((unsigned long long)res[kTile + tid] << 32)) : 0ull;
}

// Greedy pass, one warp per frame (iou3d_nms.cpp:159-176).  Lane l owns words l, l+32, ... of the "removed" set
// (WPL words per lane: 2 covers 4096 boxes, 8 covers 16 384); mask rows are prefetched a few iterations ahead.
template <int WPL>
__global__ void __launch_bounds__(32)
nms_sweep_kernel(int k, int words, const int *__restrict__ counts, const unsigned long long *__restrict__ mask,
                 int *__restrict__ keep, int *__restrict__ num_keep) {
    const int f = blockIdx.x, lane = threadIdx.x;
    const int n = counts ? min(__ldg(counts + f), k) : k;
    const unsigned long long *fm = mask + (size_t)f * k * words;
    int *fk = keep + (size_t)f * k;
    unsigned long long rem[WPL];
#pragma unroll
    for (int w = 0; w < WPL; ++w) rem[w] = 0ull;
    constexpr int PF = WPL <= 2 ? 4 : 2;
    unsigned long long r[PF][WPL];
#pragma unroll
    for (int q = 0; q < PF; ++q)
#pragma unroll
        for (int w = 0; w < WPL; ++w)
            r[q][w] = (q < n && lane + 32 * w < words) ? __ldg(fm + (size_t)q * words + lane + 32 * w) : 0ull;
    int kept = 0;
    for (int base = 0; base < n; base += PF) {
#pragma unroll
        for (int q = 0; q < PF; ++q) {
            const int i = base + q;
            unsigned long long m[WPL];
            const int nxt = i + PF;  // refill this slot for iteration i + PF
#pragma unroll
            for (int w = 0; w < WPL; ++w) {
                m[w] = r[q][w];
                r[q][w] = (nxt < n && lane + 32 * w < words) ? __ldg(fm + (size_t)nxt * words + lane + 32 * w) : 0ull;
            }
            if (i < n) {
                const int wi = i >> 6;
                unsigned long long mine = rem[0];
#pragma unroll
                for (int w = 1; w < WPL; ++w)
                    if ((wi >> 5) == w) mine = rem[w];
                const unsigned long long word = __shfl_sync(0xffffffffu, mine, wi & 31);
                if (!((word >> (i & 63)) & 1ull)) {   // warp-uniform
                    if (lane == 0) fk[kept] = i;
                    ++kept;
#pragma unroll
                    for (int w = 0; w < WPL; ++w) rem[w] |= m[w];
                }
            }
        }
    }
    for (int t = kept + lane; t < k; t += 32) fk[t] = -1;
    if (lane == 0) num_keep[f] = kept;
}

}  // namespace pdm

extern "C" {

static int pair_bev(bool iou, int na, const float *boxes_a, int nb, const float *boxes_b, float *ans, void *stream) {
    using namespace pdm;
    if (na < 0 || nb < 0) return fail(PDM_ERR_INVALID_ARG, "boxes_bev: negative size");
    if (na == 0 || nb == 0) return PDM_OK;
    if (!boxes_a || !boxes_b || !ans) return fail(PDM_ERR_INVALID_ARG, "boxes_bev: null pointer");
    dim3 grid((nb + 15) / 16, (na + 15) / 16);
    if (grid.y > 65535) return fail(PDM_ERR_UNSUPPORTED, "boxes_bev: too many boxes");
    if (iou) pair_bev_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(na, boxes_a, nb, boxes_b, ans);
    else pair_bev_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(na, boxes_a, nb, boxes_b, ans);
    count_launch();
    PDM_CHECK_LAUNCH("boxes_bev");
    return PDM_OK;
}

int pdm_boxes_iou_bev(int na, const float *boxes_a, int nb, const float *boxes_b, float *ans_iou, void *stream) {
    return pair_bev(true, na, boxes_a, nb, boxes_b, ans_iou, stream);
}

int pdm_boxes_overlap_bev(int na, const float *boxes_a, int nb, const float *boxes_b, float *ans_overlap, void *stream) {
    return pair_bev(false, na, boxes_a, nb, boxes_b, ans_overlap, stream);
}

int pdm_boxes_overlap_bev_paired(int n, const float *boxes_a, const float *boxes_b, float *ans_overlap, void *stream) {
    using namespace pdm;
    if (n < 0) return fail(PDM_ERR_INVALID_ARG, "boxes_overlap_bev_paired: negative size");
    if (n == 0) return PDM_OK;
    if (!boxes_a || !boxes_b || !ans_overlap) return fail(PDM_ERR_INVALID_ARG, "boxes_overlap_bev_paired: null pointer");
    paired_bev_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, boxes_a, boxes_b, ans_overlap);
    count_launch();
    PDM_CHECK_LAUNCH("boxes_overlap_bev_paired");
    return PDM_OK;
}

static int nms_batched(bool normal, int frames, int k, const float *boxes, const int *counts, float thresh, int *keep,
                       int *num_keep, void *stream) {
    using namespace pdm;
    if (frames < 0 || k < 0) return fail(PDM_ERR_INVALID_ARG, "nms_bev_batched: negative size");
    if (frames == 0) return PDM_OK;
    if (!keep || !num_keep || (k > 0 && !boxes)) return fail(PDM_ERR_INVALID_ARG, "nms_bev_batched: null pointer");
    if (k > 16384) return fail(PDM_ERR_UNSUPPORTED, "nms_bev_batched: at most 16384 boxes per frame (got %d)", k);
    if (frames > 65535) return fail(PDM_ERR_UNSUPPORTED, "nms_bev_batched: too many frames");
    cudaStream_t st = (cudaStream_t)stream;
    const int words = (k + kTile - 1) / kTile;
    unsigned long long *mask = nullptr;
    if (k > 0) {
        mask = static_cast<unsigned long long *>(stream_scratch(st, (size_t)frames * k * words * sizeof(unsigned long long)));
        if (!mask) return PDM_ERR_INVALID_ARG;
        dim3 grid(words, words, frames);
        if (normal) nms_mask_kernel<true><<<grid, kMaskThreads, 0, st>>>(k, words, counts, thresh, boxes, mask);
        else nms_mask_kernel<false><<<grid, kMaskThreads, 0, st>>>(k, words, counts, thresh, boxes, mask);
        count_launch();
        PDM_CHECK_LAUNCH("nms_bev_batched(mask)");
    }
    if (words <= 64) nms_sweep_kernel<2><<<frames, 32, 0, st>>>(k, words, counts, mask, keep, num_keep);
    else nms_sweep_kernel<8><<<frames, 32, 0, st>>>(k, words, counts, mask, keep, num_keep);
    count_launch();
    PDM_CHECK_LAUNCH("nms_bev_batched(sweep)");
    return PDM_OK;
}

int pdm_nms_bev_batched(int frames, int k, const float *boxes, const int *counts, float thresh, int *keep,
                        int *num_keep, void *stream) {
    return nms_batched(false, frames, k, boxes, counts, thresh, keep, num_keep, stream);
}

int pdm_nms_normal_batched(int frames, int k, const float *boxes, const int *counts, float thresh, int *keep,
                           int *num_keep, void *stream) {
    return nms_batched(true, frames, k, boxes, counts, thresh, keep, num_keep, stream);
}

}  // extern "C"
