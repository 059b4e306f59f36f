// fps_common.cuh -- helpers shared by the farthest-point-sampling kernels (fps.cu, fps_l2.cu).
#pragma once
#include "common.cuh"

namespace pdm {

constexpr unsigned kFull = 0xffffffffu;
constexpr unsigned kPadKey = 0xffffffffu;  // tiekey of a padding slot: loses every tie

// Reference tie-break (sampling_gpu.cu:93-98,143-144 unrolled, see fps.cu header): the winner among
// equal maxima is the point with the smallest tiekey.
__device__ __forceinline__ unsigned fps_tiekey(unsigned k, int p, unsigned bsmask) {
    return p == 0 ? k : (__brev(k & bsmask) | (k >> p));
}
__device__ __forceinline__ unsigned fps_tiekey_inv(unsigned tk, int p, unsigned bsmask) {
    if (p == 0) return tk;
    const unsigned lowmask = (1u << (32 - p)) - 1u;
    return ((tk & lowmask) << p) | (__brev(tk) & bsmask);
}

__device__ __forceinline__ unsigned expand10(unsigned v) {  // 10 bits -> every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
// order-preserving float <-> uint map (for redux min/max over arbitrary-sign floats)
__device__ __forceinline__ unsigned f2ord(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ bool multi_bit(unsigned ball) { return (ball & (ball - 1u)) != 0u; }

// Spatial sort code of a point.  Buckets are runs of 32 consecutive points in code order, and the
// pruning is as good as their boxes are tight.  LiDAR frames are nearly planar (KITTI: 70 x 80 x 4 m):
// there a 2-D Hilbert curve over the two long axes gives compact, jump-free runs (3.7 surviving
// buckets per sample against 5.2 for a 3-D Morton curve, measured on KITTI-shaped frames).
// Volumetric clouds (shortest extent > 1/8 of the longest) keep the 3-D Morton key.
struct FpsCurve {
    float lo[3];
    float inv, inv16;
    int ax_u, ax_v;
    bool planar;
    __device__ void init(const float *lo_, const float *hi_) {
        const float e0 = hi_[0] - lo_[0], e1 = hi_[1] - lo_[1], e2 = hi_[2] - lo_[2];
        const float ext = fmaxf(fmaxf(e0, e1), e2);
        const float emin = fminf(fminf(e0, e1), e2);
        planar = emin * 8.0f <= ext;
        const int thin = (e2 <= e0 && e2 <= e1) ? 2 : ((e1 <= e0) ? 1 : 0);   // axis left out when planar
        ax_u = thin == 0 ? 1 : 0;
        ax_v = thin == 2 ? 1 : 2;
        inv = (ext > 0.f && ext < INFINITY) ? 1023.0f / ext : 0.f;
        inv16 = (ext > 0.f && ext < INFINITY) ? 65535.0f / ext : 0.f;
        lo[0] = lo_[0]; lo[1] = lo_[1]; lo[2] = lo_[2];
    }
    __device__ unsigned code(const float *c) const {
        if (planar) {
            int iu = (int)((c[ax_u] - lo[ax_u]) * inv16), iv = (int)((c[ax_v] - lo[ax_v]) * inv16);  // NaN -> 0
            unsigned x = (unsigned)max(0, min(65535, iu)), y = (unsigned)max(0, min(65535, iv));
            unsigned d = 0u;
#pragma unroll
            for (int sft = 15; sft >= 0; --sft) {      // 16-bit 2-D Hilbert index (xy -> d)
                const unsigned rx = (x >> sft) & 1u, ry = (y >> sft) & 1u;
                d = (d << 2) | ((3u * rx) ^ ry);
                if (ry == 0u) {
                    if (rx == 1u) { x = ~x; y = ~y; }  // reflect (only the low `sft` bits matter)
                    const unsigned tswap = x; x = y; y = tswap;
                }
            }
            return d;
        }
        unsigned q[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            int qi = (int)((c[a] - lo[a]) * inv);  // NaN -> 0
            q[a] = (unsigned)max(0, min(1023, qi));
        }
        return (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    }
    // 18-bit variant (9 bits per axis on the Hilbert curve, 6 per axis on the Morton curve): with a
    // 14-bit point index it makes a 32-bit sort key.  Cells of ext/512 (16 cm on a KITTI frame) are far
    // below the size of a 32-point bucket, so the bucket boxes are as tight as with the full code.
    __device__ unsigned code18(const float *c) const {
        if (planar) {
            int iu = (int)((c[ax_u] - lo[ax_u]) * inv16), iv = (int)((c[ax_v] - lo[ax_v]) * inv16);  // NaN -> 0
            unsigned x = (unsigned)max(0, min(65535, iu)) >> 7, y = (unsigned)max(0, min(65535, iv)) >> 7;
            unsigned d = 0u;
#pragma unroll
            for (int sft = 8; sft >= 0; --sft) {
                const unsigned rx = (x >> sft) & 1u, ry = (y >> sft) & 1u;
                d = (d << 2) | ((3u * rx) ^ ry);
                if (ry == 0u) {
                    if (rx == 1u) { x = ~x; y = ~y; }
                    const unsigned tswap = x; x = y; y = tswap;
                }
            }
            return d;
        }
        unsigned q[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            int qi = (int)((c[a] - lo[a]) * inv);  // NaN -> 0
            q[a] = (unsigned)max(0, min(1023, qi)) >> 4;
        }
        return (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    }
};

// Bitonic sort of E * T 32-bit keys held E per thread in registers (element index e = tid * E + r), for a
// CTA of T threads.  The compare-exchange distance j is handled where the partner lives:
//   j < E          inside the thread (register pairs),
//   E <= j < 32 E  in another lane of the warp (shfl.xor),
//   j >= 32 E      in another warp: one transposed round trip through shared memory (conflict-free).
// xch: E * T words of shared memory (only touched by the cross-warp stages).  All T threads must call.
template <int E, int T>
__device__ __forceinline__ void fps_sort_keys(unsigned (&v)[E], unsigned *xch, int tid) {
    static_assert((E & (E - 1)) == 0 && E >= 2 && E <= 32, "keys per thread: a power of two");
    // stages whose direction depends on the register index only
#pragma unroll
    for (int k = 2; k < E; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int r = 0; r < E; ++r) {
                if ((r & j) == 0) {
                    const unsigned a = v[r], b = v[r | j];
                    const unsigned mn = min(a, b), mx = max(a, b);
                    const bool asc = (r & k) == 0;
                    v[r] = asc ? mn : mx;
                    v[r | j] = asc ? mx : mn;
                }
            }
        }
    }
    for (int k = E; k <= E * T; k <<= 1) {
        const bool asc = ((tid * E) & k) == 0;     // the same for all registers of a thread
        for (int j = k >> 1; j >= 32 * E; j >>= 1) {            // partner in another warp
            const int pt = tid ^ (j / E);
            const bool keep_min = asc == ((tid & (j / E)) == 0);
#pragma unroll
            for (int r = 0; r < E; ++r) xch[r * T + tid] = v[r];
            __syncthreads();
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const unsigned p = xch[r * T + pt];
                v[r] = keep_min ? min(v[r], p) : max(v[r], p);
            }
            __syncthreads();
        }
#pragma unroll
        for (int jl = 16; jl >= 1; jl >>= 1) {                  // partner in another lane: j = E * jl
            if (E * jl < k) {
                const bool keep_min = asc == ((tid & jl) == 0);
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    const unsigned p = __shfl_xor_sync(kFull, v[r], jl);
                    v[r] = keep_min ? min(v[r], p) : max(v[r], p);
                }
            }
        }
#pragma unroll
        for (int J = E >> 1; J >= 1; J >>= 1) {                 // partner in another register
#pragma unroll
            for (int r = 0; r < E; ++r) {
                if ((r & J) == 0) {
                    const unsigned a = v[r], b = v[r | J];
                    const unsigned mn = min(a, b), mx = max(a, b);
                    v[r] = asc ? mn : mx;
                    v[r | J] = asc ? mx : mn;
                }
            }
        }
    }
}

// t[jj] for a warp-uniform runtime jj, with t[] in registers.  A 32-way switch around the
// whole bucket update thrashed the instruction cache (20 KB loop), a flat 32-way select tree
// costs ~95 half-rate ALU instructions.  So: groups of 8 registers; a (uniform) 2-level branch
// picks the group, a 3-level select tree (7 FSEL) picks the register inside it.
template <int W>
__device__ __forceinline__ float reg_select_tree(const float *t, int jj) {
    float a[W];
#pragma unroll
    for (int i = 0; i < W; ++i) a[i] = t[i];
#pragma unroll
    for (int bit = 0; (1 << bit) < W; ++bit) {
        const bool odd = (jj >> bit) & 1;
#pragma unroll
        for (int i = 0; i < (W >> (bit + 1)); ++i) a[i] = odd ? a[2 * i + 1] : a[2 * i];
    }
    return a[0];
}
template <int BPW>
__device__ __forceinline__ float reg_select(const float (&t)[BPW], int jj) {
    if constexpr (BPW <= 8) {
        return reg_select_tree<BPW>(t, jj);
    } else {
        switch (jj >> 3) {
            case 0: return reg_select_tree<8>(&t[0], jj & 7);
            case 1: return reg_select_tree<8>(&t[8], jj & 7);
            case 2: if constexpr (BPW > 16) return reg_select_tree<8>(&t[16], jj & 7);
            default: if constexpr (BPW > 24) return reg_select_tree<8>(&t[24], jj & 7);
        }
        return 0.f;
    }
}
template <int BPW>
__device__ __forceinline__ void reg_store(float (&t)[BPW], int jj, float v) {
    if constexpr (BPW <= 8) {
#pragma unroll
        for (int q = 0; q < BPW; ++q) t[q] = (q == jj) ? v : t[q];
    } else {
        const int r = jj & 7;
        switch (jj >> 3) {
#define PDM_GRP(G)                                                         \
    case G:                                                                \
        if constexpr (BPW > 8 * G) {                                       \
            _Pragma("unroll") for (int q = 0; q < 8; ++q) t[8 * G + q] = (q == r) ? v : t[8 * G + q]; \
        }                                                                  \
        break;
            PDM_GRP(0) PDM_GRP(1) PDM_GRP(2) PDM_GRP(3)
#undef PDM_GRP
            default: break;
        }
    }
}

// fps_l2.cu: throughput-oriented variant (coordinates stay in L2, several frames per SM).
// Returns PDM_ERR_UNSUPPORTED (without recording an error) when the shape is outside its range.
int fps_l2_launch(int b, int n, int m, int p, const float *xyz, float *temp, int *idx, int *stats, cudaStream_t st);
bool fps_l2_supports(int n);

// fps_cluster.cu: one thread-block cluster per frame for 16384 < n <= 196608 (same convention).
int fps_cluster_launch(int b, int n, int m, int p, const float *xyz, float *temp, int *idx, cudaStream_t st);
bool fps_cluster_supports(int n);


// fps_cluster_bucket.cu: the same cluster layout with exact bucket pruning and multi-sample rounds (n <= 196608).
int fps_cluster_bucket_launch(int b, int n, int m, int p, const float *xyz, float *temp, int *idx, int *stats, cudaStream_t st,
                              bool small_ok = false);
bool fps_cluster_bucket_supports(int n);

}  // namespace pdm
