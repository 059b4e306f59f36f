// fps.cu -- farthest point sampling, bit-exact with sampling_gpu.cu:100-216.
//
// What has to be reproduced
// -------------------------
// round j = 1..m-1:  temp[k] = min(temp[k], d(k, old))  for all k;  old = argmax_k temp[k].
// d is rounded as sqdist_ref (common.cuh).  The reference's argmax is a tournament: thread
// tid scans k = tid, tid+bs, ... keeping the FIRST maximum (strict >), then a shared-memory
// tree in which the LEFT slot survives ties.  Unrolling that tree, the winner among equal
// maxima is the point with the smallest
//        tiekey(k) = bitreverse_p(k mod bs) * 2^(32-p)  +  (k div bs),     bs = 2^p
// where bs = the block size the reference would have launched (cuda_utils.h:10-14).
// So   argmax  ==  max over the 64-bit key  (float_bits(temp[k]) << 32) | ~tiekey(k)
// (temp >= 0, so its bit pattern orders like the float), which is decomposition-free:
// any parallel reduction order gives the reference's answer.
//
// Kernels
// -------
// fps_bucket_kernel<NW,BPW>  (n <= NW*BPW*32 <= 16384): one CTA per frame, everything
//   on-chip.  Points are Morton-sorted once (in-CTA bitonic sort) and cut into buckets of
//   32 consecutive points = one point per lane; bucket b belongs to warp b mod NW
//   (interleaved so that a spatial neighbourhood spreads over all warps).  Each bucket keeps
//   its bounding box and its current maximum of temp.  In a round a bucket can only change
//   if  lowerbound(d(box, sample)) < bucket max ;  the lower bound is sqdist_ref of the
//   per-axis gaps, which is <= d(k, sample) for every k in the box because fp32 rounding is
//   monotone -- so skipping the other buckets is EXACT, not approximate.  After the first
//   few hundred rounds only a handful of buckets survive the test, and a round costs
//   O(buckets/warp) bound checks + a couple of 32-point updates + one block-wide argmax
//   instead of a sweep over all n points.  Running minima live in registers, coordinates in
//   shared memory (SoA, conflict-free), the argmax uses redux.sync and one barrier/round.
// fps_generic_kernel: any n; same key trick, temp in global memory.
#include "common.cuh"

namespace pdm {

constexpr unsigned kFull = 0xffffffffu;
constexpr unsigned kPadKey = 0xffffffffu;  // tiekey of a padding slot: loses every tie

__device__ __forceinline__ unsigned fps_tiekey(unsigned k, int p, unsigned bsmask) {
    return p == 0 ? k : (__brev(k & bsmask) | (k >> p));
}
__device__ __forceinline__ unsigned fps_tiekey_inv(unsigned tk, int p, unsigned bsmask) {
    if (p == 0) return tk;
    const unsigned lowmask = (1u << (32 - p)) - 1u;
    return ((tk & lowmask) << p) | (__brev(tk) & bsmask);
}

// ---------------------------------------------------------------------------------------
// generic kernel
// ---------------------------------------------------------------------------------------
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
fps_generic_kernel(int n, int m, int p, const float *__restrict__ xyz, float *__restrict__ temp,
                   int *__restrict__ idxs) {
    constexpr int NWARP = THREADS / 32;
    __shared__ unsigned long long wbest[2][NWARP];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned bsmask = (1u << p) - 1u;
    const float *dataset = xyz + (size_t)blockIdx.x * n * 3;
    float *tmp = temp + (size_t)blockIdx.x * n;
    int *out = idxs + (size_t)blockIdx.x * m;
    int old = 0;
    if (tid == 0) out[0] = 0;
    for (int j = 1; j < m; ++j) {
        const float x1 = __ldg(dataset + old * 3 + 0);
        const float y1 = __ldg(dataset + old * 3 + 1);
        const float z1 = __ldg(dataset + old * 3 + 2);
        unsigned long long best = 0ull;
        for (int k = tid; k < n; k += THREADS) {
            const float d = sqdist_ref(__fsub_rn(__ldg(dataset + k * 3 + 0), x1),
                                       __fsub_rn(__ldg(dataset + k * 3 + 1), y1),
                                       __fsub_rn(__ldg(dataset + k * 3 + 2), z1));
            const float d2 = fminf(d, tmp[k]);
            tmp[k] = d2;
            const unsigned long long key =
                ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)(~fps_tiekey(k, p, bsmask));
            best = key > best ? key : best;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(kFull, best, o);
            best = other > best ? other : best;
        }
        if (lane == 0) wbest[j & 1][warp] = best;
        __syncthreads();
        best = lane < NWARP ? wbest[j & 1][lane] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(kFull, best, o);
            best = other > best ? other : best;
        }
        old = (int)fps_tiekey_inv(~(unsigned)best, p, bsmask);
        if (tid == 0) out[j] = old;
    }
}

// ---------------------------------------------------------------------------------------
// bucketed on-chip kernel
// ---------------------------------------------------------------------------------------
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

template <int NW, int BPW>
struct FpsSmem {
    static constexpr int CAP = NW * BPW * 32;
    // [0, 12*CAP): sx, sy, sz (aliased by the 8*CAP-byte sort keys during set-up)
    // then pub[2][NW] uint4, then 8 floats of frame box scratch per warp
    static constexpr size_t kPubOff = (size_t)12 * CAP;
    static constexpr size_t kBoxOff = kPubOff + sizeof(uint4) * 2 * NW;
    static constexpr size_t kBytes = kBoxOff + sizeof(float) * 6 * NW;
};

template <int NW, int BPW>
__global__ void __launch_bounds__(NW * 32, 1)
fps_bucket_kernel(int n, int m, int p, const float *__restrict__ xyz, float *__restrict__ temp,
                  int *__restrict__ idxs) {
    using L = FpsSmem<NW, BPW>;
    constexpr int CAP = L::CAP;
    constexpr int T = NW * 32;
    static_assert(BPW <= 32, "one lane per owned bucket");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sx = reinterpret_cast<float *>(smem_raw);
    float *sy = sx + CAP;
    float *sz = sy + CAP;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);
    uint4 *pub = reinterpret_cast<uint4 *>(smem_raw + L::kPubOff);
    float *box = reinterpret_cast<float *>(smem_raw + L::kBoxOff);

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const unsigned bsmask = (1u << p) - 1u;
    const float *dataset = xyz + (size_t)blockIdx.x * n * 3;
    float *tmp = temp + (size_t)blockIdx.x * n;
    int *out = idxs + (size_t)blockIdx.x * m;

    if (tid == 0) out[0] = 0;
    if (m <= 1) return;

    // ---- 1. frame bounding box ----------------------------------------------------------
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = tid; k < n; k += T) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = __ldg(dataset + k * 3 + a);
            lo[a] = fminf(lo[a], v);
            hi[a] = fmaxf(hi[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = ord2f(__reduce_min_sync(kFull, f2ord(lo[a])));
        hi[a] = ord2f(__reduce_max_sync(kFull, f2ord(hi[a])));
    }
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            box[w * 6 + a] = lo[a];
            box[w * 6 + 3 + a] = hi[a];
        }
    }
    __syncthreads();
    {
        float l2[3], h2[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            l2[a] = lane < NW ? box[lane * 6 + a] : INFINITY;
            h2[a] = lane < NW ? box[lane * 6 + 3 + a] : -INFINITY;
            lo[a] = ord2f(__reduce_min_sync(kFull, f2ord(l2[a])));
            hi[a] = ord2f(__reduce_max_sync(kFull, f2ord(h2[a])));
        }
    }
    const float ext = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
    const float inv = (ext > 0.f && ext < INFINITY) ? 1023.0f / ext : 0.f;

    // ---- 2. Morton keys -> shared, bitonic sort -------------------------------------------
    for (int k = tid; k < CAP; k += T) {
        unsigned long long key = ~0ull;
        if (k < n) {
            unsigned q[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float f = (__ldg(dataset + k * 3 + a) - lo[a]) * inv;
                int qi = (int)f;  // NaN -> 0
                qi = max(0, min(1023, qi));
                q[a] = (unsigned)qi;
            }
            const unsigned mort = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
            key = ((unsigned long long)mort << 32) | (unsigned)k;
        }
        keys[k] = key;
    }
    __syncthreads();
    for (int kk = 2; kk <= CAP; kk <<= 1) {
        for (int jj = kk >> 1; jj > 0; jj >>= 1) {
            for (int i = tid; i < CAP / 2; i += T) {
                const int l = ((i & ~(jj - 1)) << 1) | (i & (jj - 1));
                const int r = l | jj;
                const unsigned long long a = keys[l], b = keys[r];
                const bool up = (l & kk) == 0;
                if ((a > b) == up) {
                    keys[l] = b;
                    keys[r] = a;
                }
            }
            __syncthreads();
        }
    }

    // ---- 3. distribute: lane owns slot `lane` of buckets  b = j*NW + w --------------------
    unsigned tk[BPW];   // tiekey of my point in owned bucket j (kPadKey for padding)
    float t[BPW];       // its running min distance
    unsigned kk_[BPW];
#pragma unroll
    for (int j = 0; j < BPW; ++j) kk_[j] = (unsigned)keys[((j * NW + w) << 5) + lane];  // low 32 bits = k
#pragma unroll
    for (int j = 0; j < BPW; ++j) {
        // padding keys are ~0 -> low word 0xffffffff
        const bool pad = keys[((j * NW + w) << 5) + lane] == ~0ull;
        tk[j] = pad ? kPadKey : fps_tiekey(kk_[j], p, bsmask);
    }
    __syncthreads();  // keys are dead from here on; the region becomes sx/sy/sz

    // per-bucket state, held by lane j of the owning warp
    float blox = INFINITY, bloy = INFINITY, bloz = INFINITY;
    float bhix = -INFINITY, bhiy = -INFINITY, bhiz = -INFINITY;
    unsigned bmax = 0u, btk = kPadKey, bwl = 0u;
#pragma unroll
    for (int j = 0; j < BPW; ++j) {
        const int pos = ((j * NW + w) << 5) + lane;
        const bool pad = tk[j] == kPadKey;
        float x = 0.f, y = 0.f, z = 0.f;
        t[j] = 0.f;
        if (!pad) {
            x = __ldg(dataset + kk_[j] * 3 + 0);
            y = __ldg(dataset + kk_[j] * 3 + 1);
            z = __ldg(dataset + kk_[j] * 3 + 2);
            t[j] = tmp[kk_[j]];
        }
        sx[pos] = x;
        sy[pos] = y;
        sz[pos] = z;
        // padding and NaN coordinates stay out of the box (a NaN point never changes anyway:
        // its distance is NaN and fminf keeps the old minimum, exactly as in the reference)
        const bool ox = pad || x != x, oy = pad || y != y, oz = pad || z != z;
        const unsigned lx = __reduce_min_sync(kFull, ox ? 0xffffffffu : f2ord(x));
        const unsigned ly = __reduce_min_sync(kFull, oy ? 0xffffffffu : f2ord(y));
        const unsigned lz = __reduce_min_sync(kFull, oz ? 0xffffffffu : f2ord(z));
        const unsigned hx = __reduce_max_sync(kFull, ox ? 0u : f2ord(x));
        const unsigned hy = __reduce_max_sync(kFull, oy ? 0u : f2ord(y));
        const unsigned hz = __reduce_max_sync(kFull, oz ? 0u : f2ord(z));
        const unsigned tb = __float_as_uint(t[j]);
        const unsigned mx = __reduce_max_sync(kFull, tb);
        const unsigned cand = (tb == mx) ? tk[j] : kPadKey;
        const unsigned tkm = __reduce_min_sync(kFull, cand);
        const unsigned wl = __ffs(__ballot_sync(kFull, cand == tkm)) - 1;
        if (lane == j) {
            blox = ord2f(lx); bloy = ord2f(ly); bloz = ord2f(lz);
            bhix = ord2f(hx); bhiy = ord2f(hy); bhiz = ord2f(hz);
            bmax = mx; btk = tkm; bwl = wl;
        }
    }
    __syncthreads();

    // ---- 4. rounds -------------------------------------------------------------------------
    float cx = __ldg(dataset + 0), cy = __ldg(dataset + 1), cz = __ldg(dataset + 2);  // sample 0 = point 0
    unsigned wm = 0u, wtk = kPadKey, wpos = 0u;  // cached best of this warp
    bool dirty = true;
    for (int j = 1; j < m; ++j) {
        // A. which of my buckets can change?
        bool act = false;
        if (lane < BPW) {
            const float gx = fmaxf(fmaxf(__fsub_rn(blox, cx), __fsub_rn(cx, bhix)), 0.f);
            const float gy = fmaxf(fmaxf(__fsub_rn(bloy, cy), __fsub_rn(cy, bhiy)), 0.f);
            const float gz = fmaxf(fmaxf(__fsub_rn(bloz, cz), __fsub_rn(cz, bhiz)), 0.f);
            act = sqdist_ref(gx, gy, gz) < __uint_as_float(bmax);
        }
        const unsigned mask = __ballot_sync(kFull, act);
        // B. update the surviving buckets (warp-uniform branches, static register indices)
        if (mask) {
            dirty = true;
#pragma unroll
            for (int jj = 0; jj < BPW; ++jj) {
                if (mask & (1u << jj)) {
                    const int pos = ((jj * NW + w) << 5) + lane;
                    const float d = sqdist_ref(__fsub_rn(sx[pos], cx), __fsub_rn(sy[pos], cy),
                                               __fsub_rn(sz[pos], cz));
                    const float nt = fminf(d, t[jj]);
                    t[jj] = nt;
                    const unsigned tb = __float_as_uint(nt);
                    const unsigned mx = __reduce_max_sync(kFull, tb);
                    const unsigned cand = (tb == mx) ? tk[jj] : kPadKey;
                    const unsigned tkm = __reduce_min_sync(kFull, cand);
                    const unsigned wl = __ffs(__ballot_sync(kFull, cand == tkm)) - 1;
                    if (lane == jj) { bmax = mx; btk = tkm; bwl = wl; }
                }
            }
        }
        // C. block argmax over bucket maxima
        if (dirty) {
            dirty = false;
            const unsigned v = lane < BPW ? bmax : 0u;
            wm = __reduce_max_sync(kFull, v);
            const unsigned c2 = (lane < BPW && v == wm) ? btk : kPadKey;
            wtk = __reduce_min_sync(kFull, c2);
            const int src = __ffs(__ballot_sync(kFull, c2 == wtk)) - 1;  // lane 0 if nothing but padding
            const unsigned sl = __shfl_sync(kFull, bwl, src);
            wpos = (((unsigned)src * NW + w) << 5) + sl;
        }
        if (lane == 0) pub[(j & 1) * NW + w] = make_uint4(wm, wtk, wpos, 0u);
        __syncthreads();
        uint4 e = make_uint4(0u, kPadKey, 0u, 0u);
        if (lane < NW) e = pub[(j & 1) * NW + lane];
        const unsigned gm = __reduce_max_sync(kFull, e.x);
        const unsigned c3 = (e.x == gm) ? e.y : kPadKey;
        const unsigned gtk = __reduce_min_sync(kFull, c3);
        const int srcw = __ffs(__ballot_sync(kFull, c3 == gtk)) - 1;
        const unsigned gpos = __shfl_sync(kFull, e.z, srcw);
        cx = sx[gpos];
        cy = sy[gpos];
        cz = sz[gpos];
        if (tid == 0) out[j] = (int)fps_tiekey_inv(gtk, p, bsmask);
    }

    // ---- 5. leave temp as the reference does ---------------------------------------------
#pragma unroll
    for (int j = 0; j < BPW; ++j)
        if (tk[j] != kPadKey) tmp[fps_tiekey_inv(tk[j], p, bsmask)] = t[j];
}

template <int NW, int BPW>
static int launch_bucket(int b, int n, int m, int p, const float *xyz, float *temp, int *idx,
                         cudaStream_t st) {
    using L = FpsSmem<NW, BPW>;
    auto kern = fps_bucket_kernel<NW, BPW>;
    // per launch (a few hundred ns): the attribute is per device, and one process may drive several
    PDM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kBytes));
    kern<<<b, NW * 32, L::kBytes, st>>>(n, m, p, xyz, temp, idx);
    count_launch();
    PDM_CHECK_LAUNCH("farthest_point_sampling(bucket)");
    return PDM_OK;
}

}  // namespace pdm

extern "C" int pdm_farthest_point_sampling(int b, int n, int m, const float *xyz, float *temp,
                                           int *idx, void *stream) {
    using namespace pdm;
    if (b < 0 || n < 0 || m < 0) return fail(PDM_ERR_INVALID_ARG, "farthest_point_sampling: negative size");
    if (b == 0 || m == 0) return PDM_OK;
    if (n == 0) return fail(PDM_ERR_INVALID_ARG, "farthest_point_sampling: n == 0 with m > 0");
    if (!xyz || !temp || !idx) return fail(PDM_ERR_INVALID_ARG, "farthest_point_sampling: null pointer");
    if ((long long)n * 3 > 0x7fffffffLL) return fail(PDM_ERR_UNSUPPORTED, "farthest_point_sampling: n too large");
    cudaStream_t st = (cudaStream_t)stream;
    const int bs = ref_fps_block_size(n);
    int p = 0;
    while ((1 << p) < bs) ++p;
    const char *force = getenv("PDM_FPS_KERNEL");  // "generic" | unset (debug/testing knob)
    const bool generic = force && force[0] == 'g';
    if (!generic && n >= 512 && n <= 16384) {
        if (n <= 1024) return launch_bucket<32, 1>(b, n, m, p, xyz, temp, idx, st);
        if (n <= 2048) return launch_bucket<32, 2>(b, n, m, p, xyz, temp, idx, st);
        if (n <= 4096) return launch_bucket<32, 4>(b, n, m, p, xyz, temp, idx, st);
        if (n <= 8192) return launch_bucket<32, 8>(b, n, m, p, xyz, temp, idx, st);
        return launch_bucket<32, 16>(b, n, m, p, xyz, temp, idx, st);
    }
    fps_generic_kernel<1024><<<b, 1024, 0, st>>>(n, m, p, xyz, temp, idx);
    count_launch();
    PDM_CHECK_LAUNCH("farthest_point_sampling(generic)");
    return PDM_OK;
}
