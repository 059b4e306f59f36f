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
#include <stdlib.h>

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
    // then pub[2][NW] uint2 (value bits, position), pubtk[2][NW] tiekeys, frame-box scratch
    static constexpr size_t kPubOff = (size_t)12 * CAP;
    static constexpr size_t kTkOff = kPubOff + sizeof(uint2) * 2 * NW;
    static constexpr size_t kBoxOff = kTkOff + sizeof(unsigned) * 2 * NW;
    static constexpr size_t kBytes = kBoxOff + sizeof(float) * 6 * NW;
};

// Critical-path idiom (latencies measured on B200, tools/micro/lat.cu): a warp argmax that also
// needs a payload of the winning lane costs  redux(28) + redux(27) = 55 cycles as
//     mx = redux.max(v);  payload = redux.max(v == mx ? payload : 0)
// against 108 for redux + ballot + ffs + shfl.  The payload form is only valid when exactly one
// lane holds the maximum; equal maxima (duplicate points) are detected with a ballot that runs
// off the critical path and resolved by the smallest tiekey in a slow path.
__device__ __forceinline__ bool multi_bit(unsigned ball) { return (ball & (ball - 1u)) != 0u; }

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

// PROF: debug build that accumulates clock64() per phase and per warp into `prof`
// ([frame][warp][8] = A, B, C, barrier wait, D cycles, #bucket updates, #C runs, total).
template <int NW, int BPW, bool PROF = false>
__global__ void __launch_bounds__(NW * 32, 1)
fps_bucket_kernel(int n, int m, int p, const float *__restrict__ xyz, float *__restrict__ temp,
                  int *__restrict__ idxs, long long *__restrict__ prof = nullptr) {
    using L = FpsSmem<NW, BPW>;
    constexpr int CAP = L::CAP;
    constexpr int T = NW * 32;
    static_assert(BPW <= 32 && NW <= 32, "one lane per owned bucket / per warp");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sx = reinterpret_cast<float *>(smem_raw);
    float *sy = sx + CAP;
    float *sz = sy + CAP;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);
    uint2 *pub = reinterpret_cast<uint2 *>(smem_raw + L::kPubOff);
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
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float l2 = lane < NW ? box[lane * 6 + a] : INFINITY;
        const float h2 = lane < NW ? box[lane * 6 + 3 + a] : -INFINITY;
        lo[a] = ord2f(__reduce_min_sync(kFull, f2ord(l2)));
        hi[a] = ord2f(__reduce_max_sync(kFull, f2ord(h2)));
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
    // Sorted position pos = b*32 + lane.  Padding keys sort last, so pos >= n <=> padding.
    // The caller's scratch `temp` doubles as the pos -> original-index map while the kernel
    // runs (read back only for the output index and for tie-breaks); step 5 restores it.
    float t[BPW];       // running min distance of my point in owned bucket j
    unsigned kk_[BPW];
#pragma unroll
    for (int j = 0; j < BPW; ++j) kk_[j] = (unsigned)keys[((j * NW + w) << 5) + lane];  // low word = k
    __syncthreads();  // keys are dead from here on; the region becomes sx/sy/sz
#pragma unroll
    for (int j = 0; j < BPW; ++j) {
        const int pos = ((j * NW + w) << 5) + lane;
        t[j] = pos < n ? tmp[kk_[j]] : 0.f;   // padding: 0 and never the tie winner
    }
    __syncthreads();  // every initial temp value is read before the map overwrites the buffer
    unsigned *pmap = reinterpret_cast<unsigned *>(tmp);

    // per-bucket state, held by lane j of the owning warp
    float blox = INFINITY, bloy = INFINITY, bloz = INFINITY;
    float bhix = -INFINITY, bhiy = -INFINITY, bhiz = -INFINITY;
    unsigned bmax = 0u, bwl = 0u;
#pragma unroll
    for (int j = 0; j < BPW; ++j) {
        const int pos = ((j * NW + w) << 5) + lane;
        const bool pad = pos >= n;
        float x = 0.f, y = 0.f, z = 0.f;
        if (!pad) {
            x = __ldg(dataset + kk_[j] * 3 + 0);
            y = __ldg(dataset + kk_[j] * 3 + 1);
            z = __ldg(dataset + kk_[j] * 3 + 2);
            pmap[pos] = kk_[j];
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
        const unsigned cand = (tb == mx && !pad) ? fps_tiekey(kk_[j], p, bsmask) : kPadKey;
        const unsigned tkm = __reduce_min_sync(kFull, cand);
        const unsigned wl = __ffs(__ballot_sync(kFull, cand == tkm)) - 1;
        if (lane == j) {
            blox = ord2f(lx); bloy = ord2f(ly); bloz = ord2f(lz);
            bhix = ord2f(hx); bhiy = ord2f(hy); bhiz = ord2f(hz);
            bmax = mx; bwl = wl;
        }
    }
    __syncthreads();

    // ---- 4. rounds -------------------------------------------------------------------------
    float cx = __ldg(dataset + 0), cy = __ldg(dataset + 1), cz = __ldg(dataset + 2);  // sample 0 = point 0
    unsigned wm = 0u, wpos = 0u;  // cached best of this warp (value bits, sorted position)
    bool dirty = true;
    const int wbase = (w << 5) + lane;                               // my slot in owned bucket 0
    const unsigned bbase = (unsigned)(lane * (NW * 32) + (w << 5));  // first slot of owned bucket `lane`
    // tiekey of the point at sorted position pos (global read: slow paths only)
    auto tiekey_at = [&](unsigned pos) -> unsigned {
        return pos < (unsigned)n ? fps_tiekey(pmap[pos], p, bsmask) : kPadKey;
    };
    unsigned pending = 0u;  // thread 0: original index of the previous round's winner (in flight)
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tq = 0, tq0 = 0;
    if (PROF) tq0 = tq = clock64();
#define PDM_TICK(SLOT)                         \
    if (PROF) {                                \
        const long long now = clock64();       \
        pc[SLOT] += now - tq;                  \
        tq = now;                              \
    }
    for (int j = 1; j < m; ++j) {
        // A. which of my buckets can change?  exact lower bound of d over the bucket box
        // (lanes >= BPW hold an empty box: gap = +inf, never active -- no divergent branch needed)
        const float gx = fmaxf(fmaxf(__fsub_rn(blox, cx), __fsub_rn(cx, bhix)), 0.f);
        const float gy = fmaxf(fmaxf(__fsub_rn(bloy, cy), __fsub_rn(cy, bhiy)), 0.f);
        const float gz = fmaxf(fmaxf(__fsub_rn(bloz, cz), __fsub_rn(cz, bhiz)), 0.f);
        unsigned mask = __ballot_sync(kFull, sqdist_ref(gx, gy, gz) < __uint_as_float(bmax));
        PDM_TICK(0)
        // B. update the surviving buckets, one 32-point bucket per iteration
        while (mask) {
            const int jj = 31 - __clz(mask);
            mask ^= 1u << jj;
            const int pos = jj * (NW * 32) + wbase;
            const float d = sqdist_ref(__fsub_rn(sx[pos], cx), __fsub_rn(sy[pos], cy), __fsub_rn(sz[pos], cz));
            const float nt = fminf(d, reg_select<BPW>(t, jj));
            const unsigned tb = __float_as_uint(nt);
            const unsigned mx = __reduce_max_sync(kFull, tb);
            const bool hit = tb == mx;
            unsigned wl = __reduce_max_sync(kFull, hit ? (unsigned)lane : 0u);
            if (multi_bit(__ballot_sync(kFull, hit))) {  // duplicates: smallest tiekey wins
                const unsigned cand = hit ? tiekey_at(pos) : kPadKey;
                const unsigned tkm = __reduce_min_sync(kFull, cand);
                wl = __reduce_max_sync(kFull, cand == tkm ? (unsigned)lane : 0u);
            }
            if (lane == jj) { bmax = mx; bwl = wl; }
            reg_store<BPW>(t, jj, nt);
            dirty = true;
            if (PROF) pc[5] += 1;
        }
        PDM_TICK(1)
        // C. best of this warp (only when one of its buckets changed)
        if (dirty) {
            dirty = false;
            const unsigned v = lane < BPW ? bmax : 0u;
            wm = __reduce_max_sync(kFull, v);
            const bool hit = lane < BPW && v == wm;
            wpos = __reduce_max_sync(kFull, hit ? bbase + bwl : 0u);
            if (multi_bit(__ballot_sync(kFull, hit))) {  // several buckets share the maximum
                const unsigned c2 = hit ? tiekey_at(bbase + bwl) : kPadKey;
                const unsigned tkm = __reduce_min_sync(kFull, c2);
                wpos = __reduce_max_sync(kFull, (hit && c2 == tkm) ? bbase + bwl : 0u);
            }
            if (PROF) pc[6] += 1;
        }
        PDM_TICK(2)
        const int par = (j & 1) * NW;
        if (lane == 0) pub[par + w] = make_uint2(wm, wpos);
        __syncthreads();
        PDM_TICK(3)
        // D. block argmax (every warp redundantly: no second barrier)
        uint2 e = make_uint2(0u, 0u);
        if (lane < NW) e = pub[par + lane];
        const unsigned gm = __reduce_max_sync(kFull, e.x);
        const bool ghit = lane < NW && e.x == gm;
        unsigned gpos = __reduce_max_sync(kFull, ghit ? e.y : 0u);
        if (multi_bit(__ballot_sync(kFull, ghit))) {  // several warps share the maximum
            const unsigned c3 = ghit ? tiekey_at(e.y) : kPadKey;
            const unsigned gtk = __reduce_min_sync(kFull, c3);
            gpos = __reduce_max_sync(kFull, (ghit && c3 == gtk) ? e.y : 0u);
        }
        cx = sx[gpos];
        cy = sy[gpos];
        cz = sz[gpos];
        // output index: software-pipelined so the global read of the map never stalls a round
        if (tid == 0) {
            if (j > 1) out[j - 1] = (int)pending;
            pending = pmap[gpos];
        }
        if (PROF) { cx += 0.f * __uint_as_float(gpos); }  // keep the loads inside the D window
        PDM_TICK(4)
    }
    if (tid == 0) out[m - 1] = (int)pending;
#undef PDM_TICK
    if (PROF && lane == 0 && prof) {
        pc[7] = clock64() - tq0;
        for (int q = 0; q < 8; ++q) prof[((size_t)blockIdx.x * NW + w) * 8 + q] = pc[q];
    }

    // ---- 5. leave temp as the reference does: running minima in original order ------------
    __syncthreads();
    unsigned ko[BPW];
#pragma unroll
    for (int j = 0; j < BPW; ++j) {
        const int pos = ((j * NW + w) << 5) + lane;
        ko[j] = pos < n ? pmap[pos] : 0u;
    }
    __syncthreads();  // all map entries are read before any of them is overwritten
#pragma unroll
    for (int j = 0; j < BPW; ++j) {
        const int pos = ((j * NW + w) << 5) + lane;
        if (pos < n) tmp[ko[j]] = t[j];
    }
}

template <int NW, int BPW>
static int launch_bucket(int b, int n, int m, int p, const float *xyz, float *temp, int *idx,
                         cudaStream_t st) {
    using L = FpsSmem<NW, BPW>;
    auto kern = fps_bucket_kernel<NW, BPW>;
    // per launch (a few hundred ns): the attribute is per device, and one process may drive several
    PDM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kBytes));
    kern<<<b, NW * 32, L::kBytes, st>>>(n, m, p, xyz, temp, idx, nullptr);
    count_launch();
    PDM_CHECK_LAUNCH("farthest_point_sampling(bucket)");
    return PDM_OK;
}

// capacity (points) -> kernel with `nw` warps; returns PDM_ERR_UNSUPPORTED for combinations
// that do not exist (BPW must stay within 1..32)
template <int CAP>
static int launch_cap(int nw, int b, int n, int m, int p, const float *xyz, float *temp, int *idx,
                      cudaStream_t st) {
    switch (nw) {
#define PDM_NW(NWV)                                                                    \
    case NWV:                                                                          \
        if constexpr (CAP / (32 * NWV) >= 1 && CAP / (32 * NWV) <= 32)                 \
            return launch_bucket<NWV, CAP / (32 * NWV)>(b, n, m, p, xyz, temp, idx, st); \
        break;
        PDM_NW(4) PDM_NW(8) PDM_NW(16) PDM_NW(32)
#undef PDM_NW
        default: break;
    }
    return fail(PDM_ERR_UNSUPPORTED, "farthest_point_sampling: no bucket kernel for cap %d with %d warps", CAP, nw);
}

}  // namespace pdm

// Debug-only entry (not part of include/pdm_ops.h): per-phase cycle counters of the bucket
// kernel, prof = device buffer of b*16*8 long long.  16384-point and 4096-point frames only.
extern "C" int pdm_debug_fps_profile(int b, int n, int m, const float *xyz, float *temp, int *idx,
                                     long long *prof, void *stream) {
    using namespace pdm;
    const int bs = ref_fps_block_size(n);
    int p = 0;
    while ((1 << p) < bs) ++p;
    cudaStream_t st = (cudaStream_t)stream;
    if (n > 4096 && n <= 16384) {
        using L = FpsSmem<16, 32>;
        auto kern = fps_bucket_kernel<16, 32, true>;
        PDM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kBytes));
        kern<<<b, 512, L::kBytes, st>>>(n, m, p, xyz, temp, idx, prof);
    } else if (n <= 4096 && n > 2048) {
        using L = FpsSmem<16, 8>;
        auto kern = fps_bucket_kernel<16, 8, true>;
        PDM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kBytes));
        kern<<<b, 512, L::kBytes, st>>>(n, m, p, xyz, temp, idx, prof);
    } else {
        return fail(PDM_ERR_UNSUPPORTED, "debug_fps_profile: n=%d", n);
    }
    PDM_CHECK_LAUNCH("debug_fps_profile");
    return PDM_OK;
}

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
        // warps per CTA: few warps keep the per-round issue + barrier cost low, enough warps
        // keep the (few) surviving bucket updates of a round in different warps.  Tuned on B200;
        // PDM_FPS_NW overrides for experiments.
        const char *nwenv = getenv("PDM_FPS_NW");
        int nw = nwenv ? atoi(nwenv) : 0;
        if (n <= 1024) return launch_cap<1024>(nw ? nw : 16, b, n, m, p, xyz, temp, idx, st);
        if (n <= 2048) return launch_cap<2048>(nw ? nw : 16, b, n, m, p, xyz, temp, idx, st);
        if (n <= 4096) return launch_cap<4096>(nw ? nw : 16, b, n, m, p, xyz, temp, idx, st);
        if (n <= 8192) return launch_cap<8192>(nw ? nw : 16, b, n, m, p, xyz, temp, idx, st);
        return launch_cap<16384>(nw ? nw : 16, b, n, m, p, xyz, temp, idx, st);
    }
    fps_generic_kernel<1024><<<b, 1024, 0, st>>>(n, m, p, xyz, temp, idx);
    count_launch();
    PDM_CHECK_LAUNCH("farthest_point_sampling(generic)");
    return PDM_OK;
}
