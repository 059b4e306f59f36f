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
//   on-chip.  Points are sorted once along a space-filling curve (in-CTA bitonic sort) and cut into buckets of
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
//   SEVERAL samples are accepted per barrier round when that is provably what the sequential
//   algorithm would do (see "multi-sample rounds" below).
// fps_generic_kernel: any n; same key trick, temp in global memory.
#include <stdlib.h>

#include "fps_common.cuh"

namespace pdm {

// Stacked (ragged) frames, pointnet2_stack/src/sampling_gpu.cu:263-276: frame f owns rows
// [sum n_cnt[:f], +n_cnt[f]) of xyz / temp and entries [sum m_cnt[:f], +m_cnt[f]) of idx; the sampled
// indices are GLOBAL rows (local index + frame start).  n_cnt == nullptr: uniform (B,N,3) frames.
struct FpsRagged {
    const int *n_cnt = nullptr, *m_cnt = nullptr;
    int skip_le = 0;      // generic kernel only: frames with n <= skip_le were sampled by the on-chip kernel
    int flags = 0;        // bucket kernel: bit 0 = update surviving buckets two at a time (fps_pair_flag)
};
// PDM_FPS_PAIR=0 switches the paired bucket update off (A/B measurements)
static int fps_pair_flag() {
    static const int v = [] { const char *e = getenv("PDM_FPS_PAIR"); return (e && e[0] == '0') ? 0 : 1; }();
    return v;
}
__device__ __forceinline__ bool fps_ragged_frame(const FpsRagged &rg, int f, int &n, int &m, size_t &pstart,
                                                 size_t &ostart) {
    int ps = 0, os = 0;
    for (int k = 0; k < f; ++k) { ps += __ldg(rg.n_cnt + k); os += __ldg(rg.m_cnt + k); }
    n = __ldg(rg.n_cnt + f);
    m = __ldg(rg.m_cnt + f);
    pstart = (size_t)ps;
    ostart = (size_t)os;
    return n > 0 && m > 0;
}

// ---------------------------------------------------------------------------------------
// generic kernel
// ---------------------------------------------------------------------------------------
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
fps_generic_kernel(int n, int m, int p, const float *__restrict__ xyz, float *__restrict__ temp,
                   int *__restrict__ idxs, FpsRagged rg = FpsRagged{}) {
    constexpr int NWARP = THREADS / 32;
    __shared__ unsigned long long wbest[2][NWARP];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned bsmask = (1u << p) - 1u;
    size_t pstart = (size_t)blockIdx.x * n, ostart = (size_t)blockIdx.x * m;
    int obase = 0;
    if (rg.n_cnt) {   // stacked frames (pointnet2_stack): per-frame n, m; indices are global rows
        if (!fps_ragged_frame(rg, blockIdx.x, n, m, pstart, ostart)) return;
        if (n <= rg.skip_le) return;     // the on-chip kernel took this frame
        obase = (int)pstart;
    }
    const float *dataset = xyz + pstart * 3;
    float *tmp = temp + pstart;
    int *out = idxs + ostart;
    int old = 0;
    if (tid == 0) out[0] = obase;
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
        if (tid == 0) out[j] = old + obase;
    }
}

// ---------------------------------------------------------------------------------------
// bucketed on-chip kernel
// ---------------------------------------------------------------------------------------
template <int NW, int BPW, int KMAX>
struct FpsSmem {
    static constexpr int CAP = NW * BPW * 32;
    // [0, 12*CAP): sx, sy, sz (aliased by the 4*CAP-byte sort keys / exchange buffer during set-up), then
    // pub[2][2*NW] uint2 (value bits, position) -- two candidates per warp,
    // pubU[2][NW] bound on every other point of the warp, samp[NW][KMAX] float4 accepted samples
    // (one private copy per warp), frame-box scratch.
    static constexpr size_t kPubOff = (size_t)12 * CAP;
    static constexpr size_t kUOff = kPubOff + sizeof(uint2) * 2 * 2 * NW;
    static constexpr size_t kSampOff = ((kUOff + sizeof(unsigned) * 2 * NW + 15) / 16) * 16;
    static constexpr size_t kBoxOff = kSampOff + sizeof(float4) * NW * KMAX;
    static constexpr size_t kBytes = kBoxOff + sizeof(float) * 6 * NW;
};

// Critical-path idiom (latencies measured on B200, tools/micro/lat.cu): a warp argmax that also
// needs a payload of the winning lane costs  redux(28) + redux(27) = 55 cycles as
//     mx = redux.max(v);  payload = redux.max(v == mx ? payload : 0)
// against 108 for redux + ballot + ffs + shfl.  The payload form is only valid when exactly one
// lane holds the maximum; equal maxima (duplicate points) are detected with a ballot that runs
// off the critical path and resolved by the smallest tiekey in a slow path.

// (reg_select / reg_store: register-array access by a warp-uniform runtime index, fps_common.cuh)

// Multi-sample rounds
// -------------------
// A barrier round costs ~1100 cycles of dependent latency, so the kernel tries to emit several
// samples per round.  Every warp publishes TWO candidates -- the best point of its best and of
// its second-best bucket, with exact values -- and a bound U_w on all its other points (the
// maxima of its remaining buckets and the runner-up inside the two candidate buckets).  With
// U = max_w U_w, every warp then replays the sequential algorithm on the 2*NW candidates alone
// (one per lane, redundantly, no further barrier):
//   pick 1: the exact block argmax with the reference's tie-break, exactly as a 1-sample round;
//   pick i>1: update the candidates' minima with the previous pick (same sqdist_ref rounding);
//             if the largest candidate value is unique and STRICTLY greater than U it is the
//             unique global maximum -- every non-candidate is <= U because minima only decrease --
//             so it is what the reference picks next, no tie-break needed; otherwise stop.
// The accepted samples (<= KMAX) are then applied to the buckets together.  On KITTI-shaped
// frames this emits ~6 samples per barrier.
// TRACE: debug instantiation; frame 0 dumps clock64() stamps per round and warp into `trace`
// ([round][warp][8] = t_start, t_afterA, t_afterB, t_afterC(before barrier), t_afterBarrier,
//  t_afterPick1, t_afterLoop, K | nupdates<<8).
// LB: thread count promised to ptxas.  The kernel always runs NW*32 threads; promising more makes ptxas
// budget fewer registers per thread (65536 / LB), which leaves register-file room on the SM for
// CTAs of OTHER kernels (ball query, grouping) next to a resident FPS CTA -- see fps_dispatch.
template <int NW, int BPW, int KMAX, bool TRACE = false, int LB = NW * 32, bool PAIR = (BPW < 32)>
__global__ void __launch_bounds__(LB, 1)
fps_bucket_kernel(int n, int m, int p, const float *__restrict__ xyz, float *__restrict__ temp,
                  int *__restrict__ idxs, int *__restrict__ stats, long long *__restrict__ trace = nullptr,
                  FpsRagged rg = FpsRagged{}) {
    using L = FpsSmem<NW, BPW, KMAX>;
    constexpr int CAP = L::CAP;
    constexpr int T = NW * 32;
    static_assert(BPW <= 32 && NW <= 16, "one lane per owned bucket; two candidates per warp in one warp");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sx = reinterpret_cast<float *>(smem_raw);
    float *sy = sx + CAP;
    float *sz = sy + CAP;
    uint2 *pub = reinterpret_cast<uint2 *>(smem_raw + L::kPubOff);
    unsigned *pubU = reinterpret_cast<unsigned *>(smem_raw + L::kUOff);
    float4 *samp = reinterpret_cast<float4 *>(smem_raw + L::kSampOff);
    float *box = reinterpret_cast<float *>(smem_raw + L::kBoxOff);

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const unsigned bsmask = (1u << p) - 1u;
    size_t pstart = (size_t)blockIdx.x * n, ostart = (size_t)blockIdx.x * m;
    int obase = 0;
    if (rg.n_cnt) {   // stacked frames (pointnet2_stack): per-frame n, m; indices are global rows
        if (!fps_ragged_frame(rg, blockIdx.x, n, m, pstart, ostart)) return;
        if (n > CAP) return;             // left to the any-size kernel (launched next)
        obase = (int)pstart;
    }
    const float *dataset = xyz + pstart * 3;
    float *tmp = temp + pstart;
    int *out = idxs + ostart;

    if (tid == 0) out[0] = obase;
    if (m <= 1) return;
    // paired bucket updates: -2 % on frames of <= 8192 points, nothing at 16384 -- where the extra code costs registers
    // (52 -> 112 bytes of spills at 96 registers, +3 %): compiled out for the 32-buckets-per-warp instantiation
    const bool pair = PAIR && (rg.flags & 1) != 0;

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
    FpsCurve curve;   // spatial sort code (fps_common.cuh)
    curve.init(lo, hi);

    // ---- 2. sort along the curve: 32-bit keys (18-bit code | 14-bit index), BPW per thread in registers
    //         (fps_sort_keys: register / shuffle / a few shared-memory stages; the first version sorted
    //         64-bit keys in shared memory with a barrier per stage: ~0.25 ms of a 2.3 ms kernel)
    unsigned *skeys = reinterpret_cast<unsigned *>(smem_raw);
    {
        unsigned v[BPW];
#pragma unroll
        for (int r = 0; r < BPW; ++r) {
            const int k = r * T + tid;            // any initial arrangement will do: coalesced reads
            v[r] = 0xffffffffu;                   // padding sorts last
            if (k < n) {
                const float c[3] = {__ldg(dataset + k * 3 + 0), __ldg(dataset + k * 3 + 1), __ldg(dataset + k * 3 + 2)};
                v[r] = (curve.code18(c) << 14) | (unsigned)k;
            }
        }
        fps_sort_keys<BPW, T>(v, skeys, tid);
        __syncthreads();                           // nobody is still reading the exchange buffer
#pragma unroll
        for (int r = 0; r < BPW; ++r) skeys[tid * BPW + r] = v[r];   // sorted position e = tid * BPW + r
    }
    __syncthreads();

    // ---- 3. distribute: lane owns slot `lane` of buckets  b = j*NW + w --------------------
    // Sorted position pos = b*32 + lane.  Padding keys sort last, so pos >= n <=> padding.
    // The caller's scratch `temp` doubles as the pos -> original-index map while the kernel
    // runs (read back only for the output index and for tie-breaks); step 5 restores it.
    float t[BPW];       // running min distance of my point in owned bucket j
    unsigned kk_[BPW];
#pragma unroll
    for (int j = 0; j < BPW; ++j) kk_[j] = skeys[((j * NW + w) << 5) + lane] & 0x3fffu;   // low 14 bits = k
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
    unsigned bmax = 0u, bwl = 0u, bsec = 0u;  // max (bits), lane holding it, runner-up (bits)
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
        const unsigned sec = __reduce_max_sync(kFull, lane == (int)wl ? 0u : tb);
        if (lane == j) {
            blox = ord2f(lx); bloy = ord2f(ly); bloz = ord2f(lz);
            bhix = ord2f(hx); bhiy = ord2f(hy); bhiz = ord2f(hz);
            bmax = mx; bwl = wl; bsec = sec;
        }
    }
    if (lane == 0) samp[w * KMAX] = make_float4(__ldg(dataset + 0), __ldg(dataset + 1), __ldg(dataset + 2), 0.f);
    __syncthreads();

    // ---- 4. rounds -------------------------------------------------------------------------
    const int wbase = (w << 5) + lane;                               // my slot in owned bucket 0
    const unsigned bbase = (unsigned)(lane * (NW * 32) + (w << 5));  // first slot of owned bucket `lane`
    // tiekey of the point at sorted position pos (global read: slow paths only)
    auto tiekey_at = [&](unsigned pos) -> unsigned {
        return pos < (unsigned)n ? fps_tiekey(pmap[pos], p, bsmask) : kPadKey;
    };
    unsigned c1v = 0u, c1p = 0u, c2v = 0u, c2p = 0u, wU = 0u;  // cached candidates / bound of this warp
    bool dirty = true;
    int K = 1;         // samples accepted in the previous round, waiting to be applied (sample 0 first)
    int j = 1;         // samples emitted so far
    int rounds = 0;
    int pend_slot = -1;       // warp 0: output slot of the sample this lane's candidate became ...
    unsigned pend_val = 0u;   // ... and its original index (global load in flight since last round)
    long long tr[8];
    int nupd = 0;
#define PDM_STAMP(I) if (TRACE) tr[I] = clock64();
    for (;;) {
        PDM_STAMP(0)
        nupd = 0;
        // A. which of my buckets can change?  exact lower bound of d over the bucket box, for
        //    each of the K new samples (lanes >= BPW hold an empty box: gap = +inf, never active).
        //    Runtime loop over the K samples (warp-private copy in shared memory, broadcast reads):
        //    this phase is issue-bound -- every warp runs it for every sample.
        const float4 *ws = samp + w * KMAX;
        unsigned amask = 0u;  // bit k: sample k can change my bucket
        {
            const float bm = __uint_as_float(bmax);
            // two samples per iteration: the two bound chains are independent and overlap
            // (entries beyond K are stale but finite or zero; their bits are masked off)
            for (int k = 0; k < K; k += 2) {
                const float4 c = ws[k], e2 = ws[(k + 1 < KMAX) ? k + 1 : k];
                const float gx = fmaxf(fmaxf(__fsub_rn(blox, c.x), __fsub_rn(c.x, bhix)), 0.f);
                const float gy = fmaxf(fmaxf(__fsub_rn(bloy, c.y), __fsub_rn(c.y, bhiy)), 0.f);
                const float gz = fmaxf(fmaxf(__fsub_rn(bloz, c.z), __fsub_rn(c.z, bhiz)), 0.f);
                const float hx = fmaxf(fmaxf(__fsub_rn(blox, e2.x), __fsub_rn(e2.x, bhix)), 0.f);
                const float hy = fmaxf(fmaxf(__fsub_rn(bloy, e2.y), __fsub_rn(e2.y, bhiy)), 0.f);
                const float hz = fmaxf(fmaxf(__fsub_rn(bloz, e2.z), __fsub_rn(e2.z, bhiz)), 0.f);
                const unsigned a0 = sqdist_ref(gx, gy, gz) < bm ? 1u : 0u;
                const unsigned a1 = (k + 1 < K && sqdist_ref(hx, hy, hz) < bm) ? 2u : 0u;
                amask |= (a0 | a1) << k;
            }
        }
        unsigned mask = __ballot_sync(kFull, amask != 0u);
        PDM_STAMP(1)
        // B. update the surviving buckets (32 points = 32 lanes each) with the samples that
        //    reach them (usually one)
        while (mask) {
            if (pair && (mask & (mask - 1u))) {
                // two surviving buckets at once: their update chains (shuffle -> shared-memory reads -> distance ->
                // three reductions, ~300 cycles of dependent latency each) are independent and overlap
                const int j1 = 31 - __clz(mask);
                mask ^= 1u << j1;
                const int j2 = 31 - __clz(mask);
                mask ^= 1u << j2;
                const unsigned s1 = __shfl_sync(kFull, amask, j1), s2 = __shfl_sync(kFull, amask, j2);
                const int p1 = j1 * (NW * 32) + wbase, p2 = j2 * (NW * 32) + wbase;
                const float x1 = sx[p1], y1 = sy[p1], z1 = sz[p1];
                const float x2 = sx[p2], y2 = sy[p2], z2 = sz[p2];
                float n1 = reg_select<BPW>(t, j1), n2 = reg_select<BPW>(t, j2);
                unsigned su = s1 | s2;
                while (su) {
                    const int k = 31 - __clz(su);
                    su ^= 1u << k;
                    const float4 c = ws[k];
                    const float d1 = sqdist_ref(__fsub_rn(x1, c.x), __fsub_rn(y1, c.y), __fsub_rn(z1, c.z));
                    const float d2 = sqdist_ref(__fsub_rn(x2, c.x), __fsub_rn(y2, c.y), __fsub_rn(z2, c.z));
                    n1 = ((s1 >> k) & 1u) ? fminf(d1, n1) : n1;
                    n2 = ((s2 >> k) & 1u) ? fminf(d2, n2) : n2;
                }
                const unsigned tb1 = __float_as_uint(n1), tb2 = __float_as_uint(n2);
                const unsigned mx1 = __reduce_max_sync(kFull, tb1), mx2 = __reduce_max_sync(kFull, tb2);
                const bool h1 = tb1 == mx1, h2 = tb2 == mx2;
                unsigned wl1 = __reduce_max_sync(kFull, h1 ? (unsigned)lane : 0u), wl2 = __reduce_max_sync(kFull, h2 ? (unsigned)lane : 0u);
                unsigned sec1 = __reduce_max_sync(kFull, h1 ? 0u : tb1), sec2 = __reduce_max_sync(kFull, h2 ? 0u : tb2);
                const unsigned bl1 = __ballot_sync(kFull, h1), bl2 = __ballot_sync(kFull, h2);
                if (multi_bit(bl1)) {
                    const unsigned cand = h1 ? tiekey_at(p1) : kPadKey;
                    const unsigned tkm = __reduce_min_sync(kFull, cand);
                    wl1 = __reduce_max_sync(kFull, cand == tkm ? (unsigned)lane : 0u);
                    sec1 = mx1;
                }
                if (multi_bit(bl2)) {
                    const unsigned cand = h2 ? tiekey_at(p2) : kPadKey;
                    const unsigned tkm = __reduce_min_sync(kFull, cand);
                    wl2 = __reduce_max_sync(kFull, cand == tkm ? (unsigned)lane : 0u);
                    sec2 = mx2;
                }
                if (lane == j1) { bmax = mx1; bwl = wl1; bsec = sec1; }
                if (lane == j2) { bmax = mx2; bwl = wl2; bsec = sec2; }
                reg_store<BPW>(t, j1, n1);
                reg_store<BPW>(t, j2, n2);
                dirty = true;
                nupd += 2;
                continue;
            }
            const int jj = 31 - __clz(mask);
            mask ^= 1u << jj;
            unsigned smask = __shfl_sync(kFull, amask, jj);
            const int pos = jj * (NW * 32) + wbase;
            const float x = sx[pos], y = sy[pos], z = sz[pos];
            float nt = reg_select<BPW>(t, jj);
            while (smask) {
                const int k = 31 - __clz(smask);
                smask ^= 1u << k;
                const float4 c = ws[k];
                nt = fminf(sqdist_ref(__fsub_rn(x, c.x), __fsub_rn(y, c.y), __fsub_rn(z, c.z)), nt);
            }
            const unsigned tb = __float_as_uint(nt);
            const unsigned mx = __reduce_max_sync(kFull, tb);
            const bool hit = tb == mx;
            unsigned wl = __reduce_max_sync(kFull, hit ? (unsigned)lane : 0u);
            unsigned sec = __reduce_max_sync(kFull, hit ? 0u : tb);  // next distinct value ...
            if (multi_bit(__ballot_sync(kFull, hit))) {  // duplicates: smallest tiekey wins,
                const unsigned cand = hit ? tiekey_at(pos) : kPadKey;
                const unsigned tkm = __reduce_min_sync(kFull, cand);
                wl = __reduce_max_sync(kFull, cand == tkm ? (unsigned)lane : 0u);
                sec = mx;                                // ... and the runner-up equals the maximum
            }
            if (lane == jj) { bmax = mx; bwl = wl; bsec = sec; }
            reg_store<BPW>(t, jj, nt);
            dirty = true;
            ++nupd;
        }
        if (j >= m) break;
        ++rounds;
        PDM_STAMP(2)
        // C. this warp's two candidates and the bound on everything else it owns
        if (dirty) {
            dirty = false;
            const unsigned v = lane < BPW ? bmax : 0u;
            c1v = __reduce_max_sync(kFull, v);
            const bool hit1 = lane < BPW && v == c1v;
            unsigned src1 = __reduce_max_sync(kFull, hit1 ? (unsigned)lane : 0u);
            c2v = __reduce_max_sync(kFull, hit1 ? 0u : v);  // next distinct value (issued early)
            if (multi_bit(__ballot_sync(kFull, hit1))) {    // several buckets share the maximum
                const unsigned cc = hit1 ? tiekey_at(bbase + bwl) : kPadKey;
                const unsigned tkm = __reduce_min_sync(kFull, cc);
                src1 = __reduce_max_sync(kFull, (hit1 && cc == tkm) ? (unsigned)lane : 0u);
                c2v = c1v;  // an equal-valued bucket becomes the second candidate
            }
            const bool hit2 = lane < BPW && lane != (int)src1 && v == c2v;
            const unsigned src2 = __reduce_max_sync(kFull, hit2 ? (unsigned)lane : 0u);
            const bool has2 = __ballot_sync(kFull, hit2) != 0u;
            const bool mine = lane == (int)src1 || (has2 && lane == (int)src2);
            wU = __reduce_max_sync(kFull, mine ? bsec : v);
            c1p = __shfl_sync(kFull, bbase + bwl, src1);
            c2p = __shfl_sync(kFull, bbase + bwl, src2);
            if (!has2) { c2v = 0u; c2p = c1p; }  // single-bucket warp: a dead second candidate
        }
        const int par = (rounds & 1);
        if (lane == 0) {
            pub[par * 2 * NW + 2 * w] = make_uint2(c1v, c1p);
            pub[par * 2 * NW + 2 * w + 1] = make_uint2(c2v, c2p);
            pubU[par * NW + w] = wU;
        }
        PDM_STAMP(3)
        __syncthreads();
        // D. every warp replays the sequential selection on the 2*NW candidates (one per lane)
        if (w == 0 && pend_slot >= 0) out[pend_slot] = (int)pend_val + obase;
        pend_slot = -1;
        const bool live = lane < 2 * NW;
        const uint2 e = live ? pub[par * 2 * NW + lane] : make_uint2(0u, 0u);
        if (TRACE) { tr[4] = clock64() + (long long)(e.x & 0u); }
        const unsigned U = __reduce_max_sync(kFull, lane < NW ? pubU[par * NW + lane] : 0u);
        const float x = sx[e.y], y = sy[e.y], z = sz[e.y];
        float v = __uint_as_float(e.x);
        // pick 1: exact argmax over the warps' first candidates, reference tie-break
        const bool first = live && !(lane & 1);
        const unsigned gm = __reduce_max_sync(kFull, first ? e.x : 0u);
        bool ghit = first && e.x == gm;
        if (multi_bit(__ballot_sync(kFull, ghit))) {
            const unsigned c3 = ghit ? tiekey_at(e.y) : kPadKey;
            const unsigned gtk = __reduce_min_sync(kFull, c3);
            ghit = ghit && c3 == gtk;
        }
        if (TRACE) { v += 0.f * (float)(gm & 1u) + 0.f * x; }  // keep pick 1 inside its window
        PDM_STAMP(5)
        const int kmax_now = min(KMAX, m - j);
        K = 0;
        for (;;) {
            if (ghit) {
                samp[w * KMAX + K] = make_float4(x, y, z, 0.f);
                pend_slot = j + K;  // (only warp 0 reports; see below)
            }
            ++K;
            if (K >= kmax_now) break;
            __syncwarp();
            const float4 s4 = samp[w * KMAX + K - 1];
            v = fminf(sqdist_ref(__fsub_rn(x, s4.x), __fsub_rn(y, s4.y), __fsub_rn(z, s4.z)), v);
            const unsigned vb = live ? __float_as_uint(v) : 0u;
            const unsigned g2 = __reduce_max_sync(kFull, vb);
            if (!(g2 > U)) break;              // a non-candidate may be as large: stop
            ghit = live && vb == g2;
            if (multi_bit(__ballot_sync(kFull, ghit))) break;  // equal candidates: let pick 1 of the next round decide
        }
        __syncwarp();
        // Report the indices through the map one round late: ONE load instruction per round,
        // issued after the loop (a load per pick into the same register would serialise on the
        // register scoreboard at L2 latency), consumed by the store at the top of the next D.
        if (w == 0 && pend_slot >= 0) pend_val = pmap[e.y];
        if (TRACE) {
            tr[6] = clock64();
            tr[7] = K | (nupd << 8);
            if (blockIdx.x == 0 && lane == 0 && trace)
                for (int q = 0; q < 8; ++q) trace[((size_t)(rounds - 1) * NW + w) * 8 + q] = tr[q];
        }
        j += K;
        if (j >= m) {   // the very last sample is never applied (the reference stops after writing it)
            --K;
            if (K == 0) break;
        }
    }
    if (w == 0 && pend_slot >= 0) out[pend_slot] = (int)pend_val + obase;
    if (stats && tid == 0) stats[blockIdx.x] = rounds;

    // ---- 5. leave temp as the reference does: running minima in original order ------------
    __syncthreads();
    unsigned ko[BPW];
#pragma unroll
    for (int jq = 0; jq < BPW; ++jq) {
        const int pos = ((jq * NW + w) << 5) + lane;
        ko[jq] = pos < n ? pmap[pos] : 0u;
    }
    __syncthreads();  // all map entries are read before any of them is overwritten
#pragma unroll
    for (int jq = 0; jq < BPW; ++jq) {
        const int pos = ((jq * NW + w) << 5) + lane;
        if (pos < n) tmp[ko[jq]] = t[jq];
    }
}

template <int NW, int BPW, int KMAX, int LB = NW * 32>
static int launch_bucket(int b, int n, int m, int p, const float *xyz, float *temp, int *idx,
                         int *stats, cudaStream_t st) {
    using L = FpsSmem<NW, BPW, KMAX>;
    auto kern = fps_bucket_kernel<NW, BPW, KMAX, false, LB>;
    if (int rc = ensure_dynamic_smem((const void *)kern, L::kBytes)) return rc;
    prefer_max_smem((const void *)kern);
    FpsRagged rg0;
    rg0.flags = fps_pair_flag();
    kern<<<b, NW * 32, L::kBytes, st>>>(n, m, p, xyz, temp, idx, stats, nullptr, rg0);
    count_launch();
    PDM_CHECK_LAUNCH("farthest_point_sampling(bucket)");
    return PDM_OK;
}

// capacity (points) -> kernel with `nw` warps and up to `kmax` samples per round; returns
// PDM_ERR_UNSUPPORTED for combinations that are not instantiated
template <int CAP>
static int launch_cap(int nw, int kmax, int b, int n, int m, int p, const float *xyz, float *temp,
                      int *idx, int *stats, cudaStream_t st) {
#define PDM_TRY(NWV, KV)                                                                  \
    if (nw == NWV && kmax == KV) {                                                        \
        if constexpr (CAP / (32 * NWV) >= 1 && CAP / (32 * NWV) <= 32)                    \
            return launch_bucket<NWV, CAP / (32 * NWV), KV>(b, n, m, p, xyz, temp, idx, stats, st); \
    }
    PDM_TRY(16, 8) PDM_TRY(16, 1) PDM_TRY(8, 8) PDM_TRY(16, 4) PDM_TRY(16, 12)
#undef PDM_TRY
    return fail(PDM_ERR_UNSUPPORTED, "farthest_point_sampling: no bucket kernel for cap %d, %d warps, kmax %d",
                CAP, nw, kmax);
}

static std::atomic<int> g_fps_mode{PDM_FPS_MODE_AUTO};

static int fps_dispatch(int b, int n, int m, const float *xyz, float *temp, int *idx, int *stats,
                        cudaStream_t st, int mode_arg = -1) {
    const int bs = ref_fps_block_size(n);
    int p = 0;
    while ((1 << p) < bs) ++p;
    const char *force = getenv("PDM_FPS_KERNEL");  // "generic" | "l2" | "smem" | unset (debug/testing knob)
    const bool generic = force && force[0] == 'g';
    // throughput variant (fps_l2.cu): asked for by the caller (pdm_set_fps_mode) or when one launch
    // alone has more frames than the GPU has SMs
    // per call (pdm_farthest_point_sampling_ex) or the process-wide default (pdm_set_fps_mode)
    const int mode = mode_arg >= 0 ? mode_arg : g_fps_mode.load(std::memory_order_relaxed);
    const bool want_l2 = force ? force[0] == 'l' : (mode == PDM_FPS_MODE_THROUGHPUT || (mode == PDM_FPS_MODE_AUTO && b > kNumSMs));
    static const int l2_min_n = [] { const char *e = getenv("PDM_FPS_L2_MIN_N"); return e ? atoi(e) : 0; }();
    if (!generic && want_l2 && n >= l2_min_n && fps_l2_supports(n)) return fps_l2_launch(b, n, m, p, xyz, temp, idx, stats, st);
    if (force && force[0] == 'c' && force[1] == 'b') {      // experiment: a cluster per KITTI-sized frame
        const int rc = fps_cluster_bucket_launch(b, n, m, p, xyz, temp, idx, stats, st, true);
        if (rc != PDM_ERR_UNSUPPORTED) return rc;
    }
    if (!generic && n >= 512 && n <= 16384) {
        // warps per CTA / samples per round: tuned on B200; PDM_FPS_NW / PDM_FPS_KMAX override
        const char *nwenv = getenv("PDM_FPS_NW"), *kenv = getenv("PDM_FPS_KMAX");
        const int nw = nwenv ? atoi(nwenv) : 16;
        const int km = kenv ? atoi(kenv) : (n <= 8192 ? 12 : 8);      // measured: 12 is 2-3 % faster up to 8192 points, 8 at 16384
        if (n <= 1024) return launch_cap<1024>(nw, km, b, n, m, p, xyz, temp, idx, stats, st);
        if (n <= 2048) return launch_cap<2048>(nw, km, b, n, m, p, xyz, temp, idx, stats, st);
        if (n <= 4096) return launch_cap<4096>(nw, km, b, n, m, p, xyz, temp, idx, stats, st);
        if (n <= 8192) return launch_cap<8192>(nw, km, b, n, m, p, xyz, temp, idx, stats, st);
        // 16384-point frames: 96 registers per thread (LB 544) measured as fast as the 128-register
        // build and leaves a quarter of the register file to co-resident CTAs; PDM_FPS_LB=512 selects
        // the 128-register build
        const char *lbenv = getenv("PDM_FPS_LB");
        if (nw == 16 && km == 8 && !(lbenv && atoi(lbenv) == 512))
            return launch_bucket<16, 32, 8, 544>(b, n, m, p, xyz, temp, idx, stats, st);
        return launch_cap<16384>(nw, km, b, n, m, p, xyz, temp, idx, stats, st);
    }
    // frames larger than one SM: a cluster of CTAs per frame, bucket-pruned (PDM_FPS_KERNEL=cluster: the full-sweep version)
    if (!generic && !(force && force[0] == 'c' && force[1] == 'l') && fps_cluster_bucket_supports(n)) {
        const int rc = fps_cluster_bucket_launch(b, n, m, p, xyz, temp, idx, stats, st);
        if (rc != PDM_ERR_UNSUPPORTED) return rc;
    }
    if (!generic && fps_cluster_supports(n)) {
        const int rc = fps_cluster_launch(b, n, m, p, xyz, temp, idx, st);
        if (rc != PDM_ERR_UNSUPPORTED) return rc;
    }
    fps_generic_kernel<1024><<<b, 1024, 0, st>>>(n, m, p, xyz, temp, idx);
    count_launch();
    PDM_CHECK_LAUNCH("farthest_point_sampling(generic)");
    return PDM_OK;
}

}  // namespace pdm

namespace pdm {
// stack_farthest_point_sampling_kernel_launcher (pointnet2_stack/src/sampling_gpu.cu:335-348): always the
// <1024> instantiation, so the tie-break uses block size 1024 (p = 10) whatever the frame sizes are.
static int fps_stack_dispatch(int n_total, int batch, const float *xyz, float *temp, const int *xyz_cnt, int *idx,
                              const int *m_cnt, cudaStream_t st) {
    const int p = 10;
    FpsRagged rg;
    rg.n_cnt = xyz_cnt;
    rg.m_cnt = m_cnt;
    rg.flags = fps_pair_flag();
    const char *force = getenv("PDM_FPS_KERNEL");
    const bool generic = force && force[0] == 'g';
    int cap = 0;
    if (!generic) {
#define PDM_STACK_BUCKET(NWV, BPWV, LBV)                                                            \
    {                                                                                               \
        using L = FpsSmem<NWV, BPWV, 8>;                                                            \
        auto kern = fps_bucket_kernel<NWV, BPWV, 8, false, LBV>;                                    \
        if (int rc = ensure_dynamic_smem((const void *)kern, L::kBytes)) return rc;                 \
        prefer_max_smem((const void *)kern);                                                        \
        kern<<<batch, NWV * 32, L::kBytes, st>>>(0, 0, p, xyz, temp, idx, nullptr, nullptr, rg);    \
        count_launch();                                                                             \
        PDM_CHECK_LAUNCH("stack_farthest_point_sampling(bucket)");                                  \
        cap = L::CAP;                                                                               \
    }
        // capacity from the total row count (an upper bound of every frame): small stacks get the small kernels
        if (n_total <= 2048) PDM_STACK_BUCKET(16, 4, 512)
        else if (n_total <= 4096) PDM_STACK_BUCKET(16, 8, 512)
        else if (n_total <= 8192) PDM_STACK_BUCKET(16, 16, 512)
        else PDM_STACK_BUCKET(16, 32, 544)
#undef PDM_STACK_BUCKET
    }
    if (n_total > cap) {    // some frame may be larger than the on-chip capacity: the any-size kernel takes those
        rg.skip_le = cap;
        fps_generic_kernel<1024><<<batch, 1024, 0, st>>>(0, 0, p, xyz, temp, idx, rg);
        count_launch();
        PDM_CHECK_LAUNCH("stack_farthest_point_sampling(generic)");
    }
    return PDM_OK;
}
}  // namespace pdm

extern "C" int pdm_stack_farthest_point_sampling(int n_total, int batch, const float *xyz, float *temp,
                                                 const int *xyz_batch_cnt, int *idx, const int *num_sampled_points,
                                                 void *stream) {
    using namespace pdm;
    if (n_total < 0 || batch < 0) return fail(PDM_ERR_INVALID_ARG, "stack_farthest_point_sampling: negative size");
    if (batch == 0 || n_total == 0) return PDM_OK;
    if (!xyz || !temp || !idx || !xyz_batch_cnt || !num_sampled_points)
        return fail(PDM_ERR_INVALID_ARG, "stack_farthest_point_sampling: null pointer");
    if ((long long)n_total * 3 > 0x7fffffffLL) return fail(PDM_ERR_UNSUPPORTED, "stack_farthest_point_sampling: too many rows");
    return fps_stack_dispatch(n_total, batch, xyz, temp, xyz_batch_cnt, idx, num_sampled_points, (cudaStream_t)stream);
}

// Debug-only entry: timestamp trace of frame 0 (16384-point frames, <16,32,8> kernel).
extern "C" int pdm_debug_fps_trace(int b, int n, int m, const float *xyz, float *temp, int *idx,
                                   int *stats, long long *trace, void *stream) {
    using namespace pdm;
    if (n <= 8192 || n > 16384) return fail(PDM_ERR_UNSUPPORTED, "debug_fps_trace: n=%d", n);
    int p = 0;
    while ((1 << p) < ref_fps_block_size(n)) ++p;
    using L = FpsSmem<16, 32, 8>;
    auto kern = fps_bucket_kernel<16, 32, 8, true>;
    if (int rc = ensure_dynamic_smem((const void *)kern, L::kBytes)) return rc;
    kern<<<b, 512, L::kBytes, (cudaStream_t)stream>>>(n, m, p, xyz, temp, idx, stats, trace, FpsRagged{});
    PDM_CHECK_LAUNCH("debug_fps_trace");
    return PDM_OK;
}

// Debug-only entry (not part of include/pdm_ops.h): like pdm_farthest_point_sampling, and also
// writes the number of barrier rounds each frame needed into stats[b] (device ints).
extern "C" int pdm_debug_fps_rounds(int b, int n, int m, const float *xyz, float *temp, int *idx,
                                    int *stats, void *stream) {
    return pdm::fps_dispatch(b, n, m, xyz, temp, idx, stats, (cudaStream_t)stream);
}

extern "C" int pdm_set_fps_mode(int mode) {
    if (mode < PDM_FPS_MODE_AUTO || mode > PDM_FPS_MODE_THROUGHPUT)
        return pdm::fail(PDM_ERR_INVALID_ARG, "set_fps_mode: unknown mode %d", mode);
    pdm::g_fps_mode.store(mode);
    return PDM_OK;
}

static int fps_entry(int b, int n, int m, const float *xyz, float *temp, int *idx, int mode, void *stream);

extern "C" int pdm_farthest_point_sampling(int b, int n, int m, const float *xyz, float *temp,
                                           int *idx, void *stream) {
    return fps_entry(b, n, m, xyz, temp, idx, -1, stream);
}

extern "C" int pdm_farthest_point_sampling_ex(int b, int n, int m, const float *xyz, float *temp,
                                              int *idx, int mode, void *stream) {
    if (mode < PDM_FPS_MODE_AUTO || mode > PDM_FPS_MODE_THROUGHPUT)
        return pdm::fail(PDM_ERR_INVALID_ARG, "farthest_point_sampling_ex: unknown mode %d", mode);
    return fps_entry(b, n, m, xyz, temp, idx, mode, stream);
}

static int fps_entry(int b, int n, int m, const float *xyz, float *temp, int *idx, int mode, void *stream) {
    using namespace pdm;
    if (b < 0 || n < 0 || m < 0) return fail(PDM_ERR_INVALID_ARG, "farthest_point_sampling: negative size");
    if (b == 0 || m == 0) return PDM_OK;
    if (n == 0) return fail(PDM_ERR_INVALID_ARG, "farthest_point_sampling: n == 0 with m > 0");
    if (!xyz || !temp || !idx) return fail(PDM_ERR_INVALID_ARG, "farthest_point_sampling: null pointer");
    if ((long long)n * 3 > 0x7fffffffLL) return fail(PDM_ERR_UNSUPPORTED, "farthest_point_sampling: n too large");
    return fps_dispatch(b, n, m, xyz, temp, idx, nullptr, (cudaStream_t)stream, mode);
}
