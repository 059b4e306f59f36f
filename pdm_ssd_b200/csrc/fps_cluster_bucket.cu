// fps_cluster_bucket.cu -- farthest point sampling for frames larger than one SM (16384 < n <= 196608, the Waymo-scale
// frames of BASELINE configs[4]) with EXACT BUCKET PRUNING and MULTI-SAMPLE ROUNDS on a thread-block cluster.
//
// fps_cluster.cu gives a frame a cluster of CTAs but still sweeps all of its points every round (2.4 us per sample at
// 163840 points).  Here every CTA of the cluster runs the bucket algorithm of fps.cu on its chunk of the frame:
//   * the chunk is sorted once along a space-filling curve and cut into buckets of 32 points (one per lane, bucket b of
//     warp b mod 16) with a bounding box and a running maximum; a new sample only touches the buckets whose box
//     lower bound is below their maximum (exact: fp32 rounding is monotone, fps.cu header);
//   * running minima live in registers, coordinates and the position -> frame-index map in shared memory (16 B/point);
//   * per round every warp publishes two exact candidates + a bound on all its other points; warp 0 reduces the CTA's
//     32 entries to the CTA's two best candidates + a bound on everything else in the CTA and writes that record
//     (coordinates, value, tiekey) into EVERY CTA of the cluster through distributed shared memory; after ONE cluster
//     barrier every warp of every CTA replays the sequential selection on the 2 x cluster-size candidates: pick 1 is the
//     exact argmax with the reference's tie-break, further picks are accepted while the largest updated candidate is
//     unique and strictly above the bound (so it is the unique global maximum, fps.cu "multi-sample rounds").
// All CTAs replay the same selection, so they agree on the samples and on the number of rounds without further traffic.
// Exactly the reference's sample sequence and final `temp` (bit-exact, ties included).
#include <cooperative_groups.h>
#include <stdlib.h>

#include <map>
#include <mutex>

#include "fps_common.cuh"

namespace cg = cooperative_groups;

namespace pdm {

constexpr int kCbMaxCl = 16;
constexpr int kCbNW = 16;
constexpr int kCbT = kCbNW * 32;
constexpr int cb_pow2(int v) { return v <= 1 ? 1 : 2 * cb_pow2((v + 1) / 2); }   // keys sorted per thread: next power of two >= BPW

// SMAP: the position -> frame-index map lives in shared memory (16 B/point: 12288 points per CTA); otherwise in the
// caller's scratch `temp` (12 B/point on chip: 16384 points per CTA, so a 163840-point frame fits a cluster of 10 and a
// batch of 8 such frames is resident at once -- this B200 keeps 11 clusters of 10 CTAs but only 7 of 11..16).
template <int BPW, int KMAX, bool SMAP>
struct CbSmem {
    static constexpr int CAP = kCbNW * BPW * 32;
    static constexpr int kSortE = cb_pow2(BPW) < 2 ? 2 : cb_pow2(BPW);               // keys sorted per thread (padding sorts last)
    static_assert((size_t)CAP * 12 >= (size_t)kSortE * kCbT * 4, "the sort scratch aliases the coordinate arrays");
    static constexpr size_t kMapOff = (size_t)12 * CAP;                          // pmap[CAP]: frame index of a sorted position
    static constexpr size_t kPubOff = (size_t)(SMAP ? 16 : 12) * CAP;            // pub[2][2 NW] uint2 (value bits, position)
    static constexpr size_t kUOff = kPubOff + sizeof(uint2) * 2 * 2 * kCbNW;     // pubU[2][NW]
    static constexpr size_t kSampOff = ((kUOff + sizeof(unsigned) * 2 * kCbNW + 15) / 16) * 16;   // samp[NW][KMAX] float4
    static constexpr size_t kBoxOff = kSampOff + sizeof(float4) * kCbNW * KMAX;  // box scratch [6 NW]
    static constexpr size_t kCpubOff = ((kBoxOff + sizeof(float) * 6 * kCbNW + 15) / 16) * 16;   // cpub[2][16][3] float4
    static constexpr size_t kBytes = kCpubOff + sizeof(float4) * 2 * kCbMaxCl * 3;
};

template <int BPW, int KMAX, bool SMAP>
__global__ void __launch_bounds__(kCbT, 1)
fps_cluster_bucket_kernel(int n, int m, int p, int chunk /*points per CTA*/, const float *__restrict__ xyz,
                          float *__restrict__ temp, int *__restrict__ idxs, int *__restrict__ stats) {
    using L = CbSmem<BPW, KMAX, SMAP>;
    constexpr int CAP = L::CAP, NW = kCbNW, T = kCbT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sx = reinterpret_cast<float *>(smem_raw);
    float *sy = sx + CAP;
    float *sz = sy + CAP;
    unsigned *smap = reinterpret_cast<unsigned *>(smem_raw + L::kMapOff);      // SMAP only
    uint2 *pub = reinterpret_cast<uint2 *>(smem_raw + L::kPubOff);
    unsigned *pubU = reinterpret_cast<unsigned *>(smem_raw + L::kUOff);
    float4 *samp = reinterpret_cast<float4 *>(smem_raw + L::kSampOff);
    float *box = reinterpret_cast<float *>(smem_raw + L::kBoxOff);
    float4 *cpub = reinterpret_cast<float4 *>(smem_raw + L::kCpubOff);

    cg::cluster_group cluster = cg::this_cluster();
    const int cl = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int frame = blockIdx.x / cl;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const unsigned bsmask = (1u << p) - 1u;
    const float *frame_xyz = xyz + (size_t)frame * n * 3;
    float *tmp = temp + (size_t)frame * n;
    int *out = idxs + (size_t)frame * m;
    const int k0 = rank * chunk;
    const int cnt = max(0, min(chunk, n - k0));       // my points: frame indices [k0, k0 + cnt)
    const float *dataset = frame_xyz + (size_t)k0 * 3;
    unsigned *gmap = reinterpret_cast<unsigned *>(tmp + k0);       // !SMAP: my part of the scratch doubles as the map
    auto map_get = [&](unsigned pos) -> unsigned { if constexpr (SMAP) return smap[pos]; else return gmap[pos]; };

    if (rank == 0 && tid == 0) out[0] = 0;
    if (m <= 1) return;                                // (every CTA of the cluster takes this exit together)

    // ---- 1. chunk bounding box ------------------------------------------------------------------
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = tid; k < cnt; k += T) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = __ldg(dataset + (size_t)k * 3 + a);
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
    FpsCurve curve;
    curve.init(lo, hi);

    // ---- 2. sort the chunk along the curve: E keys per thread (18-bit code | 14-bit local index), padding last
    unsigned *skeys = reinterpret_cast<unsigned *>(smem_raw);
    {
        constexpr int E = L::kSortE;
        unsigned v[E];
#pragma unroll
        for (int r = 0; r < E; ++r) {
            const int k = r * T + tid;
            v[r] = 0xffffffffu;
            if (k < cnt) {
                const float c[3] = {__ldg(dataset + (size_t)k * 3 + 0), __ldg(dataset + (size_t)k * 3 + 1), __ldg(dataset + (size_t)k * 3 + 2)};
                v[r] = (curve.code18(c) << 14) | (unsigned)k;
            }
        }
        fps_sort_keys<E, T>(v, skeys, tid);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < E; ++r) skeys[tid * E + r] = v[r];      // sorted position e = tid * E + r
    }
    __syncthreads();

    // ---- 3. distribute: lane owns slot `lane` of buckets b = j * NW + w; sorted position pos = b * 32 + lane
    //         (cnt <= chunk <= CAP, and padding sorts last: every real point has pos < cnt <= CAP)
    float t[BPW];
    unsigned kk_[BPW];
#pragma unroll
    for (int j = 0; j < BPW; ++j) kk_[j] = skeys[((j * NW + w) << 5) + lane] & 0x3fffu;
    __syncthreads();   // keys are dead from here on; the region becomes sx / sy / sz / pmap
#pragma unroll
    for (int j = 0; j < BPW; ++j) {
        const int pos = ((j * NW + w) << 5) + lane;
        t[j] = pos < cnt ? tmp[k0 + kk_[j]] : 0.f;     // padding: 0 and never the tie winner
    }
    if constexpr (!SMAP) __syncthreads();              // every initial value is read before the map overwrites the scratch
    float blox = INFINITY, bloy = INFINITY, bloz = INFINITY;
    float bhix = -INFINITY, bhiy = -INFINITY, bhiz = -INFINITY;
    unsigned bmax = 0u, bwl = 0u, bsec = 0u;
#pragma unroll
    for (int j = 0; j < BPW; ++j) {
        const int pos = ((j * NW + w) << 5) + lane;
        const bool pad = pos >= cnt;
        float x = 0.f, y = 0.f, z = 0.f;
        if (!pad) {
            x = __ldg(dataset + (size_t)kk_[j] * 3 + 0);
            y = __ldg(dataset + (size_t)kk_[j] * 3 + 1);
            z = __ldg(dataset + (size_t)kk_[j] * 3 + 2);
        }
        if constexpr (SMAP) smap[pos] = pad ? 0u : (unsigned)(k0 + kk_[j]);
        else if (!pad) gmap[pos] = (unsigned)(k0 + kk_[j]);
        sx[pos] = x;
        sy[pos] = y;
        sz[pos] = z;
        const bool ox = pad || x != x, oy = pad || y != y, oz = pad || z != z;
        const unsigned lx = __reduce_min_sync(kFull, ox ? 0xffffffffu : f2ord(x));
        const unsigned ly = __reduce_min_sync(kFull, oy ? 0xffffffffu : f2ord(y));
        const unsigned lz = __reduce_min_sync(kFull, oz ? 0xffffffffu : f2ord(z));
        const unsigned hx = __reduce_max_sync(kFull, ox ? 0u : f2ord(x));
        const unsigned hy = __reduce_max_sync(kFull, oy ? 0u : f2ord(y));
        const unsigned hz = __reduce_max_sync(kFull, oz ? 0u : f2ord(z));
        const unsigned tb = __float_as_uint(t[j]);
        const unsigned mx = __reduce_max_sync(kFull, tb);
        const unsigned cand = (tb == mx && !pad) ? fps_tiekey((unsigned)(k0 + kk_[j]), p, bsmask) : kPadKey;
        const unsigned tkm = __reduce_min_sync(kFull, cand);
        const unsigned wl = __ffs(__ballot_sync(kFull, cand == tkm)) - 1;
        const unsigned sec = __reduce_max_sync(kFull, lane == (int)wl ? 0u : tb);
        if (lane == j) {
            blox = ord2f(lx); bloy = ord2f(ly); bloz = ord2f(lz);
            bhix = ord2f(hx); bhiy = ord2f(hy); bhiz = ord2f(hz);
            bmax = mx; bwl = wl; bsec = sec;
        }
    }
    if (lane == 0) samp[w * KMAX] = make_float4(__ldg(frame_xyz + 0), __ldg(frame_xyz + 1), __ldg(frame_xyz + 2), 0.f);
    __syncthreads();
    cluster.sync();    // every CTA of the cluster is resident before anyone writes into its shared memory

    // ---- 4. rounds ----------------------------------------------------------------------------------
    const int wbase = (w << 5) + lane;
    const unsigned bbase = (unsigned)(lane * (NW * 32) + (w << 5));
    auto tiekey_at = [&](unsigned pos) -> unsigned {
        return pos < (unsigned)cnt ? fps_tiekey(map_get(pos), p, bsmask) : kPadKey;
    };
    unsigned c1v = 0u, c1p = 0u, c2v = 0u, c2p = 0u, wU = 0u;
    bool dirty = true;
    int K = 1;
    int j = 1;
    int rounds = 0;
    for (;;) {
        // A. which of my buckets can change?  (fps.cu phase A)
        const float4 *ws = samp + w * KMAX;
        unsigned amask = 0u;
        {
            const float bm = __uint_as_float(bmax);
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
        // B. update the surviving buckets (fps.cu phase B)
        while (mask) {
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
            unsigned sec = __reduce_max_sync(kFull, hit ? 0u : tb);
            if (multi_bit(__ballot_sync(kFull, hit))) {
                const unsigned cand = hit ? tiekey_at(pos) : kPadKey;
                const unsigned tkm = __reduce_min_sync(kFull, cand);
                wl = __reduce_max_sync(kFull, cand == tkm ? (unsigned)lane : 0u);
                sec = mx;
            }
            if (lane == jj) { bmax = mx; bwl = wl; bsec = sec; }
            reg_store<BPW>(t, jj, nt);
            dirty = true;
        }
        if (j >= m) break;
        ++rounds;
        // C. this warp's two candidates and the bound on everything else it owns (fps.cu phase C)
        if (dirty) {
            dirty = false;
            const unsigned v = lane < BPW ? bmax : 0u;
            c1v = __reduce_max_sync(kFull, v);
            const bool hit1 = lane < BPW && v == c1v;
            unsigned src1 = __reduce_max_sync(kFull, hit1 ? (unsigned)lane : 0u);
            c2v = __reduce_max_sync(kFull, hit1 ? 0u : v);
            if (multi_bit(__ballot_sync(kFull, hit1))) {
                const unsigned cc = hit1 ? tiekey_at(bbase + bwl) : kPadKey;
                const unsigned tkm = __reduce_min_sync(kFull, cc);
                src1 = __reduce_max_sync(kFull, (hit1 && cc == tkm) ? (unsigned)lane : 0u);
                c2v = c1v;
            }
            const bool hit2 = lane < BPW && lane != (int)src1 && v == c2v;
            const unsigned src2 = __reduce_max_sync(kFull, hit2 ? (unsigned)lane : 0u);
            const bool has2 = __ballot_sync(kFull, hit2) != 0u;
            const bool mine = lane == (int)src1 || (has2 && lane == (int)src2);
            wU = __reduce_max_sync(kFull, mine ? bsec : v);
            c1p = __shfl_sync(kFull, bbase + bwl, src1);
            c2p = __shfl_sync(kFull, bbase + bwl, src2);
            if (!has2) { c2v = 0u; c2p = c1p; }
        }
        const int par = (rounds & 1);
        if (lane == 0) {
            pub[par * 2 * NW + 2 * w] = make_uint2(c1v, c1p);
            pub[par * 2 * NW + 2 * w + 1] = make_uint2(c2v, c2p);
            pubU[par * NW + w] = wU;
        }
        __syncthreads();
        // D1. warp 0: the CTA's two best candidates + a bound on everything else in the CTA, written into every CTA
        float4 *mycp = cpub + (size_t)par * kCbMaxCl * 3;
        if (w == 0) {
            const uint2 e = pub[par * 2 * NW + lane];            // 2 NW = 32 entries, one per lane
            const bool first = !(lane & 1);
            const unsigned gm = __reduce_max_sync(kFull, first ? e.x : 0u);
            bool h1 = first && e.x == gm;
            const unsigned mytk = tiekey_at(e.y);
            if (multi_bit(__ballot_sync(kFull, h1))) {
                const unsigned tkm = __reduce_min_sync(kFull, h1 ? mytk : kPadKey);
                h1 = h1 && mytk == tkm;
            }
            const int l1 = __ffs(__ballot_sync(kFull, h1)) - 1;
            const unsigned v2 = __reduce_max_sync(kFull, lane == l1 ? 0u : e.x);
            const int l2 = 31 - __clz(__ballot_sync(kFull, lane != l1 && e.x == v2));
            const unsigned rest = __reduce_max_sync(kFull, (lane == l1 || lane == l2) ? 0u : e.x);
            const unsigned uw = __reduce_max_sync(kFull, lane < NW ? pubU[par * NW + lane] : 0u);
            const float x = sx[e.y], y = sy[e.y], z = sz[e.y];
            const float4 ra = make_float4(__shfl_sync(kFull, x, l1), __shfl_sync(kFull, y, l1), __shfl_sync(kFull, z, l1), __uint_as_float(gm));
            const float4 rb = make_float4(__shfl_sync(kFull, x, l2), __shfl_sync(kFull, y, l2), __shfl_sync(kFull, z, l2), __uint_as_float(v2));
            const float4 rc = make_float4(__uint_as_float(__shfl_sync(kFull, mytk, l1)), __uint_as_float(__shfl_sync(kFull, mytk, l2)),
                                          __uint_as_float(max(rest, uw)), 0.f);
            if (lane < cl) {
                float4 *dst = cluster.map_shared_rank(mycp + rank * 3, lane);
                dst[0] = ra;
                dst[1] = rb;
                dst[2] = rc;
            }
        }
        // one cluster barrier per round; only the publishing warp needs release semantics
        if (w == 0) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        else asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        // D2. every warp replays the sequential selection on the 2 * cl candidates (one per lane)
        const bool live = lane < 2 * cl;
        const int which = lane & 1;
        const float4 rec = live ? mycp[(lane >> 1) * 3 + which] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 aux = live ? mycp[(lane >> 1) * 3 + 2] : make_float4(0.f, 0.f, 0.f, 0.f);
        const unsigned tkc = live ? __float_as_uint(which ? aux.y : aux.x) : kPadKey;
        const unsigned U = __reduce_max_sync(kFull, (live && !which) ? __float_as_uint(aux.z) : 0u);
        const float x = rec.x, y = rec.y, z = rec.z;
        float v = rec.w;
        const unsigned vb0 = live ? __float_as_uint(v) : 0u;
        const bool first = live && !which;
        const unsigned gm = __reduce_max_sync(kFull, first ? vb0 : 0u);
        bool ghit = first && vb0 == gm;
        if (multi_bit(__ballot_sync(kFull, ghit))) {
            const unsigned gtk = __reduce_min_sync(kFull, ghit ? tkc : kPadKey);
            ghit = ghit && tkc == gtk;
        }
        const int kmax_now = min(KMAX, m - j);
        K = 0;
        for (;;) {
            if (ghit) {
                samp[w * KMAX + K] = make_float4(x, y, z, 0.f);
                if (rank == 0 && w == 0) out[j + K] = (int)fps_tiekey_inv(tkc, p, bsmask);
            }
            ++K;
            if (K >= kmax_now) break;
            __syncwarp();
            const float4 s4 = samp[w * KMAX + K - 1];
            v = fminf(sqdist_ref(__fsub_rn(x, s4.x), __fsub_rn(y, s4.y), __fsub_rn(z, s4.z)), v);
            const unsigned vb = live ? __float_as_uint(v) : 0u;
            const unsigned g2 = __reduce_max_sync(kFull, vb);
            if (!(g2 > U)) break;
            ghit = live && vb == g2;
            if (multi_bit(__ballot_sync(kFull, ghit))) break;
        }
        __syncwarp();
        j += K;
        if (j >= m) {   // the very last sample is never applied (the reference stops after writing it)
            --K;
            if (K == 0) break;
        }
    }
    if (stats && tid == 0 && rank == 0) stats[frame] = rounds;

    // ---- 5. leave temp as the reference does: running minima in original order
    unsigned ko[BPW];
#pragma unroll
    for (int jq = 0; jq < BPW; ++jq) {
        const int pos = ((jq * NW + w) << 5) + lane;
        ko[jq] = pos < cnt ? map_get(pos) : 0u;
    }
    if constexpr (!SMAP) __syncthreads();              // all map entries are read before any of them is overwritten
#pragma unroll
    for (int jq = 0; jq < BPW; ++jq) {
        const int pos = ((jq * NW + w) << 5) + lane;
        if (pos < cnt) tmp[ko[jq]] = t[jq];
    }
    cluster.sync();   // nobody exits while a peer may still address its shared memory
}

constexpr int kCbKMAX = 8;
using CbKern = void (*)(int, int, int, int, const float *, float *, int *, int *);
struct CbVariant {
    CbKern kern;
    int cap;          // points per CTA
    size_t smem;
    bool smap;        // index map on chip
};
// 16 warps x BPW buckets x 32 points per CTA.  The small ones serve frames of <= 16384 points spread over a cluster
// (latency mode of the KITTI-sized layers, PDM_FPS_KERNEL=cb); 24: map on chip, 32: map in the caller's scratch.
static const CbVariant kCbVariants[] = {
    {fps_cluster_bucket_kernel<4, kCbKMAX, true>, CbSmem<4, kCbKMAX, true>::CAP, CbSmem<4, kCbKMAX, true>::kBytes, true},
    {fps_cluster_bucket_kernel<8, kCbKMAX, true>, CbSmem<8, kCbKMAX, true>::CAP, CbSmem<8, kCbKMAX, true>::kBytes, true},
    {fps_cluster_bucket_kernel<24, kCbKMAX, true>, CbSmem<24, kCbKMAX, true>::CAP, CbSmem<24, kCbKMAX, true>::kBytes, true},
    {fps_cluster_bucket_kernel<32, kCbKMAX, false>, CbSmem<32, kCbKMAX, false>::CAP, CbSmem<32, kCbKMAX, false>::kBytes, false},
};
constexpr int kCbNumVariants = 4;

bool fps_cluster_bucket_supports(int n) { return n > 16384 && n <= kCbMaxCl * kCbVariants[kCbNumVariants - 1].cap; }

static int cb_max_active_clusters(int variant, int cl) {
    static std::mutex mu;
    static std::map<int, int> cache;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    std::lock_guard<std::mutex> lock(mu);
    const int key = (dev * 64 + cl) * 8 + variant;
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    const CbVariant &V = kCbVariants[variant];
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)cl);
    cfg.blockDim = dim3(kCbT);
    cfg.dynamicSmemBytes = V.smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int v = 0;
    if (cl > 8 && cudaFuncSetAttribute((const void *)V.kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
        (void)cudaGetLastError();
    if (cudaOccupancyMaxActiveClusters(&v, (const void *)V.kern, &cfg) != cudaSuccess) {
        (void)cudaGetLastError();
        v = 0;
    }
    cache[key] = v;
    return v;
}

// Returns PDM_ERR_UNSUPPORTED (no error recorded) when the shape is out of range or the device cannot co-schedule a
// cluster of the needed size; the caller then uses fps_cluster_launch / the any-size kernel.  `small_ok`: also take
// frames of <= 16384 points (one CTA would do; the cluster shortens a round).
int fps_cluster_bucket_launch(int b, int n, int m, int p, const float *xyz, float *temp, int *idx, int *stats, cudaStream_t st,
                              bool small_ok) {
    if (!(fps_cluster_bucket_supports(n) || (small_ok && n >= 2048 && n <= 16384))) return PDM_ERR_UNSUPPORTED;
    // (variant, cluster size): fewest waves first (a batch resident in ONE wave halves the time), then the on-chip map,
    // then the most CTAs per frame (fewer buckets per warp).  PDM_FPS_CLUSTER=<size> / PDM_FPS_CLUSTER_MAP=smem|global
    // force a choice (testing).
    const char *fe = getenv("PDM_FPS_CLUSTER"), *me = getenv("PDM_FPS_CLUSTER_MAP");
    const int forced = fe ? atoi(fe) : 0;
    const bool dbg = getenv("PDM_DEBUG_CLUSTER") != nullptr;
    int cl = 0, cap = 0, var = -1, best_waves = 1 << 30;
    for (int pass = 0; pass < 2; ++pass) {          // on-chip map first
        for (int v = 0; v < kCbNumVariants; ++v) {
            const CbVariant &V = kCbVariants[v];
            if (V.smap != (pass == 0)) continue;
            if (me && ((me[0] == 's') != V.smap)) continue;
            for (int c = kCbMaxCl; c >= 2; --c) {
                if (forced && c != forced) continue;
                const int chunk = ((n + c - 1) / c + 31) / 32 * 32;
                if (chunk > V.cap) break;
                if (v > 0 && chunk <= kCbVariants[v - 1].cap && kCbVariants[v - 1].smap == V.smap) continue;   // a smaller variant holds it
                if ((long long)(c - 1) * chunk >= n) continue;      // the last CTA would be empty
                if (ensure_dynamic_smem((const void *)V.kern, V.smem) != PDM_OK) continue;
                const int act = cb_max_active_clusters(v, c);
                if (dbg) fprintf(stderr, "[pdm]   cluster-bucket variant %d of %d: %d points per CTA, %d active clusters\n", v, c, chunk, act);
                if (act < 1) continue;
                const int waves = (b + act - 1) / act;
                if (waves < best_waves) { best_waves = waves; cl = c; cap = chunk; var = v; }
            }
        }
    }
    if (cl == 0) return PDM_ERR_UNSUPPORTED;
    const CbVariant &V = kCbVariants[var];
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(b * cl));
    cfg.blockDim = dim3(kCbT);
    cfg.dynamicSmemBytes = V.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (dbg) fprintf(stderr, "[pdm] fps cluster-bucket: b=%d n=%d variant=%d cl=%d chunk=%d waves=%d\n", b, n, var, cl, cap, best_waves);
    const cudaError_t e = cudaLaunchKernelEx(&cfg, V.kern, n, m, p, cap, xyz, temp, idx, stats);
    if (e != cudaSuccess) return fail((int)e, "farthest_point_sampling(cluster-bucket of %d): %s", cl, cudaGetErrorString(e));
    count_launch();
    return PDM_OK;
}

}  // namespace pdm
