// fps_l2.cu -- farthest point sampling, throughput-oriented variant (bit-exact, same algorithm as
// fps_bucket_kernel in fps.cu; read that file's header first).
//
// Why a second kernel.  fps_bucket_kernel keeps the frame's coordinates in shared memory (12 B/point:
// 197 KB for 16384 points) and the running minima in registers (128 or 96 per thread), so ONE frame
// owns an SM for its whole 2.3 ms -- and a round is a chain of dependent latencies that uses ~15 % of
// the SM's issue slots.  A batch pipeline (12+ batches in flight) is then bound by SMs x frame time:
// 16.2 us per frame whatever else is tuned.  Here the per-point state that stays on chip is only the
// running minimum (4 B/point in shared memory); the sorted coordinates live in a global scratch
// buffer that stays resident in the 126 MB L2 (a surviving bucket costs one coalesced 3 x 128 B read,
// ~250 cycles, issued for all surviving buckets of the round before the first one is consumed).
// Shared memory drops to ~70 KB and registers to <= 64 per thread, so TWO OR THREE frames share an SM
// and fill each other's latency bubbles.
//
// Launch sequence per call:
//   fps_prepare_kernel  (CTA per frame): bounding box, space-filling-curve sort (in shared memory),
//                        writes sorted SoA coordinates, sorted initial minima, and the
//                        sorted-position -> original-index map to scratch;
//   fps_l2_kernel       (CTA per frame): the rounds; at exit scatters the running minima back into the
//                        caller's `temp` in original order (what the reference leaves there).
#include <stdlib.h>

#include <type_traits>

#include "fps_common.cuh"

namespace pdm {

// ---------------------------------------------------------------------------------------------------
// prepare
// ---------------------------------------------------------------------------------------------------
// Sort of CAP 32-bit keys (18-bit curve code | 14-bit point index), 16 keys per thread in REGISTERS:
// a bitonic network whose compare-exchange distance j is handled where the partner lives --
//   j < 16        inside the thread (register pairs),
//   16 <= j < 512 in another lane of the warp (shfl.xor),
//   j >= 512      in another warp: one transposed round trip through shared memory (conflict-free).
// Of the 105 stages of a 16384-key sort only 15 touch shared memory (the first version kept 64-bit
// keys in shared memory and paid two barriers and 256 KB of shared-memory traffic for every stage:
// 247 us per frame, a fifth of the SM-time of sampling in throughput mode).
template <int CAP>
__global__ void __launch_bounds__(CAP / 16)
fps_prepare_kernel(int n, const float *__restrict__ xyz, const float *__restrict__ temp, float *__restrict__ sorted,
                   float *__restrict__ tinit, unsigned *__restrict__ pmap) {
    constexpr int E = 16, T = CAP / E, NWARP = T / 32;
    static_assert(CAP >= 1024 && CAP <= 16384, "14-bit point index, at least two warps");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned *xch = reinterpret_cast<unsigned *>(smem_raw);   // [E][T], transposed exchange buffer
    __shared__ float box[NWARP * 6];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const float *dataset = xyz + (size_t)blockIdx.x * n * 3;
    const float *tmp = temp + (size_t)blockIdx.x * n;
    float *gx = sorted + (size_t)blockIdx.x * 3 * CAP, *gy = gx + CAP, *gz = gy + CAP;
    float *ti = tinit + (size_t)blockIdx.x * CAP;
    unsigned *pm = pmap + (size_t)blockIdx.x * CAP;

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
        lo[a] = ord2f(__reduce_min_sync(kFull, f2ord(lane < NWARP ? box[lane * 6 + a] : INFINITY)));
        hi[a] = ord2f(__reduce_max_sync(kFull, f2ord(lane < NWARP ? box[lane * 6 + 3 + a] : -INFINITY)));
    }
    FpsCurve curve;
    curve.init(lo, hi);
    // keys: any initial arrangement will do, so thread t takes points r * T + t (coalesced reads)
    unsigned v[E];
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const int k = r * T + tid;
        v[r] = 0xffffffffu;   // padding sorts last (a real key cannot be all ones while padding exists)
        if (k < n) {
            const float c[3] = {__ldg(dataset + k * 3 + 0), __ldg(dataset + k * 3 + 1), __ldg(dataset + k * 3 + 2)};
            v[r] = (curve.code18(c) << 14) | (unsigned)k;
        }
    }
    fps_sort_keys<E, T>(v, xch, tid);
    // thread t now holds sorted positions t*16 .. t*16+15 (padding keys last: position >= n <=> padding)
#pragma unroll
    for (int r0 = 0; r0 < E; r0 += 4) {     // four positions at a time: 16-byte stores, few live registers
        float ox[4], oy[4], oz[4], ot[4];
        unsigned op[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int pos = tid * E + r0 + q;
            const unsigned k = v[r0 + q] & 0x3fffu;
            const bool pad = pos >= n;
            ox[q] = pad ? 0.f : __ldg(dataset + k * 3 + 0);
            oy[q] = pad ? 0.f : __ldg(dataset + k * 3 + 1);
            oz[q] = pad ? 0.f : __ldg(dataset + k * 3 + 2);
            ot[q] = pad ? 0.f : __ldg(tmp + k);   // padding: 0 and never the tie winner
            op[q] = pad ? 0xffffffffu : k;
        }
        const int pos = tid * E + r0;
        *reinterpret_cast<float4 *>(gx + pos) = make_float4(ox[0], ox[1], ox[2], ox[3]);
        *reinterpret_cast<float4 *>(gy + pos) = make_float4(oy[0], oy[1], oy[2], oy[3]);
        *reinterpret_cast<float4 *>(gz + pos) = make_float4(oz[0], oz[1], oz[2], oz[3]);
        *reinterpret_cast<float4 *>(ti + pos) = make_float4(ot[0], ot[1], ot[2], ot[3]);
        *reinterpret_cast<uint4 *>(pm + pos) = make_uint4(op[0], op[1], op[2], op[3]);
    }
}

// ---------------------------------------------------------------------------------------------------
// rounds
// ---------------------------------------------------------------------------------------------------
template <int NW, int BPW, int KMAX>
struct FpsL2Smem {
    static constexpr int CAP = NW * BPW * 32;
    // ts[CAP] running minima; pubA[2][2*NW] uint4 (value bits, position, x bits, y bits) and
    // pubZ[2][2*NW] -- two candidates per warp with their coordinates; pubU[2][NW] bound on every other
    // point of the warp; samp[KMAX] float4 accepted samples + their count (written by warp 0)
    static constexpr size_t kPubOff = (size_t)4 * CAP;
    static constexpr size_t kPubZOff = kPubOff + sizeof(uint4) * 2 * 2 * NW;
    static constexpr size_t kUOff = kPubZOff + sizeof(float) * 2 * 2 * NW;
    static constexpr size_t kSampOff = ((kUOff + sizeof(unsigned) * 2 * NW + 15) / 16) * 16;
    static constexpr size_t kKOff = kSampOff + sizeof(float4) * KMAX;
    static constexpr size_t kBytes = kKOff + 16;
};

template <int NW, int BPW, int KMAX, int OCC>
__global__ void __launch_bounds__(NW * 32, OCC)
fps_l2_kernel(int n, int m, int p, const float *__restrict__ xyz, const float *__restrict__ sorted,
              const float *__restrict__ tinit, const unsigned *__restrict__ pmap, float *__restrict__ temp,
              int *__restrict__ idxs, int *__restrict__ stats) {
    using L = FpsL2Smem<NW, BPW, KMAX>;
    constexpr int CAP = L::CAP;
    static_assert(BPW <= 32 && NW <= 16, "one lane per owned bucket; two candidates per warp in one warp");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *ts = reinterpret_cast<float *>(smem_raw);
    uint4 *pubA = reinterpret_cast<uint4 *>(smem_raw + L::kPubOff);
    float *pubZ = reinterpret_cast<float *>(smem_raw + L::kPubZOff);
    unsigned *pubU = reinterpret_cast<unsigned *>(smem_raw + L::kUOff);
    float4 *samp = reinterpret_cast<float4 *>(smem_raw + L::kSampOff);
    int *ksh = reinterpret_cast<int *>(smem_raw + L::kKOff);

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const unsigned bsmask = (1u << p) - 1u;
    const float *gx = sorted + (size_t)blockIdx.x * 3 * CAP, *gy = gx + CAP, *gz = gy + CAP;
    const unsigned *pm = pmap + (size_t)blockIdx.x * CAP;
    int *out = idxs + (size_t)blockIdx.x * m;

    if (tid == 0) out[0] = 0;
    if (m <= 1) return;   // temp untouched: the reference's loop body never runs either

    // tiekey of the point at sorted position pos (global read: set-up and slow paths only)
    auto tiekey_at = [&](unsigned pos) -> unsigned {
        const unsigned k = pos < (unsigned)CAP ? __ldg(pm + pos) : 0xffffffffu;
        return k != 0xffffffffu ? fps_tiekey(k, p, bsmask) : kPadKey;
    };

    // ---- set-up: lane j of warp w holds the state of bucket  b = j*NW + w  (32 consecutive sorted points)
    float blox = INFINITY, bloy = INFINITY, bloz = INFINITY;
    float bhix = -INFINITY, bhiy = -INFINITY, bhiz = -INFINITY;
    unsigned bmax = 0u, bsec = 0u;          // max (bits) and runner-up (bits) of the bucket's minima
    unsigned bwl = 0u;                      // lane holding the max
    float bwx = 0.f, bwy = 0.f, bwz = 0.f;  // ... and its coordinates (published with the candidate)
#pragma unroll 4
    for (int j = 0; j < BPW; ++j) {
        const int pos = ((j * NW + w) << 5) + lane;
        const bool pad = pos >= n;
        const float x = __ldg(gx + pos), y = __ldg(gy + pos), z = __ldg(gz + pos);
        const float tj = __ldg(tinit + (size_t)blockIdx.x * CAP + pos);
        ts[pos] = tj;
        // padding and NaN coordinates stay out of the box (a NaN point never changes anyway:
        // its distance is NaN and fminf keeps the old minimum, exactly as in the reference)
        const bool ox = pad || x != x, oy = pad || y != y, oz = pad || z != z;
        const unsigned lx = __reduce_min_sync(kFull, ox ? 0xffffffffu : f2ord(x));
        const unsigned ly = __reduce_min_sync(kFull, oy ? 0xffffffffu : f2ord(y));
        const unsigned lz = __reduce_min_sync(kFull, oz ? 0xffffffffu : f2ord(z));
        const unsigned hx = __reduce_max_sync(kFull, ox ? 0u : f2ord(x));
        const unsigned hy = __reduce_max_sync(kFull, oy ? 0u : f2ord(y));
        const unsigned hz = __reduce_max_sync(kFull, oz ? 0u : f2ord(z));
        const unsigned tb = __float_as_uint(tj);
        const unsigned mx = __reduce_max_sync(kFull, tb);
        const unsigned cand = (tb == mx && !pad) ? tiekey_at(pos) : kPadKey;
        const unsigned tkm = __reduce_min_sync(kFull, cand);
        const unsigned wl = __ffs(__ballot_sync(kFull, cand == tkm)) - 1;
        const unsigned sec = __reduce_max_sync(kFull, lane == (int)wl ? 0u : tb);
        const float wx = __shfl_sync(kFull, x, wl), wy = __shfl_sync(kFull, y, wl), wz = __shfl_sync(kFull, z, wl);
        if (lane == j) {
            blox = ord2f(lx); bloy = ord2f(ly); bloz = ord2f(lz);
            bhix = ord2f(hx); bhiy = ord2f(hy); bhiz = ord2f(hz);
            bmax = mx; bwl = wl; bsec = sec;
            bwx = wx; bwy = wy; bwz = wz;
        }
    }
    {
        const float *d0 = xyz + (size_t)blockIdx.x * n * 3;
        if (tid == 0) samp[0] = make_float4(__ldg(d0 + 0), __ldg(d0 + 1), __ldg(d0 + 2), 0.f);
    }
    __syncthreads();

    // ---- rounds (phases A-D as in fps_bucket_kernel) ------------------------------------------------
    const int wbase = (w << 5) + lane;                               // my slot in owned bucket 0
    const unsigned bbase = (unsigned)(lane * (NW * 32) + (w << 5));  // first slot of owned bucket `lane`
    unsigned c1v = 0u, c1p = 0u, c2v = 0u, c2p = 0u, wU = 0u;        // cached candidates / bound of this warp
    float c1x = 0.f, c1y = 0.f, c1z = 0.f, c2x = 0.f, c2y = 0.f, c2z = 0.f;
    bool dirty = true;
    int K = 1;         // samples accepted in the previous round, waiting to be applied (sample 0 first)
    int j = 1;         // samples emitted so far
    int rounds = 0;
    int pend_slot = -1;       // warp 0: output slot of the sample this lane's candidate became ...
    unsigned pend_val = 0u;   // ... and its original index (global load in flight since last round)
    K = 1;

    // one bucket update: new minima of my point against the samples in smask, bucket summary
    struct Upd { unsigned mx, wl, sec; float wx, wy, wz; };
    auto finish = [&](int pos, float x, float y, float z, float nt) -> Upd {
        Upd u;
        const unsigned tb = __float_as_uint(nt);
        u.mx = __reduce_max_sync(kFull, tb);
        const bool hit = tb == u.mx;
        u.wl = __reduce_max_sync(kFull, hit ? (unsigned)lane : 0u);
        u.sec = __reduce_max_sync(kFull, hit ? 0u : tb);      // next distinct value ...
        if (multi_bit(__ballot_sync(kFull, hit))) {            // duplicates: smallest tiekey wins,
            const unsigned cand = hit ? tiekey_at(pos) : kPadKey;
            const unsigned tkm = __reduce_min_sync(kFull, cand);
            u.wl = __reduce_max_sync(kFull, cand == tkm ? (unsigned)lane : 0u);
            u.sec = u.mx;                                      // ... and the runner-up equals the maximum
        }
        u.wx = __shfl_sync(kFull, x, u.wl);
        u.wy = __shfl_sync(kFull, y, u.wl);
        u.wz = __shfl_sync(kFull, z, u.wl);
        return u;
    };

    for (;;) {
        // A. which of my buckets can change?  exact lower bound of d over the bucket box
        const float4 *ws = samp;
        unsigned amask = 0u;  // bit k: sample k can change my bucket
        {
            const float bm = __uint_as_float(bmax);
            for (int k = 0; k < K; k += 2) {
                const float4 c = ws[k], e2 = ws[(k + 1 < KMAX) ? k + 1 : k];
                const float gx_ = fmaxf(fmaxf(__fsub_rn(blox, c.x), __fsub_rn(c.x, bhix)), 0.f);
                const float gy_ = fmaxf(fmaxf(__fsub_rn(bloy, c.y), __fsub_rn(c.y, bhiy)), 0.f);
                const float gz_ = fmaxf(fmaxf(__fsub_rn(bloz, c.z), __fsub_rn(c.z, bhiz)), 0.f);
                const float hx = fmaxf(fmaxf(__fsub_rn(blox, e2.x), __fsub_rn(e2.x, bhix)), 0.f);
                const float hy = fmaxf(fmaxf(__fsub_rn(bloy, e2.y), __fsub_rn(e2.y, bhiy)), 0.f);
                const float hz = fmaxf(fmaxf(__fsub_rn(bloz, e2.z), __fsub_rn(e2.z, bhiz)), 0.f);
                const unsigned a0 = sqdist_ref(gx_, gy_, gz_) < bm ? 1u : 0u;
                const unsigned a1 = (k + 1 < K && sqdist_ref(hx, hy, hz) < bm) ? 2u : 0u;
                amask |= (a0 | a1) << k;
            }
        }
        unsigned mask = __ballot_sync(kFull, amask != 0u);
        // B. update the surviving buckets, two per iteration: their coordinate loads (L2) and
        //    reduction chains are independent and overlap
        while (mask) {
            const int j0 = 31 - __clz(mask);
            mask ^= 1u << j0;
            const bool two = mask != 0u;
            const int j1 = two ? 31 - __clz(mask) : j0;
            mask &= ~(1u << j1);
            const int pos0 = j0 * (NW * 32) + wbase, pos1 = j1 * (NW * 32) + wbase;
            const float x0 = __ldg(gx + pos0), y0 = __ldg(gy + pos0), z0 = __ldg(gz + pos0);
            const float x1 = __ldg(gx + pos1), y1 = __ldg(gy + pos1), z1 = __ldg(gz + pos1);
            const float o0 = ts[pos0], o1 = ts[pos1];
            float n0 = o0, n1 = o1;
            unsigned s0 = __shfl_sync(kFull, amask, j0), s1 = __shfl_sync(kFull, amask, j1);
            while (s0) {
                const int k = 31 - __clz(s0);
                s0 ^= 1u << k;
                const float4 c = ws[k];
                n0 = fminf(sqdist_ref(__fsub_rn(x0, c.x), __fsub_rn(y0, c.y), __fsub_rn(z0, c.z)), n0);
            }
            while (s1) {
                const int k = 31 - __clz(s1);
                s1 ^= 1u << k;
                const float4 c = ws[k];
                n1 = fminf(sqdist_ref(__fsub_rn(x1, c.x), __fsub_rn(y1, c.y), __fsub_rn(z1, c.z)), n1);
            }
            // the box test is conservative: a bucket can survive it without any of its points coming
            // closer to a new sample -- then its summary is still valid and the reductions are skipped
            const bool ch0 = __any_sync(kFull, n0 != o0), ch1 = two && __any_sync(kFull, n1 != o1);
            if (!ch0 && !ch1) continue;
            const Upd u0 = finish(pos0, x0, y0, z0, n0);
            const Upd u1 = finish(pos1, x1, y1, z1, n1);   // (j1 == j0 when single: same result, harmless)
            if (ch0) ts[pos0] = n0;
            if (ch1) ts[pos1] = n1;
            if (ch0 && lane == j0) { bmax = u0.mx; bwl = u0.wl; bsec = u0.sec; bwx = u0.wx; bwy = u0.wy; bwz = u0.wz; }
            if (ch1 && lane == j1) { bmax = u1.mx; bwl = u1.wl; bsec = u1.sec; bwx = u1.wx; bwy = u1.wy; bwz = u1.wz; }
            dirty = true;
        }
        if (j >= m) break;
        ++rounds;
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
            c1x = __shfl_sync(kFull, bwx, src1); c1y = __shfl_sync(kFull, bwy, src1); c1z = __shfl_sync(kFull, bwz, src1);
            c2p = __shfl_sync(kFull, bbase + bwl, src2);
            c2x = __shfl_sync(kFull, bwx, src2); c2y = __shfl_sync(kFull, bwy, src2); c2z = __shfl_sync(kFull, bwz, src2);
            if (!has2) { c2v = 0u; c2p = c1p; c2x = c1x; c2y = c1y; c2z = c1z; }  // single-bucket warp: a dead second candidate
        }
        const int par = (rounds & 1);
        if (lane == 0) {
            pubA[par * 2 * NW + 2 * w] = make_uint4(c1v, c1p, __float_as_uint(c1x), __float_as_uint(c1y));
            pubA[par * 2 * NW + 2 * w + 1] = make_uint4(c2v, c2p, __float_as_uint(c2x), __float_as_uint(c2y));
            pubZ[par * 2 * NW + 2 * w] = c1z;
            pubZ[par * 2 * NW + 2 * w + 1] = c2z;
            pubU[par * NW + w] = wU;
        }
        __syncthreads();
        // D. warp 0 replays the sequential selection on the 2*NW candidates (one per lane); the other
        //    warps sleep at the second barrier.  (fps_bucket_kernel lets EVERY warp replay it to save that
        //    barrier; that is 45 % of its instructions, which is what co-resident frames compete for.)
        if (w == 0) {
            if (pend_slot >= 0) out[pend_slot] = (int)pend_val;
            pend_slot = -1;
            const bool live = lane < 2 * NW;
            const uint4 e = live ? pubA[par * 2 * NW + lane] : make_uint4(0u, 0u, 0u, 0u);
            const float z = live ? pubZ[par * 2 * NW + lane] : 0.f;
            const float x = __uint_as_float(e.z), y = __uint_as_float(e.w);
            const unsigned U = __reduce_max_sync(kFull, lane < NW ? pubU[par * NW + lane] : 0u);
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
            const int kmax_now = min(KMAX, m - j);
            int kk = 0;
            for (;;) {
                // coordinates of the pick, broadcast from the winning lane (exactly one lane has ghit)
                const float px = __uint_as_float(__reduce_max_sync(kFull, ghit ? __float_as_uint(x) : 0u));
                const float py = __uint_as_float(__reduce_max_sync(kFull, ghit ? __float_as_uint(y) : 0u));
                const float pz = __uint_as_float(__reduce_max_sync(kFull, ghit ? __float_as_uint(z) : 0u));
                if (ghit) pend_slot = j + kk;
                if (lane == 0) samp[kk] = make_float4(px, py, pz, 0.f);
                ++kk;
                if (kk >= kmax_now) break;
                v = fminf(sqdist_ref(__fsub_rn(x, px), __fsub_rn(y, py), __fsub_rn(z, pz)), v);
                const unsigned vb = live ? __float_as_uint(v) : 0u;
                const unsigned g2 = __reduce_max_sync(kFull, vb);
                if (!(g2 > U)) break;              // a non-candidate may be as large: stop
                ghit = live && vb == g2;
                if (multi_bit(__ballot_sync(kFull, ghit))) break;  // equal candidates: pick 1 of the next round decides
            }
            if (lane == 0) *ksh = kk;
            // index of the accepted samples through the map, one round late (one load per round)
            if (pend_slot >= 0) pend_val = __ldg(pm + e.y);
        }
        __syncthreads();
        K = *ksh;
        j += K;
        if (j >= m) {   // the very last sample is never applied (the reference stops after writing it)
            --K;
            if (K == 0) break;
        }
    }
    if (w == 0 && pend_slot >= 0) out[pend_slot] = (int)pend_val;
    if (stats && tid == 0) stats[blockIdx.x] = rounds;

    // ---- leave temp as the reference does: running minima in original order -------------------------
    __syncthreads();
    float *tmp = temp + (size_t)blockIdx.x * n;
    for (int pos = tid; pos < n; pos += NW * 32) tmp[__ldg(pm + pos)] = ts[pos];
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
template <int CAP, int OCC>
static int launch_l2(int b, int n, int m, int p, const float *xyz, float *temp, int *idx, int *stats, cudaStream_t st) {
    constexpr int NW = 16, BPW = CAP / (32 * NW), KMAX = 8;
    using L = FpsL2Smem<NW, BPW, KMAX>;
    const size_t per_frame = (size_t)CAP * 5 * sizeof(float);   // 3 coordinates + initial minima + map
    char *scratch = static_cast<char *>(stream_scratch(st, per_frame * b));
    if (!scratch) return PDM_ERR_INVALID_ARG;  // message recorded by stream_scratch
    float *sorted = reinterpret_cast<float *>(scratch);
    float *tinit = sorted + (size_t)b * 3 * CAP;
    unsigned *pmap = reinterpret_cast<unsigned *>(tinit + (size_t)b * CAP);
    auto prep = fps_prepare_kernel<CAP>;
    auto kern = fps_l2_kernel<NW, BPW, KMAX, OCC>;
    const size_t prep_smem = (size_t)CAP * sizeof(unsigned);
    if (int rc = ensure_dynamic_smem((const void *)prep, prep_smem)) return rc;
    if (int rc = ensure_dynamic_smem((const void *)kern, L::kBytes)) return rc;
    prefer_max_smem((const void *)prep);
    prefer_max_smem((const void *)kern);
    prep<<<b, CAP / 16, prep_smem, st>>>(n, xyz, temp, sorted, tinit, pmap);
    count_launch();
    PDM_CHECK_LAUNCH("farthest_point_sampling(prepare)");
    kern<<<b, NW * 32, L::kBytes, st>>>(n, m, p, xyz, sorted, tinit, pmap, temp, idx, stats);
    count_launch();
    PDM_CHECK_LAUNCH("farthest_point_sampling(l2)");
    return PDM_OK;
}

bool fps_l2_supports(int n) { return n >= 512 && n <= 16384; }

int fps_l2_launch(int b, int n, int m, int p, const float *xyz, float *temp, int *idx, int *stats, cudaStream_t st) {
    if (!fps_l2_supports(n)) return PDM_ERR_UNSUPPORTED;
    static const int occ = [] { const char *e = getenv("PDM_FPS_OCC"); return e ? atoi(e) : 2; }();
    if (n <= 1024) return launch_l2<1024, 2>(b, n, m, p, xyz, temp, idx, stats, st);
    if (n <= 2048) return launch_l2<2048, 2>(b, n, m, p, xyz, temp, idx, stats, st);
    if (n <= 4096) return launch_l2<4096, 2>(b, n, m, p, xyz, temp, idx, stats, st);
    if (n <= 8192) return launch_l2<8192, 2>(b, n, m, p, xyz, temp, idx, stats, st);
    if (occ == 3) return launch_l2<16384, 3>(b, n, m, p, xyz, temp, idx, stats, st);
    if (occ == 1) return launch_l2<16384, 1>(b, n, m, p, xyz, temp, idx, stats, st);
    return launch_l2<16384, 2>(b, n, m, p, xyz, temp, idx, stats, st);
}

}  // namespace pdm
