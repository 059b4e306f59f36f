// fps_cluster.cu -- farthest point sampling for frames that do not fit one SM (16384 < n <= 196608,
// e.g. the Waymo-scale frames of BASELINE configs[4]): one THREAD-BLOCK CLUSTER per frame.
//
// The any-size kernel (fps_generic_kernel) gives a frame one CTA and walks all n points from L2 every
// round: 163840 points = 160 per thread per round, 32 us per round, 534 ms for 16384 samples.  Here the
// frame is cut into `cl` contiguous chunks, one per CTA of a cluster of up to 16 CTAs (SMs of one
// GPC); a CTA keeps its chunk's coordinates AND running minima in shared memory (16 B/point, up to
// 14336 points) for the whole kernel, so a round touches no global memory at all:
//   1. every thread updates its <= 12 points against the last sample and keeps its best
//      (value bits, tiekey) -- the reference's argmax with its tie-break (fps.cu header);
//   2. warp argmax with two redux.sync, CTA argmax through 32 shared-memory slots;
//   3. warp 0 writes the CTA's candidate (key, coordinates) into slot [round parity][rank] of EVERY
//      CTA of the cluster through distributed shared memory;
//   4. one cluster barrier (release/acquire);
//   5. every warp reduces the `cl` candidates (two redux.sync): the next sample and its coordinates.
// Exactly the reference's sequence of samples and its final `temp` (bit-exact, ties included).
#include <cooperative_groups.h>
#include <stdlib.h>

#include <map>
#include <mutex>

#include "fps_common.cuh"

namespace cg = cooperative_groups;

namespace pdm {

constexpr int kClThreads = 1024;
constexpr int kClMax = 16;
constexpr int kClMaxChunk = 14336;  // points per CTA: 16 B each = 224 KB of the 227 KB of shared memory

__global__ void __launch_bounds__(kClThreads, 1)
fps_cluster_kernel(int n, int m, int p, int cap /*points per CTA, a multiple of 1024*/, const float *__restrict__ xyz,
                   float *__restrict__ temp, int *__restrict__ idxs) {
    extern __shared__ __align__(16) float cl_smem[];
    __shared__ unsigned long long wbest[kClThreads / 32];
    __shared__ unsigned long long cand_key[2][kClMax];
    __shared__ float4 cand_xyz[2][kClMax];
    cg::cluster_group cluster = cg::this_cluster();
    const int cl = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int frame = blockIdx.x / cl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned bsmask = (1u << p) - 1u;
    float *sx = cl_smem, *sy = sx + cap, *sz = sy + cap, *st = sz + cap;
    const float *dataset = xyz + (size_t)frame * n * 3;
    float *tmp = temp + (size_t)frame * n;
    int *out = idxs + (size_t)frame * m;
    const int chunk = cap;            // a multiple of 1024 (see step 1)
    const int k0 = rank * chunk;
    const int cnt = max(0, min(chunk, n - k0));
    const int ppt = chunk / kClThreads;

    for (int i = tid; i < cnt; i += kClThreads) {
        sx[i] = __ldg(dataset + (size_t)(k0 + i) * 3 + 0);
        sy[i] = __ldg(dataset + (size_t)(k0 + i) * 3 + 1);
        sz[i] = __ldg(dataset + (size_t)(k0 + i) * 3 + 2);
        st[i] = tmp[k0 + i];
    }
    float x1 = __ldg(dataset + 0), y1 = __ldg(dataset + 1), z1 = __ldg(dataset + 2);
    if (rank == 0 && tid == 0) out[0] = 0;
    cluster.sync();   // every CTA of the cluster is resident before anyone writes into its shared memory

    for (int j = 1; j < m; ++j) {
        // 1. my points against the last sample.  Chunks start at multiples of the reference block size
        //    (1024 for every n >= 1024), so a thread's points tid, tid + 1024, ... share the bit-reversed
        //    part of the tiekey and come in ascending tiekey order: keeping the FIRST maximum (strict >) is
        //    the reference's tie-break inside the thread -- its own strided scan does exactly this.
        unsigned bvb = 0u;
        int bi = -1;
#pragma unroll 2
        for (int q = 0; q < ppt; ++q) {
            const int i = tid + q * kClThreads;
            if (i < cnt) {
                const float d = sqdist_ref(__fsub_rn(sx[i], x1), __fsub_rn(sy[i], y1), __fsub_rn(sz[i], z1));
                const float t = st[i];
                const float d2 = fminf(d, t);
                if (d2 != t) st[i] = d2;
                const unsigned ub = __float_as_uint(d2);
                const bool better = ub > bvb || bi < 0;
                bvb = better ? ub : bvb;
                bi = better ? i : bi;
            }
        }
        const unsigned btk = bi >= 0 ? fps_tiekey((unsigned)(k0 + bi), p, bsmask) : kPadKey;
        // 2. warp, then CTA
        {
            const unsigned mx = __reduce_max_sync(kFull, bvb);
            const unsigned tkm = __reduce_min_sync(kFull, (bi >= 0 && bvb == mx) ? btk : kPadKey);
            if (lane == 0) wbest[warp] = ((unsigned long long)mx << 32) | (unsigned)(~tkm);
        }
        __syncthreads();
        const int par = j & 1;
        if (warp == 0) {
            const unsigned long long wb = wbest[lane];
            const unsigned hi = (unsigned)(wb >> 32), tk = ~(unsigned)wb;
            const unsigned mx = __reduce_max_sync(kFull, hi);
            const unsigned tkm = __reduce_min_sync(kFull, hi == mx ? tk : kPadKey);
            // 3. publish to every CTA of the cluster (lane r writes into CTA r); a CTA without points
            //    publishes key 0 / tiekey 0xffffffff, which loses against any real candidate
            if (lane < cl) {
                float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
                if (tkm != kPadKey) {
                    const int li = (int)fps_tiekey_inv(tkm, p, bsmask) - k0;
                    c = make_float4(sx[li], sy[li], sz[li], 0.f);
                }
                unsigned long long *rk = cluster.map_shared_rank(&cand_key[par][rank], lane);
                float4 *rc = cluster.map_shared_rank(&cand_xyz[par][rank], lane);
                *rk = ((unsigned long long)mx << 32) | (unsigned)(~tkm);
                *rc = c;
            }
        }
        // 4. one barrier per round: candidates of this round are visible, everyone is done with the
        //    slots of the previous round's parity.  Only the publishing warp needs release semantics
        //    (cg::cluster_group::sync() makes all 32 warps execute the fence: 26 % of the samples in ncu).
        if (warp == 0) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        else asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        // 5. the sample: largest value, smallest tiekey
        {
            const unsigned long long ck = lane < cl ? cand_key[par][lane] : 0ull;
            const unsigned hi = (unsigned)(ck >> 32), tk = lane < cl ? ~(unsigned)ck : kPadKey;
            const unsigned mx = __reduce_max_sync(kFull, hi);
            const unsigned tkm = __reduce_min_sync(kFull, (lane < cl && hi == mx) ? tk : kPadKey);
            const int wr = __ffs(__ballot_sync(kFull, lane < cl && hi == mx && tk == tkm)) - 1;
            const float4 c = cand_xyz[par][wr];
            x1 = c.x; y1 = c.y; z1 = c.z;
            if (rank == 0 && tid == 0) out[j] = (int)fps_tiekey_inv(tkm, p, bsmask);
        }
    }
    // the reference leaves the running minima in temp
    for (int i = tid; i < cnt; i += kClThreads) tmp[k0 + i] = st[i];
    cluster.sync();   // nobody exits while a peer may still address its shared memory
}

bool fps_cluster_supports(int n) { return n > 16384 && n <= kClMax * kClMaxChunk; }   // (n > 16384: p = 10)

// Clusters of `cl` CTAs of this kernel the device can keep resident at once (a 16-CTA cluster needs a
// GPC with 16 free SMs: 7 on this B200, so a batch of 8 frames would run in two waves).
static int max_active_clusters(int cl, size_t smem) {
    static std::mutex mu;
    static std::map<std::pair<int, size_t>, int> cache;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(dev * 64 + cl, smem);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)cl);
    cfg.blockDim = dim3(kClThreads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int v = 0;
    // clusters of more than 8 CTAs are "non-portable": allowed per device, once (first query on that device)
    if (cl > 8 && cudaFuncSetAttribute((const void *)fps_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
        (void)cudaGetLastError();
    if (cudaOccupancyMaxActiveClusters(&v, (const void *)fps_cluster_kernel, &cfg) != cudaSuccess) {
        (void)cudaGetLastError();
        v = 0;
    }
    cache[key] = v;
    return v;
}

// Returns PDM_ERR_UNSUPPORTED (no error recorded) when the shape is out of range or the device cannot
// co-schedule a cluster of the needed size; the caller then uses the any-size kernel.
int fps_cluster_launch(int b, int n, int m, int p, const float *xyz, float *temp, int *idx, cudaStream_t st) {
    if (!fps_cluster_supports(n)) return PDM_ERR_UNSUPPORTED;
    auto kern = fps_cluster_kernel;
    if (int rc = ensure_dynamic_smem((const void *)kern, (size_t)kClMaxChunk * 16)) return rc;
    // cluster size: the one with the least modelled time  waves(b, cl) x (1.0 + 0.2 x points per thread) us per
    // round (constants measured on B200: 2.95 us at 10 points per thread, 2.27 us at 6)
    int cl = 0, cap = 0;
    double best = 1e30;
    for (int c = kClMax; c >= 2; --c) {
        const int chunk = ((n + c - 1) / c + kClThreads - 1) / kClThreads * kClThreads;
        if (chunk > kClMaxChunk) break;
        if ((long long)(c - 1) * chunk >= n) continue;   // the last CTA would be empty: a smaller cluster does the same
        const int act = max_active_clusters(c, (size_t)chunk * 16);
        if (getenv("PDM_DEBUG_CLUSTER")) fprintf(stderr, "[pdm]   cluster of %d: %d points per CTA, %d active clusters\n", c, chunk, act);
        if (act < 1) continue;
        const int waves = (b + act - 1) / act;
        const double t = waves * (1.0 + 0.2 * (chunk / kClThreads));
        if (t < best) { best = t; cl = c; cap = chunk; }
    }
    if (cl == 0) return PDM_ERR_UNSUPPORTED;
    const size_t smem = (size_t)cap * 16;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(b * cl));
    cfg.blockDim = dim3(kClThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static const bool dbg = getenv("PDM_DEBUG_CLUSTER") != nullptr;
    if (dbg) fprintf(stderr, "[pdm] fps cluster: b=%d n=%d cl=%d smem=%zu max_active_clusters=%d\n", b, n, cl, smem, max_active_clusters(cl, smem));
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, n, m, p, cap, xyz, temp, idx);
    if (e != cudaSuccess) return fail((int)e, "farthest_point_sampling(cluster of %d): %s", cl, cudaGetErrorString(e));
    count_launch();
    return PDM_OK;
}

}  // namespace pdm
