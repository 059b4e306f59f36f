// fps_cluster.cu -- farthest point sampling for frames that do not fit one SM (16384 < n <= 196608,
// e.g. the Waymo-scale frames of BASELINE configs[4]): one THREAD-BLOCK CLUSTER per frame.
//
// The any-size kernel (fps_generic_kernel) gives a frame one CTA and walks all n points from L2 every
// round: 163840 points = 160 per thread per round, 32 us per round, 534 ms for 16384 samples.  Here the
// frame is cut into `cl` contiguous chunks, one per CTA of a cluster of up to 16 CTAs (16 SMs of one
// GPC); a CTA keeps its chunk's coordinates AND running minima in shared memory (16 B/point, up to
// 12288 points) for the whole kernel, so a round touches no global memory at all:
//   1. every thread updates its <= 12 points against the last sample and keeps its best
//      (value bits, tiekey) -- the reference's argmax with its tie-break (fps.cu header);
//   2. warp argmax with two redux.sync, CTA argmax through 32 shared-memory slots;
//   3. warp 0 writes the CTA's candidate (key, coordinates) into slot [round parity][rank] of EVERY
//      CTA of the cluster through distributed shared memory;
//   4. one cluster barrier (release/acquire);
//   5. every warp reduces the `cl` candidates (two redux.sync): the next sample and its coordinates.
// Exactly the reference's sequence of samples and its final `temp` (bit-exact, ties included).
#include <cooperative_groups.h>

#include "fps_common.cuh"

namespace cg = cooperative_groups;

namespace pdm {

constexpr int kClThreads = 1024;
constexpr int kClMaxPPT = 12;      // points per thread: 12288 points = 192 KB of shared memory per CTA
constexpr int kClMax = 16;

__global__ void __launch_bounds__(kClThreads, 1)
fps_cluster_kernel(int n, int m, int p, int ppt, const float *__restrict__ xyz, float *__restrict__ temp,
                   int *__restrict__ idxs) {
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
    const int cap = kClThreads * ppt;
    float *sx = cl_smem, *sy = sx + cap, *sz = sy + cap, *st = sz + cap;
    const float *dataset = xyz + (size_t)frame * n * 3;
    float *tmp = temp + (size_t)frame * n;
    int *out = idxs + (size_t)frame * m;
    const int chunk = (n + cl - 1) / cl;
    const int k0 = rank * chunk;
    const int cnt = max(0, min(chunk, n - k0));

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
        // 1. my points against the last sample; best = largest value, smallest tiekey among equals
        unsigned bvb = 0u, btk = kPadKey;
        int bi = -1;
        for (int q = 0; q < ppt; ++q) {
            const int i = tid + q * kClThreads;
            if (i < cnt) {
                const float d = sqdist_ref(__fsub_rn(sx[i], x1), __fsub_rn(sy[i], y1), __fsub_rn(sz[i], z1));
                const float t = st[i];
                const float d2 = fminf(d, t);
                if (d2 != t) st[i] = d2;
                const unsigned ub = __float_as_uint(d2);
                if (ub > bvb || bi < 0) {
                    bvb = ub; bi = i; btk = kPadKey;          // tiekey computed lazily
                } else if (ub == bvb) {
                    if (btk == kPadKey) btk = fps_tiekey((unsigned)(k0 + bi), p, bsmask);
                    const unsigned tk = fps_tiekey((unsigned)(k0 + i), p, bsmask);
                    if (tk < btk) { btk = tk; bi = i; }
                }
            }
        }
        if (bi >= 0 && btk == kPadKey) btk = fps_tiekey((unsigned)(k0 + bi), p, bsmask);
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
        //    slots of the previous round's parity
        cluster.sync();
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

bool fps_cluster_supports(int n) { return n > 16384 && n <= kClMax * kClThreads * kClMaxPPT; }

// Returns PDM_ERR_UNSUPPORTED (no error recorded) when the shape is out of range or the device cannot
// co-schedule a cluster of the needed size; the caller then uses the any-size kernel.
int fps_cluster_launch(int b, int n, int m, int p, const float *xyz, float *temp, int *idx, cudaStream_t st) {
    if (!fps_cluster_supports(n)) return PDM_ERR_UNSUPPORTED;
    int cl = 4;
    while (cl < kClMax && (long long)cl * kClThreads * kClMaxPPT < n) cl <<= 1;
    // more CTAs per frame while the batch leaves SMs idle and a thread keeps >= 4 points
    while (cl < kClMax && (long long)b * cl * 2 <= kNumSMs && n / (cl * 2 * kClThreads) >= 4) cl <<= 1;
    const int chunk = (n + cl - 1) / cl;
    const int ppt = (chunk + kClThreads - 1) / kClThreads;
    const size_t smem = (size_t)kClThreads * ppt * 16;
    auto kern = fps_cluster_kernel;
    if (int rc = ensure_dynamic_smem((const void *)kern, smem)) return rc;
    if (cl > 8) {
        static std::atomic<int> allowed{0};   // 0 unknown, 1 ok, -1 refused
        if (allowed.load() == 0)
            allowed.store(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess ? 1 : -1);
        if (allowed.load() < 0) { (void)cudaGetLastError(); return PDM_ERR_UNSUPPORTED; }
    }
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
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, (const void *)kern, &cfg) != cudaSuccess || max_clusters < 1) {
        (void)cudaGetLastError();
        return PDM_ERR_UNSUPPORTED;
    }
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, n, m, p, ppt, xyz, temp, idx);
    if (e != cudaSuccess) return fail((int)e, "farthest_point_sampling(cluster of %d): %s", cl, cudaGetErrorString(e));
    count_launch();
    return PDM_OK;
}

}  // namespace pdm
