// pdm_neck.cu -- the PDM neck of SPEC_PDM.md: point dilation (n1), spherical-harmonic x Gaussian
// feature filling (n2), multi-centre fusion + height compression into a dense BEV map (n3).
//
// There is no reference code for this stage; SPEC_PDM.md is the spec and
// oracle/pdm_neck_oracle.py its executable form.  The arithmetic of n1/n2 is spelled with
// round-to-nearest intrinsics in exactly the operation order of the spec, so the cell
// coordinates / keys are bit-identical to the oracle's and the weights differ only through expf.
//
// Pipeline (one stream, no host sync, no atomics on floating-point data):
//   pdm_emit_kernel     thread per (centre, offset): cell, key3, weight w; histogram of key3
//   pdm_scan_kernel     CTA per 4096-cell chunk: exclusive scan of the per-cell counts (chunk bases
//                       from per-chunk counts the emit kernel accumulates), frame totals
//   pdm_scatter_kernel  entry ids bucketed by cell (order inside a cell is arbitrary here ...)
//   pdm_bev_kernel      CTA per (frame, y, 32-wide x tile), warp per pillar: walks the pillar's
//                       cells in ascending z and each cell's entries in ASCENDING ENTRY ID
//                       (... restored by a min-selection), so every sum is the serial in-order
//                       fp32 sum the spec prescribes: F = sum(w f) / (sum|w| + eps), BEV = sum_z F.
//                       Lanes own channels (coalesced 128-byte reads of feature rows); the tile is
//                       transposed through shared memory so the (B,C,Y,X) output is written once,
//                       in full 128-byte rows, zeros included (no memset pass, HBM-write bound).
#include "common.cuh"

namespace pdm {

struct NeckCfg {
    float rmin[3];
    float voxel[3];
    int grid[3];   // X, Y, Z
    int dil[3];    // kx, ky, kz
    int degree;    // SH degree 0..2
    float two_sigma2;
    float eps;
};

constexpr int kScanChunkFwd = 4096;
constexpr float kSH0 = 0.28209479177387814f, kSH1 = 0.4886025119029199f, kSH2 = 1.0925484305920792f,
                kSH3 = 0.31539156525252005f, kSH4 = 0.5462742152960396f;

__global__ void __launch_bounds__(256)
pdm_emit_kernel(int p_total, int K, int nsh, NeckCfg cfg, const float *__restrict__ coords,
                const float *__restrict__ coef, int *__restrict__ keys, float *__restrict__ wts,
                int *__restrict__ count, int *__restrict__ chunk_count, int chunks_per_frame) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)p_total * K) return;
    const int p = (int)(e / K), o = (int)(e % K);
    const int ny = 2 * cfg.dil[1] + 1, nz = 2 * cfg.dil[2] + 1;
    const int ox = o / (ny * nz) - cfg.dil[0], oy = (o / nz) % ny - cfg.dil[1], oz = o % nz - cfg.dil[2];
    const float *pc = coords + (size_t)p * 4;
    const int b = (int)__ldg(pc);
    const float xyz[3] = {__ldg(pc + 1), __ldg(pc + 2), __ldg(pc + 3)};
    const int off[3] = {ox, oy, oz};
    int cell[3];
    bool ok = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int c0 = (int)floorf(__fdiv_rn(__fsub_rn(xyz[a], cfg.rmin[a]), cfg.voxel[a]));
        ok = ok && c0 >= 0 && c0 < cfg.grid[a];
        cell[a] = c0 + off[a];
        ok = ok && cell[a] >= 0 && cell[a] < cfg.grid[a];
    }
    int key = -1;
    float w = 0.f;
    if (ok) {
        key = ((b * cfg.grid[0] + cell[0]) * cfg.grid[1] + cell[1]) * cfg.grid[2] + cell[2];
        float d[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float ctr = __fadd_rn(__fmul_rn(__fadd_rn((float)cell[a], 0.5f), cfg.voxel[a]), cfg.rmin[a]);
            d[a] = __fsub_rn(ctr, xyz[a]);
        }
        const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2]));
        const float nrm = __fsqrt_rn(__fadd_rn(r2, cfg.eps));
        const float ux = __fdiv_rn(d[0], nrm), uy = __fdiv_rn(d[1], nrm), uz = __fdiv_rn(d[2], nrm);
        const float *cf = coef + (size_t)p * nsh;
        float acc = __fmul_rn(__ldg(cf), kSH0);
        if (cfg.degree >= 1) {
            acc = __fadd_rn(acc, __fmul_rn(__ldg(cf + 1), __fmul_rn(kSH1, uy)));
            acc = __fadd_rn(acc, __fmul_rn(__ldg(cf + 2), __fmul_rn(kSH1, uz)));
            acc = __fadd_rn(acc, __fmul_rn(__ldg(cf + 3), __fmul_rn(kSH1, ux)));
        }
        if (cfg.degree >= 2) {
            acc = __fadd_rn(acc, __fmul_rn(__ldg(cf + 4), __fmul_rn(kSH2, __fmul_rn(ux, uy))));
            acc = __fadd_rn(acc, __fmul_rn(__ldg(cf + 5), __fmul_rn(kSH2, __fmul_rn(uy, uz))));
            acc = __fadd_rn(acc, __fmul_rn(__ldg(cf + 6),
                                           __fmul_rn(kSH3, __fsub_rn(__fmul_rn(3.0f, __fmul_rn(uz, uz)), 1.0f))));
            acc = __fadd_rn(acc, __fmul_rn(__ldg(cf + 7), __fmul_rn(kSH2, __fmul_rn(ux, uz))));
            acc = __fadd_rn(acc, __fmul_rn(__ldg(cf + 8),
                                           __fmul_rn(kSH4, __fsub_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy)))));
        }
        w = __fmul_rn(acc, expf(__fdiv_rn(-r2, cfg.two_sigma2)));
        atomicAdd(&count[key], 1);
        const int cpf = cfg.grid[0] * cfg.grid[1] * cfg.grid[2];
        atomicAdd(&chunk_count[b * chunks_per_frame + (key - b * cpf) / kScanChunkFwd], 1);
    }
    keys[e] = key;
    wts[e] = w;
}

// Exclusive scan of the per-cell counts, in place and per frame.  The emit kernel also counts the
// entries of every chunk of kScanChunk cells (chunk_count), so each CTA can scan one chunk
// independently: its base is the sum of the earlier chunks of its frame.  CTA (chunk, frame).
constexpr int kScanThreads = 256;
constexpr int kScanChunk = kScanChunkFwd;   // cells per CTA: 16 per thread
__global__ void __launch_bounds__(kScanThreads)
pdm_scan_kernel(int cells_per_frame, int chunks_per_frame, int *__restrict__ count,
                const int *__restrict__ chunk_count, int *__restrict__ totals) {
    __shared__ int wsum[kScanThreads / 32];
    __shared__ int base_s;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int chunk = blockIdx.x, frame = blockIdx.y;
    const int *cc = chunk_count + (size_t)frame * chunks_per_frame;
    if (w == 0) {  // base = entries of the earlier chunks of this frame (and the frame total, once)
        int s = 0, all = 0;
        for (int q = lane; q < chunks_per_frame; q += 32) {
            const int v = __ldg(cc + q);
            all += v;
            if (q < chunk) s += v;
        }
        s = __reduce_add_sync(0xffffffffu, s);
        all = __reduce_add_sync(0xffffffffu, all);
        if (lane == 0) {
            base_s = s;
            if (chunk == 0) totals[frame] = all;
        }
    }
    int *c = count + (size_t)frame * cells_per_frame + (size_t)chunk * kScanChunk;
    const int ncell = min(kScanChunk, cells_per_frame - chunk * kScanChunk);
    constexpr int PER = kScanChunk / kScanThreads;
    int v[PER];
    int tsum = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        v[q] = (tid * PER + q < ncell) ? c[tid * PER + q] : 0;
        tsum += v[q];
    }
    int incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    int run = base_s + incl - tsum;
    for (int q = 0; q < w; ++q) run += wsum[q];
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        if (tid * PER + q < ncell) c[tid * PER + q] = run;
        run += v[q];
    }
}

__global__ void __launch_bounds__(256)
pdm_scatter_kernel(long long n_entries, int cells_per_frame, const int *__restrict__ keys,
                   int *__restrict__ cursor, const int *__restrict__ totals, int *__restrict__ sorted) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    const int key = __ldg(keys + e);
    if (key < 0) return;
    const int b = key / cells_per_frame;
    int base = 0;
    for (int q = 0; q < b; ++q) base += __ldg(totals + q);
    sorted[base + atomicAdd(&cursor[key], 1)] = (int)e;
}

constexpr int kBevWarps = 8;
constexpr int kBevTileX = 32;
constexpr int kCPL = 8;  // channels per lane per pass (256 channels per pass)

__global__ void __launch_bounds__(kBevWarps * 32)
pdm_bev_kernel(int c_total, int K, NeckCfg cfg, const float *__restrict__ feats, const float *__restrict__ wts,
               const int *__restrict__ cend /*per-frame local exclusive ends after the scatter*/,
               const int *__restrict__ totals, const int *__restrict__ sorted, float *__restrict__ bev) {
    __shared__ float tile[kCPL * 32][kBevTileX + 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int X = cfg.grid[0], Y = cfg.grid[1], Z = cfg.grid[2];
    const int b = blockIdx.z, cy = blockIdx.y, x0 = blockIdx.x * kBevTileX;
    const int cells_per_frame = X * Y * Z;
    int base = 0;
    for (int q = 0; q < b; ++q) base += __ldg(totals + q);
    const int *ce = cend + (size_t)b * cells_per_frame;
    const int *srt = sorted + base;

    for (int cb = 0; cb < c_total; cb += kCPL * 32) {  // 256 channels per pass
        for (int px = w; px < kBevTileX; px += kBevWarps) {
            const int cx = x0 + px;
            float acc[kCPL];
#pragma unroll
            for (int i = 0; i < kCPL; ++i) acc[i] = 0.f;
            if (cx < X) {
                const int cell0 = (cx * Y + cy) * Z;  // frame-local index of the pillar's z = 0 cell
                // the pillar's Z cell ends are contiguous: fetched 32 at a time by the lanes in one
                // coalesced load (a serial walk is Z dependent L2 round trips even for empty pillars)
                int s = cell0 == 0 ? 0 : __ldg(ce + cell0 - 1);
                for (int zb = 0; zb < Z; zb += 32) {
                  const int zn = min(32, Z - zb);
                  const int ends = lane < zn ? __ldg(ce + cell0 + zb + lane) : 0;
                  const int last_end = __shfl_sync(0xffffffffu, ends, zn - 1);
                  if (last_end == s) continue;          // nothing in these cells (most pillars)
                  for (int zi = 0; zi < zn; ++zi) {
                    const int e_end = __shfl_sync(0xffffffffu, ends, zi);
                    if (e_end > s) {  // warp-uniform
                        float num[kCPL];
#pragma unroll
                        for (int i = 0; i < kCPL; ++i) num[i] = 0.f;
                        float den = 0.f;
                        const int cnt = e_end - s;
                        if (cnt <= 32) {
                            // usual case: one entry per lane; a bitonic network over the warp puts the
                            // entry ids in ascending order (padding = 0xffffffff sorts last), then the
                            // weights and row indices are fetched by all lanes at once and the rows
                            // are streamed in order (their addresses are known up front, so the loads
                            // of consecutive entries overlap; only the fp32 adds are serial).
                            unsigned v = lane < cnt ? (unsigned)__ldg(srt + s + lane) : 0xffffffffu;
#pragma unroll
                            for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
                                for (int j = k >> 1; j > 0; j >>= 1) {
                                    const unsigned o = __shfl_xor_sync(0xffffffffu, v, j);
                                    const bool keep_min = ((lane & k) == 0) == ((lane & j) == 0);
                                    v = keep_min ? min(v, o) : max(v, o);
                                }
                            }
                            const float wq = lane < cnt ? __ldg(wts + v) : 0.f;
                            const unsigned pq = lane < cnt ? v / (unsigned)K : 0u;
#pragma unroll 2
                            for (int it = 0; it < cnt; ++it) {
                                const float wv = __shfl_sync(0xffffffffu, wq, it);
                                const float *f = feats + (size_t)__shfl_sync(0xffffffffu, pq, it) * c_total + cb + lane;
#pragma unroll
                                for (int i = 0; i < kCPL; ++i)
                                    if (cb + lane + 32 * i < c_total) num[i] = __fadd_rn(num[i], __fmul_rn(wv, __ldg(f + 32 * i)));
                                den = __fadd_rn(den, fabsf(wv));
                            }
                        } else {
                            // crowded cell: repeated min-selection over the unsorted list
                            unsigned last = 0u;
                            bool first = true;
                            for (int it = s; it < e_end; ++it) {
                                unsigned cand = 0xffffffffu;
                                for (int q = s + lane; q < e_end; q += 32) {
                                    const unsigned id = (unsigned)__ldg(srt + q);
                                    if ((first || id > last) && id < cand) cand = id;
                                }
                                const unsigned id = __reduce_min_sync(0xffffffffu, cand);
                                last = id;
                                first = false;
                                const float wv = __ldg(wts + id);
                                const float *f = feats + (size_t)(id / (unsigned)K) * c_total + cb + lane;
#pragma unroll
                                for (int i = 0; i < kCPL; ++i)
                                    if (cb + lane + 32 * i < c_total) num[i] = __fadd_rn(num[i], __fmul_rn(wv, __ldg(f + 32 * i)));
                                den = __fadd_rn(den, fabsf(wv));
                            }
                        }
                        const float dn = __fadd_rn(den, cfg.eps);
#pragma unroll
                        for (int i = 0; i < kCPL; ++i) acc[i] = __fadd_rn(acc[i], __fdiv_rn(num[i], dn));
                    }
                    s = e_end;
                  }
                }
            }
#pragma unroll
            for (int i = 0; i < kCPL; ++i) tile[lane + 32 * i][px] = acc[i];
        }
        __syncthreads();
        // coalesced store: one 32-float row of the tile per (channel) -> bev[b, c, cy, x0 .. x0+31]
        for (int r = w; r < kCPL * 32; r += kBevWarps) {
            const int c = cb + r;
            const int cx = x0 + lane;
            if (c < c_total && cx < X) st_cs_f1(bev + (((size_t)b * c_total + c) * Y + cy) * X + cx, tile[r][lane]);
        }
        __syncthreads();
    }
}

}  // namespace pdm

extern "C" int pdm_neck_forward(int batch, int p, int c, const float *point_coords,
                                const float *point_features, const float *coef,
                                const float *range_min, const float *voxel, const int *grid,
                                const int *dilation, int sh_degree, float sigma, float eps,
                                float *spatial_features, int *dbg_keys, float *dbg_w, void *stream) {
    using namespace pdm;
    if (batch < 0 || p < 0 || c < 0) return fail(PDM_ERR_INVALID_ARG, "neck_forward: negative size");
    if (!range_min || !voxel || !grid || !dilation) return fail(PDM_ERR_INVALID_ARG, "neck_forward: null config");
    if (sh_degree < 0 || sh_degree > 2) return fail(PDM_ERR_UNSUPPORTED, "neck_forward: SH degree %d", sh_degree);
    NeckCfg cfg;
    for (int a = 0; a < 3; ++a) {
        cfg.rmin[a] = range_min[a];
        cfg.voxel[a] = voxel[a];
        cfg.grid[a] = grid[a];
        cfg.dil[a] = dilation[a];
        if (grid[a] <= 0 || dilation[a] < 0 || !(voxel[a] > 0.f))
            return fail(PDM_ERR_INVALID_ARG, "neck_forward: bad grid/dilation/voxel on axis %d", a);
    }
    cfg.degree = sh_degree;
    cfg.two_sigma2 = 2.0f * sigma * sigma;
    cfg.eps = eps;
    const long long cells_per_frame = (long long)grid[0] * grid[1] * grid[2];
    if (cells_per_frame * (long long)(batch > 0 ? batch : 1) >= 0x7fffffffLL)
        return fail(PDM_ERR_UNSUPPORTED, "neck_forward: B*X*Y*Z must fit int32");
    if (batch == 0 || c == 0) return PDM_OK;
    if (!spatial_features || (p > 0 && (!point_coords || !point_features || !coef)))
        return fail(PDM_ERR_INVALID_ARG, "neck_forward: null pointer");
    if (grid[1] > 65535 || batch > 65535) return fail(PDM_ERR_UNSUPPORTED, "neck_forward: Y or batch > 65535");
    cudaStream_t st = (cudaStream_t)stream;
    const int K = (2 * dilation[0] + 1) * (2 * dilation[1] + 1) * (2 * dilation[2] + 1);
    const int nsh = (sh_degree + 1) * (sh_degree + 1);
    const long long n_entries = (long long)p * K;
    if (n_entries >= 0x7fffffffLL) return fail(PDM_ERR_UNSUPPORTED, "neck_forward: too many entries");

    auto align = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t sz_keys = align((size_t)n_entries * 4 + 4), sz_w = sz_keys, sz_sorted = sz_keys;
    const size_t sz_count = align((size_t)cells_per_frame * batch * 4);
    const int chunks_per_frame = (int)((cells_per_frame + kScanChunk - 1) / kScanChunk);
    const size_t sz_tot = align((size_t)batch * 4) + align((size_t)batch * chunks_per_frame * 4);
    char *ws = static_cast<char *>(stream_scratch(st, sz_keys + sz_w + sz_sorted + sz_count + sz_tot));
    if (!ws) return PDM_ERR_INVALID_ARG;  // message recorded by stream_scratch
    int *keys = dbg_keys ? dbg_keys : reinterpret_cast<int *>(ws);
    float *wts = dbg_w ? dbg_w : reinterpret_cast<float *>(ws + sz_keys);
    int *sorted = reinterpret_cast<int *>(ws + sz_keys + sz_w);
    int *count = reinterpret_cast<int *>(ws + sz_keys + sz_w + sz_sorted);
    int *totals = reinterpret_cast<int *>(ws + sz_keys + sz_w + sz_sorted + sz_count);
    int *chunk_count = reinterpret_cast<int *>(ws + sz_keys + sz_w + sz_sorted + sz_count + align((size_t)batch * 4));
    cudaError_t err = cudaMemsetAsync(count, 0, sz_count + sz_tot, st);
    if (err == cudaSuccess && n_entries > 0) {
        pdm_emit_kernel<<<(unsigned)((n_entries + 255) / 256), 256, 0, st>>>(p, K, nsh, cfg, point_coords, coef, keys,
                                                                           wts, count, chunk_count, chunks_per_frame);
        count_launch();
        err = cudaGetLastError();
    }
    if (err == cudaSuccess) {
        pdm_scan_kernel<<<dim3(chunks_per_frame, batch), kScanThreads, 0, st>>>((int)cells_per_frame, chunks_per_frame,
                                                                               count, chunk_count, totals);
        count_launch();
        err = cudaGetLastError();
    }
    if (err == cudaSuccess && n_entries > 0) {
        pdm_scatter_kernel<<<(unsigned)((n_entries + 255) / 256), 256, 0, st>>>(n_entries, (int)cells_per_frame, keys,
                                                                              count, totals, sorted);
        count_launch();
        err = cudaGetLastError();
    }
    if (err == cudaSuccess) {
        dim3 g((grid[0] + kBevTileX - 1) / kBevTileX, grid[1], batch);
        pdm_bev_kernel<<<g, kBevWarps * 32, 0, st>>>(c, K, cfg, point_features, wts, count, totals, sorted,
                                                     spatial_features);
        count_launch();
        err = cudaGetLastError();
    }
    if (err != cudaSuccess) return fail((int)err, "neck_forward: %s", cudaGetErrorString(err));
    return PDM_OK;
}
