/*
 * oracle/pdm_stack_oracle.c -- CPU restatement of the reference's pointnet2_stack ops (stacked / ragged batches).
 *
 * TEST INFRASTRUCTURE ONLY (same rules as pdm_oracle.c: never imported, linked or called by pdm_ssd_b200/).
 *
 * Parity status: PINNED.  Checked against golden vectors produced by the reference's own pointnet2_stack kernels,
 * compiled unmodified from /root/reference into oracle/_ref/ and run on a B200 (tests/golden/make_golden_stack.py ->
 * tests/golden/stack_*.npz), and against the live reference extension in the GPU tests.
 *
 * Arithmetic (reference SASS, nvcc 12.9 default contraction, cuobjdump of the objects under oracle/_ref/stack): every
 * a*a + b*b + c*c is  t = rn(b*b); t = fma(a,a,t); fma(c,c,t);  w0*f0 + w1*f1 + w2*f2 is
 * fma(w2,f2, fma(w0,f0, rn(w1*f1)));  the vector-pool cell index uses an IEEE division.  Built with -ffp-contract=off.
 * Paths below are relative to pcdet/ops/pointnet2/pointnet2_stack/src/.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline float sq3(float dx, float dy, float dz) {
    float t = dy * dy;
    t = fmaf(dx, dx, t);
    return fmaf(dz, dz, t);
}

/* frame of row r given per-frame counts (the loop every kernel opens with, e.g. ball_query_gpu.cu:25-31) */
static int frame_of(const int *cnt, int b, int r) {
    int bs = 0, acc = cnt[0];
    for (int k = 1; k < b; ++k) {
        if (r < acc) break;
        acc += cnt[k];
        bs = k;
    }
    return bs;
}
static int start_of(const int *cnt, int f) {
    int s = 0;
    for (int k = 0; k < f; ++k) s += cnt[k];
    return s;
}

/* ball_query_gpu.cu:15-70 ball_query_kernel_stack; idx (M,nsample) pre-zeroed by the caller (pointnet2_utils.py:31) */
void oracle_stack_ball_query(int b, int m, float radius, int nsample, const float *new_xyz, const int *new_cnt,
                             const float *xyz, const int *xyz_cnt, int *idx) {
    const float radius2 = radius * radius;
    for (int pt = 0; pt < m; ++pt) {
        const int f = frame_of(new_cnt, b, pt);
        const float *pts = xyz + (size_t)start_of(xyz_cnt, f) * 3;
        const int n = xyz_cnt[f];
        const float nx = new_xyz[pt * 3], ny = new_xyz[pt * 3 + 1], nz = new_xyz[pt * 3 + 2];
        int *o = idx + (size_t)pt * nsample;
        int cnt = 0;
        for (int k = 0; k < n; ++k) {
            const float d2 = sq3(nx - pts[k * 3], ny - pts[k * 3 + 1], nz - pts[k * 3 + 2]);
            if (d2 < radius2) {
                if (cnt == 0)
                    for (int l = 0; l < nsample; ++l) o[l] = k;
                o[cnt] = k;
                ++cnt;
                if (cnt >= nsample) break;
            }
        }
        if (cnt == 0) o[0] = -1;
    }
}

/* sampling_gpu.cu:263-333 stack_farthest_point_sampling_kernel<1024>: literal simulation of the 1024 virtual threads
 * and the shared-memory tournament, per frame; global row indices out */
void oracle_stack_fps(int batch, const float *xyz, float *temp, const int *xyz_cnt, int *idxs, const int *m_cnt) {
    enum { BS = 1024 };
    static float dists[BS];
    static int dists_i[BS];
    for (int f = 0; f < batch; ++f) {
        const int xs = start_of(xyz_cnt, f), os = start_of(m_cnt, f);
        const float *dataset = xyz + (size_t)xs * 3;
        float *tmp = temp + xs;
        int *out = idxs + os;
        const int n = xyz_cnt[f], m = m_cnt[f];
        if (m <= 0) continue;
        int old = 0;
        out[0] = xs;
        for (int j = 1; j < m; ++j) {
            const float x1 = dataset[old * 3], y1 = dataset[old * 3 + 1], z1 = dataset[old * 3 + 2];
            for (int t = 0; t < BS; ++t) { dists[t] = -1.0f; dists_i[t] = 0; }
            for (int k = 0; k < n; ++k) {
                const int t = k & (BS - 1);
                const float d = sq3(dataset[k * 3] - x1, dataset[k * 3 + 1] - y1, dataset[k * 3 + 2] - z1);
                const float d2 = fminf(d, tmp[k]);
                tmp[k] = d2;
                if (d2 > dists[t]) { dists[t] = d2; dists_i[t] = k; }
            }
            for (int half = BS >> 1; half >= 1; half >>= 1)
                for (int t = 0; t < half; ++t) {
                    const float v1 = dists[t], v2 = dists[t + half];
                    const int i1 = dists_i[t], i2 = dists_i[t + half];
                    dists[t] = fmaxf(v1, v2);
                    dists_i[t] = v2 > v1 ? i2 : i1;
                }
            old = dists_i[0];
            out[j] = old + xs;
        }
    }
}

/* group_points_gpu.cu:67-101 */
void oracle_stack_group_points(int b, int m, int c, int nsample, const float *features, const int *features_cnt,
                               const int *idx, const int *idx_cnt, float *out) {
    for (int pt = 0; pt < m; ++pt) {
        const int f = frame_of(idx_cnt, b, pt);
        const float *feat = features + (size_t)start_of(features_cnt, f) * c;
        for (int ch = 0; ch < c; ++ch)
            for (int s = 0; s < nsample; ++s)
                out[((size_t)pt * c + ch) * nsample + s] = feat[(size_t)idx[(size_t)pt * nsample + s] * c + ch];
    }
}

/* group_points_gpu.cu:14-44 (serial order = ascending (pt, c, s); the GPU's atomicAdd order is unspecified) */
void oracle_stack_group_points_grad(int b, int m, int c, int nsample, const float *grad_out, const int *idx,
                                    const int *idx_cnt, const int *features_cnt, float *grad_features) {
    for (int pt = 0; pt < m; ++pt) {
        const int f = frame_of(idx_cnt, b, pt);
        float *gf = grad_features + (size_t)start_of(features_cnt, f) * c;
        for (int ch = 0; ch < c; ++ch)
            for (int s = 0; s < nsample; ++s)
                gf[(size_t)idx[(size_t)pt * nsample + s] * c + ch] += grad_out[((size_t)pt * c + ch) * nsample + s];
    }
}

/* interpolate_gpu.cu:17-75 three_nn_kernel_stack (double bests init 1e40, stored as float) */
void oracle_stack_three_nn(int b, int n, const float *unknown, const int *unknown_cnt, const float *known,
                           const int *known_cnt, float *dist2, int *idx) {
    for (int pt = 0; pt < n; ++pt) {
        const int f = frame_of(unknown_cnt, b, pt);
        const int ks = start_of(known_cnt, f), kn = known_cnt[f];
        const float *kp = known + (size_t)ks * 3;
        const float ux = unknown[pt * 3], uy = unknown[pt * 3 + 1], uz = unknown[pt * 3 + 2];
        double b1 = 1e40, b2 = 1e40, b3 = 1e40;
        int i1 = 0, i2 = 0, i3 = 0;
        for (int k = 0; k < kn; ++k) {
            const float d = sq3(ux - kp[k * 3], uy - kp[k * 3 + 1], uz - kp[k * 3 + 2]);
            if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
            else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
            else if (d < b3) { b3 = d; i3 = k; }
        }
        dist2[pt * 3] = (float)b1; dist2[pt * 3 + 1] = (float)b2; dist2[pt * 3 + 2] = (float)b3;
        idx[pt * 3] = i1 + ks; idx[pt * 3 + 1] = i2 + ks; idx[pt * 3 + 2] = i3 + ks;
    }
}

/* interpolate_gpu.cu:100-120 */
void oracle_stack_three_interpolate(int n, int c, const float *features, const int *idx, const float *weight, float *out) {
    for (int pt = 0; pt < n; ++pt)
        for (int ch = 0; ch < c; ++ch) {
            const float w0 = weight[pt * 3], w1 = weight[pt * 3 + 1], w2 = weight[pt * 3 + 2];
            const float f0 = features[(size_t)idx[pt * 3] * c + ch], f1 = features[(size_t)idx[pt * 3 + 1] * c + ch],
                        f2 = features[(size_t)idx[pt * 3 + 2] * c + ch];
            float t = w1 * f1;
            t = fmaf(w0, f0, t);
            out[(size_t)pt * c + ch] = fmaf(w2, f2, t);
        }
}

/* interpolate_gpu.cu:137-160 (serial order: ascending (pt, k) per channel) */
void oracle_stack_three_interpolate_grad(int n, int c, const float *grad_out, const int *idx, const float *weight,
                                         float *grad_features) {
    for (int pt = 0; pt < n; ++pt)
        for (int k = 0; k < 3; ++k)
            for (int ch = 0; ch < c; ++ch)
                grad_features[(size_t)idx[pt * 3 + k] * c + ch] += grad_out[(size_t)pt * c + ch] * weight[pt * 3 + k];
}

/* voxel_query_gpu.cu:10-88 */
void oracle_stack_voxel_query(int m, int r1, int r2, int r3, int nsample, float radius, int z_range, int y_range,
                              int x_range, const float *new_xyz, const float *xyz, const int *new_coords,
                              const int *point_indices, int *idx) {
    const float radius2 = radius * radius;
    for (int pt = 0; pt < m; ++pt) {
        const float nx = new_xyz[pt * 3], ny = new_xyz[pt * 3 + 1], nz = new_xyz[pt * 3 + 2];
        const int bi = new_coords[pt * 4], cz = new_coords[pt * 4 + 1], cy = new_coords[pt * 4 + 2], cx = new_coords[pt * 4 + 3];
        int *o = idx + (size_t)pt * nsample;
        int cnt = 0;
        for (int dz = -z_range; dz <= z_range; ++dz) {
            const int zc = cz + dz;
            if (zc < 0 || zc >= r1) continue;
            for (int dy = -y_range; dy <= y_range; ++dy) {
                const int yc = cy + dy;
                if (yc < 0 || yc >= r2) continue;
                for (int dx = -x_range; dx <= x_range; ++dx) {
                    const int xc = cx + dx;
                    if (xc < 0 || xc >= r3) continue;
                    const int k = point_indices[(((size_t)bi * r1 + zc) * r2 + yc) * r3 + xc];
                    if (k < 0) continue;
                    const float d2 = sq3(xyz[k * 3] - nx, xyz[k * 3 + 1] - ny, xyz[k * 3 + 2] - nz);
                    if (d2 > radius2) continue;
                    if (cnt < nsample) {
                        if (cnt == 0)
                            for (int l = 0; l < nsample; ++l) o[l] = k;
                        o[cnt] = k;
                        ++cnt;
                    }
                }
            }
        }
        if (cnt == 0) o[0] = -1;
    }
}

static int vp_in_range(float lx, float ly, float lz, float dmax, float r2, int neighbor_type) {
    if (neighbor_type == 1) return !(sq3(lx, ly, lz) > r2);
    return !((fabsf(lx) > dmax) | (fabsf(ly) > dmax) | (fabsf(lz) > dmax));
}

/* vector_pool_gpu.cu:98-160 query_stacked_local_neighbor_idxs_kernel; serial order of centres (a valid outcome of the
 * kernel's atomicAdd hand-out).  Returns the total (what the kernel leaves in *cumsum). */
int oracle_stack_local_neighbor_idxs(int b, int m, const float *support_xyz, const int *xyz_cnt, const float *new_xyz,
                                     const int *new_cnt, int *stack_neighbor_idxs, int *start_len, int avg_length,
                                     float dmax, int nsample, int neighbor_type) {
    const float r2 = dmax * dmax;
    const long long max_thresh = (long long)avg_length * m;
    int cumsum = 0;
    int tmp[1000];
    for (int pt = 0; pt < m; ++pt) {
        const int f = frame_of(new_cnt, b, pt);
        const int xs = start_of(xyz_cnt, f), n = xyz_cnt[f];
        const float *pts = support_xyz + (size_t)xs * 3;
        const float nx = new_xyz[pt * 3], ny = new_xyz[pt * 3 + 1], nz = new_xyz[pt * 3 + 2];
        int cnt = 0;
        for (int k = 0; k < n; ++k) {
            if (!vp_in_range(pts[k * 3] - nx, pts[k * 3 + 1] - ny, pts[k * 3 + 2] - nz, dmax, r2, neighbor_type)) continue;
            if (cnt < 1000) tmp[cnt] = k; else break;
            ++cnt;
            if (nsample > 0 && cnt >= nsample) break;
        }
        const int start = cumsum;
        cumsum += cnt;
        start_len[pt * 2] = start;
        start_len[pt * 2 + 1] = cnt;
        if (start >= max_thresh) continue;
        int w = cnt;
        if ((long long)start + cnt >= max_thresh) w = (int)(max_thresh - start);
        for (int k = 0; k < w; ++k) stack_neighbor_idxs[start + k] = tmp[k] + xs;
    }
    return cumsum;
}

/* vector_pool_gpu.cu:19-77 */
void oracle_stack_three_nn_local(int m, int g, const float *support_xyz, const float *grid_centers, int *grid_idxs,
                                 float *grid_dist2, const int *stack_neighbor_idxs, const int *start_len) {
    for (int pt = 0; pt < m; ++pt)
        for (int gi = 0; gi < g; ++gi) {
            const size_t e = (size_t)pt * g + gi;
            const float cx = grid_centers[e * 3], cy = grid_centers[e * 3 + 1], cz = grid_centers[e * 3 + 2];
            const int *nb = stack_neighbor_idxs + start_len[pt * 2];
            const int len = start_len[pt * 2 + 1];
            double b1 = 1e40, b2 = 1e40, b3 = 1e40;
            int i1 = -1, i2 = -1, i3 = -1;
            for (int k = 0; k < len; ++k) {
                const int q = nb[k];
                const float d = sq3(cx - support_xyz[q * 3], cy - support_xyz[q * 3 + 1], cz - support_xyz[q * 3 + 2]);
                if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = q; }
                else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = q; }
                else if (d < b3) { b3 = d; i3 = q; }
            }
            if (i2 == -1) { i2 = i1; b2 = b1; }
            if (i3 == -1) { i3 = i1; b3 = b1; }
            grid_dist2[e * 3] = (float)b1; grid_dist2[e * 3 + 1] = (float)b2; grid_dist2[e * 3 + 2] = (float)b3;
            grid_idxs[e * 3] = i1; grid_idxs[e * 3 + 1] = i2; grid_idxs[e * 3 + 2] = i3;
        }
}

/* vector_pool_gpu.cu:183-373 vector_pool_kernel_stack + launcher (grid sizes :312-314); serial centre order.
 * Outputs pre-zeroed by the caller.  Returns cum_sum. */
int oracle_stack_vector_pool(int b, int m, int c_in, int c_out, int g, const float *support_xyz, const int *xyz_cnt,
                             const float *support_features, const float *new_xyz, const int *new_cnt, float *new_features,
                             float *new_local_xyz, int *point_cnt_of_grid, int *grouped_idxs, int ngx, int ngy, int ngz,
                             float dmax, int use_xyz, int num_max_sum_points, int nsample, int neighbor_type,
                             int pooling_type) {
    const int ceg = c_out / g;
    const float gsx = dmax * 2 / ngx, gsy = dmax * 2 / ngy, gsz = dmax * 2 / ngz;
    const float r2 = dmax * dmax;
    int cum_sum = 0;
    for (int pt = 0; pt < m; ++pt) {
        const int f = frame_of(new_cnt, b, pt);
        const int xs = start_of(xyz_cnt, f), n = xyz_cnt[f];
        const float *pts = support_xyz + (size_t)xs * 3;
        const float *feat = support_features + (size_t)xs * c_in;
        float *nf = new_features + (size_t)pt * c_out;
        float *nl = new_local_xyz + (size_t)pt * 3 * g;
        int *pc = point_cnt_of_grid + (size_t)pt * g;
        const float nx = new_xyz[pt * 3], ny = new_xyz[pt * 3 + 1], nz = new_xyz[pt * 3 + 2];
        int sample_cnt = 0;
        for (int k = 0; k < n; ++k) {
            const float lx = pts[k * 3] - nx, ly = pts[k * 3 + 1] - ny, lz = pts[k * 3 + 2] - nz;
            if (!vp_in_range(lx, ly, lz, dmax, r2, neighbor_type)) continue;
            const int gxi = (int)floorf((lx + dmax) / gsx), gyi = (int)floorf((ly + dmax) / gsy),
                      gzi = (int)floorf((lz + dmax) / gsz);
            int gi = gxi * ngy * ngz + gyi * ngz + gzi;
            gi = gi < 0 ? 0 : (gi > g - 1 ? g - 1 : gi);
            if (pooling_type == 0) {
                pc[gi]++;
                for (int i = 0; i < c_in; ++i) nf[gi * ceg + i % ceg] += feat[(size_t)k * c_in + i];
                if (use_xyz) { nl[gi * 3] += lx; nl[gi * 3 + 1] += ly; nl[gi * 3 + 2] += lz; }
                const int cnt = cum_sum++;
                if (cnt >= num_max_sum_points) continue;
                grouped_idxs[cnt * 3] = xs + k; grouped_idxs[cnt * 3 + 1] = pt; grouped_idxs[cnt * 3 + 2] = gi;
                sample_cnt++;
                if (nsample > 0 && sample_cnt >= nsample) break;
            } else if (pooling_type == 1) {
                if (pc[gi] == 0) {
                    pc[gi]++;
                    for (int i = 0; i < c_in; ++i) nf[gi * ceg + i % ceg] = feat[(size_t)k * c_in + i];
                    if (use_xyz) { nl[gi * 3] = lx; nl[gi * 3 + 1] = ly; nl[gi * 3 + 2] = lz; }
                    const int cnt = cum_sum++;
                    if (cnt >= num_max_sum_points) continue;
                    grouped_idxs[cnt * 3] = xs + k; grouped_idxs[cnt * 3 + 1] = pt; grouped_idxs[cnt * 3 + 2] = gi;
                    sample_cnt++;
                    if ((nsample > 0 && sample_cnt >= nsample) || sample_cnt >= g) break;
                }
            }
        }
    }
    return cum_sum;
}

/* vector_pool_gpu.cu:376-401 (serial order: ascending (entry, channel)) */
void oracle_stack_vector_pool_grad(int c_out, int c_in, int g, int entries, const float *grad_new_features,
                                   const int *point_cnt_of_grid, const int *grouped_idxs, float *grad_support_features) {
    const int ceg = c_out / g;
    for (int e = 0; e < entries; ++e)
        for (int ch = 0; ch < c_in; ++ch) {
            const int is = grouped_idxs[e * 3], in = grouped_idxs[e * 3 + 1], ig = grouped_idxs[e * 3 + 2];
            const float cur = 1 / fmaxf((float)point_cnt_of_grid[(size_t)in * g + ig], 1.0f);
            grad_support_features[(size_t)is * c_in + ch] += grad_new_features[(size_t)in * c_out + (size_t)ig * ceg + ch % ceg] * cur;
        }
}
