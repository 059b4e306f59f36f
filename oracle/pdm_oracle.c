/*
 * oracle/pdm_oracle.c -- CPU restatement of the reference's pointnet2_batch ops.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package (pdm_ssd_b200/) may
 * import, link or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker and as
 * the reported CPU baseline.
 *
 * Parity status: PINNED.  Every function below is checked bit-for-bit (ints) or
 * to the stated tolerance against the reference's own CUDA kernels, compiled
 * unmodified from /root/reference into oracle/_ref/ (oracle/build_ref.py) and
 * executed on a B200; the resulting golden vectors live in tests/golden/ with
 * the generating script tests/golden/make_golden.py.
 *
 * Arithmetic facts (read from the SASS of the reference built with nvcc 12.9
 * for sm_100, cuobjdump -sass on oracle/_ref/ objects): every squared distance
 *      (ax-bx)*(ax-bx) + (ay-by)*(ay-by) + (az-bz)*(az-bz)
 * is emitted as  t = rn(dy*dy); t = fma(dx,dx,t); d = fma(dz,dz,t)
 * and the 3-point interpolation as  fma(w2,p2, fma(w0,p0, rn(w1*p1))).
 * This file spells those with fmaf() and must be built with
 * -ffp-contract=off so the compiler adds no contraction of its own.
 *
 * Each function cites the reference file:line it follows
 * (paths relative to pcdet/ops/pointnet2/pointnet2_batch/src/).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>

/* squared distance with the reference's contraction pattern (see header) */
static inline float sqdist_ref(float dx, float dy, float dz) {
    float t = dy * dy;
    t = fmaf(dx, dx, t);
    return fmaf(dz, dz, t);
}

/* cuda_utils.h:10-14 opt_n_threads: 2^floor(log2 n) clipped to [1,1024], computed
 * through double log() exactly as the reference host code does. */
int oracle_opt_n_threads(int work_size) {
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int v = 1 << pow_2;
    if (v > 1024) v = 1024;
    if (v < 1) v = 1;
    return v;
}

/* ---- tiny pthread parallel-for (libgomp is not in this image) -------------------
 * Work items are handed out through an atomic counter; results do not depend on the
 * thread count because every item writes a disjoint output range. */
static int g_threads = 1;
void oracle_set_threads(int nthreads) { g_threads = nthreads > 0 ? nthreads : 1; }
int oracle_max_threads(void) { return g_threads; }

typedef void (*item_fn)(int item, void *ctx);
typedef struct { item_fn fn; void *ctx; int n; int next; } par_t;
static void *par_worker(void *arg) {
    par_t *p = (par_t *)arg;
    for (;;) {
        int i = __atomic_fetch_add(&p->next, 1, __ATOMIC_RELAXED);
        if (i >= p->n) break;
        p->fn(i, p->ctx);
    }
    return NULL;
}
static void par_for(int n, item_fn fn, void *ctx) {
    int nt = g_threads < n ? g_threads : n;
    par_t p = {fn, ctx, n, 0};
    if (nt <= 1) { par_worker(&p); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nt);
    for (int t = 1; t < nt; ++t) pthread_create(&th[t], NULL, par_worker, &p);
    par_worker(&p);
    for (int t = 1; t < nt; ++t) pthread_join(th[t], NULL);
    free(th);
}

/*
 * sampling_gpu.cu:100-216 farthest_point_sampling_kernel<block_size>, launched
 * <<<b, opt_n_threads(n)>>> (sampling_gpu.cu:218-260).
 *
 * Literal simulation: `bs` virtual threads, each owning k = tid, tid+bs, ...
 * with strict `>` (first k wins inside a thread, :143-144), then the shared
 * memory tree of __update (:93-98) where the LEFT slot survives ties.
 * temp is read/written exactly as the kernel does (caller pre-fills 1e10,
 * pointnet2_utils.py:26).
 */
typedef struct { int n, m, bs; const float *xyz; float *temp; int *idxs; } fps_ctx;
static void fps_item(int bi, void *vc) {
    const fps_ctx *c = (const fps_ctx *)vc;
    const int n = c->n, m = c->m, bs = c->bs;
    const float *dataset = c->xyz + (size_t)bi * n * 3;
    float *tmp = c->temp + (size_t)bi * n;
    int *out = c->idxs + (size_t)bi * m;
    float *dists = (float *)malloc(sizeof(float) * bs);
    int *dists_i = (int *)malloc(sizeof(int) * bs);
    int old = 0;
    out[0] = old;
    for (int j = 1; j < m; ++j) {
        const float x1 = dataset[old * 3 + 0];
        const float y1 = dataset[old * 3 + 1];
        const float z1 = dataset[old * 3 + 2];
        /* virtual thread tid owns k = tid, tid+bs, ...; walking k upward visits each
         * thread's points in the same ascending order as the kernel's strided loop */
        for (int tid = 0; tid < bs; ++tid) { dists[tid] = -1.0f; dists_i[tid] = 0; }
        for (int k = 0; k < n; ++k) {
            const int tid = k & (bs - 1);
            const float x2 = dataset[k * 3 + 0];
            const float y2 = dataset[k * 3 + 1];
            const float z2 = dataset[k * 3 + 2];
            const float d = sqdist_ref(x2 - x1, y2 - y1, z2 - z1);
            const float d2 = fminf(d, tmp[k]);
            tmp[k] = d2;
            if (d2 > dists[tid]) { dists[tid] = d2; dists_i[tid] = k; }
        }
        for (int half = bs >> 1; half >= 1; half >>= 1) {
            for (int tid = 0; tid < half; ++tid) {
                const float v1 = dists[tid], v2 = dists[tid + half];
                const int i1 = dists_i[tid], i2 = dists_i[tid + half];
                dists[tid] = fmaxf(v1, v2);
                dists_i[tid] = v2 > v1 ? i2 : i1;
            }
        }
        old = dists_i[0];
        out[j] = old;
    }
    free(dists);
    free(dists_i);
}
void oracle_fps(int b, int n, int m, const float *xyz, float *temp, int *idxs) {
    if (m <= 0) return;
    fps_ctx c = {n, m, oracle_opt_n_threads(n), xyz, temp, idxs};
    par_for(b, fps_item, &c);
}

/* sampling_gpu.cu:15-31 gather_points_kernel_fast: out[b,c,m] = points[b,c,idx[b,m]] */
void oracle_gather_points(int b, int c, int n, int m, const float *points, const int *idx, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *p = points + ((size_t)bi * c + ci) * n;
            const int *id = idx + (size_t)bi * m;
            float *o = out + ((size_t)bi * c + ci) * m;
            for (int j = 0; j < m; ++j) o[j] = p[id[j]];
        }
}

/* sampling_gpu.cu:53-70 gather_points_grad_kernel_fast: grad_points[b,c,idx[b,m]] += grad_out[b,c,m]
 * (reference uses atomicAdd => order undefined; the oracle sums in ascending m). */
void oracle_gather_points_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                               float *grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *g = grad_out + ((size_t)bi * c + ci) * m;
            const int *id = idx + (size_t)bi * m;
            float *gp = grad_points + ((size_t)bi * c + ci) * n;
            for (int j = 0; j < m; ++j) gp[id[j]] += g[j];
        }
}

/*
 * ball_query_gpu.cu:15-51 ball_query_kernel_fast.  One virtual thread per
 * centre; scan k ascending; d2 = (new-x)^2.. with the contraction pattern;
 * first hit floods all nsample slots (:41-45); stop at nsample hits.
 * idx is NOT cleared here: the caller zero-fills (pointnet2_utils.py:218),
 * rows without any hit keep whatever the caller put there.
 */
typedef struct { int n, m, nsample; float radius; const float *new_xyz, *xyz; int *idx; } bq_ctx;
#define BQ_CHUNK 64
static void bq_item(int item, void *vc) {
    const bq_ctx *c = (const bq_ctx *)vc;
    const int n = c->n, m = c->m, nsample = c->nsample;
    const int chunks = (m + BQ_CHUNK - 1) / BQ_CHUNK;
    const int bi = item / chunks, p0 = (item % chunks) * BQ_CHUNK;
    const int p1 = p0 + BQ_CHUNK < m ? p0 + BQ_CHUNK : m;
    const float radius2 = c->radius * c->radius;
    const float *pts = c->xyz + (size_t)bi * n * 3;
    for (int pi = p0; pi < p1; ++pi) {
        const float *q = c->new_xyz + ((size_t)bi * m + pi) * 3;
        int *o = c->idx + ((size_t)bi * m + pi) * nsample;
        const float nx = q[0], ny = q[1], nz = q[2];
        int cnt = 0;
        for (int k = 0; k < n; ++k) {
            const float d2 = sqdist_ref(nx - pts[k * 3 + 0], ny - pts[k * 3 + 1], nz - pts[k * 3 + 2]);
            if (d2 < radius2) {
                if (cnt == 0)
                    for (int l = 0; l < nsample; ++l) o[l] = k;
                o[cnt] = k;
                ++cnt;
                if (cnt >= nsample) break;
            }
        }
    }
}
void oracle_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                       const float *xyz, int *idx) {
    bq_ctx c = {n, m, nsample, radius, new_xyz, xyz, idx};
    par_for(b * ((m + BQ_CHUNK - 1) / BQ_CHUNK), bq_item, &c);
}

/* group_points_gpu.cu:53-72 group_points_kernel_fast: out[b,c,p,s] = points[b,c,idx[b,p,s]] */
typedef struct { int c, n, npoints, nsample; const float *points; const int *idx; float *out; } grp_ctx;
static void grp_item(int item, void *vc) {
    const grp_ctx *g = (const grp_ctx *)vc;
    const int bi = item / g->c, ci = item % g->c;
    const size_t ms = (size_t)g->npoints * g->nsample;
    const float *p = g->points + ((size_t)bi * g->c + ci) * g->n;
    const int *id = g->idx + (size_t)bi * ms;
    float *o = g->out + ((size_t)bi * g->c + ci) * ms;
    for (size_t j = 0; j < ms; ++j) o[j] = p[id[j]];
}
void oracle_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                         const int *idx, float *out) {
    grp_ctx g = {c, n, npoints, nsample, points, idx, out};
    par_for(b * c, grp_item, &g);
}

/* group_points_gpu.cu:14-31 group_points_grad_kernel_fast (atomicAdd in the reference;
 * ascending (p,s) order here). */
void oracle_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                              const int *idx, float *grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *g = grad_out + ((size_t)bi * c + ci) * npoints * nsample;
            const int *id = idx + (size_t)bi * npoints * nsample;
            float *gp = grad_points + ((size_t)bi * c + ci) * n;
            for (int j = 0; j < npoints * nsample; ++j) gp[id[j]] += g[j];
        }
}

/*
 * interpolate_gpu.cu:16-59 three_nn_kernel_fast.  Bests are doubles initialised
 * to 1e40 (:36), the float distance is compared after promotion, strict `<`
 * (earliest index wins ties), results are narrowed to float on store (:57)
 * (1e40 -> +inf when fewer than 3 known points exist).
 */
typedef struct { int n, m; const float *unknown, *known; float *dist2; int *idx; } nn_ctx;
static void nn_item(int item, void *vc) {
    const nn_ctx *c = (const nn_ctx *)vc;
    const int n = c->n, m = c->m;
    const int chunks = (n + BQ_CHUNK - 1) / BQ_CHUNK;
    const int bi = item / chunks, p0 = (item % chunks) * BQ_CHUNK;
    const int p1 = p0 + BQ_CHUNK < n ? p0 + BQ_CHUNK : n;
    const float *kn = c->known + (size_t)bi * m * 3;
    for (int pi = p0; pi < p1; ++pi) {
        const float *u = c->unknown + ((size_t)bi * n + pi) * 3;
        const float ux = u[0], uy = u[1], uz = u[2];
        double best1 = 1e40, best2 = 1e40, best3 = 1e40;
        int besti1 = 0, besti2 = 0, besti3 = 0;
        for (int k = 0; k < m; ++k) {
            const float d = sqdist_ref(ux - kn[k * 3 + 0], uy - kn[k * 3 + 1], uz - kn[k * 3 + 2]);
            if (d < best1) {
                best3 = best2; besti3 = besti2;
                best2 = best1; besti2 = besti1;
                best1 = d; besti1 = k;
            } else if (d < best2) {
                best3 = best2; besti3 = besti2;
                best2 = d; besti2 = k;
            } else if (d < best3) {
                best3 = d; besti3 = k;
            }
        }
        float *od = c->dist2 + ((size_t)bi * n + pi) * 3;
        int *oi = c->idx + ((size_t)bi * n + pi) * 3;
        od[0] = (float)best1; od[1] = (float)best2; od[2] = (float)best3;
        oi[0] = besti1; oi[1] = besti2; oi[2] = besti3;
    }
}
void oracle_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                     int *idx) {
    nn_ctx c = {n, m, unknown, known, dist2, idx};
    par_for(b * ((n + BQ_CHUNK - 1) / BQ_CHUNK), nn_item, &c);
}

/* interpolate_gpu.cu:84-104 three_interpolate_kernel_fast:
 * out[b,c,n] = w0*p[i0] + w1*p[i1] + w2*p[i2] as fma(w2,p2, fma(w0,p0, rn(w1*p1))). */
void oracle_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                              const float *weight, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *p = points + ((size_t)bi * c + ci) * m;
            float *o = out + ((size_t)bi * c + ci) * n;
            for (int j = 0; j < n; ++j) {
                const float *w = weight + ((size_t)bi * n + j) * 3;
                const int *id = idx + ((size_t)bi * n + j) * 3;
                float t = w[1] * p[id[1]];
                t = fmaf(w[0], p[id[0]], t);
                o[j] = fmaf(w[2], p[id[2]], t);
            }
        }
}

/* interpolate_gpu.cu:127-149 three_interpolate_grad_kernel_fast (atomicAdd in the
 * reference; ascending n then k order here). */
void oracle_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                   const float *weight, float *grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *g = grad_out + ((size_t)bi * c + ci) * n;
            float *gp = grad_points + ((size_t)bi * c + ci) * m;
            for (int j = 0; j < n; ++j) {
                const float *w = weight + ((size_t)bi * n + j) * 3;
                const int *id = idx + ((size_t)bi * n + j) * 3;
                gp[id[0]] += g[j] * w[0];
                gp[id[1]] += g[j] * w[1];
                gp[id[2]] += g[j] * w[2];
            }
        }
}

/* =============================================================================================
 * Rotated BEV IoU and greedy NMS -- restatement of pcdet/ops/iou3d_nms
 * (paths relative to pcdet/ops/iou3d_nms/src/).  PINNED against the reference's own kernels
 * (oracle/_ref/iou3d_nms_cuda_ref.so) on a B200: IoU to 1e-5 absolute (the device cos/sin/atan2 and
 * nvcc's FMA contraction differ from libm in the last bits), keep lists identical on the
 * committed golden cases (tests/golden/nms_*.npz).
 * ============================================================================================= */
typedef struct { float x, y; } pt2;

/* iou3d_nms_kernel.cu:39-41 */
static inline float nms_cross3(pt2 p1, pt2 p2, pt2 p0) {
    return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}

/* iou3d_nms_kernel.cu:51-60 check_in_box2d: point inside the box grown by a 1e-2 margin */
static int nms_in_box(const float *box, pt2 p) {
    const float MARGIN = 1e-2f;
    const float c = cosf(-box[6]), s = sinf(-box[6]);
    const float rx = (p.x - box[0]) * c + (p.y - box[1]) * (-s);
    const float ry = (p.x - box[0]) * s + (p.y - box[1]) * c;
    return fabsf(rx) < box[3] / 2 + MARGIN && fabsf(ry) < box[4] / 2 + MARGIN;
}

/* iou3d_nms_kernel.cu:62-91 intersection (with check_rect_cross :43-49) */
static int nms_intersection(pt2 p1, pt2 p0, pt2 q1, pt2 q0, pt2 *ans) {
    const float EPS = 1e-8f;
    if (!(fminf(p0.x, p1.x) <= fmaxf(q0.x, q1.x) && fminf(q0.x, q1.x) <= fmaxf(p0.x, p1.x) &&
          fminf(p0.y, p1.y) <= fmaxf(q0.y, q1.y) && fminf(q0.y, q1.y) <= fmaxf(p0.y, p1.y)))
        return 0;
    const float s1 = nms_cross3(q0, p1, p0), s2 = nms_cross3(p1, q1, p0);
    const float s3 = nms_cross3(p0, q1, q0), s4 = nms_cross3(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
    const float s5 = nms_cross3(q1, p1, p0);
    if (fabsf(s5 - s1) > EPS) {
        ans->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        ans->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        const float D = a0 * b1 - a1 * b0;
        ans->x = (b0 * c1 - b1 * c0) / D;
        ans->y = (a1 * c0 - a0 * c1) / D;
    }
    return 1;
}

static void nms_corners(const float *box, pt2 *c) {
    /* iou3d_nms_kernel.cu:104-147: axis-aligned corners, then rotate_around_center (:93-97) */
    const float hx = box[3] / 2, hy = box[4] / 2;
    const float cs = cosf(box[6]), sn = sinf(box[6]);
    const float px[4] = {box[0] - hx, box[0] + hx, box[0] + hx, box[0] - hx};
    const float py[4] = {box[1] - hy, box[1] - hy, box[1] + hy, box[1] + hy};
    for (int k = 0; k < 4; ++k) {
        c[k].x = (px[k] - box[0]) * cs + (py[k] - box[1]) * (-sn) + box[0];
        c[k].y = (px[k] - box[0]) * sn + (py[k] - box[1]) * cs + box[1];
    }
    c[4] = c[0];
}

/* iou3d_nms_kernel.cu:99-222 box_overlap */
static float nms_box_overlap(const float *a, const float *b) {
    pt2 ca[5], cb[5], pts[16], ctr = {0.f, 0.f};
    int cnt = 0;
    nms_corners(a, ca);
    nms_corners(b, cb);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (nms_intersection(ca[i + 1], ca[i], cb[j + 1], cb[j], &pts[cnt])) {
                ctr.x += pts[cnt].x; ctr.y += pts[cnt].y;
                ++cnt;
            }
    for (int k = 0; k < 4; ++k) {
        if (nms_in_box(a, cb[k])) { ctr.x += cb[k].x; ctr.y += cb[k].y; pts[cnt++] = cb[k]; }
        if (nms_in_box(b, ca[k])) { ctr.x += ca[k].x; ctr.y += ca[k].y; pts[cnt++] = ca[k]; }
    }
    ctr.x /= cnt; ctr.y /= cnt;
    for (int j = 0; j < cnt - 1; ++j)        /* bubble sort by atan2 around the centroid (:189-199) */
        for (int i = 0; i < cnt - j - 1; ++i)
            if (atan2f(pts[i].y - ctr.y, pts[i].x - ctr.x) > atan2f(pts[i + 1].y - ctr.y, pts[i + 1].x - ctr.x)) {
                pt2 t = pts[i]; pts[i] = pts[i + 1]; pts[i + 1] = t;
            }
    float area = 0.f;
    for (int k = 0; k < cnt - 1; ++k) {
        const float ux = pts[k].x - pts[0].x, uy = pts[k].y - pts[0].y;
        const float vx = pts[k + 1].x - pts[0].x, vy = pts[k + 1].y - pts[0].y;
        area += ux * vy - uy * vx;
    }
    return fabsf(area) / 2.0f;
}

/* iou3d_nms_kernel.cu:224-232 iou_bev */
static float nms_iou_bev(const float *a, const float *b) {
    const float sa = a[3] * a[4], sb = b[3] * b[4];
    const float s = nms_box_overlap(a, b);
    return s / fmaxf(sa + sb - s, 1e-8f);
}

/* iou3d_nms_kernel.cu:279-293 boxes_iou_bev_kernel */
void oracle_boxes_iou_bev(int na, const float *a, int nb, const float *b, float *ans) {
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) ans[(size_t)i * nb + j] = nms_iou_bev(a + i * 7, b + j * 7);
}

/* nms_kernel (iou3d_nms_kernel.cu:295-341) + host greedy loop (iou3d_nms.cpp:159-176):
 * boxes are already sorted by score; returns the number kept, keep[] = positions in that order. */
int oracle_nms_bev(int n, const float *boxes, float thresh, int *keep) {
    unsigned char *removed = (unsigned char *)calloc(n > 0 ? n : 1, 1);
    int kept = 0;
    for (int i = 0; i < n; ++i) {
        if (removed[i]) continue;
        keep[kept++] = i;
        for (int j = i + 1; j < n; ++j)
            if (!removed[j] && nms_iou_bev(boxes + i * 7, boxes + j * 7) > thresh) removed[j] = 1;
    }
    free(removed);
    return kept;
}
