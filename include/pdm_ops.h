/*
 * pdm_ops.h -- C ABI of libpdmops.so, the B200 (sm_100a) implementation of the
 * PDM-SSD / OpenPCDet point-processing hot path.
 *
 * Every entry point takes plain device pointers + sizes + the CUDA stream to launch
 * on (a cudaStream_t passed as void*; NULL = legacy default stream, which is what
 * the reference launches on).  No torch types cross this boundary.  All tensors are
 * dense, row-major ("contiguous" in torch terms), fp32 / int32, resident on the
 * CURRENT CUDA device of the calling thread.  Outputs (and the FPS scratch `temp`)
 * are allocated by the caller, exactly as in the reference (pointnet2_utils.py:25-26,
 * 55,94-95,128,172,218); the library only writes into them.
 *
 * Return value: 0 on success, otherwise non-zero (a cudaError_t value, or
 * PDM_ERR_* below); pdm_last_error() then returns a thread-local message.  The
 * library never calls exit() (the reference does: sampling_gpu.cu:46-50).
 *
 * Each declaration cites the reference binding it replaces: the pybind entry in
 * pcdet/ops/pointnet2/pointnet2_batch/src/pointnet2_api.cpp and the C++ wrapper
 * behind it.  INTEGRATION.md shows the stub a pcdet maintainer adds to bind these.
 */
#ifndef PDM_OPS_H_
#define PDM_OPS_H_

#ifdef __cplusplus
extern "C" {
#endif

#define PDM_OK 0
#define PDM_ERR_INVALID_ARG (-1)   /* negative size, NULL pointer with non-empty tensor, ... */
#define PDM_ERR_UNSUPPORTED (-2)   /* shape outside what the kernels implement */

/* ABI version, bumped on any signature change. */
int pdm_abi_version(void);
/* Message for the last non-zero return on this thread ("" if none). */
const char *pdm_last_error(void);
/* Number of kernel launches issued by this library since load / since the last reset
 * (bench.py reports it as gpu_launches). */
long long pdm_launch_count(void);
void pdm_reset_launch_count(void);

/* ---- set-abstraction ops (reference: pointnet2_batch) ----------------------------- */

/* farthest_point_sampling_wrapper (pointnet2_api.cpp:18, sampling.cpp:37-46,
 * sampling_gpu.cu:100-260).  xyz (B,N,3) -> idx (B,M); temp (B,N) is caller scratch
 * pre-filled with 1e10; on return it holds the running min distances exactly as the
 * reference leaves them.  idx[:,0] = 0; ties resolved like the reference's
 * block_size-dependent shared-memory tournament (DESIGN.md "FPS tie-break"). */
int pdm_farthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx,
                                void *stream);

/* Scheduling hint for pdm_farthest_point_sampling (process-wide default; no reference counterpart -- the
 * reference has one kernel).  Results are bit-identical in every mode.
 *   AUTO:       on-chip kernel (one frame per SM, lowest latency) unless one call has more frames than SMs
 *   LATENCY:    always the on-chip kernel
 *   THROUGHPUT: coordinates stay in L2, 2-3 frames share an SM: ~20 % longer per frame, about twice the
 *               frames per second when many batches are in flight on several streams */
#define PDM_FPS_MODE_AUTO 0
#define PDM_FPS_MODE_LATENCY 1
#define PDM_FPS_MODE_THROUGHPUT 2
int pdm_set_fps_mode(int mode);
/* The same sampling with the scheduling mode given per call (no process-wide state: two pipelines or threads in
 * one process can use different modes); pdm_set_fps_mode only sets the default of the plain entry above. */
int pdm_farthest_point_sampling_ex(int b, int n, int m, const float *xyz, float *temp, int *idx, int mode,
                                   void *stream);

/* gather_points_wrapper (pointnet2_api.cpp:16, sampling.cpp:14-23, sampling_gpu.cu:15-51).
 * points (B,C,N), idx (B,M) -> out (B,C,M). */
int pdm_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx,
                      float *out, void *stream);

/* gather_points_grad_wrapper (pointnet2_api.cpp:17, sampling.cpp:25-35, sampling_gpu.cu:53-91).
 * grad_out (B,C,M), idx (B,M) -> grad_points (B,C,N) accumulated (+=). */
int pdm_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int *idx,
                           float *grad_points, void *stream);

/* ball_query_wrapper (pointnet2_api.cpp:13, ball_query.cpp:29-39, ball_query_gpu.cu:15-73).
 * new_xyz (B,M,3), xyz (B,N,3) -> idx (B,M,nsample): first nsample indices k (ascending)
 * with d2 < radius*radius, padded with the first hit; rows without a hit are left
 * untouched (the caller zero-fills, pointnet2_utils.py:218). */
int pdm_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                   const float *xyz, int *idx, void *stream);
/* The same query with the scheduling mode of the caller given per call (PDM_FPS_MODE_*; results do not depend on it):
 * LATENCY / AUTO build the lookup grid with a cluster of CTAs per frame (shortest time for a batch alone),
 * THROUGHPUT with one CTA per frame (no co-scheduling constraint: best when many batches are in flight). */
int pdm_ball_query_ex(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                      const float *xyz, int *idx, int mode, void *stream);

/* group_points_wrapper (pointnet2_api.cpp:14, group_points.cpp:27-37, group_points_gpu.cu:53-92).
 * points (B,C,N), idx (B,npoints,nsample) -> out (B,C,npoints,nsample). */
int pdm_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                     const int *idx, float *out, void *stream);

/* group_points_grad_wrapper (pointnet2_api.cpp:15, group_points.cpp:15-25,
 * group_points_gpu.cu:14-51).  grad_out (B,C,npoints,nsample) -> grad_points (B,C,N) (+=). */
int pdm_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                          const int *idx, float *grad_points, void *stream);

/* three_nn_wrapper (pointnet2_api.cpp:19, interpolate.cpp:18-27, interpolate_gpu.cu:16-81).
 * unknown (B,N,3), known (B,M,3) -> dist2 (B,N,3) squared distances, idx (B,N,3). */
int pdm_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                 int *idx, void *stream);

/* three_interpolate_wrapper (pointnet2_api.cpp:20, interpolate.cpp:29-42,
 * interpolate_gpu.cu:84-124).  points (B,C,M), idx (B,N,3), weight (B,N,3) -> out (B,C,N). */
int pdm_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                          const float *weight, float *out, void *stream);

/* three_interpolate_grad_wrapper (pointnet2_api.cpp:21, interpolate.cpp:44-56,
 * interpolate_gpu.cu:127-169).  grad_out (B,C,N) -> grad_points (B,C,M) (+=). */
int pdm_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                               const float *weight, float *grad_points, void *stream);

/* Deterministic forms of the three gradient kernels above (same arguments, same "+=" contract).  The reference
 * accumulates with atomicAdd (sampling_gpu.cu:53-70, group_points_gpu.cu:14-31, interpolate_gpu.cu:127-149), so the
 * fp32 summation order changes from run to run; here every target sums its contributions in ascending source position
 * (stable sort by target, then a serial in-order sum per target): bit-identical results on every run. */
int pdm_gather_points_grad_det(int b, int c, int n, int npoints, const float *grad_out, const int *idx,
                               float *grad_points, void *stream);
int pdm_group_points_grad_det(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                              const int *idx, float *grad_points, void *stream);
int pdm_three_interpolate_grad_det(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                   const float *weight, float *grad_points, void *stream);

/* QueryAndGroup.forward after the ball query in one pass (pointnet2_utils.py:250-257: xyz^T copy,
 * group xyz, subtract centres, group features, cat).  xyz (B,N,3), new_xyz (B,npoints,3),
 * features (B,C,N) or NULL when c == 0, idx (B,npoints,nsample) ->
 * out (B, 3*use_xyz + C, npoints, nsample) with channel order [xyz - centre, features]. */
int pdm_query_and_group(int b, int c, int n, int npoints, int nsample, int use_xyz, const float *xyz,
                        const float *new_xyz, const float *features, const int *idx, float *out,
                        void *stream);

/* ---- fused set-abstraction scale ---------------------------------------------------------- */

/* One scale of _PointnetSAModuleBase.forward in a single kernel (inference): replaces
 * QueryAndGroup.forward after the ball query (pointnet2_utils.py:250-257: group xyz, subtract the
 * centre, group features, cat), the shared MLP `self.mlps[i]` (Conv2d 1x1 bias=False + BatchNorm2d
 * (eval) + ReLU per layer, pointnet2_modules.py:40,90-97) and the max-pool over nsample
 * (pointnet2_modules.py:41-52).
 *   xyz (B,N,3), features (B,c_feat,N) or NULL when c_feat == 0, new_xyz (B,M,3),
 *   idx (B,M,nsample) from pdm_ball_query; widths[n_layers+1] host array with
 *   widths[0] = 3*use_xyz + c_feat; packed = per layer l the BN-folded weights transposed
 *   Wt[k][pad4(widths[l+1])] (k < widths[l], zero-padded columns) followed by the folded bias
 *   [pad4(widths[l+1])], all layers back to back (device, fp32)
 *   packed_tc (optional, may be NULL) = the same weights prepared for the tcgen05 tensor-core
 *   kernel: bias[n_layers][128], then per layer and per block of <= 64 output columns
 *   W_hi[kpad/4][ns][4], W_lo[kpad/4][ns][4] (tf32-exact high part and fp32 remainder, K padded
 *   to 8, N to 16, K-major core-matrix layout); when given and the scale fits, the layers run as
 *   tcgen05.mma.kind::tf32 with hi*hi + lo*hi + hi*lo (fp32 accuracy), else on the CUDA cores
 *   -> out (B, widths[n_layers], M).
 * Limits: n_layers <= 4, every width <= 128, nsample a power of two in 4..128. */
int pdm_sa_fused_forward(int b, int n, int m, int c_feat, int nsample, int use_xyz, const float *xyz,
                         const float *features, const float *new_xyz, const int *idx, int n_layers,
                         const int *widths, const float *packed, const float *packed_tc, float *out,
                         void *stream);

/* pdm_sa_fused_forward with the optional operands of the faster kernels (all may be NULL):
 *   features_pm  the same features point-major (B, N, c_feat): the gather reads a row's channels from one place
 *   packed_tc3 / bias_tc3  operands of the persistent tcgen05 kernel (csrc/sa_tc.cu): per layer the BN-folded
 *                weights W'[n][k] as two bf16 planes hi|lo (hi + lo = w to 2^-18), each plane
 *                [kpad/8][npad][8] with kpad, npad = widths rounded up to 16, layers back to back; bias fp32 [n_layers][128]
 *   out_pm       the result also written point-major (B, M, widths[n_layers]) -- what the next layer's gather and the
 *                detector's `point_features` (pointnet2_backbone.py:91-92) read
 * Kernel choice per scale: thread-per-row (narrow MLPs, csrc/sa_rows.cu), persistent tcgen05 (nsample 32, hidden widths
 * <= 64, output <= 128), round 1's tcgen05 kernel, CUDA cores. */
int pdm_sa_fused_forward_v2(int b, int n, int m, int c_feat, int nsample, int use_xyz, const float *xyz,
                            const float *features, const float *features_pm, const float *new_xyz, const int *idx,
                            int n_layers, const int *widths, const float *packed, const float *packed_tc,
                            const void *packed_tc3, const float *bias_tc3, float *out, float *out_pm, void *stream);

/* ---- PDM neck (SPEC_PDM.md) ---------------------------------------------------------- */

/* Point dilation + SH/Gaussian feature filling + multi-centre fusion + height compression.
 * The reference tree has NO code for this stage, so there is no reference
 * binding to cite; the entry implements SPEC_PDM.md and fills the pcdet `map_to_bev` slot
 * (detector3d_template.py:85-95; output contract of height_compression.py:22-25).
 *   point_coords (P,4) [b,x,y,z] grouped by b ascending, point_features (P,C),
 *   coef (P,(sh_degree+1)^2) per-centre SH coefficients,
 *   range_min[3], voxel[3], grid[3]=(X,Y,Z), dilation[3]=(kx,ky,kz) host arrays,
 *   -> spatial_features (B,C,Y,X), every element written (zeros for empty pillars).
 * dbg_keys / dbg_w (optional, may be NULL): (P,K) int32 cell keys (-1 = dropped) and fp32 weights
 * of every (centre, offset) entry, K = (2kx+1)(2ky+1)(2kz+1), for parity tests. */
int pdm_neck_forward(int batch, int p, int c, const float *point_coords, const float *point_features,
                     const float *coef, const float *range_min, const float *voxel, const int *grid,
                     const int *dilation, int sh_degree, float sigma, float eps,
                     float *spatial_features, int *dbg_keys, float *dbg_w, void *stream);

/* ---- rotated BEV IoU / NMS (reference: pcdet/ops/iou3d_nms) ----------------------------------- */

/* boxes_iou_bev_gpu (iou3d_nms_api.cpp:15, iou3d_nms.cpp:113-135, iou3d_nms_kernel.cu:279-293).
 * boxes_a (Na,7), boxes_b (Nb,7) [x,y,z,dx,dy,dz,heading] -> ans_iou (Na,Nb). */
int pdm_boxes_iou_bev(int na, const float *boxes_a, int nb, const float *boxes_b, float *ans_iou,
                      void *stream);

/* boxes_overlap_bev_gpu (iou3d_nms_api.cpp:13, iou3d_nms.cpp:68-90, iou3d_nms_kernel.cu:236-249):
 * overlap area instead of IoU, same shapes (used by boxes_iou3d_gpu, iou3d_nms_utils.py:47-82). */
int pdm_boxes_overlap_bev(int na, const float *boxes_a, int nb, const float *boxes_b,
                          float *ans_overlap, void *stream);

/* nms_gpu (iou3d_nms_api.cpp:16, iou3d_nms.cpp:137-183, iou3d_nms_kernel.cu:295-341) for ALL frames
 * of a batch in two launches and without the reference's per-call cudaMalloc, blocking D2H copy
 * and host greedy loop: the greedy pass runs on the device.
 *   boxes (frames,k,7) already sorted by descending score within each frame (the reference sorts
 *   in torch first, iou3d_nms_utils.py:128-133); counts (frames) = valid boxes per frame, or NULL
 *   when every frame has k; -> keep (frames,k) int32: positions of the kept boxes in ascending
 *   order (= the reference's keep[:num_out]), padded with -1; num_keep (frames).
 * Limit: k <= 16384 (the reference takes any N; stock configs use NMS_PRE_MAXSIZE <= 9000). */
int pdm_nms_bev_batched(int frames, int k, const float *boxes, const int *counts, float thresh,
                        int *keep, int *num_keep, void *stream);

/* nms_normal_gpu (iou3d_nms_api.cpp:17, iou3d_nms.cpp:186-233, iou3d_nms_kernel.cu:341-398): the same greedy
 * suppression on the axis-aligned BEV IoU (heading ignored); arguments as pdm_nms_bev_batched. */
int pdm_nms_normal_batched(int frames, int k, const float *boxes, const int *counts, float thresh,
                           int *keep, int *num_keep, void *stream);

/* paired_boxes_overlap_bev_gpu / boxes_aligned_overlap_bev_gpu (iou3d_nms_api.cpp:12,14, iou3d_nms.cpp:42-66,92-111,
 * iou3d_nms_kernel.cu:251-277): overlap area of box i of boxes_a with box i of boxes_b, (N,7) x (N,7) -> (N). */
int pdm_boxes_overlap_bev_paired(int n, const float *boxes_a, const float *boxes_b, float *ans_overlap, void *stream);

/* ---- dense BEV layers on the tensor cores (csrc/conv_tc.cu) ------------------------------------ */

/* Activation layout between the dense layers, "split NHWC8": a bf16 tensor (2, B, Y, C/8, X, 8), plane 0 =
 * bf16(v), plane 1 = bf16(v - plane0) (hi + lo carries an fp32 value to 2^-17).  C must be a multiple of 8.
 * pdm_act_split_bytes gives the buffer size; the two converters are for API edges and tests (inside the
 * detector the producing kernels write the split form themselves). */
int pdm_act_split_bytes(int b, int c, int y, int x, long long *bytes);
int pdm_act_split_from_nchw(int b, int c, int y, int x, const float *in, void *out_split, void *stream);
int pdm_act_split_to_nchw(int b, int c, int y, int x, const void *in_split, float *out, void *stream);

/* Conv2d(k x k, stride 1, padding (k-1)/2) + folded eval-mode BatchNorm + activation as a tcgen05 implicit
 * GEMM with TMA-staged operands; replaces the cuDNN stacks of BaseBEVBackbone's stride-1 block
 * (base_bev_backbone.py:27-47: ZeroPad2d + Conv2d 3x3 bias=False + BatchNorm2d(eps 1e-3) + ReLU, repeated) and of
 * SeparateHead / the shared conv of CenterHead (center_head.py:12-46, 63-72: Conv2d 3x3 + BN + ReLU ... Conv2d 3x3
 * bias=True) at inference.
 *   in_split: split NHWC8 activation (b, cin, y, x); cin a multiple of 32
 *   w_packed: bf16 [cin/32][k*k][hi|lo][4][npad][8] = BN-folded weights W'[n][c][ky][kx] split like the
 *             activations, c = 32*i0 + 8*i3 + i5, tap = ky*k + kx, npad = cout rounded up to 16/32/64/128 (zero rows)
 *   bias:     fp32 [npad] folded bias;  act: 0 none, 1 ReLU, 2 sigmoid
 *   out_split (optional): split NHWC8 (b, cout, y, x), needs cout % 8 == 0;  out_nchw (optional): fp32 (b, cout, y, x)
 * Products are formed as hi*hi + lo*hi + hi*lo with fp32 accumulation (about 1e-5 relative to an fp32 convolution).
 * Limits: k in {1, 3}, cout <= 128. */
int pdm_conv_tc_forward(int b, int y, int x, int cin, int cout, int ksize, const void *in_split,
                        const void *w_packed, const float *bias, int act, void *out_split, float *out_nchw,
                        void *stream);

/* pdm_neck_forward with the BEV map written in the split NHWC8 layout for the tensor-core convolutions
 * (spatial_split, may be NULL) and/or as fp32 (B,C,Y,X) (spatial_features, may be NULL); at least one. */
int pdm_neck_forward_split(int batch, int p, int c, const float *point_coords, const float *point_features,
                           const float *coef, const float *range_min, const float *voxel, const int *grid,
                           const int *dilation, int sh_degree, float sigma, float eps,
                           float *spatial_features, void *spatial_split, void *stream);

/* out (P, nout) = x (P, C) W^T + b, nout <= 16: nn.Linear for a handful of outputs per point (the neck's
 * spherical-harmonic coefficients, SPEC_PDM.md "coef").  w (nout, C) row-major as in nn.Linear.weight; b may be NULL. */
int pdm_linear_rows(int p, int c, int nout, const float *x, const float *w, const float *b, float *out, void *stream);

/* The per-point half of the hybrid head (SPEC_HEAD.md steps 3-6) in one kernel.  Replaces, at inference:
 * the FC stacks built by PointHeadTemplate.make_fc_layers (point_head_template.py:36-47) as used in
 * PointHeadBox.forward (point_head_box.py:85-86: cls_layers / box_layers), the class maximum of
 * generate_predicted_boxes (point_head_template.py:166-184) and PointResidualCoder.decode_torch
 * (box_coder_utils.py:189-222, use_mean_size=True), plus the hybrid head's feature fusion and score calibration.
 *   point_coords (P,4) [b,x,y,z]; point_features (P,c_point); bev_split: split NHWC8 (batch, c_bev, y, x);
 *   heatmap fp32 (batch, n_class, y, x), already sigmoid-ed; range_min_xy[2], voxel_xy[2] host arrays;
 *   w1t (c_point + c_bev, hidden_cls + hidden_box): first Linear of both stacks, BN folded, transposed and
 *   concatenated [cls | box]; b1 (hidden_cls + hidden_box); w2_cls (n_class, hidden_cls), b2_cls; w2_box (8, hidden_box),
 *   b2_box; mean_size (n_class, 3)
 *   -> scores (P, n_class) = sigmoid(cls) * sqrt(heatmap at the point's pillar); best_score (P), best_label (P) int32
 *      (first maximum); boxes (P,7) decoded with the best class; cls_raw (P, n_class) / box_raw (P,8) optional logits.
 * Limits: n_class <= 8, hidden widths multiples of 4 with sum <= 256, c_point % 4 == 0, c_bev % 8 == 0. */
int pdm_point_head_forward(int p, int batch, int c_point, int c_bev, int y, int x, int n_class, int hidden_cls,
                           int hidden_box, const float *range_min_xy, const float *voxel_xy,
                           const float *point_coords, const float *point_features, const void *bev_split,
                           const float *heatmap, const float *w1t, const float *b1, const float *w2_cls,
                           const float *b2_cls, const float *w2_box, const float *b2_box, const float *mean_size,
                           float *scores, float *boxes, float *best_score, int *best_label, float *cls_raw,
                           float *box_raw, void *stream);

/* ---- input staging (csrc/sample_points.cu) ----------------------------------------------------------- */

/* DataProcessor.sample_points (pcdet/datasets/processor/data_processor.py:182-212) for every frame of a batch, plus the
 * `points` part of DatasetTemplate.collate_batch (pcdet/datasets/dataset.py:237-244), on the device.
 *   points (total_points, c) fp32: the raw frames back to back, columns x, y, z, features...; counts (b) int32 DEVICE
 *   array with the rows of each frame (their sum must equal total_points)
 *   -> out (b * num_points, 1 + c) fp32 = [batch index, x, y, z, features...]: per frame exactly num_points rows chosen
 *      and shuffled as the reference does (all far points kept, near points subsampled without replacement; short frames
 *      padded with a random choice), the randomness defined by a counter-based hash of (seed, frame, index) -- see the
 *      file header; oracle/sample_points_oracle.py is the numpy restatement.  choice (optional, b * num_points int32):
 *      row of the source frame each output row was copied from (-1 for an empty frame).
 * No host synchronisation; safe to capture in a CUDA graph after one eager call. */
int pdm_sample_points(int b, int total_points, int c, int num_points, unsigned seed, const float *points,
                      const int *counts, float *out, int *choice, void *stream);

/* ---- stacked (ragged-batch) operator family (reference: pcdet/ops/pointnet2/pointnet2_stack) ----------------------
 * The 15 pybind entries of pointnet2_stack/src/pointnet2_api.cpp:12-31.  Stacked tensors
 * hold the frames of a batch back to back: xyz (N1+N2+.., 3), features (N1+N2+.., C) channel-last, and a DEVICE int32
 * array *_batch_cnt (batch) = [N1, N2, ..].  No entry reads a count on the host.  The plain
 * `farthest_point_sampling_wrapper` of that module (pointnet2_api.cpp:16) is pdm_farthest_point_sampling above. */

/* ball_query_wrapper_stack (pointnet2_api.cpp:13, ball_query.cpp:29-45, ball_query_gpu.cu:15-70).
 * new_xyz (M,3), xyz (N,3) -> idx (M,nsample): indices LOCAL to the centre's frame, first nsample hits (d2 < r*r) in
 * ascending order padded with the first; a centre without a hit gets idx[0] = -1 and the rest of its row untouched
 * (the caller zero-fills, pointnet2_utils.py:31). */
int pdm_stack_ball_query(int b, int m_total, int n_total, float radius, int nsample, const float *new_xyz,
                         const int *new_xyz_batch_cnt, const float *xyz, const int *xyz_batch_cnt, int *idx,
                         void *stream);

/* voxel_query_wrapper_stack (pointnet2_api.cpp:14, voxel_query.cpp:21-41, voxel_query_gpu.cu:10-88).
 * new_coords (M,4) int32 [batch, z, y, x]; point_indices (B, r1, r2, r3) int32 = row of the point held by a voxel or -1;
 * xyz (N,3) GLOBAL rows -> idx (M,nsample): the first nsample voxels of the (z,y,x)-ordered neighbourhood whose point
 * has d2 <= r*r, padded with the first; idx[0] = -1 when there is none. */
int pdm_stack_voxel_query(int m, int r1, int r2, int r3, int nsample, float radius, int z_range, int y_range,
                          int x_range, const float *new_xyz, const float *xyz, const int *new_coords,
                          const int *point_indices, int *idx, void *stream);

/* stack_farthest_point_sampling_wrapper (pointnet2_api.cpp:17, sampling.cpp:37-56, sampling_gpu.cu:263-348).
 * xyz (N,3), temp (N) pre-filled with 1e10, xyz_batch_cnt (batch), num_sampled_points (batch) ->
 * idx (sum of num_sampled_points): GLOBAL rows, frame after frame; ties as the reference's 1024-thread tournament. */
int pdm_stack_farthest_point_sampling(int n_total, int batch, const float *xyz, float *temp, const int *xyz_batch_cnt,
                                      int *idx, const int *num_sampled_points, void *stream);

/* group_points_wrapper_stack (pointnet2_api.cpp:19, group_points.cpp:50-67, group_points_gpu.cu:67-122).
 * features (N,C), idx (M,nsample) local indices -> out (M,C,nsample). */
int pdm_stack_group_points(int b, int m, int c, int nsample, const float *features, const int *features_batch_cnt,
                           const int *idx, const int *idx_batch_cnt, float *out, void *stream);

/* group_points_grad_wrapper_stack (pointnet2_api.cpp:20, group_points.cpp:29-48, group_points_gpu.cu:14-65).
 * grad_out (M,C,nsample) -> grad_features (N,C) (+=).  deterministic != 0: in-order sums instead of atomicAdd. */
int pdm_stack_group_points_grad(int b, int m, int c, int n, int nsample, const float *grad_out, const int *idx,
                                const int *idx_batch_cnt, const int *features_batch_cnt, float *grad_features,
                                int deterministic, void *stream);

/* three_nn_wrapper_stack (pointnet2_api.cpp:22, interpolate.cpp:32-60, interpolate_gpu.cu:17-96).
 * unknown (N,3), known (M,3) -> dist2 (N,3) squared distances, idx (N,3) GLOBAL rows of `known`. */
int pdm_stack_three_nn(int b, int n, int m, const float *unknown, const int *unknown_batch_cnt, const float *known,
                       const int *known_batch_cnt, float *dist2, int *idx, void *stream);

/* three_interpolate_wrapper_stack (pointnet2_api.cpp:23, interpolate.cpp:63-83, interpolate_gpu.cu:100-134).
 * features (M,C), idx (N,3), weight (N,3) -> out (N,C). */
int pdm_stack_three_interpolate(int n, int c, const float *features, const int *idx, const float *weight, float *out,
                                void *stream);

/* three_interpolate_grad_wrapper_stack (pointnet2_api.cpp:24, interpolate.cpp:86-106, interpolate_gpu.cu:137-194).
 * grad_out (N,C) -> grad_features (M,C) (+=). */
int pdm_stack_three_interpolate_grad(int n, int c, int m, const float *grad_out, const int *idx, const float *weight,
                                     float *grad_features, int deterministic, void *stream);

/* query_stacked_local_neighbor_idxs_wrapper_stack (pointnet2_api.cpp:26, vector_pool.cpp:27-52,
 * vector_pool_gpu.cu:98-180).  Per centre the first min(1000, nsample > 0 ? nsample : inf) support points of its frame
 * inside the ball (neighbor_type 1) or cube of half-size max_neighbour_distance, as GLOBAL rows appended to
 * stack_neighbor_idxs (capacity avg_length * M); start_len (M,2) = (start, count); *cumsum (device, caller-zeroed) +=
 * all counts.  Block order in the list is unspecified (atomics, as in the reference). */
int pdm_stack_query_local_neighbor_idxs(int b, int m, const float *support_xyz, const int *xyz_batch_cnt,
                                        const float *new_xyz, const int *new_xyz_batch_cnt, int *stack_neighbor_idxs,
                                        int *start_len, int *cumsum, int avg_length_of_neighbor_idxs,
                                        float max_neighbour_distance, int nsample, int neighbor_type, void *stream);

/* query_three_nn_by_stacked_local_idxs_wrapper_stack (pointnet2_api.cpp:27, vector_pool.cpp:55-78,
 * vector_pool_gpu.cu:19-95).  new_xyz_grid_centers (M,G,3) -> idxs (M,G,3) GLOBAL rows (-1: empty list; a missing 2nd /
 * 3rd neighbour repeats the 1st), dist2 (M,G,3). */
int pdm_stack_query_three_nn_by_local_idxs(int m, int num_total_grids, const float *support_xyz,
                                           const float *new_xyz_grid_centers, int *new_xyz_grid_idxs,
                                           float *new_xyz_grid_dist2, const int *stack_neighbor_idxs,
                                           const int *start_len, void *stream);

/* vector_pool_wrapper_stack (pointnet2_api.cpp:29, vector_pool.cpp:81-120, vector_pool_gpu.cu:183-373).
 * Outputs (caller zero-fills, pointnet2_utils.py:397-402): new_features (M, num_c_out) per-cell channel SUMS in
 * ascending point order, new_local_xyz (M, 3G), point_cnt_of_grid (M,G), grouped_idxs (num_max_sum_points,3) =
 * (support row, centre row, cell) in unspecified order.  The total number of list entries is left in *cum_sum_device
 * (a device int the caller reads back -- the reference's launcher does the cudaMemcpy itself, :360). */
int pdm_stack_vector_pool(int b, int n, int m, int num_c_in, int num_c_out, int num_total_grids,
                          const float *support_xyz, const int *xyz_batch_cnt, const float *support_features,
                          const float *new_xyz, const int *new_xyz_batch_cnt, float *new_features,
                          float *new_local_xyz, int *point_cnt_of_grid, int *grouped_idxs, int num_grid_x,
                          int num_grid_y, int num_grid_z, float max_neighbour_distance, int use_xyz,
                          int num_max_sum_points, int nsample, int neighbor_type, int pooling_type,
                          int *cum_sum_device, void *stream);

/* vector_pool_grad_wrapper_stack (pointnet2_api.cpp:30, vector_pool.cpp:123-147, vector_pool_gpu.cu:376-424).
 * grad_new_features (M, num_c_out) -> grad_support_features (N, num_c_in) (+=). */
int pdm_stack_vector_pool_grad(int m, int num_c_out, int n, int num_c_in, int num_total_grids, int num_entries,
                               const float *grad_new_features, const int *point_cnt_of_grid,
                               const int *grouped_idxs, float *grad_support_features, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PDM_OPS_H_ */
