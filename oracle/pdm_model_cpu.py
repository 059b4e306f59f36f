"""The full PDM-SSD detector on the HOST CPU: bench.py's `--impl reference` arm and `cpu_baseline` leg.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (never imported by the product package).  BASELINE.json north_star
prescribes this arm: "the reference's PyTorch CPU path on the box's own host cores ... with the CUDA-only pointnet2
ops replaced by a torch-CPU reference of identical semantics".  Concretely:
  * the detector's own torch modules (reference-shaped path: group / conv / bn / relu / max-pool as separate torch
    ops, pointnet2_modules.py:19-55 order) run on CPU tensors;
  * the nine `pointnet2_batch_cuda` entry points (pointnet2_api.cpp:10-24), which exist only as CUDA kernels in the
    reference, are served by the C oracle (oracle/pdm_oracle.c, multi-threaded) through `CpuPointnet2Backend`;
  * the PDM neck is the pure-torch oracle (oracle/pdm_neck_oracle.py = BASELINE configs[0]);
  * rotated NMS is the C oracle's (post-processing of SPEC_HEAD.md step 7).
"""
import numpy as np
import torch

import oracle
import pdm_head_oracle as ho
import pdm_neck_oracle as no


def _np(t):
    return t.detach().contiguous().numpy()


class CpuPointnet2Backend:
    """The nine wrapper names and positional signatures of pointnet2_api.cpp:10-24 on CPU tensors (caller allocates)."""

    @staticmethod
    def farthest_point_sampling_wrapper(b, n, m, points, temp, idx):
        i, t = oracle.fps(_np(points).reshape(b, n, 3), m, return_temp=True)
        idx.copy_(torch.from_numpy(i))
        temp.copy_(torch.from_numpy(t))
        return 1

    @staticmethod
    def gather_points_wrapper(b, c, n, npoints, points, idx, out):
        out.copy_(torch.from_numpy(oracle.gather_points(_np(points), _np(idx))))
        return 1

    @staticmethod
    def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
        grad_points.add_(torch.from_numpy(oracle.gather_points_grad(_np(grad_out), _np(idx), n)))
        return 1

    @staticmethod
    def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
        idx.copy_(torch.from_numpy(oracle.ball_query(radius, nsample, _np(xyz), _np(new_xyz))))
        return 1

    @staticmethod
    def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
        out.copy_(torch.from_numpy(oracle.group_points(_np(points), _np(idx))))
        return 1

    @staticmethod
    def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
        grad_points.add_(torch.from_numpy(oracle.group_points_grad(_np(grad_out), _np(idx), n)))
        return 1

    @staticmethod
    def three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
        d, i = oracle.three_nn(_np(unknown), _np(known))
        dist2.copy_(torch.from_numpy(d))
        idx.copy_(torch.from_numpy(i))

    @staticmethod
    def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
        out.copy_(torch.from_numpy(oracle.three_interpolate(_np(points), _np(idx), _np(weight))))

    @staticmethod
    def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
        grad_points.add_(torch.from_numpy(oracle.three_interpolate_grad(_np(grad_out), _np(idx), _np(weight), m)))


@torch.no_grad()
def cpu_forward(model, points, batch_size, threads=None):
    """model: a `pdm_ssd_b200.detector.PDMSSD` on the CPU (eval mode); points (B*N, 5) CPU tensor.
    -> detections (B, K, 9), num (B,)."""
    from pdm_ssd_b200 import pointnet2_utils as pu
    if threads:
        oracle.set_threads(threads)
        torch.set_num_threads(threads)
    cfg = model.cfg
    bd = {"batch_size": batch_size, "points": points}
    with pu.use_backend(CpuPointnet2Backend):
        bd = model.backbone_3d(bd)
    neck = model.map_to_bev_module
    feats, coords = bd["point_features"].contiguous(), bd["point_coords"].contiguous()
    bd["spatial_features"] = no.neck_forward(coords, feats, neck.coef(feats), batch_size, neck.point_cloud_range, neck.voxel_size,
                                             neck.dilation, neck.sh_degree, neck.sigma, neck.eps)
    bd = model.backbone_2d(bd)
    head = model.dense_head
    saved, head.post_cfg = head.post_cfg, None          # device-side NMS is GPU-only: post-process with the oracle below
    try:
        bd = head(bd)
    finally:
        head.post_cfg = saved
    best, label = bd["batch_cls_preds"].max(dim=1)
    post = cfg.POST_PROCESSING
    return ho.post_process(bd["batch_box_preds"], best, label, batch_size, post.SCORE_THRESH, post.NMS_CONFIG.NMS_THRESH,
                           post.NMS_CONFIG.NMS_PRE_MAXSIZE, post.NMS_CONFIG.NMS_POST_MAXSIZE)
