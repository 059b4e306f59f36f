"""Whole-batch, device-resident post-processing on top of the iou3d_nms extension (csrc/nms.cu).

The reference's per-frame helpers (pcdet/ops/iou3d_nms/iou3d_nms_utils.py, pcdet/models/model_utils/
model_nms_utils.py) are NOT mirrored here: they run unchanged on our extension module
(`pdm_ssd_b200.iou3d_nms_cuda` fills the `iou3d_nms_cuda` slot; tests/test_dropin_gpu.py).
What this module adds is what the reference lacks: every frame (and class) of a batch in one pass, fixed shapes,
no host synchronisation -- CUDA-graph friendly.  Semantics per frame are those of `class_agnostic_nms`
(model_nms_utils.py:6-25) / `multi_classes_nms` (model_nms_utils.py:28-66) as called from
Detector3DTemplate.post_processing (detector3d_template.py:199-254).
"""
import torch

from . import iou3d_nms_cuda


def batched_nms_gpu(boxes, scores, thresh, pre_maxsize, post_maxsize, score_thresh=None, nms_type="nms_gpu"):
    """boxes (F,M,7+), scores (F,M) -> selected (F,post_maxsize) int64 indices into M in descending score
    order, -1 padded, and num (F,) int32.  Per frame: keep scores >= score_thresh, the `pre_maxsize` best,
    greedy rotated (nms_gpu) or axis-aligned (nms_normal_gpu) suppression at `thresh`, first `post_maxsize`."""
    F, M = scores.shape
    k = min(int(pre_maxsize), M)
    top, order = scores.topk(k, dim=1)                       # descending
    cand = torch.gather(boxes[..., :7], 1, order[..., None].expand(F, k, 7)).contiguous()
    counts = (top >= score_thresh).sum(dim=1).to(torch.int32) if score_thresh is not None else None
    keep = torch.empty((F, k), dtype=torch.int32, device=boxes.device)
    num = torch.empty((F,), dtype=torch.int32, device=boxes.device)
    iou3d_nms_cuda.nms_bev_batched(cand, counts, thresh, keep, num, normal=(nms_type == "nms_normal_gpu"))
    p = min(int(post_maxsize), k)
    kept = keep[:, :p].to(torch.int64)
    selected = torch.where(kept >= 0, torch.gather(order, 1, kept.clamp(min=0)), kept)
    if p < post_maxsize:
        selected = torch.nn.functional.pad(selected, (0, post_maxsize - p), value=-1)
    return selected, num.clamp(max=post_maxsize)


def batched_multi_classes_nms_gpu(boxes, cls_scores, thresh, pre_maxsize, post_maxsize, score_thresh=None,
                                  nms_type="nms_gpu"):
    """Per-class suppression for a whole batch: boxes (F,M,7+), cls_scores (F,M,C) ->
    selected (F,C,post_maxsize) int64 (-1 padded) and num (F,C) int32; class k of frame f is suppressed on its own
    score column exactly as one iteration of the reference's class loop (model_nms_utils.py:38-56).  The (frame,
    class) pairs go through the batched kernels as F*C independent lists."""
    F, M, C = cls_scores.shape
    sc = cls_scores.permute(0, 2, 1).reshape(F * C, M)
    bx = boxes[..., :7].unsqueeze(1).expand(F, C, M, 7).reshape(F * C, M, 7)
    sel, num = batched_nms_gpu(bx, sc, thresh, pre_maxsize, post_maxsize, score_thresh, nms_type)
    return sel.view(F, C, -1), num.view(F, C)
