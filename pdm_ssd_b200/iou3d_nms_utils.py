"""Mirror of pcdet/ops/iou3d_nms/iou3d_nms_utils.py (GPU functions on the inference path), same
names, arguments and return values, plus `batched_nms_gpu` for whole batches without host syncs."""
import torch

from . import iou3d_nms_cuda


def boxes_iou_bev(boxes_a, boxes_b):
    """(N,7), (M,7) [x, y, z, dx, dy, dz, heading] -> (N, M) rotated BEV IoU (iou3d_nms_utils.py:31-44)."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    ans_iou = torch.zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    iou3d_nms_cuda.boxes_iou_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), ans_iou)
    return ans_iou


def boxes_iou3d_gpu(boxes_a, boxes_b):
    """(N,7), (M,7) -> (N, M) 3-D IoU = BEV overlap x height overlap (iou3d_nms_utils.py:47-82)."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    boxes_a_height_max = (boxes_a[:, 2] + boxes_a[:, 5] / 2).view(-1, 1)
    boxes_a_height_min = (boxes_a[:, 2] - boxes_a[:, 5] / 2).view(-1, 1)
    boxes_b_height_max = (boxes_b[:, 2] + boxes_b[:, 5] / 2).view(1, -1)
    boxes_b_height_min = (boxes_b[:, 2] - boxes_b[:, 5] / 2).view(1, -1)
    overlaps_bev = torch.zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    iou3d_nms_cuda.boxes_overlap_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), overlaps_bev)
    max_of_min = torch.max(boxes_a_height_min, boxes_b_height_min)
    min_of_max = torch.min(boxes_a_height_max, boxes_b_height_max)
    overlaps_h = torch.clamp(min_of_max - max_of_min, min=0)
    overlaps_3d = overlaps_bev * overlaps_h
    vol_a = (boxes_a[:, 3] * boxes_a[:, 4] * boxes_a[:, 5]).view(-1, 1)
    vol_b = (boxes_b[:, 3] * boxes_b[:, 4] * boxes_b[:, 5]).view(1, -1)
    return overlaps_3d / torch.clamp(vol_a + vol_b - overlaps_3d, min=1e-6)


def nms_gpu(boxes, scores, thresh, pre_maxsize=None, **kwargs):
    """(N,7), (N,) -> (kept indices into `boxes` in descending-score order, None)
    (iou3d_nms_utils.py:120-135)."""
    assert boxes.shape[1] == 7
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    boxes = boxes[order].contiguous()
    keep = torch.zeros(boxes.size(0), dtype=torch.int64)
    num_out = iou3d_nms_cuda.nms_gpu(boxes, keep, thresh)
    return order[keep[:num_out].to(boxes.device)].contiguous(), None


def batched_nms_gpu(boxes, scores, thresh, pre_maxsize, post_maxsize, score_thresh=None):
    """All frames at once, device-resident, fixed shapes (CUDA-graph friendly).

    boxes (F,M,7), scores (F,M) -> selected (F,post_maxsize) int64 indices into M (descending score,
    -1 padded) and num (F,) int32.  Per frame this is `class_agnostic_nms`
    (model_nms_utils.py:6-25): score threshold, top-`pre_maxsize`, rotated NMS, first `post_maxsize`.
    """
    F, M = scores.shape
    k = min(pre_maxsize, M)
    top, order = scores.topk(k, dim=1)                       # sorted descending (torch.topk default)
    sorted_boxes = torch.gather(boxes[..., :7], 1, order[..., None].expand(F, k, 7)).contiguous()
    counts = None
    if score_thresh is not None:
        counts = (top >= score_thresh).sum(dim=1).to(torch.int32)
    keep = torch.empty((F, k), dtype=torch.int32, device=boxes.device)
    num = torch.empty((F,), dtype=torch.int32, device=boxes.device)
    iou3d_nms_cuda.nms_bev_batched(sorted_boxes, counts, thresh, keep, num)
    p = min(post_maxsize, k)
    kept = keep[:, :p].to(torch.int64)
    selected = torch.where(kept >= 0, torch.gather(order, 1, kept.clamp(min=0)), kept)
    if p < post_maxsize:
        selected = torch.nn.functional.pad(selected, (0, post_maxsize - p), value=-1)
    return selected, num.clamp(max=post_maxsize)
