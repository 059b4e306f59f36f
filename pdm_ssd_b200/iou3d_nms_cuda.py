"""Drop-in for the reference extension module `iou3d_nms_cuda` (GPU entries used at inference).

Same names and positional signatures as pcdet/ops/iou3d_nms/src/iou3d_nms_api.cpp:12-17 for
`boxes_iou_bev_gpu`, `boxes_overlap_bev_gpu` and `nms_gpu`, on top of the C ABI in libpdmops.so
(current stream, RuntimeError instead of exit(-1)).  `nms_bev_batched` is the entry the reference
lacks: every frame of a batch in two launches, keep lists left on the device.
"""
import torch

from . import _lib
from .pointnet2_batch_cuda import _F32, _I32, _chk, _stream


def _boxes(t, name):
    p = _chk(t, name, _F32)
    if t.dim() != 2 or t.shape[1] != 7:
        raise RuntimeError("%s must have shape (N, 7), got %s" % (name, tuple(t.shape)))
    return p


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    lib = _lib.load()
    a, b = _boxes(boxes_a, "boxes_a"), _boxes(boxes_b, "boxes_b")
    o = _chk(ans_iou, "ans_iou", _F32)
    if ans_iou.numel() < boxes_a.shape[0] * boxes_b.shape[0]:
        raise RuntimeError("ans_iou is too small")
    with torch.cuda.device(boxes_a.device):
        _lib.check(lib.pdm_boxes_iou_bev(boxes_a.shape[0], a, boxes_b.shape[0], b, o, _stream(boxes_a)), "boxes_iou_bev")
    return 1


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    lib = _lib.load()
    a, b = _boxes(boxes_a, "boxes_a"), _boxes(boxes_b, "boxes_b")
    o = _chk(ans_overlap, "ans_overlap", _F32)
    if ans_overlap.numel() < boxes_a.shape[0] * boxes_b.shape[0]:
        raise RuntimeError("ans_overlap is too small")
    with torch.cuda.device(boxes_a.device):
        _lib.check(lib.pdm_boxes_overlap_bev(boxes_a.shape[0], a, boxes_b.shape[0], b, o, _stream(boxes_a)), "boxes_overlap_bev")
    return 1


def nms_bev_batched(boxes, counts, thresh, keep, num_keep):
    """boxes (F,K,7) sorted by descending score per frame, counts (F,) int32 or None ->
    keep (F,K) int32 (kept positions ascending, -1 padded), num_keep (F,) int32; no host sync."""
    lib = _lib.load()
    p = _chk(boxes, "boxes", _F32)
    if boxes.dim() != 3 or boxes.shape[2] != 7:
        raise RuntimeError("boxes must have shape (F, K, 7), got %s" % (tuple(boxes.shape),))
    F, K = boxes.shape[0], boxes.shape[1]
    c = _chk(counts, "counts", _I32) if counts is not None else None
    k = _chk(keep, "keep", _I32)
    n = _chk(num_keep, "num_keep", _I32)
    if keep.numel() < F * K or num_keep.numel() < F or (counts is not None and counts.numel() < F):
        raise RuntimeError("keep / num_keep / counts are too small for %d frames x %d boxes" % (F, K))
    with torch.cuda.device(boxes.device):
        _lib.check(lib.pdm_nms_bev_batched(F, K, p, c, float(thresh), k, n, _stream(boxes)), "nms_bev_batched")
    return 1


def nms_gpu(boxes, keep, nms_overlap_thresh):
    """Reference signature (iou3d_nms.cpp:137): boxes (N,7) CUDA sorted by score, keep (N,) CPU int64
    filled with the kept positions; returns their number.  (The copy to the host tensor `keep`
    synchronises, as the reference does; the batched entry above does not.)"""
    _boxes(boxes, "boxes")
    if keep.is_cuda or keep.dtype != torch.int64 or not keep.is_contiguous():
        raise RuntimeError("keep must be a contiguous CPU int64 tensor")
    n = boxes.shape[0]
    if n == 0:
        return 0
    dkeep = torch.empty((1, n), dtype=_I32, device=boxes.device)
    dnum = torch.empty((1,), dtype=_I32, device=boxes.device)
    nms_bev_batched(boxes.view(1, n, 7), None, nms_overlap_thresh, dkeep, dnum)
    num = int(dnum.item())
    keep[:num] = dkeep[0, :num].to(torch.int64).cpu()
    return num
