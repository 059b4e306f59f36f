"""Drop-in for the reference extension module `iou3d_nms_cuda` (GPU entries used at inference).

Same names and positional signatures as the six GPU entries of pcdet/ops/iou3d_nms/src/iou3d_nms_api.cpp:12-17
(`boxes_aligned_overlap_bev_gpu`, `boxes_overlap_bev_gpu`, `paired_boxes_overlap_bev_gpu`, `boxes_iou_bev_gpu`,
`nms_gpu`, `nms_normal_gpu`), on top of the C ABI in libpdmops.so (current stream, RuntimeError instead of
exit(-1)), so the reference's own iou3d_nms_utils.py / model_nms_utils.py run on it unchanged
(tests/test_dropin_gpu.py).  The two CPU entries (`boxes_iou_bev_cpu`, `boxes_aligned_iou_bev_cpu`) are host code
outside the GPU path and are not provided.  `nms_bev_batched` / `nms_normal_batched` are the entries the reference
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


def _paired(boxes_a, boxes_b, ans_overlap, what):
    lib = _lib.load()
    a, b = _boxes(boxes_a, "boxes_a"), _boxes(boxes_b, "boxes_b")
    o = _chk(ans_overlap, "ans_overlap", _F32)
    if boxes_a.shape[0] != boxes_b.shape[0] or ans_overlap.numel() < boxes_a.shape[0]:
        raise RuntimeError("%s: boxes_a and boxes_b must pair up and ans_overlap must hold one value per pair" % what)
    with torch.cuda.device(boxes_a.device):
        _lib.check(lib.pdm_boxes_overlap_bev_paired(boxes_a.shape[0], a, b, o, _stream(boxes_a)), what)
    return 1


def paired_boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    """iou3d_nms.cpp:92-111: overlap of box i of boxes_a with box i of boxes_b -> ans_overlap (N,1)."""
    return _paired(boxes_a, boxes_b, ans_overlap, "paired_boxes_overlap_bev")


def boxes_aligned_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    """iou3d_nms.cpp:42-66: same pairing (the reference has two identical kernels)."""
    return _paired(boxes_a, boxes_b, ans_overlap, "boxes_aligned_overlap_bev")


def nms_bev_batched(boxes, counts, thresh, keep, num_keep, normal=False):
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
    fn = lib.pdm_nms_normal_batched if normal else lib.pdm_nms_bev_batched
    with torch.cuda.device(boxes.device):
        _lib.check(fn(F, K, p, c, float(thresh), k, n, _stream(boxes)), "nms_bev_batched")
    return 1


def nms_normal_batched(boxes, counts, thresh, keep, num_keep):
    """Axis-aligned variant of nms_bev_batched (heading ignored, iou3d_nms_kernel.cu:341-398)."""
    return nms_bev_batched(boxes, counts, thresh, keep, num_keep, normal=True)


def _single(boxes, keep, thresh, normal):
    _boxes(boxes, "boxes")
    if keep.is_cuda or keep.dtype != torch.int64 or not keep.is_contiguous():
        raise RuntimeError("keep must be a contiguous CPU int64 tensor")
    n = boxes.shape[0]
    if n == 0:
        return 0
    dkeep = torch.empty((1, n), dtype=_I32, device=boxes.device)
    dnum = torch.empty((1,), dtype=_I32, device=boxes.device)
    nms_bev_batched(boxes.view(1, n, 7), None, thresh, dkeep, dnum, normal=normal)
    num = int(dnum.item())
    keep[:num] = dkeep[0, :num].to(torch.int64).cpu()
    return num


def nms_gpu(boxes, keep, nms_overlap_thresh):
    """Reference signature (iou3d_nms.cpp:137): boxes (N,7) CUDA sorted by score, keep (N,) CPU int64
    filled with the kept positions; returns their number.  (The copy to the host tensor `keep`
    synchronises, as the reference does; the batched entry above does not.)"""
    return _single(boxes, keep, nms_overlap_thresh, False)


def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    """Reference signature (iou3d_nms.cpp:186): as nms_gpu on the axis-aligned IoU."""
    return _single(boxes, keep, nms_overlap_thresh, True)
