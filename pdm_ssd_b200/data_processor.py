"""Input staging on the GPU: `DataProcessor.sample_points` (pcdet/datasets/processor/data_processor.py:182-212) and the
`points` part of `DatasetTemplate.collate_batch` (pcdet/datasets/dataset.py:237-244) for a whole batch in one call.

The reference samples, pads and shuffles each frame with numpy on the host and collates on the host; at several thousand
frames per second that work and the pageable copy behind it bound the pipeline.  Here the raw,
ragged frames are uploaded as they are and `sample_points` produces the `(B * NUM_POINTS, 1 + C)` tensor the detector's
`batch_dict['points']` expects -- batch index in column 0, equal point counts per frame -- on the device, with the same
selection rules; the random choices come from a counter-based hash instead of numpy's generator (csrc/sample_points.cu).
"""
import torch

from . import _lib


def sample_points(points, counts, num_points, seed=0, return_choice=False):
    """points (sum(counts), C) CUDA fp32: raw frames back to back [x, y, z, features...]; counts (B,) CUDA int32.
    -> (B * num_points, 1 + C) CUDA fp32 [batch_idx, x, y, z, features...] [, choice (B, num_points) int32]."""
    if not (points.is_cuda and points.dtype == torch.float32 and points.is_contiguous() and points.dim() == 2):
        raise RuntimeError("points must be a contiguous CUDA float32 (rows, C) tensor")
    if not (counts.is_cuda and counts.dtype == torch.int32 and counts.is_contiguous() and counts.dim() == 1):
        raise RuntimeError("counts must be a contiguous CUDA int32 (B,) tensor")
    B, C = counts.shape[0], points.shape[1]
    out = torch.empty((B * num_points, 1 + C), dtype=torch.float32, device=points.device)
    choice = torch.empty((B, num_points), dtype=torch.int32, device=points.device) if return_choice else None
    with torch.cuda.device(points.device):
        rc = _lib.load().pdm_sample_points(B, points.shape[0], C, int(num_points), int(seed) & 0xffffffff, points.data_ptr(),
                                           counts.data_ptr(), out.data_ptr(), choice.data_ptr() if choice is not None else None,
                                           torch.cuda.current_stream(points.device).cuda_stream)
    _lib.check(rc, "pdm_sample_points")
    return (out, choice) if return_choice else out
