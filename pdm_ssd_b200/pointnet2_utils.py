"""Host-side operator layer: the autograd Functions and grouping modules of
pcdet/ops/pointnet2/pointnet2_batch/pointnet2_utils.py, re-implemented on top of the
B200 kernels in libpdmops.so.

Public names, argument order, tensor layouts, dtypes and the "caller allocates" convention
are the reference's (file:line cited per class), so code written against
`pcdet.ops.pointnet2.pointnet2_batch.pointnet2_utils` runs unchanged.  The native backend
is a module exposing the nine `*_wrapper` functions; by default that is our
`pointnet2_batch_cuda` shim.  Tests and bench.py can swap in the reference's own compiled
extension with `use_backend(...)` to run the identical host code over the reference kernels
(that is the "reference CUDA-op pipeline" timing arm and the bit-exact GPU oracle).
"""
import contextlib
from typing import Tuple

import torch
import torch.nn as nn
from torch.autograd import Function

from . import pointnet2_batch_cuda as _native

_backend = _native


def get_backend():
    return _backend


@contextlib.contextmanager
def use_backend(module):
    """Temporarily route the nine native calls to `module` (e.g. the reference .so)."""
    global _backend
    prev, _backend = _backend, module
    try:
        yield module
    finally:
        _backend = prev


def _require_contiguous(**tensors):
    for name, t in tensors.items():
        if not t.is_contiguous():
            raise AssertionError("%s must be contiguous" % name)


class FarthestPointSampling(Function):
    """pointnet2_utils.py:10-36.  xyz (B,N,3) f32 -> (B,npoint) i32 sample indices."""

    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        _require_contiguous(xyz=xyz)
        B, N, _ = xyz.shape
        idx = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
        running_min = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
        _backend.farthest_point_sampling_wrapper(B, N, npoint, xyz, running_min, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, grad=None):
        return None, None


farthest_point_sample = furthest_point_sample = FarthestPointSampling.apply


class GatherOperation(Function):
    """pointnet2_utils.py:39-73.  features (B,C,N), idx (B,npoint) -> (B,C,npoint)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        _require_contiguous(features=features, idx=idx)
        B, npoint = idx.shape
        _, C, N = features.shape
        out = torch.empty((B, C, npoint), dtype=torch.float32, device=features.device)
        _backend.gather_points_wrapper(B, C, N, npoint, features, idx, out)
        ctx.save_for_backward(idx)
        ctx.shape_cn = (C, N)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        C, N = ctx.shape_cn
        B, npoint = idx.shape
        grad_features = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
        _backend.gather_points_grad_wrapper(B, C, N, npoint, grad_out.detach().contiguous(), idx, grad_features)
        return grad_features, None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    """pointnet2_utils.py:76-105.  unknown (B,n,3), known (B,m,3) -> (dist (B,n,3) L2, idx (B,n,3))."""

    @staticmethod
    def forward(ctx, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        _require_contiguous(unknown=unknown, known=known)
        B, n, _ = unknown.shape
        m = known.shape[1]
        dist2 = torch.empty((B, n, 3), dtype=torch.float32, device=unknown.device)
        idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknown.device)
        _backend.three_nn_wrapper(B, n, m, unknown, known, dist2, idx)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    """pointnet2_utils.py:108-153.  features (B,C,m), idx (B,n,3), weight (B,n,3) -> (B,C,n)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        _require_contiguous(features=features, idx=idx, weight=weight)
        B, C, m = features.shape
        n = idx.shape[1]
        out = torch.empty((B, C, n), dtype=torch.float32, device=features.device)
        _backend.three_interpolate_wrapper(B, C, m, n, features, idx, weight, out)
        ctx.save_for_backward(idx, weight)
        ctx.m = m
        return out

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, weight = ctx.saved_tensors
        B, C, n = grad_out.shape
        grad_features = torch.zeros((B, C, ctx.m), dtype=torch.float32, device=grad_out.device)
        _backend.three_interpolate_grad_wrapper(B, C, n, ctx.m, grad_out.detach().contiguous(), idx, weight,
                                                grad_features)
        return grad_features, None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    """pointnet2_utils.py:156-197.  features (B,C,N), idx (B,npoint,nsample) -> (B,C,npoint,nsample)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        _require_contiguous(features=features, idx=idx)
        B, npoint, nsample = idx.shape
        _, C, N = features.shape
        out = torch.empty((B, C, npoint, nsample), dtype=torch.float32, device=features.device)
        _backend.group_points_wrapper(B, C, N, npoint, nsample, features, idx, out)
        ctx.save_for_backward(idx)
        ctx.n = N
        return out

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        (idx,) = ctx.saved_tensors
        B, C, npoint, nsample = grad_out.shape
        grad_features = torch.zeros((B, C, ctx.n), dtype=torch.float32, device=grad_out.device)
        _backend.group_points_grad_wrapper(B, C, ctx.n, npoint, nsample, grad_out.detach().contiguous(), idx,
                                           grad_features)
        return grad_features, None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    """pointnet2_utils.py:200-228.  (radius, nsample, xyz (B,N,3), new_xyz (B,npoint,3)) ->
    idx (B,npoint,nsample) i32, zero-initialised so that centres without any hit read 0."""

    @staticmethod
    def forward(ctx, radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
        _require_contiguous(new_xyz=new_xyz, xyz=xyz)
        B, N, _ = xyz.shape
        npoint = new_xyz.shape[1]
        idx = torch.zeros((B, npoint, nsample), dtype=torch.int32, device=xyz.device)
        _backend.ball_query_wrapper(B, N, npoint, radius, nsample, new_xyz, xyz, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class QueryAndGroup(nn.Module):
    """pointnet2_utils.py:231-264: ball query, then grouped (xyz - centre) and grouped features,
    concatenated as [xyz(3), features(C)] -> (B, 3+C, npoint, nsample)."""

    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None) -> torch.Tensor:
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        if (_backend is _native and not torch.is_grad_enabled() and xyz.is_cuda
                and (features is not None or self.use_xyz)):
            return _query_and_group_fused(xyz, new_xyz, features, idx, self.use_xyz)
        channels_first_xyz = xyz.transpose(1, 2).contiguous()
        local_xyz = grouping_operation(channels_first_xyz, idx)
        local_xyz -= new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return local_xyz
        grouped = grouping_operation(features, idx)
        return torch.cat([local_xyz, grouped], dim=1) if self.use_xyz else grouped


def _query_and_group_fused(xyz, new_xyz, features, idx, use_xyz):
    """One kernel for group-xyz / subtract-centre / group-features / cat (inference, native backend).
    Bit-identical to the unfused sequence: the same gathers and the same fp32 subtraction."""
    from . import _lib
    B, N, _ = xyz.shape
    _, M, S = idx.shape
    C = 0 if features is None else features.shape[1]
    feats = features.contiguous() if features is not None else None
    out = torch.empty((B, (3 if use_xyz else 0) + C, M, S), dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        rc = _lib.load().pdm_query_and_group(
            B, C, N, M, S, 1 if use_xyz else 0, xyz.contiguous().data_ptr(), new_xyz.contiguous().data_ptr(),
            feats.data_ptr() if feats is not None else None, idx.data_ptr(), out.data_ptr(),
            torch.cuda.current_stream(xyz.device).cuda_stream)
    _lib.check(rc, "pdm_query_and_group")
    return out


class GroupAll(nn.Module):
    """pointnet2_utils.py:267-290: a single group holding every point -> (B, 3+C, 1, N)."""

    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        all_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return all_xyz
        all_feat = features.unsqueeze(2)
        return torch.cat([all_xyz, all_feat], dim=1) if self.use_xyz else all_feat
