"""Set-abstraction / feature-propagation modules with the reference's plugin API
(pcdet/ops/pointnet2/pointnet2_batch/pointnet2_modules.py): same class names, keyword
arguments, attribute names (`groupers`, `mlps`, `mlp`, `npoint`, `pool_method`) and therefore
the same state_dict keys (`mlps.{i}.{3k}.weight` conv, `mlps.{i}.{3k+1}.*` batch-norm), so
reference checkpoints and configs load as they are.
"""
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointnet2_utils


def _shared_mlp(widths: List[int]) -> nn.Sequential:
    """1x1 Conv2d(bias=False) + BatchNorm2d + ReLU per hop (pointnet2_modules.py:90-97,132-139)."""
    layers = []
    for c_in, c_out in zip(widths[:-1], widths[1:]):
        layers += [nn.Conv2d(c_in, c_out, kernel_size=1, bias=False), nn.BatchNorm2d(c_out), nn.ReLU()]
    return nn.Sequential(*layers)


class _PointnetSAModuleBase(nn.Module):
    """pointnet2_modules.py:10-55."""

    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None
        self.pool_method = 'max_pool'

    def _pool(self, x: torch.Tensor) -> torch.Tensor:
        window = [1, x.size(3)]
        if self.pool_method == 'max_pool':
            return F.max_pool2d(x, kernel_size=window)
        if self.pool_method == 'avg_pool':
            return F.avg_pool2d(x, kernel_size=window)
        raise NotImplementedError

    def forward(self, xyz: torch.Tensor, features: Optional[torch.Tensor] = None,
                new_xyz: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """xyz (B,N,3), features (B,C,N) -> new_xyz (B,npoint,3), new_features (B,sum C_out,npoint)."""
        if new_xyz is None and self.npoint is not None:
            sample_idx = pointnet2_utils.farthest_point_sample(xyz, self.npoint)
            channels_first = xyz.transpose(1, 2).contiguous()
            new_xyz = pointnet2_utils.gather_operation(channels_first, sample_idx).transpose(1, 2).contiguous()
        pooled = []
        for grouper, mlp in zip(self.groupers, self.mlps):
            grouped = grouper(xyz, new_xyz, features)          # (B, C_in, npoint, nsample)
            pooled.append(self._pool(mlp(grouped)).squeeze(-1))  # (B, C_out, npoint)
        return new_xyz, torch.cat(pooled, dim=1)


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    """Multi-scale-grouping set abstraction (pointnet2_modules.py:58-99)."""

    def __init__(self, *, npoint: int, radii: List[float], nsamples: List[int], mlps: List[List[int]],
                 bn: bool = True, use_xyz: bool = True, pool_method='max_pool'):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint = npoint
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        for radius, nsample, widths in zip(radii, nsamples, mlps):
            if npoint is not None:
                self.groupers.append(pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz))
            else:
                self.groupers.append(pointnet2_utils.GroupAll(use_xyz))
            if use_xyz:
                widths[0] += 3  # in place, like the reference (:87-88): the caller's list sees it too
            self.mlps.append(_shared_mlp(widths))
        self.pool_method = pool_method


class PointnetSAModule(PointnetSAModuleMSG):
    """Single-scale set abstraction (pointnet2_modules.py:102-119)."""

    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None, nsample: int = None,
                 bn: bool = True, use_xyz: bool = True, pool_method='max_pool'):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], bn=bn,
                         use_xyz=use_xyz, pool_method=pool_method)


class PointnetFPModule(nn.Module):
    """Feature propagation: inverse-distance 3-NN interpolation + shared MLP
    (pointnet2_modules.py:122-170)."""

    def __init__(self, *, mlp: List[int], bn: bool = True):
        super().__init__()
        self.mlp = _shared_mlp(mlp)

    def forward(self, unknown: torch.Tensor, known: torch.Tensor, unknow_feats: torch.Tensor,
                known_feats: torch.Tensor) -> torch.Tensor:
        if known is None:
            carried = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))
        else:
            dist, idx = pointnet2_utils.three_nn(unknown, known)
            inv = 1.0 / (dist + 1e-8)
            weight = inv / torch.sum(inv, dim=2, keepdim=True)
            carried = pointnet2_utils.three_interpolate(known_feats, idx, weight)
        stacked = carried if unknow_feats is None else torch.cat([carried, unknow_feats], dim=1)
        return self.mlp(stacked.unsqueeze(-1)).squeeze(-1)
