"""Point backbones with pcdet's backbone_3d plugin API (detector3d_template.py:68-83:
`Backbone(model_cfg=..., input_channels=..., grid_size=..., voxel_size=..., point_cloud_range=...)`,
attribute `num_point_features`, `forward(batch_dict) -> batch_dict`).

* `PointNet2MSG` mirrors pcdet/models/backbones_3d/pointnet2_backbone.py:9-94 (SA chain + FP chain,
  same config keys SA_CONFIG.{NPOINTS,RADIUS,NSAMPLE,MLPS,USE_XYZ}, FP_MLPS, same attribute names
  -> same state_dict keys) on top of our SA/FP modules.
* `PDMSSDBackbone` is the single-stage (SSD-style) variant the PDM-SSD pipeline needs: the SA chain
  only, handing the LAST layer's centres and features to the neck ("use a
  PointNet2MSG-shaped SA chain without FP").

Unlike the reference (`pointnet2_backbone.py:72-76`: one `.sum()` per frame and a `min()==max()`
assert, i.e. batch_size+2 host syncs), equal per-frame point counts are derived from the tensor
shape: no device synchronisation in forward.
"""
import torch
import torch.nn as nn

from . import pointnet2_modules


# Device-side (asynchronous, no host sync) check that `points[:, 0]` really is 0..B-1 in equal blocks.  Off by default:
# it adds two small kernels per forward; tests and debugging switch it on.
CHECK_BATCH_INDEX = False


# PDMSSDBackbone: sample the next layer's centres on a side stream while the current layer runs its MLP (inference).
OVERLAP_SAMPLING = True


class AttrDict(dict):
    """Minimal EasyDict stand-in (pcdet/config.py uses easydict): attribute access + .get, nested."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = AttrDict(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    __setattr__ = dict.__setitem__


def _split_points(points, batch_size):
    """pcdet `points` (B*N, 1+3+C) -> batch_idx (B*N,), xyz (B,N,3), features (B,C,N) or None."""
    if points.shape[0] % batch_size != 0:
        raise RuntimeError("points rows (%d) are not a multiple of batch_size (%d): frames must carry "
                           "the same number of points (data_processor.sample_points)" % (points.shape[0], batch_size))
    batch_idx = points[:, 0]
    if CHECK_BATCH_INDEX:
        # the reference checks per-frame counts with one .sum() per frame (pointnet2_backbone.py:72-76: batch_size + 2
        # host syncs); here a device-side assertion: rows must be grouped by frame, `rows / batch_size` per frame
        want = torch.arange(batch_size, device=points.device, dtype=points.dtype).repeat_interleave(points.shape[0] // batch_size)
        torch._assert_async((batch_idx == want).all(), "points are not grouped by frame with equal counts")
    xyz = points[:, 1:4].contiguous().view(batch_size, -1, 3)
    feats = None
    if points.size(-1) > 4:
        feats = points[:, 4:].contiguous().view(batch_size, -1, points.size(-1) - 4).permute(0, 2, 1).contiguous()
    return batch_idx, xyz, feats


def _build_sa_chain(sa_cfg, channel_in):
    modules, skip = nn.ModuleList(), [channel_in]
    for k in range(len(sa_cfg.NPOINTS)):
        mlps = [[channel_in] + list(m) for m in sa_cfg.MLPS[k]]
        channel_out = sum(m[-1] for m in mlps)
        modules.append(pointnet2_modules.PointnetSAModuleMSG(
            npoint=sa_cfg.NPOINTS[k], radii=list(sa_cfg.RADIUS[k]), nsamples=list(sa_cfg.NSAMPLE[k]),
            mlps=mlps, use_xyz=sa_cfg.get('USE_XYZ', True)))
        skip.append(channel_out)
        channel_in = channel_out
    return modules, skip, channel_in


class PointNet2MSG(nn.Module):
    def __init__(self, model_cfg, input_channels, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.SA_modules, skip, channel_out = _build_sa_chain(model_cfg.SA_CONFIG, input_channels - 3)
        self.FP_modules = nn.ModuleList()
        fp = model_cfg.FP_MLPS
        for k in range(len(fp)):
            pre = fp[k + 1][-1] if k + 1 < len(fp) else channel_out
            self.FP_modules.append(pointnet2_modules.PointnetFPModule(mlp=[pre + skip[k]] + list(fp[k])))
        self.num_point_features = fp[0][-1]

    def forward(self, batch_dict):
        batch_idx, xyz, feats = _split_points(batch_dict['points'], batch_dict['batch_size'])
        l_xyz, l_feats = [xyz], [feats]
        for sa in self.SA_modules:
            nx, nf = sa(l_xyz[-1], l_feats[-1])
            l_xyz.append(nx)
            l_feats.append(nf)
        for i in range(-1, -(len(self.FP_modules) + 1), -1):
            l_feats[i - 1] = self.FP_modules[i](l_xyz[i - 1], l_xyz[i], l_feats[i - 1], l_feats[i])
        pf = l_feats[0].permute(0, 2, 1).contiguous()
        batch_dict['point_features'] = pf.view(-1, pf.shape[-1])
        batch_dict['point_coords'] = torch.cat((batch_idx[:, None].float(), l_xyz[0].view(-1, 3)), dim=1)
        return batch_dict


class PDMSSDBackbone(nn.Module):
    def __init__(self, model_cfg, input_channels, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.SA_modules, _, channel_out = _build_sa_chain(model_cfg.SA_CONFIG, input_channels - 3)
        self.num_point_features = channel_out

    def _sample(self, sa, xyz):
        """FPS + gather of one SA layer (pointnet2_modules.py:28-36), so that a layer can be handed its centres."""
        from . import pointnet2_utils
        idx = pointnet2_utils.farthest_point_sample(xyz, sa.npoint)
        return pointnet2_utils.gather_operation(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()

    def forward(self, batch_dict):
        B = batch_dict['batch_size']
        _, xyz, feats = _split_points(batch_dict['points'], B)
        if xyz.is_cuda and not self.training and OVERLAP_SAMPLING and all(sa.npoint is not None for sa in self.SA_modules):
            # Layer l+1's sampling needs only layer l's CENTRES, not its features: it runs on a side stream while layer l
            # does its ball query and shared MLP (a batch's sampling keeps 16 SMs busy, the MLP needs the other 132).
            # The modules take the centres through their `new_xyz` argument (pointnet2_modules.py:19-23).
            main = torch.cuda.current_stream(xyz.device)
            side = self.__dict__.setdefault('_side_stream', {}).get(xyz.device)
            if side is None:
                side = self.__dict__['_side_stream'][xyz.device] = torch.cuda.Stream(device=xyz.device)
            centres = self._sample(self.SA_modules[0], xyz)
            for li, sa in enumerate(self.SA_modules):
                nxt = None
                if li + 1 < len(self.SA_modules):
                    side.wait_stream(main)
                    with torch.cuda.stream(side):
                        nxt = self._sample(self.SA_modules[li + 1], centres)
                    nxt.record_stream(main)
                _, feats = sa(xyz, feats, new_xyz=centres)
                xyz = centres
                if nxt is not None:
                    main.wait_stream(side)
                    centres = nxt
        else:
            for sa in self.SA_modules:
                xyz, feats = sa(xyz, feats)
        M = xyz.shape[1]
        bcol = torch.arange(B, device=xyz.device, dtype=torch.float32).repeat_interleave(M)[:, None]
        batch_dict['point_coords'] = torch.cat((bcol, xyz.reshape(-1, 3)), dim=1)
        pm = getattr(feats, '_pdm_point_major', None)          # (B, M, C) written by the fused SA kernel
        if pm is not None and pm.shape == (B, M, feats.shape[1]):
            batch_dict['point_features'] = pm.view(B * M, -1)
        else:
            batch_dict['point_features'] = feats.permute(0, 2, 1).reshape(B * M, -1).contiguous()
        return batch_dict
