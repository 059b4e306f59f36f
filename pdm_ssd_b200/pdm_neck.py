"""PDM neck (SPEC_PDM.md) as a pcdet `map_to_bev` plugin.

Constructor and forward follow the slot's contract (detector3d_template.py:85-95): built as
`PDMNeck(model_cfg=..., grid_size=...)`, exposes `num_bev_features`, `forward(batch_dict)` reads
`point_coords` / `point_features` / `batch_size` and writes `spatial_features (B,C,Y,X)`.
The compute is one C-ABI call (`pdm_neck_forward[_split]`, csrc/pdm_neck.cu); the per-centre SH
coefficients come from `pdm_linear_rows` (csrc/point_head.cu) at inference, from `nn.Linear` otherwise.
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from . import _lib


def _get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def neck_forward(point_coords, point_features, coef, batch_size, point_cloud_range, voxel_size, grid,
                 dilation=(1, 1, 1), sh_degree=2, sigma=0.8, eps=1e-6, return_debug=False, output="nchw"):
    """Functional form.  All tensors CUDA fp32 contiguous.
    output="nchw": returns spatial_features (B,C,Y,X) [, keys (P,K) int32, weights (P,K) fp32];
    output="split": returns the map as a `conv_tc.SplitAct` (what the tensor-core convolutions read);
    output="both": (spatial_features, SplitAct)."""
    lib = _lib.load()
    for name, t in (("point_coords", point_coords), ("point_features", point_features), ("coef", coef)):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError("%s must be a contiguous CUDA float32 tensor" % name)
    P, C = point_features.shape
    nsh = (sh_degree + 1) ** 2
    if point_coords.shape != (P, 4) or coef.shape != (P, nsh):
        raise RuntimeError("shape mismatch: point_coords %s, coef %s" % (tuple(point_coords.shape), tuple(coef.shape)))
    X, Y, Z = (int(g) for g in grid)
    K = int(np.prod([2 * int(k) + 1 for k in dilation]))
    dev = point_features.device
    f3, i3 = ctypes.c_float * 3, ctypes.c_int * 3
    cfg = (f3(*[float(v) for v in point_cloud_range[:3]]), f3(*[float(v) for v in voxel_size]),
           i3(X, Y, Z), i3(*[int(k) for k in dilation]), int(sh_degree), float(sigma), float(eps))
    if output != "nchw":
        if return_debug:
            raise RuntimeError("return_debug needs output='nchw'")
        from .conv_tc import SplitAct
        split = SplitAct.empty(batch_size, C, Y, X, dev)
        out = torch.empty((batch_size, C, Y, X), dtype=torch.float32, device=dev) if output == "both" else None
        with torch.cuda.device(dev):
            rc = lib.pdm_neck_forward_split(
                int(batch_size), P, C, point_coords.data_ptr(), point_features.data_ptr(), coef.data_ptr(), *cfg,
                out.data_ptr() if out is not None else None, split.data.data_ptr(),
                torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "pdm_neck_forward_split")
        return (out, split) if output == "both" else split
    out = torch.empty((batch_size, C, Y, X), dtype=torch.float32, device=dev)
    keys = wts = None
    if return_debug:
        keys = torch.empty((P, K), dtype=torch.int32, device=out.device)
        wts = torch.empty((P, K), dtype=torch.float32, device=out.device)
    with torch.cuda.device(out.device):
        rc = lib.pdm_neck_forward(
            int(batch_size), P, C, point_coords.data_ptr(), point_features.data_ptr(), coef.data_ptr(), *cfg,
            out.data_ptr(), keys.data_ptr() if keys is not None else None,
            wts.data_ptr() if wts is not None else None,
            torch.cuda.current_stream(out.device).cuda_stream)
    _lib.check(rc, "pdm_neck_forward")
    return (out, keys, wts) if return_debug else out


def linear_rows(x, weight, bias=None):
    """nn.Linear for a handful of outputs per row on our kernel (csrc/point_head.cu): x (P,C) -> (P,nout)."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
        raise RuntimeError("linear_rows needs a contiguous CUDA float32 input")
    P, C = x.shape
    w = weight.detach().contiguous()
    b = bias.detach().contiguous() if bias is not None else None
    out = torch.empty((P, w.shape[0]), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().pdm_linear_rows(P, C, w.shape[0], x.data_ptr(), w.data_ptr(), b.data_ptr() if b is not None else None,
                                         out.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "pdm_linear_rows")
    return out


class PDMNeck(nn.Module):
    def __init__(self, model_cfg, grid_size=None, voxel_size=None, point_cloud_range=None, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_bev_features = int(_get(model_cfg, "NUM_BEV_FEATURES"))
        self.voxel_size = [float(v) for v in (voxel_size if voxel_size is not None else _get(model_cfg, "VOXEL_SIZE"))]
        self.point_cloud_range = [float(v) for v in (point_cloud_range if point_cloud_range is not None
                                                     else _get(model_cfg, "POINT_CLOUD_RANGE"))]
        if grid_size is None:
            r = np.asarray(self.point_cloud_range, dtype=np.float64)
            grid_size = np.round((r[3:] - r[:3]) / np.asarray(self.voxel_size, dtype=np.float64))
        self.grid_size = [int(g) for g in grid_size]
        self.dilation = tuple(int(k) for k in _get(model_cfg, "DILATION", (1, 1, 1)))
        self.sh_degree = int(_get(model_cfg, "SH_DEGREE", 2))
        self.sigma = float(_get(model_cfg, "SIGMA", 0.8))
        self.eps = float(_get(model_cfg, "EPS", 1e-6))
        self.coef = nn.Linear(self.num_bev_features, (self.sh_degree + 1) ** 2)

    def forward(self, batch_dict):
        feats = batch_dict["point_features"].contiguous()
        coords = batch_dict["point_coords"].contiguous()
        # inference on a GPU: the coefficient Linear runs on our kernel and the map is written in the layout the
        # tensor-core convolutions read; `spatial_features` (fp32, the slot's contract) is written too unless the
        # caller says nobody will read it (batch_dict['pdm_fused_dense'], set by the PDMSSD detector)
        native = feats.is_cuda and not self.training and not torch.is_grad_enabled() and feats.dtype == torch.float32
        coef = linear_rows(feats, self.coef.weight, self.coef.bias) if native and self.coef.out_features <= 16 \
            else self.coef(feats).contiguous()
        args = (coords, feats, coef, int(batch_dict["batch_size"]), self.point_cloud_range, self.voxel_size,
                self.grid_size, self.dilation, self.sh_degree, self.sigma, self.eps)
        if native and self.num_bev_features % 8 == 0 and batch_dict.get("pdm_fused_dense", False):
            batch_dict["spatial_features_split"] = neck_forward(*args, output="split")
        elif native and self.num_bev_features % 8 == 0 and batch_dict.get("pdm_want_split", False):
            batch_dict["spatial_features"], batch_dict["spatial_features_split"] = neck_forward(*args, output="both")
        else:
            batch_dict["spatial_features"] = neck_forward(*args)
        batch_dict["spatial_features_stride"] = 1
        return batch_dict
