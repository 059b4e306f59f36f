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

from . import _lib, pointnet2_utils

# Inference fast path: ball query + ONE fused kernel per scale (grouping + shared MLP + max-pool,
# csrc/sa_fused.cu) instead of group / sub / cat / conv / bn / relu / pool.  Used only in eval mode
# without autograd, with our native backend, and when the scale fits the kernel's limits;
# otherwise the reference-shaped path below runs.  Set to False to force that path.
ENABLE_FUSED_SA = True


def _fold_mlp(mlp: nn.Sequential):
    """[(W' (out,in), b' (out))] with eval-mode BatchNorm folded into the 1x1 conv, or None when the
    Sequential is not the Conv2d(1x1,bias=False)+BatchNorm2d+ReLU pattern of pointnet2_modules.py:90-97."""
    layers = list(mlp)
    if len(layers) == 0 or len(layers) % 3 != 0:
        return None
    folded = []
    for conv, bn, act in zip(layers[0::3], layers[1::3], layers[2::3]):
        if not (isinstance(conv, nn.Conv2d) and isinstance(bn, nn.BatchNorm2d) and isinstance(act, nn.ReLU)):
            return None
        if conv.kernel_size != (1, 1) or conv.bias is not None or not bn.track_running_stats or not bn.affine:
            return None
        scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        folded.append((conv.weight[:, :, 0, 0] * scale[:, None], bn.bias - bn.running_mean * scale))
    return folded


def _pack_folded(folded):
    """Layout expected by pdm_sa_fused_forward: per layer Wt[k][pad4(out)] then bias[pad4(out)]."""
    chunks, widths = [], [folded[0][0].shape[1]]
    for w, b in folded:
        out_c, pad = w.shape[0], (w.shape[0] + 3) // 4 * 4
        wt = w.new_zeros((w.shape[1], pad))
        wt[:, :out_c] = w.t()
        bp = b.new_zeros(pad)
        bp[:out_c] = b
        chunks += [wt.reshape(-1), bp]
        widths.append(out_c)
    return torch.cat(chunks).contiguous(), widths


def _pack_folded_tc(folded):
    """Operands of the tcgen05 kernel (layout documented at SATCParams in csrc/sa_fused.cu):
    bias[L][128]; per layer, per block of <= 64 output columns: W_hi[K/4][ns][4], W_lo[K/4][ns][4] with
    K padded to 8 and N to 16, hi = the 19 leading bits (what a tf32 operand keeps), lo = w - hi."""
    L = len(folded)
    bias = folded[0][0].new_zeros((L, 128))
    chunks = []
    for l, (w, b) in enumerate(folded):           # w (N, K)
        n_out, k_in = w.shape
        kpad, npad = (k_in + 7) // 8 * 8, (n_out + 15) // 16 * 16
        wp = w.new_zeros((npad, kpad))
        wp[:n_out, :k_in] = w
        bias[l, :n_out] = b
        hi = (wp.contiguous().view(torch.int32) & -8192).view(torch.float32)    # 0xffffe000
        lo = wp - hi
        for n0 in range(0, npad, 64):
            for part in (hi, lo):
                blk = part[n0:n0 + 64]                                         # (ns, kpad)
                chunks.append(blk.reshape(blk.shape[0], kpad // 4, 4).permute(1, 0, 2).reshape(-1))
    return torch.cat([bias.reshape(-1)] + chunks).contiguous()


def _pack_folded_tc3(folded):
    """Operands of the persistent tcgen05 kernel (csrc/sa_tc.cu): per layer the folded weights W'[n][k] as two
    bf16 planes hi|lo (hi + lo = w to 2^-18), each plane [kpad/8][npad][8] with kpad, npad = widths rounded
    up to 16 -- the K-major no-swizzle core-matrix layout, so a layer is one straight bulk copy into shared memory;
    bias fp32 (L, 128).  Returns (planes as a bf16 tensor, bias) or None when the stack does not fit the kernel."""
    if len(folded) < 2 or len(folded) > 3:
        return None
    bias = folded[0][0].new_zeros((len(folded), 128))
    chunks = []
    for l, (w, b) in enumerate(folded):           # w (N, K)
        n_out, k_in = w.shape
        if n_out > 128 or k_in > 96:
            return None
        kpad, npad = (k_in + 15) // 16 * 16, (n_out + 15) // 16 * 16
        wp = w.new_zeros((npad, kpad))
        wp[:n_out, :k_in] = w
        bias[l, :n_out] = b
        hi = wp.to(torch.bfloat16)
        lo = (wp - hi.float()).to(torch.bfloat16)
        planes = torch.stack([hi, lo]).reshape(2, npad, kpad // 8, 8).permute(0, 2, 1, 3)          # plane, k8, n, 8
        chunks.append(planes.reshape(-1))
    return torch.cat(chunks).contiguous(), bias.contiguous()


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
        for si, (grouper, mlp) in enumerate(zip(self.groupers, self.mlps)):
            fused = self._fused_scale(si, grouper, mlp, xyz, new_xyz, features)
            if fused is not None:
                pooled.append(fused)
                continue
            grouped = grouper(xyz, new_xyz, features)          # (B, C_in, npoint, nsample)
            pooled.append(self._pool(mlp(grouped)).squeeze(-1))  # (B, C_out, npoint)
        return new_xyz, (pooled[0] if len(pooled) == 1 else torch.cat(pooled, dim=1))

    def _fused_scale(self, si, grouper, mlp, xyz, new_xyz, features):
        """(B, C_out, npoint) through csrc/sa_fused.cu, or None when the fast path does not apply."""
        if not ENABLE_FUSED_SA or self.training or torch.is_grad_enabled() or self.pool_method != 'max_pool':
            return None
        if not isinstance(grouper, pointnet2_utils.QueryAndGroup) or not xyz.is_cuda:
            return None
        if pointnet2_utils.get_backend() is not pointnet2_utils._native:
            return None
        S = grouper.nsample
        if S < 4 or S > 128 or (S & (S - 1)) != 0 or (features is None and not grouper.use_xyz):
            return None
        # cache key: storage address + in-place version of every tensor of the stack (a replaced Parameter changes the
        # address, an in-place update the version)
        version = tuple((t.data_ptr(), t._version) for t in list(mlp.parameters()) + list(mlp.buffers()))
        cache = self.__dict__.setdefault('_fused_cache', {})
        hit = cache.get(si)
        if hit is None or hit[0] != version or hit[1].device != xyz.device:
            folded = _fold_mlp(mlp)
            if folded is None or len(folded) > 4 or max(max(w.shape) for w, _ in folded) > 128:
                cache[si] = (version, xyz.new_zeros(1), None)
            else:
                fl = [(w.detach().float(), b.detach().float()) for w, b in folded]
                packed, widths = _pack_folded(fl)
                tc3 = _pack_folded_tc3(fl)
                cache[si] = (version, packed.to(xyz.device), widths, _pack_folded_tc(fl).to(xyz.device),
                             tc3[0].to(xyz.device) if tc3 is not None else None, tc3[1].to(xyz.device) if tc3 is not None else None)
            hit = cache[si]
        packed, widths, packed_tc = hit[1], hit[2], (hit[3] if len(hit) > 3 else None)
        packed_tc3, bias_tc3 = (hit[4], hit[5]) if len(hit) > 5 else (None, None)
        if widths is None:
            return None
        B, N, _ = xyz.shape
        M = new_xyz.shape[1]
        c_feat = 0 if features is None else features.shape[1]
        if widths[0] != (3 if grouper.use_xyz else 0) + c_feat:
            return None
        if features is not None and (features.dtype != torch.float32 or not features.is_cuda):
            return None                                   # fp16 / fp64 / CPU features: the reference-shaped path handles them
        idx = pointnet2_utils.ball_query(grouper.radius, S, xyz, new_xyz)
        out = torch.empty((B, widths[-1], M), dtype=torch.float32, device=xyz.device)
        out_pm = torch.empty((B, M, widths[-1]), dtype=torch.float32, device=xyz.device)
        feats = features.contiguous() if features is not None else None
        # the same features point-major (B, N, C), when the producing SA layer left them (see below)
        feats_pm = getattr(features, '_pdm_point_major', None) if features is not None else None
        if feats_pm is not None and (feats_pm.shape != (B, N, c_feat) or not feats_pm.is_contiguous()):
            feats_pm = None
        import ctypes
        warr = (ctypes.c_int * len(widths))(*widths)
        with torch.cuda.device(xyz.device):
            rc = _lib.load().pdm_sa_fused_forward_v2(
                B, N, M, c_feat, S, 1 if grouper.use_xyz else 0, xyz.data_ptr(),
                feats.data_ptr() if feats is not None else None, feats_pm.data_ptr() if feats_pm is not None else None,
                new_xyz.data_ptr(), idx.data_ptr(), len(widths) - 1, warr, packed.data_ptr(),
                packed_tc.data_ptr() if packed_tc is not None else None,
                packed_tc3.data_ptr() if packed_tc3 is not None else None,
                bias_tc3.data_ptr() if bias_tc3 is not None else None, out.data_ptr(), out_pm.data_ptr(),
                torch.cuda.current_stream(xyz.device).cuda_stream)
        _lib.check(rc, "pdm_sa_fused_forward")
        # the kernels write the result twice: (B, C, M) for the API and (B, M, C) for whoever gathers rows next
        # (the next SA layer, the backbone's `point_features`); the copy rides along as an attribute
        out._pdm_point_major = out_pm
        return out


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
