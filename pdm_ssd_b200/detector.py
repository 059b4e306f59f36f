"""PDM-SSD single-stage detector assembled from pcdet-style plugins (inference path).

    points -> PDMSSDBackbone (SA chain, our CUDA ops) -> PDMNeck (CUDA, SPEC_PDM.md)
           -> BEV context convs -> HybridHead (BEV heatmap + per-point vote/box head, fused scores)
           -> fixed-shape top-K detections

The module list / batch_dict protocol is pcdet's (`Detector3DTemplate`, detector3d_template.py:14-50,
point_rcnn.py:9-22).  The hybrid head has NO reference code: this is a compact
builder-defined reading of the abstract -- "scene heatmap is predicted to complement the voting point
set ... target probability of detected boxes are calibrated through feature fusion" -- built from the
closest in-tree pieces: CenterHead's shared 3x3 conv + heatmap branch with -2.19 bias
(center_head.py:12-46,57-99) and PointHeadBox's FC stacks + PointResidualCoder.decode_torch
(point_head_template.py:36-47, point_head_box.py:71-115, box_coder_utils.py:189-222).
Post-processing is pcdet's class-agnostic rotated NMS (detector3d_template.py:199-254,
model_nms_utils.py:6-25) for all frames at once on the device (`iou3d_nms_utils.batched_nms_gpu`).
SPEC_HEAD.md is the specification of the BEV context block and the hybrid head.  At inference on a GPU the
dense layers run on our own kernels: the 3x3 conv + BN + ReLU stacks as tcgen05 implicit GEMMs
(csrc/conv_tc.cu, activations handed from layer to layer in the split NHWC8 layout), the per-point FC
stacks + score fusion + box decode as one fused kernel (csrc/point_head.cu).  The same modules keep a plain
torch path (training, CPU, `ENABLE_FUSED_DENSE = False`) which is also what the parity tests compare with.
"""
import ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .backbone import AttrDict, PDMSSDBackbone
from .conv_tc import ACT_RELU, ACT_SIGMOID, PackedConv
from .iou3d_nms_utils import batched_multi_classes_nms_gpu, batched_nms_gpu
from .pdm_neck import PDMNeck

# Inference fast path for the dense layers (see module docstring).  False forces the torch modules.
ENABLE_FUSED_DENSE = True


def _param_key(module):
    """Cache key of a module's parameters and buffers: storage address + in-place version of every tensor."""
    return tuple((t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers()))


def _fused_dense_ok(module, x):
    return (ENABLE_FUSED_DENSE and x is not None and not module.training and not torch.is_grad_enabled())

KITTI_MEAN_SIZE = [[3.9, 1.6, 1.56], [0.8, 0.6, 1.73], [1.76, 0.6, 1.73]]  # Car, Pedestrian, Cyclist (l,w,h)


def default_cfg(num_points=16384):
    """KITTI 3-class configuration (the reference's YAMLs are git-ignored)."""
    return AttrDict(
        CLASS_NAMES=['Car', 'Pedestrian', 'Cyclist'],
        POINT_CLOUD_RANGE=[0.0, -40.0, -3.0, 70.4, 40.0, 1.0],
        VOXEL_SIZE=[0.4, 0.4, 0.4],
        BACKBONE_3D=dict(NAME='PDMSSDBackbone', SA_CONFIG=dict(
            NPOINTS=[num_points // 4, num_points // 16], RADIUS=[[0.8], [1.6]], NSAMPLE=[[32], [32]],
            MLPS=[[[16, 16, 32]], [[64, 64, 128]]], USE_XYZ=True)),
        MAP_TO_BEV=dict(NAME='PDMNeck', NUM_BEV_FEATURES=128, DILATION=[1, 1, 1], SH_DEGREE=2, SIGMA=0.8),
        BACKBONE_2D=dict(NUM_FILTERS=128, LAYER_NUM=2),
        DENSE_HEAD=dict(NAME='HybridHead', SHARED_CONV_CHANNEL=64, CLS_FC=[128], REG_FC=[128], MAX_OBJ_PER_SAMPLE=100),
        # key names of detector3d_template.py:195-275 / model_nms_utils.py:15-20
        POST_PROCESSING=dict(SCORE_THRESH=0.1, NMS_CONFIG=dict(
            MULTI_CLASSES_NMS=False, NMS_TYPE='nms_gpu', NMS_THRESH=0.1, NMS_PRE_MAXSIZE=4096, NMS_POST_MAXSIZE=100)),
    )


class BEVContext(nn.Module):
    """'Context learning' block of docs/workflow.svg: 3x3 conv + BN(eps 1e-3) + ReLU stack, the
    first block of BaseBEVBackbone (base_bev_backbone.py:27-47) at stride 1."""

    def __init__(self, model_cfg, input_channels):
        super().__init__()
        c = model_cfg.NUM_FILTERS
        layers, cin = [], input_channels
        for _ in range(model_cfg.LAYER_NUM):
            layers += [nn.Conv2d(cin, c, 3, padding=1, bias=False), nn.BatchNorm2d(c, eps=1e-3, momentum=0.01), nn.ReLU()]
            cin = c
        self.blocks = nn.Sequential(*layers)
        self.num_bev_features = c

    def _packed(self, device):
        key = (_param_key(self), str(device))
        hit = self.__dict__.get('_packed_cache')
        if hit is None or hit[0] != key:
            mods = list(self.blocks)
            ok = all(m.in_channels % 32 == 0 and m.out_channels % 8 == 0 and m.out_channels <= 128 for m in mods[0::3])
            layers = [PackedConv(conv, bn, act=ACT_RELU, device=device) for conv, bn in zip(mods[0::3], mods[1::3])] if ok else None
            hit = (key, layers)
            self.__dict__['_packed_cache'] = hit
        return hit[1]

    def forward(self, batch_dict):
        split = batch_dict.get('spatial_features_split')
        if _fused_dense_ok(self, split):
            layers = self._packed(split.data.device)
            if layers is not None:
                for layer in layers:
                    split, _ = layer(split, want_split=True)
                batch_dict['spatial_features_2d_split'] = split
                if not batch_dict.get('pdm_fused_dense', False):
                    batch_dict['spatial_features_2d'] = split.to_nchw()      # the slot's fp32 contract
                return batch_dict
        batch_dict['spatial_features_2d'] = self.blocks(batch_dict['spatial_features'])
        return batch_dict


def _fc(cin, widths, cout):
    layers = []
    for w in widths:
        layers += [nn.Linear(cin, w, bias=False), nn.BatchNorm1d(w), nn.ReLU()]
        cin = w
    layers.append(nn.Linear(cin, cout, bias=True))
    return nn.Sequential(*layers)


class HybridHead(nn.Module):
    def __init__(self, model_cfg, input_channels, point_channels, num_class, point_cloud_range, voxel_size,
                 post_cfg=None):
        super().__init__()
        self.post_cfg = post_cfg          # POST_PROCESSING block; None = score top-K without NMS
        self.num_class = num_class
        self.range = point_cloud_range
        self.voxel = voxel_size
        self.topk = model_cfg.MAX_OBJ_PER_SAMPLE
        sc = model_cfg.SHARED_CONV_CHANNEL
        self.shared_conv = nn.Sequential(nn.Conv2d(input_channels, sc, 3, padding=1, bias=True), nn.BatchNorm2d(sc), nn.ReLU())
        self.hm = nn.Sequential(nn.Conv2d(sc, sc, 3, padding=1, bias=True), nn.BatchNorm2d(sc), nn.ReLU(),
                                nn.Conv2d(sc, num_class, 3, padding=1, bias=True))
        self.hm[-1].bias.data.fill_(-2.19)  # center_head.py:38-39
        fused = point_channels + sc              # point feature + the BEV context under the point
        self.cls_layers = _fc(fused, model_cfg.CLS_FC, num_class)
        self.box_layers = _fc(fused, model_cfg.REG_FC, 8)   # PointResidualCoder: 8 = xyz, lwh, sin, cos
        self.register_buffer('mean_size', torch.tensor(KITTI_MEAN_SIZE[:num_class], dtype=torch.float32))

    def decode(self, enc, points, cls_idx):
        """box_coder_utils.py:189-222 (PointResidualCoder.decode_torch, use_mean_size=True)."""
        xt, yt, zt, dxt, dyt, dzt, cost, sint = torch.split(enc, 1, dim=-1)
        xa, ya, za = torch.split(points, 1, dim=-1)
        ms = self.mean_size[cls_idx]
        dxa, dya, dza = torch.split(ms, 1, dim=-1)
        diag = torch.sqrt(dxa ** 2 + dya ** 2)
        xg, yg, zg = xt * diag + xa, yt * diag + ya, zt * dza + za
        dxg, dyg, dzg = torch.exp(dxt) * dxa, torch.exp(dyt) * dya, torch.exp(dzt) * dza
        return torch.cat([xg, yg, zg, dxg, dyg, dzg, torch.atan2(sint, cost)], dim=-1)

    def _packed(self, device):
        key = (_param_key(self), str(device))
        hit = self.__dict__.get('_packed_cache')
        if hit is None or hit[0] != key:
            pk = None
            cls, box = list(self.cls_layers), list(self.box_layers)
            sc = self.shared_conv[0]
            if (len(cls) == 4 and len(box) == 4 and sc.in_channels % 32 == 0 and sc.out_channels % 32 == 0 and sc.out_channels <= 128
                    and cls[0].out_features % 4 == 0 and box[0].out_features % 4 == 0
                    and cls[0].out_features + box[0].out_features <= 256 and self.num_class <= 8
                    and (cls[0].in_features - sc.out_channels) % 4 == 0):
                def fold(lin, bn):
                    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
                    return lin.weight.detach().float() * scale[:, None], bn.bias.detach().float() - bn.running_mean.detach().float() * scale
                wc, bc = fold(cls[0], cls[1])
                wb, bb = fold(box[0], box[1])
                pk = AttrDict()
                pk.shared = PackedConv(self.shared_conv[0], self.shared_conv[1], act=ACT_RELU, device=device)
                pk.hm1 = PackedConv(self.hm[0], self.hm[1], act=ACT_RELU, device=device)
                pk.hm2 = PackedConv(self.hm[3], None, act=ACT_SIGMOID, device=device)
                pk.w1t = torch.cat([wc, wb], 0).t().contiguous().to(device)
                pk.b1 = torch.cat([bc, bb]).contiguous().to(device)
                pk.w2c, pk.b2c = cls[3].weight.detach().float().contiguous().to(device), cls[3].bias.detach().float().contiguous().to(device)
                pk.w2b, pk.b2b = box[3].weight.detach().float().contiguous().to(device), box[3].bias.detach().float().contiguous().to(device)
                pk.hc, pk.hb = cls[0].out_features, box[0].out_features
                pk.mean_size = self.mean_size.detach().float().contiguous().to(device)
            hit = (key, pk)
            self.__dict__['_packed_cache'] = hit
        return hit[1]

    def _point_head_fused(self, pk, coords, pf, x_split, hm):
        """SPEC_HEAD steps 2-6 in one kernel (csrc/point_head.cu)."""
        P, Cp = pf.shape
        dev = pf.device
        score = torch.empty((P, self.num_class), dtype=torch.float32, device=dev)
        boxes = torch.empty((P, 7), dtype=torch.float32, device=dev)
        best = torch.empty((P,), dtype=torch.float32, device=dev)
        label = torch.empty((P,), dtype=torch.int32, device=dev)
        f2 = ctypes.c_float * 2
        with torch.cuda.device(dev):
            rc = _lib.load().pdm_point_head_forward(
                P, x_split.B, Cp, x_split.C, x_split.Y, x_split.X, self.num_class, pk.hc, pk.hb,
                f2(float(self.range[0]), float(self.range[1])), f2(float(self.voxel[0]), float(self.voxel[1])),
                coords.data_ptr(), pf.data_ptr(), x_split.data.data_ptr(), hm.data_ptr(), pk.w1t.data_ptr(), pk.b1.data_ptr(),
                pk.w2c.data_ptr(), pk.b2c.data_ptr(), pk.w2b.data_ptr(), pk.b2b.data_ptr(), pk.mean_size.data_ptr(),
                score.data_ptr(), boxes.data_ptr(), best.data_ptr(), label.data_ptr(), None, None,
                torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "pdm_point_head_forward")
        return score, boxes, best, label.long()

    def forward(self, batch_dict):
        B = batch_dict['batch_size']
        coords, pf = batch_dict['point_coords'].contiguous(), batch_dict['point_features'].contiguous()
        s2d = batch_dict.get('spatial_features_2d_split')
        pk = self._packed(pf.device) if _fused_dense_ok(self, s2d) and pf.dtype == torch.float32 else None
        if pk is not None:
            xs, _ = pk.shared(s2d, want_split=True)
            hs, _ = pk.hm1(xs, want_split=True)
            _, hm = pk.hm2(hs, want_split=False, want_nchw=True)            # sigmoid in the epilogue
            score, boxes, best, label = self._point_head_fused(pk, coords, pf, xs, hm)
        else:
            x = self.shared_conv(batch_dict['spatial_features_2d'])
            hm = torch.sigmoid(self.hm(x))                                   # (B, num_class, Y, X) scene heatmap
            # pillar of every centre (same floor((p - min)/v) convention as the neck)
            Y, X = hm.shape[2], hm.shape[3]
            cx = torch.floor((coords[:, 1] - self.range[0]) / self.voxel[0]).long().clamp_(0, X - 1)
            cy = torch.floor((coords[:, 2] - self.range[1]) / self.voxel[1]).long().clamp_(0, Y - 1)
            b = coords[:, 0].long()
            bev_at_pt = x[b, :, cy, cx]                                       # (P, sc) feature fusion
            hm_at_pt = hm[b, :, cy, cx]                                       # (P, num_class)
            fused = torch.cat([pf, bev_at_pt], dim=1)
            cls = self.cls_layers(fused)
            box = self.box_layers(fused)
            score = torch.sigmoid(cls) * hm_at_pt.sqrt()                      # point score calibrated by the heatmap
            best, label = score.max(dim=1)
            boxes = self.decode(box, coords[:, 1:4], label)
        batch_dict.update(batch_cls_preds=score, batch_box_preds=boxes, batch_index=coords[:, 0],
                          cls_preds_normalized=True, heatmap=hm)
        # fixed-shape detections (B, K, 9) = box7, score, label -- ready for one NCCL gather
        M = best.numel() // B
        if self.post_cfg is None:
            k = min(self.topk, M)
            top, idx = best.view(B, M).topk(k, dim=1)
            valid = torch.ones_like(idx, dtype=torch.bool)
        elif self.post_cfg.NMS_CONFIG.get('MULTI_CLASSES_NMS', False):
            # per-class suppression (detector3d_template.py:219-233 -> multi_classes_nms, model_nms_utils.py:28-66):
            # class-major rows like the reference, (B, C * NMS_POST_MAXSIZE, 9)
            nms = self.post_cfg.NMS_CONFIG
            C, p = self.num_class, nms.NMS_POST_MAXSIZE
            sel, num = batched_multi_classes_nms_gpu(boxes.view(B, M, 7), score.view(B, M, C), nms.NMS_THRESH, nms.NMS_PRE_MAXSIZE,
                                                     p, score_thresh=self.post_cfg.SCORE_THRESH, nms_type=nms.get('NMS_TYPE', 'nms_gpu'))
            valid = (sel >= 0).view(B, C * p)
            idx = sel.clamp(min=0).view(B, C * p)
            cls_of_row = torch.arange(C, device=idx.device).repeat_interleave(p)[None].expand(B, C * p)
            top = torch.gather(score.view(B, M, C), 1, idx[..., None].expand(B, C * p, C)).gather(2, cls_of_row[..., None]).squeeze(2)
            gather = (idx + (torch.arange(B, device=idx.device) * M)[:, None]).flatten()
            det = torch.cat([boxes[gather].view(B, C * p, 7), top[..., None], cls_of_row[..., None].float() + 1], dim=2)
            batch_dict['num_detections'] = num.sum(dim=1).to(torch.int32)
            batch_dict['num_detections_per_class'] = num
            batch_dict['detections'] = det * valid[..., None]
            return batch_dict
        else:
            # Detector3DTemplate.post_processing (detector3d_template.py:199-254) with class-agnostic NMS,
            # for every frame in one pass and without leaving the device
            nms = self.post_cfg.NMS_CONFIG
            k = nms.NMS_POST_MAXSIZE
            idx, num = batched_nms_gpu(boxes.view(B, M, 7), best.view(B, M), nms.NMS_THRESH, nms.NMS_PRE_MAXSIZE,
                                       k, score_thresh=self.post_cfg.SCORE_THRESH, nms_type=nms.get('NMS_TYPE', 'nms_gpu'))
            valid = idx >= 0
            idx = idx.clamp(min=0)
            top = torch.gather(best.view(B, M), 1, idx)
            batch_dict['num_detections'] = num
        gather = (idx + (torch.arange(B, device=idx.device) * M)[:, None]).flatten()
        det = torch.cat([boxes[gather].view(B, k, 7), top[..., None], label[gather].view(B, k, 1).float() + 1], dim=2)
        batch_dict['detections'] = det * valid[..., None]      # rows past num_detections are zero (label 0)
        return batch_dict


class PDMSSD(nn.Module):
    """module_list protocol of Detector3DTemplate.forward loops (point_rcnn.py:10-11)."""

    def __init__(self, cfg=None, input_channels=4):
        super().__init__()
        cfg = cfg or default_cfg()
        self.cfg = cfg
        self.backbone_3d = PDMSSDBackbone(cfg.BACKBONE_3D, input_channels)
        neck_cfg = AttrDict(cfg.MAP_TO_BEV, NUM_BEV_FEATURES=self.backbone_3d.num_point_features,
                            VOXEL_SIZE=cfg.VOXEL_SIZE, POINT_CLOUD_RANGE=cfg.POINT_CLOUD_RANGE)
        self.map_to_bev_module = PDMNeck(neck_cfg)
        self.backbone_2d = BEVContext(cfg.BACKBONE_2D, self.map_to_bev_module.num_bev_features)
        self.dense_head = HybridHead(cfg.DENSE_HEAD, self.backbone_2d.num_bev_features, self.backbone_3d.num_point_features,
                                     len(cfg.CLASS_NAMES), cfg.POINT_CLOUD_RANGE, cfg.VOXEL_SIZE,
                                     post_cfg=cfg.get('POST_PROCESSING'))
        self.module_list = [self.backbone_3d, self.map_to_bev_module, self.backbone_2d, self.dense_head]

    @torch.no_grad()
    def forward(self, batch_dict):
        # in eval mode nobody between the neck and the head reads the fp32 BEV maps: keep them in the split layout
        # of the tensor-core convolutions only (set batch_dict['pdm_fused_dense'] = False to get them back)
        if ENABLE_FUSED_DENSE and not self.training and batch_dict['points'].is_cuda:
            batch_dict.setdefault('pdm_fused_dense', True)
        for m in self.module_list:
            batch_dict = m(batch_dict)
        return batch_dict


def gather_detections(det, group=None):
    """(frames_per_rank, K, 9) on every rank -> (world*frames_per_rank, K, 9) on every rank; the one
    collective of the sharded pipeline (replaces the reference's pickle-file merge,
    common_utils.py:229-250).  Identity without an initialised process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return det
    out = det.new_empty((dist.get_world_size(group) * det.shape[0],) + tuple(det.shape[1:]))
    dist.all_gather_into_tensor(out, det.contiguous(), group=group)
    return out
