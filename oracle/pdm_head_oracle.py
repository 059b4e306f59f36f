"""Pure-torch CPU restatement of SPEC_HEAD.md: BEV context block + hybrid (heatmap + point) head.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED for the assembly (the reference tree has no hybrid head,
SURVEY.md section 0.1); every step follows the in-tree code it is built from, cited per function.
`decode` is pinned to the reference's own PointResidualCoder.decode_torch by tests/test_head_cpu.py
(imported from /root/reference here, from oracle/_ref/py on the GPU box).  Written functionally on a
plain state_dict so that it shares no code with pdm_ssd_b200/detector.py.
"""
import numpy as np
import torch
import torch.nn.functional as F

KITTI_MEAN_SIZE = ((3.9, 1.6, 1.56), (0.8, 0.6, 1.73), (1.76, 0.6, 1.73))


def _bn(x, sd, prefix, eps):
    """eval-mode BatchNorm{1,2}d on running statistics"""
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training=False, eps=eps)


def bev_context(sd, spatial_features, num_layers, prefix="blocks"):
    """base_bev_backbone.py:27-47 at stride 1: [Conv2d 3x3 pad 1 bias=False, BatchNorm2d(eps 1e-3), ReLU] x L."""
    x = spatial_features
    for i in range(num_layers):
        x = F.conv2d(x, sd["%s.%d.weight" % (prefix, 3 * i)], None, padding=1)
        x = F.relu(_bn(x, sd, "%s.%d" % (prefix, 3 * i + 1), 1e-3))
    return x


def heatmap_branch(sd, spatial_features_2d):
    """center_head.py:63-72 (shared conv) and :12-46 (SeparateHead 'hm': conv-bn-relu, conv with bias)."""
    x = F.relu(_bn(F.conv2d(spatial_features_2d, sd["shared_conv.0.weight"], sd["shared_conv.0.bias"], padding=1),
                   sd, "shared_conv.1", 1e-5))
    h = F.relu(_bn(F.conv2d(x, sd["hm.0.weight"], sd["hm.0.bias"], padding=1), sd, "hm.1", 1e-5))
    return x, torch.sigmoid(F.conv2d(h, sd["hm.3.weight"], sd["hm.3.bias"], padding=1))


def fc_stack(sd, prefix, x):
    """point_head_template.py:36-47 with one hidden layer: Linear(no bias) + BatchNorm1d + ReLU + Linear(bias)."""
    h = F.relu(_bn(F.linear(x, sd[prefix + ".0.weight"]), sd, prefix + ".1", 1e-5))
    return F.linear(h, sd[prefix + ".3.weight"], sd[prefix + ".3.bias"])


def point_pillars(point_coords, point_cloud_range, voxel_size, X, Y):
    """SPEC_HEAD step 2 (dynamic_voxel_vfe.py:60-71 convention, clamped)."""
    rmin = torch.tensor(point_cloud_range[:2], dtype=torch.float32)
    v = torch.tensor(voxel_size[:2], dtype=torch.float32)
    c = torch.floor((point_coords[:, 1:3].float() - rmin) / v).long()
    return c[:, 0].clamp(0, X - 1), c[:, 1].clamp(0, Y - 1)


def decode(box_enc, points, label, mean_size=KITTI_MEAN_SIZE):
    """box_coder_utils.py:189-222, use_mean_size=True; `label` is 0-based (the reference passes label + 1 and subtracts)."""
    ms = torch.tensor(mean_size, dtype=torch.float32)[label]
    xt, yt, zt, dxt, dyt, dzt, cost, sint = [box_enc[:, i] for i in range(8)]
    dxa, dya, dza = ms[:, 0], ms[:, 1], ms[:, 2]
    diag = torch.sqrt(dxa ** 2 + dya ** 2)
    return torch.stack([xt * diag + points[:, 0], yt * diag + points[:, 1], zt * dza + points[:, 2],
                        torch.exp(dxt) * dxa, torch.exp(dyt) * dya, torch.exp(dzt) * dza, torch.atan2(sint, cost)], dim=1)


def point_head(sd, spatial_features_2d, point_coords, point_features, point_cloud_range, voxel_size, num_class=3):
    """SPEC_HEAD steps 1-6.  Returns dict(x, heatmap, cls, box, scores, best, label, boxes)."""
    x, hm = heatmap_branch(sd, spatial_features_2d)
    Y, X = hm.shape[2], hm.shape[3]
    cx, cy = point_pillars(point_coords, point_cloud_range, voxel_size, X, Y)
    b = point_coords[:, 0].long()
    fused = torch.cat([point_features, x[b, :, cy, cx]], dim=1)
    cls = fc_stack(sd, "cls_layers", fused)
    box = fc_stack(sd, "box_layers", fused)
    scores = torch.sigmoid(cls) * torch.sqrt(hm[b, :, cy, cx])
    best, label = scores.max(dim=1)
    boxes = decode(box, point_coords[:, 1:4], label, KITTI_MEAN_SIZE[:num_class])
    return dict(x=x, heatmap=hm, cls=cls, box=box, scores=scores, best=best, label=label, boxes=boxes)


def post_process(boxes, best, label, batch_size, score_thresh, nms_thresh, pre_max, post_max):
    """SPEC_HEAD step 7 on the C oracle's rotated NMS (detector3d_template.py:199-254, model_nms_utils.py:6-25).
    -> detections (B, post_max, 9), num (B)."""
    import oracle
    M = best.numel() // batch_size
    det = torch.zeros((batch_size, post_max, 9), dtype=torch.float32)
    num = torch.zeros((batch_size,), dtype=torch.int32)
    for f in range(batch_size):
        s, bx, lb = best[f * M:(f + 1) * M], boxes[f * M:(f + 1) * M], label[f * M:(f + 1) * M]
        idx = torch.nonzero(s >= score_thresh).view(-1)
        if idx.numel() == 0:
            continue
        top, order = torch.topk(s[idx], k=min(pre_max, idx.numel()))
        cand = idx[order]
        keep = oracle.nms_bev(bx[cand].numpy(), nms_thresh)[:post_max]
        sel = cand[torch.from_numpy(keep.astype(np.int64))]
        n = sel.numel()
        det[f, :n, :7] = bx[sel]
        det[f, :n, 7] = s[sel]
        det[f, :n, 8] = lb[sel].float() + 1
        num[f] = n
    return det, num
