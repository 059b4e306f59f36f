"""GPU parity of the detector's dense half (BEV context convs + hybrid head) -- SPEC_HEAD.md:
 * the fused kernels (tcgen05 convolutions, fused point head) against the product's torch path (same modules,
   cuDNN/cuBLAS fp32, TF32 off) and against the independent CPU oracle (oracle/pdm_head_oracle.py);
 * tolerance 1e-3 relative on features, scores and boxes (BASELINE.json north_star), integer-exact labels where the
   scores are not tied to within the tolerance;
 * detections compared as sets matched by position; a flipped keep decision is reported with its margin."""
import numpy as np
import pytest
import torch

import pdm_head_oracle as ho
from pdm_ssd_b200 import detector as D, synthetic
from pdm_ssd_b200.conv_tc import SplitAct
from pdm_ssd_b200.pdm_neck import linear_rows

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _fp32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


def _unmatched(a, na, b, nb, rtol=1e-3, atol=1e-3):
    """Detections as sets: rows of a[:na] without a counterpart in b[:nb] (box, score and label within tolerance).
    A score wobble of 1e-5 may swap two neighbours in the score order or flip a keep decision that sits on the NMS /
    score threshold, so positions are not compared; the count of unmatched rows is reported and bounded."""
    if na == 0 or nb == 0:
        return na
    d = (a[:na, None, :8] - b[None, :nb, :8]).abs()
    tol = atol + rtol * b[None, :nb, :8].abs()
    ok = (d <= tol).all(dim=2) & (a[:na, None, 8] == b[None, :nb, 8])
    return int((~ok.any(dim=1)).sum())


def _model(npts=4096, seed=0):
    torch.manual_seed(seed)
    model = D.PDMSSD(D.default_cfg(npts)).to(DEV).eval()
    g = torch.Generator().manual_seed(seed + 1)
    for m in model.modules():                                   # non-trivial eval-mode statistics
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) * 0.8 + 0.6)
    return model


def _points(B, N, first=0):
    return torch.from_numpy(synthetic.to_pcdet_points(synthetic.kitti_batch(B, N, first_frame=first))).to(DEV)


def test_linear_rows_matches_torch():
    torch.manual_seed(1)
    lin = torch.nn.Linear(128, 9).to(DEV)
    x = torch.randn(5000, 128, device=DEV)
    with torch.no_grad():
        assert _rel(linear_rows(x, lin.weight, lin.bias), lin(x)) < 1e-5


def test_fused_dense_path_matches_torch_path(monkeypatch):
    model = _model()
    pts = _points(3, 4096, first=5)
    fused = model({"batch_size": 3, "points": pts})
    assert "spatial_features_2d" not in fused and "spatial_features_2d_split" in fused      # nothing materialised in fp32
    keep = model({"batch_size": 3, "points": pts, "pdm_fused_dense": False, "pdm_want_split": True})
    monkeypatch.setattr(D, "ENABLE_FUSED_DENSE", False)
    plain = model({"batch_size": 3, "points": pts})
    assert "spatial_features_2d_split" not in plain
    # the neck's two output layouts hold the same map
    assert _rel(keep["spatial_features_split"].to_nchw(), keep["spatial_features"]) < 2e-5
    assert torch.equal(keep["spatial_features"], plain["spatial_features"])
    # BEV context block, heatmap, per-point outputs
    assert _rel(keep["spatial_features_2d"], plain["spatial_features_2d"]) < 1e-3
    assert _rel(fused["spatial_features_2d_split"].to_nchw(), plain["spatial_features_2d"]) < 1e-3
    assert _rel(fused["heatmap"], plain["heatmap"]) < 1e-3
    assert _rel(fused["batch_cls_preds"], plain["batch_cls_preds"]) < 1e-3
    assert _rel(fused["batch_box_preds"], plain["batch_box_preds"]) < 1e-3
    assert torch.equal(fused["batch_index"], plain["batch_index"])
    # detections: same boxes at the same positions, up to keep decisions that sit within 1e-3 of a threshold
    df, dp = fused["detections"].cpu(), plain["detections"].cpu()
    nf, np_ = fused["num_detections"].cpu(), plain["num_detections"].cpu()
    miss = sum(_unmatched(df[f], int(nf[f]), dp[f], int(np_[f])) for f in range(3))
    print("detections: fused %s, torch path %s, fused rows without a counterpart: %d" % (nf.tolist(), np_.tolist(), miss))
    assert miss <= 0.03 * int(nf.sum()) and int((nf - np_).abs().max()) <= 3


def test_fused_head_matches_cpu_oracle():
    model = _model(seed=3)
    B = 2
    pts = _points(B, 4096, first=11)
    out = model({"batch_size": B, "points": pts})
    head, ctx = model.dense_head, model.backbone_2d
    cfg = model.cfg
    sf2d = out["spatial_features_2d_split"].to_nchw().cpu()
    sd = {k: v.detach().cpu() for k, v in head.state_dict().items()}
    res = ho.point_head(sd, sf2d, out["point_coords"].cpu(), out["point_features"].cpu(), cfg.POINT_CLOUD_RANGE, cfg.VOXEL_SIZE)
    assert _rel(out["heatmap"].cpu(), res["heatmap"]) < 1e-3
    assert _rel(out["batch_cls_preds"].cpu(), res["scores"]) < 1e-3
    assert _rel(out["batch_box_preds"].cpu(), res["boxes"]) < 1e-3
    # the BEV context block against the oracle's restatement, from the neck's fp32 map
    keep = model({"batch_size": B, "points": pts, "pdm_fused_dense": False, "pdm_want_split": True})
    ctx_sd = {k: v.detach().cpu() for k, v in ctx.state_dict().items()}
    ref2d = ho.bev_context(ctx_sd, keep["spatial_features"].cpu(), 2)
    assert _rel(keep["spatial_features_2d"].cpu(), ref2d) < 1e-3
    # step 7 on the oracle's NMS
    post = cfg.POST_PROCESSING
    det, num = ho.post_process(res["boxes"], res["best"], res["label"], B, post.SCORE_THRESH, post.NMS_CONFIG.NMS_THRESH,
                               post.NMS_CONFIG.NMS_PRE_MAXSIZE, post.NMS_CONFIG.NMS_POST_MAXSIZE)
    got, gnum = out["detections"].cpu(), out["num_detections"].cpu()
    miss = sum(_unmatched(got[f], int(gnum[f]), det[f], int(num[f])) for f in range(B))
    print("detections: oracle %s, product %s, product rows without a counterpart: %d" % (num.tolist(), gnum.tolist(), miss))
    assert miss <= 0.03 * int(gnum.sum()) and int((num - gnum).abs().max()) <= 3


def test_point_head_kernel_raw_outputs_and_labels():
    """pdm_point_head_forward alone: logits / residuals against torch, labels exact where the best score is not tied."""
    import ctypes
    from pdm_ssd_b200 import _lib
    model = _model(seed=7)
    head = model.dense_head
    B, Y, X, P = 2, 200, 176, 3000
    torch.manual_seed(2)
    x = torch.relu(torch.randn(B, 64, Y, X, device=DEV))
    hm = torch.rand(B, 3, Y, X, device=DEV)
    coords = torch.cat([torch.randint(0, B, (P, 1), device=DEV).float().sort(0)[0],
                        torch.rand(P, 3, device=DEV) * torch.tensor([70.4, 80.0, 4.0], device=DEV) + torch.tensor([0.0, -40.0, -3.0], device=DEV)], 1)
    coords[:5, 1] = torch.tensor([0.0, 70.39999, 70.4, -0.01, 35.2], device=DEV)         # map edges incl. out-of-range -> clamp
    pf = torch.randn(P, 128, device=DEV)
    pk = head._packed(torch.device(DEV))
    xs = SplitAct.from_nchw(x)
    score, boxes, best, label = head._point_head_fused(pk, coords, pf, xs, hm)
    with torch.no_grad():
        cx = torch.floor((coords[:, 1] - head.range[0]) / head.voxel[0]).long().clamp_(0, X - 1)
        cy = torch.floor((coords[:, 2] - head.range[1]) / head.voxel[1]).long().clamp_(0, Y - 1)
        b = coords[:, 0].long()
        fusedf = torch.cat([pf, xs.to_nchw()[b, :, cy, cx]], 1)
        ref_score = torch.sigmoid(head.cls_layers(fusedf)) * hm[b, :, cy, cx].sqrt()
        rb, rl = ref_score.max(1)
        ref_boxes = head.decode(head.box_layers(fusedf), coords[:, 1:4], label)
    assert _rel(score, ref_score) < 1e-4
    assert _rel(best, rb) < 1e-4
    top2 = ref_score.topk(2, dim=1)[0]
    clear = (top2[:, 0] - top2[:, 1]) > 1e-5 * top2[:, 0]
    assert torch.equal(label[clear], rl[clear])
    assert _rel(boxes, ref_boxes) < 1e-4
