"""CPU tests of the hybrid-head oracle (oracle/pdm_head_oracle.py, SPEC_HEAD.md) and of the product's torch path:
the box decode is pinned to the reference's PointResidualCoder.decode_torch (box_coder_utils.py:189-222) through a
golden vector produced by the reference file itself (tests/golden/make_golden_head.py) and, when the reference
Python files are present (here: /root/reference or oracle/_ref/py), through a live call."""
import os

import numpy as np
import pytest
import torch

import pdm_head_oracle as ho
from pdm_ssd_b200.detector import HybridHead, BEVContext, PDMSSD, default_cfg


def _golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "head_decode.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def test_oracle_decode_matches_reference_golden(golden_dir):
    g = _golden(golden_dir)
    out = ho.decode(g["enc"], g["points"], g["pred_classes"] - 1)
    assert torch.equal(out, g["boxes"])


def test_product_decode_matches_reference_golden(golden_dir):
    g = _golden(golden_dir)
    cfg = default_cfg()
    head = HybridHead(cfg.DENSE_HEAD, 128, 128, 3, cfg.POINT_CLOUD_RANGE, cfg.VOXEL_SIZE)
    out = head.decode(g["enc"], g["points"], g["pred_classes"] - 1)
    assert torch.equal(out, g["boxes"])


def test_decode_against_live_reference_file(golden_dir):
    import build_ref
    ns = build_ref.load_reference_tree("refpy_cpu", None, None)
    if ns is None:
        pytest.skip("reference python files not vendored (oracle/_ref/py)")
    coder = object.__new__(ns.box_coder_utils.PointResidualCoder)       # __init__ calls .cuda()
    coder.code_size, coder.use_mean_size = 8, True
    coder.mean_size = torch.tensor(ho.KITTI_MEAN_SIZE, dtype=torch.float32)
    g = torch.Generator().manual_seed(3)
    enc, pts, cls = torch.randn(777, 8, generator=g), torch.randn(777, 3, generator=g) * 20, torch.randint(1, 4, (777,), generator=g)
    assert torch.equal(coder.decode_torch(enc, pts, cls), ho.decode(enc, pts, cls - 1))


def test_oracle_matches_product_torch_path_on_cpu():
    """SPEC_HEAD steps 1-6: the oracle (functional, on the state_dict) and the product's torch modules agree on CPU
    to fp32 rounding.  (The GPU test holds the fused kernels to the same oracle.)"""
    torch.manual_seed(0)
    cfg = default_cfg()
    ctx = BEVContext(cfg.BACKBONE_2D, 128).eval()
    head = HybridHead(cfg.DENSE_HEAD, 128, 128, 3, cfg.POINT_CLOUD_RANGE, cfg.VOXEL_SIZE, post_cfg=None).eval()   # NMS is GPU-only
    for m in list(ctx.modules()) + list(head.modules()):
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.6, 1.4)
    B, Y, X, P = 1, 24, 32, 64
    sf = torch.randn(B, 128, Y, X) * (torch.rand(B, 1, Y, X) < 0.3)
    coords = torch.cat([torch.zeros(P, 1), torch.rand(P, 3) * torch.tensor([X * 0.4, Y * 0.4, 4.0]) + torch.tensor([0.0, -40.0, -3.0])], 1)
    pf = torch.randn(P, 128)
    with torch.no_grad():
        bd = ctx({"spatial_features": sf, "batch_size": B})
        bd.update(point_coords=coords, point_features=pf)
        out = head(bd)
    sf2d = ho.bev_context(ctx.state_dict(), sf, 2)
    torch.testing.assert_close(sf2d, out["spatial_features_2d"], rtol=1e-5, atol=1e-5)
    res = ho.point_head(head.state_dict(), sf2d, coords, pf, cfg.POINT_CLOUD_RANGE, cfg.VOXEL_SIZE)
    torch.testing.assert_close(res["heatmap"], out["heatmap"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(res["scores"], out["batch_cls_preds"], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(res["boxes"], out["batch_box_preds"], rtol=1e-4, atol=1e-4)
    # step 7 on the C oracle's NMS: well-formed, sorted, thresholded (the GPU test compares it with the product's)
    det, num = ho.post_process(res["boxes"], res["best"], res["label"], B, 0.1, 0.1, 4096, 100)
    n = int(num[0])
    assert (det[0, :n, 7] >= 0.1).all() and (det[0, n:] == 0).all() and (det[0, :max(n - 1, 0), 7] >= det[0, 1:max(n, 1), 7]).all()


def test_state_dict_keys_follow_spec():
    model = PDMSSD(default_cfg())
    keys = set(model.state_dict().keys())
    for k in ("dense_head.shared_conv.0.weight", "dense_head.hm.3.bias", "dense_head.cls_layers.0.weight",
              "dense_head.cls_layers.1.running_var", "dense_head.box_layers.3.bias", "backbone_2d.blocks.0.weight",
              "backbone_2d.blocks.4.running_mean", "map_to_bev_module.coef.weight", "backbone_3d.SA_modules.0.mlps.0.0.weight"):
        assert k in keys, k
    assert float(model.dense_head.hm[3].bias[0]) == pytest.approx(-2.19)          # center_head.py:38-39
