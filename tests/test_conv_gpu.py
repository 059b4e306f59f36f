"""tcgen05 implicit-GEMM convolution (csrc/conv_tc.cu) against torch fp32 convolutions (TF32 off).

Reference semantics: Conv2d(3x3, padding 1) + eval-mode BatchNorm2d + ReLU as stacked in
pcdet/models/backbones_2d/base_bev_backbone.py:27-47 and pcdet/models/dense_heads/center_head.py:12-46.
Tolerance: 1e-3 relative (BASELINE.json north_star); the bf16x3 split measures ~1e-5.
"""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def test_split_layout_round_trip():
    from pdm_ssd_b200.conv_tc import SplitAct
    x = torch.randn(2, 64, 19, 37, device="cuda") * 3
    s = SplitAct.from_nchw(x)
    assert s.data.shape == (2, 2, 19, 8, 37, 8)
    y = s.to_nchw()
    assert _rel(y, x) < 2e-5
    # plane 0 is exactly bf16(x) in NHWC8 order
    hi = s.data[0].permute(0, 2, 4, 1, 3).reshape(2, 64, 19, 37)
    assert torch.equal(hi, x.to(torch.bfloat16))


@pytest.mark.parametrize("B,Cin,Cout,Y,X,k,act", [
    (1, 32, 16, 16, 16, 3, 1),       # one unit, one chunk
    (2, 64, 64, 40, 48, 3, 1),       # several units, two accumulator sets
    (1, 128, 128, 33, 50, 3, 1),     # ragged edges, full-width N, one accumulator set
    (3, 64, 3, 24, 40, 3, 2),        # heatmap-style output: 3 channels, sigmoid, fp32 NCHW only
    (2, 128, 64, 200, 176, 3, 1),    # KITTI BEV map
    (1, 64, 32, 20, 24, 1, 0),       # 1x1, no activation
])
def test_conv_matches_torch_fp32(B, Cin, Cout, Y, X, k, act):
    from pdm_ssd_b200.conv_tc import PackedConv, SplitAct
    torch.manual_seed(Cin * 1000 + Cout + Y)
    conv = nn.Conv2d(Cin, Cout, k, padding=k // 2, bias=(act != 1)).cuda()
    bn = nn.BatchNorm2d(Cout, eps=1e-3).cuda().eval() if act == 1 else None
    if bn is not None:
        with torch.no_grad():
            bn.running_mean.normal_(0, 0.2)
            bn.running_var.uniform_(0.5, 1.5)
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.normal_(0, 0.2)
    x = torch.randn(B, Cin, Y, X, device="cuda")
    with torch.no_grad():
        ref = conv(x)
        if bn is not None:
            ref = bn(ref)
        ref = F.relu(ref) if act == 1 else torch.sigmoid(ref) if act == 2 else ref
    layer = PackedConv(conv, bn, act=act)
    want_split = Cout % 8 == 0
    out_s, out_f = layer(SplitAct.from_nchw(x), want_split=want_split, want_nchw=True)
    torch.cuda.synchronize()
    assert _rel(out_f, ref) < 1e-3, _rel(out_f, ref)
    assert _rel(out_f, ref) < 1e-4            # what the split actually delivers
    if want_split:
        assert _rel(out_s.to_nchw(), ref) < 1e-4


def test_conv_stack_matches_torch_fp32():
    """Five layers chained in the split layout (the detector's dense part): error does not build up."""
    from pdm_ssd_b200.conv_tc import PackedConv, SplitAct
    torch.manual_seed(5)
    chans = [128, 128, 128, 64, 64, 3]
    mods = []
    for i in range(5):
        mods.append((nn.Conv2d(chans[i], chans[i + 1], 3, padding=1, bias=True).cuda(),
                     nn.BatchNorm2d(chans[i + 1], eps=1e-3).cuda().eval() if i < 4 else None))
    x = torch.randn(2, 128, 56, 72, device="cuda")
    with torch.no_grad():
        ref = x
        for conv, bn in mods:
            ref = conv(ref)
            ref = F.relu(bn(ref)) if bn is not None else torch.sigmoid(ref)
    cur = SplitAct.from_nchw(x)
    for i, (conv, bn) in enumerate(mods):
        last = i == 4
        cur, out_f = PackedConv(conv, bn, act=2 if last else 1)(cur, want_split=not last, want_nchw=last)
    assert _rel(out_f, ref) < 1e-3, _rel(out_f, ref)


def test_conv_rejects_bad_arguments():
    from pdm_ssd_b200 import _lib
    from pdm_ssd_b200.conv_tc import SplitAct
    lib = _lib.load()
    x = SplitAct.from_nchw(torch.randn(1, 32, 16, 16, device="cuda"))
    p = x.data.data_ptr()
    assert lib.pdm_conv_tc_forward(1, 16, 16, 24, 16, 3, p, p, p, 1, p, None, None) != 0     # cin not a multiple of 32
    assert lib.pdm_conv_tc_forward(1, 16, 16, 32, 16, 5, p, p, p, 1, p, None, None) != 0     # 5x5
    assert b"conv_tc_forward" in lib.pdm_last_error()
