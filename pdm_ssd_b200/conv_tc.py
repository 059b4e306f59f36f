"""Host side of csrc/conv_tc.cu: the tcgen05 implicit-GEMM convolution (3x3 / 1x1 + folded eval-mode
BatchNorm + activation) that replaces the cuDNN Conv2d/BatchNorm2d/ReLU stacks of the BEV "context
learning" block (pcdet/models/backbones_2d/base_bev_backbone.py:27-47) and of the heatmap branch
(pcdet/models/dense_heads/center_head.py:12-46) at inference.

Activations travel between the layers in the "split NHWC8" layout (see conv_tc.cu): a bf16 tensor
`(2, B, Y, C/8, X, 8)` whose plane 0 is bf16(v) and plane 1 is bf16(v - plane0); `SplitAct` carries
it together with its logical shape.  Weights are folded (BN), split the same way and packed once
per module in the order the kernel consumes them.
"""
import torch
import torch.nn as nn

from . import _lib

ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


class SplitAct:
    """A (B, C, Y, X) fp32 activation stored as two bf16 planes in NHWC8 order."""

    __slots__ = ("data", "B", "C", "Y", "X")

    def __init__(self, data, B, C, Y, X):
        self.data, self.B, self.C, self.Y, self.X = data, int(B), int(C), int(Y), int(X)

    @staticmethod
    def empty(B, C, Y, X, device):
        if C % 8:
            raise RuntimeError("split activations need a multiple of 8 channels, got %d" % C)
        return SplitAct(torch.empty((2, B, Y, C // 8, X, 8), dtype=torch.bfloat16, device=device), B, C, Y, X)

    @staticmethod
    def from_nchw(x):
        if not (x.is_cuda and x.dtype == torch.float32):
            raise RuntimeError("from_nchw needs a CUDA float32 tensor")
        x = x.contiguous()
        B, C, Y, X = x.shape
        out = SplitAct.empty(B, C, Y, X, x.device)
        with torch.cuda.device(x.device):
            rc = _lib.load().pdm_act_split_from_nchw(B, C, Y, X, x.data_ptr(), out.data.data_ptr(), _stream(x))
        _lib.check(rc, "pdm_act_split_from_nchw")
        return out

    def to_nchw(self):
        out = torch.empty((self.B, self.C, self.Y, self.X), dtype=torch.float32, device=self.data.device)
        with torch.cuda.device(out.device):
            rc = _lib.load().pdm_act_split_to_nchw(self.B, self.C, self.Y, self.X, self.data.data_ptr(), out.data_ptr(),
                                                   _stream(out))
        _lib.check(rc, "pdm_act_split_to_nchw")
        return out


def fold_conv_bn(conv: nn.Conv2d, bn=None):
    """(W', b') of conv followed by eval-mode BatchNorm (running statistics)."""
    w = conv.weight.detach().float()
    b = conv.bias.detach().float() if conv.bias is not None else w.new_zeros(w.shape[0])
    if bn is not None:
        scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
        w = w * scale[:, None, None, None]
        b = (b - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()
    return w, b


def npad_of(cout):
    return 16 if cout <= 16 else 32 if cout <= 32 else 64 if cout <= 64 else 128


def ncat_of(npad):
    """Layers of at most 64 (padded) output channels use N-concatenated weight blocks (conv_tc.cu, ConvTCParams.ncat);
    PDM_CONV_NCAT=0 switches that off on both sides (A/B measurements)."""
    import os
    return npad <= 64 and os.environ.get("PDM_CONV_NCAT", "1")[:1] != "0"


def pack_conv_weight(w, b):
    """w (Cout, Cin, k, k) fp32, b (Cout,) -> (bf16 blocks [Cin/32][k*k][hi|lo][4][npad][8] -- or, N-concatenated,
    [Cin/32][k*k][4][hi|lo][npad][8] --, bias fp32 (npad,))."""
    cout, cin, k, k2 = w.shape
    if k != k2 or k not in (1, 3) or cin % 32 or cout > 128:
        raise RuntimeError("unsupported convolution shape %s for the tcgen05 path" % (tuple(w.shape),))
    npad = npad_of(cout)
    wp = w.new_zeros((npad, cin, k, k))
    wp[:cout] = w
    hi = wp.to(torch.bfloat16)
    lo = (wp - hi.float()).to(torch.bfloat16)
    planes = torch.stack([hi, lo]).reshape(2, npad, cin // 32, 4, 8, k * k)       # plane, n, kc, c8, e, tap
    if ncat_of(npad):
        packed = planes.permute(2, 5, 3, 0, 1, 4).contiguous()                   # kc, tap, c8, plane, n, e
    else:
        packed = planes.permute(2, 5, 0, 3, 1, 4).contiguous()                   # kc, tap, plane, c8, n, e
    bp = b.new_zeros(npad)
    bp[:cout] = b
    return packed, bp.contiguous()


class PackedConv:
    """Folded + packed weights of one conv(+BN) layer on a device."""

    def __init__(self, conv, bn=None, act=ACT_RELU, device=None):
        w, b = fold_conv_bn(conv, bn)
        self.cout, self.cin, self.ksize = int(w.shape[0]), int(w.shape[1]), int(w.shape[2])
        packed, bias = pack_conv_weight(w, b)
        dev = device if device is not None else conv.weight.device
        self.packed, self.bias, self.act = packed.to(dev), bias.to(dev), int(act)

    def __call__(self, x: SplitAct, want_split=True, want_nchw=False):
        return conv_forward(x, self, want_split, want_nchw)


def conv_forward(x: SplitAct, layer: PackedConv, want_split=True, want_nchw=False):
    """-> (SplitAct or None, fp32 (B,Cout,Y,X) or None)."""
    if x.C != layer.cin:
        raise RuntimeError("conv expects %d input channels, got %d" % (layer.cin, x.C))
    dev = x.data.device
    out_s = SplitAct.empty(x.B, layer.cout, x.Y, x.X, dev) if want_split else None
    out_f = torch.empty((x.B, layer.cout, x.Y, x.X), dtype=torch.float32, device=dev) if want_nchw else None
    with torch.cuda.device(dev):
        rc = _lib.load().pdm_conv_tc_forward(
            x.B, x.Y, x.X, layer.cin, layer.cout, layer.ksize, x.data.data_ptr(), layer.packed.data_ptr(),
            layer.bias.data_ptr(), layer.act, out_s.data.data_ptr() if out_s is not None else None,
            out_f.data_ptr() if out_f is not None else None, _stream(x.data))
    _lib.check(rc, "pdm_conv_tc_forward")
    return out_s, out_f
