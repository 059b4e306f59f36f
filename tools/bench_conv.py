"""Time the tcgen05 convolution layers of the detector's dense part (KITTI BEV map 200 x 176, batch 16) next to
torch/cuDNN fp32 (TF32 off).  CUDA events on the launch stream; inputs larger than L2 (288 MB per tensor).

    python tools/bench_conv.py [--batch 16] [--iters 20]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pdm_ssd_b200.conv_tc import PackedConv, SplitAct  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--no-torch", action="store_true")
a = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda")
B, Y, X = a.batch, 200, 176
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
peak_tf = float(peaks.get("bf16_tflops", 1590.0))


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


rows = []
for name, cin, cout, act in (("bev_ctx 128->128", 128, 128, 1), ("shared 128->64", 128, 64, 1), ("hm1 64->64", 64, 64, 1), ("hm2 64->3", 64, 3, 2)):
    torch.manual_seed(0)
    conv = nn.Conv2d(cin, cout, 3, padding=1, bias=True).to(dev)
    bn = nn.BatchNorm2d(cout, eps=1e-3).to(dev).eval() if act == 1 else None
    x = torch.randn(B, cin, Y, X, device=dev)
    xs = SplitAct.from_nchw(x)
    layer = PackedConv(conv, bn, act=act)
    split_out = cout % 8 == 0
    ms = timeit(lambda: layer(xs, want_split=split_out, want_nchw=not split_out), a.iters)
    flops = 2.0 * B * Y * X * cout * cin * 9
    row = {"layer": name, "ms": ms, "algorithmic_tflops": flops / ms / 1e9, "executed_tflops": 3 * flops / ms / 1e9,
           "frac_of_measured_bf16_peak_executed": 3 * flops / ms / 1e9 / peak_tf}
    if not a.no_torch:
        with torch.no_grad():
            seq = nn.Sequential(conv, bn, nn.ReLU()) if bn is not None else nn.Sequential(conv, nn.Sigmoid())
            row["torch_cudnn_fp32_ms"] = timeit(lambda: seq(x), max(3, a.iters // 4))
    rows.append(row)
    print(json.dumps(row), flush=True)
    del x, xs
print(json.dumps({"total_ms_model_dense_part": 2 * rows[0]["ms"] + rows[1]["ms"] + rows[2]["ms"] + rows[3]["ms"]}))
