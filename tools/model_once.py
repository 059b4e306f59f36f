"""A few eager forwards of the full detector (batch 16) -- the command the ncu launch list is taken from.

    python tools/model_once.py [--iters 3]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
model = bench.build_model(dev)
pts = torch.from_numpy(bench.host_points(0, 0)).to(dev)
marker = torch.zeros(1, dtype=torch.float64, device=dev)      # FillFunctor<double>: marks the start of a forward in launch lists
with torch.no_grad():
    for i in range(a.iters):
        marker.fill_(float(i))
        torch.cuda.nvtx.range_push("forward%d" % i)
        det = model({"batch_size": bench.BATCH, "points": pts})["detections"]
        torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
print("ok", tuple(det.shape), float(det[..., 7].sum()))
