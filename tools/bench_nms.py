"""Rotated NMS for a batch of frames: our two-launch device-resident path vs the reference's
per-frame nms_gpu (cudaMalloc + blocking D2H + host greedy loop, iou3d_nms.cpp:137-183).

    python tools/bench_nms.py [--frames 16] [--boxes 1024] [--out gpurun_out/nms_r1.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from pdm_ssd_b200 import iou3d_nms_cuda as ours, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=16)
ap.add_argument("--boxes", type=int, default=1024)
ap.add_argument("--clusters", type=int, default=30)
ap.add_argument("--thresh", type=float, default=0.1)
ap.add_argument("--iters", type=int, default=50)
ap.add_argument("--out", default="")
a = ap.parse_args()
dev = torch.device("cuda:0")
boxes = torch.from_numpy(np.stack([synthetic.random_boxes(a.boxes, seed=100 + f, clusters=a.clusters) for f in range(a.frames)])).to(dev)
keep = torch.empty((a.frames, a.boxes), dtype=torch.int32, device=dev)
num = torch.empty((a.frames,), dtype=torch.int32, device=dev)


def ev_time(fn, iters):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


res = {"frames": a.frames, "boxes_per_frame": a.boxes, "thresh": a.thresh,
       "ours_batched_ms": ev_time(lambda: ours.nms_bev_batched(boxes, None, a.thresh, keep, num), a.iters),
       "kept_per_frame": float(num.float().mean().item())}
import build_ref  # noqa: E402
ref = build_ref.load_ref_nms()
if ref is not None:
    hk = torch.zeros(a.boxes, dtype=torch.int64)

    def ref_all():
        return [ref.nms_gpu(boxes[f], hk, a.thresh) for f in range(a.frames)]
    for _ in range(3):
        ref_all()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        nk = ref_all()
    torch.cuda.synchronize()
    res["reference_per_frame_loop_ms"] = (time.perf_counter() - t0) / 10 * 1e3   # wall clock: the loop blocks on the host
    res["same_counts"] = bool(np.array_equal(np.array(nk), num.cpu().numpy()))
    res["speedup"] = res["reference_per_frame_loop_ms"] / res["ours_batched_ms"]
print(json.dumps(res))
if a.out:
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
