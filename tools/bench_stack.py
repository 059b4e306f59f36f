"""The stacked (pointnet2_stack) operator family on a PV-RCNN-shaped workload, ours next to the reference's own kernels
recompiled for sm_100 (oracle/_ref/pointnet2_stack_cuda_ref.so): per-op CUDA-event times + algorithmic bytes, one JSON line.

    python tools/bench_stack.py [--frames 4] [--points 16384] [--keypoints 2048]

Workload: `frames` KITTI-shaped frames of `points` points stacked back to back (ragged: frame i keeps points - 512 i of
them), `keypoints` FPS keypoints per frame, ball query r = 0.8 / nsample 16, grouping of 32-channel features, 3-NN +
interpolation of keypoint features back to the points, voxel query on a 0.4 m grid, vector pool 3 x 3 x 3.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from pdm_ssd_b200 import pointnet2_stack_cuda as ours, synthetic  # noqa: E402
import build_ref  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=4)
ap.add_argument("--points", type=int, default=16384)
ap.add_argument("--keypoints", type=int, default=2048)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda:0")
ref = build_ref.load_ref_stack()
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = float(peaks.get("hbm_gbs", 6530.0))

cnt = np.asarray([a.points - 512 * i for i in range(a.frames)], np.int32)
xyz_np = np.concatenate([synthetic.kitti_batch(1, a.points, first_frame=i)[0, :cnt[i], :3] for i in range(a.frames)], 0).astype(np.float32)
N, C = int(cnt.sum()), 32
mcnt = np.full((a.frames,), a.keypoints, np.int32)
M = int(mcnt.sum())
x, xc, mc = torch.from_numpy(xyz_np).to(dev), torch.from_numpy(cnt).to(dev), torch.from_numpy(mcnt).to(dev)
feat = torch.randn(N, C, device=dev)


def timed(fn, iters=a.iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out = {"workload": "pointnet2_stack family: %d ragged frames (%s points), %d keypoints each, C = %d" % (a.frames, cnt.tolist(), a.keypoints, C),
       "ops": {}}
for name, ext in (("ours", ours), ("reference_cuda", ref)):
    if ext is None:
        continue
    temp = torch.empty(N, device=dev)
    fidx = torch.zeros(M, dtype=torch.int32, device=dev)

    def fps():
        temp.fill_(1e10)
        ext.stack_farthest_point_sampling_wrapper(x, temp, xc, fidx, mc)
    t_fps = timed(fps, 3 if name != "ours" else a.iters)
    q = x[fidx.long()].contiguous()
    bidx = torch.zeros(M, 16, dtype=torch.int32, device=dev)

    def bq():
        bidx.zero_()
        ext.ball_query_wrapper(a.frames, M, 0.8, 16, q, mc, x, xc, bidx)
    t_bq = timed(bq)
    gi = bidx.clone()
    gi[gi[:, 0] == -1] = 0
    grouped = torch.empty(M, C, 16, device=dev)
    t_grp = timed(lambda: ext.group_points_wrapper(a.frames, M, C, 16, feat, xc, gi, mc, grouped))
    d2 = torch.zeros(N, 3, device=dev)
    nn = torch.zeros(N, 3, dtype=torch.int32, device=dev)
    t_nn = timed(lambda: ext.three_nn_wrapper(x, xc, q, mc, d2, nn))
    w = torch.rand(N, 3, device=dev)
    kf = torch.randn(M, C, device=dev)
    interp = torch.empty(N, C, device=dev)
    t_int = timed(lambda: ext.three_interpolate_wrapper(kf, nn, w, interp))
    # voxel query on a 0.4 m grid
    lo = x.min(0).values
    c = torch.floor((x - lo) / 0.4).int()
    X, Y, Z = (c.max(0).values + 1).tolist()
    bcol = torch.repeat_interleave(torch.arange(a.frames, device=dev, dtype=torch.int32), xc.long())
    pi = torch.full((a.frames, Z, Y, X), -1, dtype=torch.int32, device=dev)
    pi[bcol.long(), c[:, 2].long(), c[:, 1].long(), c[:, 0].long()] = torch.arange(N, device=dev, dtype=torch.int32)
    coords = torch.stack([bcol, c[:, 2], c[:, 1], c[:, 0]], 1).contiguous()[fidx.long()].contiguous()
    vq = torch.zeros(M, 16, dtype=torch.int32, device=dev)
    t_vq = timed(lambda: ext.voxel_query_wrapper(M, Z, Y, X, 16, 1.6, 2, 2, 2, q, x, coords, pi, vq))
    # vector pool 3 x 3 x 3, 8 channels per cell (C = 32 folds onto them)
    g, ceg = 27, 8
    nf = torch.zeros(M, g * ceg, device=dev)
    nl = torch.zeros(M, 3 * g, device=dev)
    pc = torch.zeros(M, g, dtype=torch.int32, device=dev)
    grp = torch.zeros(200 * M, 3, dtype=torch.int32, device=dev)

    def vp():
        nf.zero_(); nl.zero_(); pc.zero_()
        return ext.vector_pool_wrapper(x, xc, feat, q, mc, nf, nl, pc, grp, 3, 3, 3, 1.2, 1, 200 * M, -1, 0, 0)
    t_vp = timed(vp, 5)
    out["ops"][name] = {"stack_fps_ms": t_fps, "ball_query_ms": t_bq, "group_points_ms": t_grp, "three_nn_ms": t_nn,
                        "three_interpolate_ms": t_int, "voxel_query_ms": t_vq, "vector_pool_ms": t_vp}
if "ours" in out["ops"]:
    o = out["ops"]["ours"]
    grp_bytes = M * 16 * 4 + M * C * 16 * 4 * 2          # idx + gathered rows (read) + output (write)
    int_bytes = N * 3 * 8 + 3 * N * C * 4 + N * C * 4
    out["hbm"] = {"group_points_gbs": grp_bytes / (o["group_points_ms"] * 1e-3) / 1e9, "three_interpolate_gbs": int_bytes / (o["three_interpolate_ms"] * 1e-3) / 1e9,
                  "peak_gbs": hbm}
    if "reference_cuda" in out["ops"]:
        r = out["ops"]["reference_cuda"]
        out["speedup_vs_reference_cuda"] = {k: r[k] / o[k] for k in o}
print(json.dumps(out))
