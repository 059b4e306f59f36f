"""Run one op of the SA chain a few times (for ncu captures / quick timing).

    python tools/run_op.py fps1|fps2|bq1|bq2|grp2 [--batch 16] [--iters 3]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pdm_ssd_b200 import pointnet2_batch_cuda as ours, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("op")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda:0")
B = a.batch
N, M, r, S = (16384, 4096, 0.8, 32) if a.op.endswith("1") else (4096, 1024, 1.6, 32)
xyz = torch.from_numpy(synthetic.kitti_batch(B, N)[..., :3].copy()).to(dev)
temp = torch.empty(B, N, device=dev)
idx = torch.empty(B, M, dtype=torch.int32, device=dev)
temp.fill_(1e10)
ours.farthest_point_sampling_wrapper(B, N, M, xyz, temp, idx)
new_xyz = torch.gather(xyz, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
bidx = torch.zeros(B, M, S, dtype=torch.int32, device=dev)
ours.ball_query_wrapper(B, N, M, r, S, new_xyz, xyz, bidx)
feat = torch.randn(B, 64, N, device=dev)
out = torch.empty(B, 64, M, S, device=dev)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    if a.op.startswith("fps"):
        temp.fill_(1e10)
        ours.farthest_point_sampling_wrapper(B, N, M, xyz, temp, idx)
    elif a.op.startswith("bq"):
        ours.ball_query_wrapper(B, N, M, r, S, new_xyz, xyz, bidx)
    else:
        ours.group_points_wrapper(B, 64, N, M, S, feat, bidx, out)
e1.record()
torch.cuda.synchronize()
print(a.op, "ms/iter", e0.elapsed_time(e1) / a.iters)
