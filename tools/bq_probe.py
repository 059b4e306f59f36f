"""Ball-query probe: KITTI SA1 / SA2 and Waymo SA1 shapes, eager launches (run under ncu for per-kernel times) + event timing.
    python tools/bq_probe.py [kitti|waymo]"""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pdm_ssd_b200 import pointnet2_batch_cuda as ours, synthetic
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "kitti"
cases = []
if which == "kitti":
    fr = torch.from_numpy(synthetic.kitti_batch(16, 16384)).to(dev)[..., :3].contiguous()
    cases = [("kitti_sa1", fr, 4096, 0.8, 32), ("kitti_sa2", None, 1024, 1.6, 32)]
else:
    fr = torch.from_numpy(synthetic.waymo_batch(8, 163840)).to(dev)[..., :3].contiguous()
    cases = [("waymo_sa1", fr, 16384, 0.8, 32), ("waymo_sa2", None, 4096, 1.6, 32)]
res = {}
cur = fr
for name, x, m, r, ns in cases:
    x = cur
    B, N, _ = x.shape
    temp = torch.full((B, N), 1e10, device=dev)
    idx = torch.zeros((B, m), dtype=torch.int32, device=dev)
    ours.farthest_point_sampling_wrapper(B, N, m, x, temp, idx)
    q = torch.gather(x, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    out = torch.zeros((B, m, ns), dtype=torch.int32, device=dev)
    for _ in range(3):
        ours.ball_query_wrapper(B, N, m, r, ns, q, x, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ours.ball_query_wrapper(B, N, m, r, ns, q, x, out)
    e1.record(); torch.cuda.synchronize()
    cnt = (out != out[..., :1]).sum(-1) + 1
    res[name] = {"ms": e0.elapsed_time(e1) / 10, "full_rows_frac": float((cnt >= ns).float().mean())}
    cur = q
print(json.dumps(res))
