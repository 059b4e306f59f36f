"""One large-frame FPS call (cluster kernel) for ncu captures: python tools/run_fps_large.py [n] [m] [batch]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pdm_ssd_b200 import pointnet2_batch_cuda as ours, synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 163840
m = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
B = int(sys.argv[3]) if len(sys.argv) > 3 else 8
dev = torch.device("cuda:0")
xyz = torch.from_numpy(synthetic.waymo_batch(B, n)[..., :3].copy()).to(dev)
temp = torch.full((B, n), 1e10, device=dev)
idx = torch.empty(B, m, dtype=torch.int32, device=dev)
for it in range(2):
    temp.fill_(1e10)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ours.farthest_point_sampling_wrapper(B, n, m, xyz, temp, idx)
    e1.record(); torch.cuda.synchronize()
    print("n=%d m=%d B=%d: %.3f ms (%.2f us/round)" % (n, m, B, e0.elapsed_time(e1), 1e3 * e0.elapsed_time(e1) / (m - 1)))
