import os, sys, torch, numpy as np
sys.path.insert(0, os.getcwd())
from pdm_ssd_b200 import pointnet2_batch_cuda as ours, synthetic
dev = torch.device("cuda:0")
B, N, M = 8, 163840, 16384
x = torch.from_numpy(synthetic.waymo_batch(B, N)).to(dev)[..., :3].contiguous()
temp = torch.full((B, N), 1e10, device=dev); idx = torch.zeros((B, M), dtype=torch.int32, device=dev)
for cl in (0, 16, 15, 14):
    if cl: os.environ["PDM_FPS_CLUSTER"] = str(cl)
    ts = []
    for _ in range(3):
        temp.fill_(1e10); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ours.farthest_point_sampling_wrapper(B, N, M, x, temp, idx); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print("cluster", cl, "ms", min(ts), flush=True)
