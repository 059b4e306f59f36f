"""Sweep the FPS bucket kernel's warps-per-CTA knob (PDM_FPS_NW) and check each variant
bit-for-bit against the reference extension."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from pdm_ssd_b200 import pointnet2_batch_cuda as ours, synthetic  # noqa: E402
import build_ref  # noqa: E402

ref = build_ref.load_ref()
dev = torch.device("cuda:0")
B = 16
for N, M in ((16384, 4096), (4096, 1024), (8192, 2048), (2048, 512), (1024, 256)):
    xyz = torch.from_numpy(synthetic.kitti_batch(B, N)[..., :3].copy()).to(dev)
    temp = torch.empty(B, N, device=dev)
    idx = torch.empty(B, M, dtype=torch.int32, device=dev)
    want = torch.empty(B, M, dtype=torch.int32, device=dev)
    if ref is not None:
        temp.fill_(1e10)
        ref.farthest_point_sampling_wrapper(B, N, M, xyz, temp, want)
    for nw in (4, 8, 16, 32):
        os.environ["PDM_FPS_NW"] = str(nw)
        try:
            ts = []
            for it in range(6):
                temp.fill_(1e10)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ours.farthest_point_sampling_wrapper(B, N, M, xyz, temp, idx)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ok = bool(torch.equal(idx, want)) if ref is not None else None
            print("N=%5d M=%4d NW=%2d  %.4f ms  (%.0f ns/round)  exact=%s" % (N, M, nw, min(ts[1:]), min(ts[1:]) * 1e6 / (M - 1), ok))
        except Exception as ex:  # unsupported combination
            print("N=%5d M=%4d NW=%2d  -- %s" % (N, M, nw, str(ex)[:80]))
