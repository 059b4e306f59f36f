"""Sweep the FPS bucket kernel's knobs (PDM_FPS_NW warps per CTA, PDM_FPS_KMAX samples per
round), check each variant bit-for-bit against the reference extension and report the number of
barrier rounds per frame (debug entry pdm_debug_fps_rounds)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from pdm_ssd_b200 import _lib, pointnet2_batch_cuda as ours, synthetic  # noqa: E402
import build_ref  # noqa: E402

ref = build_ref.load_ref()
lib = _lib.load()
dbg = lib.pdm_debug_fps_rounds
dbg.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 5
dev = torch.device("cuda:0")
B = 16
combos = [(16, 8), (16, 1), (8, 8), (16, 4), (16, 12)]
for N, M in ((16384, 4096), (4096, 1024), (8192, 2048), (1024, 256)):
    xyz = torch.from_numpy(synthetic.kitti_batch(B, N)[..., :3].copy()).to(dev)
    temp = torch.empty(B, N, device=dev)
    idx = torch.empty(B, M, dtype=torch.int32, device=dev)
    want = torch.empty(B, M, dtype=torch.int32, device=dev)
    wtemp = torch.full((B, N), 1e10, device=dev)
    stats = torch.zeros(B, dtype=torch.int32, device=dev)
    if ref is not None:
        ref.farthest_point_sampling_wrapper(B, N, M, xyz, wtemp, want)
    for nw, km in combos:
        os.environ["PDM_FPS_NW"], os.environ["PDM_FPS_KMAX"] = str(nw), str(km)
        try:
            ts = []
            for it in range(5):
                temp.fill_(1e10)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ours.farthest_point_sampling_wrapper(B, N, M, xyz, temp, idx)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ok = bool(torch.equal(idx, want) and torch.equal(temp, wtemp)) if ref is not None else None
            temp.fill_(1e10)
            rc = dbg(B, N, M, xyz.data_ptr(), temp.data_ptr(), idx.data_ptr(), stats.data_ptr(), None)
            torch.cuda.synchronize()
            r = stats.float().mean().item()
            print("N=%5d M=%4d NW=%2d KMAX=%2d  %.4f ms  rounds/frame %.0f (%.2f samples/round, %.0f ns/round)  exact=%s"
                  % (N, M, nw, km, min(ts[1:]), r, (M - 1) / max(r, 1), min(ts[1:]) * 1e6 / max(r, 1), ok))
        except Exception as ex:  # unsupported combination
            print("N=%5d M=%4d NW=%2d KMAX=%2d  -- %s" % (N, M, nw, km, str(ex)[:70]))
