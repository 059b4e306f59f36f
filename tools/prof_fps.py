"""Per-phase cycle breakdown of the FPS bucket kernel (debug entry pdm_debug_fps_profile)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pdm_ssd_b200 import _lib, synthetic
lib = _lib.load()
fn = lib.pdm_debug_fps_profile
fn.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 5
dev = torch.device("cuda:0")
B = 16
for N, M in ((16384, 4096), (4096, 1024)):
    xyz = torch.from_numpy(synthetic.kitti_batch(B, N)[..., :3].copy()).to(dev)
    temp = torch.full((B, N), 1e10, device=dev)
    idx = torch.empty(B, M, dtype=torch.int32, device=dev)
    prof = torch.zeros(B, 16, 8, dtype=torch.int64, device=dev)
    rc = fn(B, N, M, xyz.data_ptr(), temp.data_ptr(), idx.data_ptr(), prof.data_ptr(), None)
    torch.cuda.synchronize()
    assert rc == 0, lib.pdm_last_error()
    pr = prof.cpu().numpy().astype(np.float64)
    rounds = M - 1
    print("N=%d M=%d   per-round averages over %d frames x 16 warps (cycles)" % (N, M, B))
    names = ["A bound+ballot", "B bucket updates", "C warp best", "publish+barrier", "D block argmax"]
    for q, nm in enumerate(names):
        print("  %-18s mean %7.1f   min-warp %7.1f  max-warp %7.1f" % (nm, pr[..., q].mean() / rounds, pr[..., q].min() / rounds, pr[..., q].max() / rounds))
    print("  bucket updates/round/warp %.3f  (sum over warps %.2f)   C runs/round/warp %.3f" % (pr[..., 5].mean() / rounds, pr[..., 5].sum(1).mean() / rounds, pr[..., 6].mean() / rounds))
    print("  cycles per B update %.1f   cycles per C run %.1f" % (pr[..., 1].sum() / max(pr[..., 5].sum(), 1), pr[..., 2].sum() / max(pr[..., 6].sum(), 1)))
    print("  total loop cycles/round %.1f" % (pr[..., 7].mean() / rounds))
