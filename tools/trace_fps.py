"""Critical-path breakdown of the multi-sample FPS kernel from per-round clock64 stamps."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pdm_ssd_b200 import _lib, synthetic
lib = _lib.load()
fn = lib.pdm_debug_fps_trace
fn.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 6
dev = torch.device("cuda:0")
B, N, M = 16, 16384, 4096
xyz = torch.from_numpy(synthetic.kitti_batch(B, N)[..., :3].copy()).to(dev)
temp = torch.full((B, N), 1e10, device=dev)
idx = torch.empty(B, M, dtype=torch.int32, device=dev)
stats = torch.zeros(B, dtype=torch.int32, device=dev)
trace = torch.zeros(M, 16, 8, dtype=torch.int64, device=dev)
rc = fn(B, N, M, xyz.data_ptr(), temp.data_ptr(), idx.data_ptr(), stats.data_ptr(), trace.data_ptr(), None)
torch.cuda.synchronize()
assert rc == 0, lib.pdm_last_error()
R = int(stats[0].item())
t = trace[:R].cpu().numpy().astype(np.float64)   # (R,16,8)
K = (t[:, 0, 7].astype(np.int64) & 0xff)
nupd = (t[:, :, 7].astype(np.int64) >> 8)
print("rounds", R, "mean K %.2f" % K.mean(), "updates/round sum %.1f max-per-warp %.2f" % (nupd.sum(1).mean(), nupd.max(1).mean()))
start = t[:, :, 0]; aA = t[:, :, 1]; aB = t[:, :, 2]; aC = t[:, :, 3]; aBar = t[:, :, 4]; aP1 = t[:, :, 5]; aL = t[:, :, 6]
print("per-warp mean durations (cycles): A %.0f  B %.0f  C %.0f  wait@barrier %.0f  pick1 %.0f  loop %.0f  total %.0f" % (
    (aA - start).mean(), (aB - aA).mean(), (aC - aB).mean(), (aBar - aC).mean(), (aP1 - aBar).mean(), (aL - aP1).mean(), (aL - start).mean()))
crit = aC.argmax(1)  # last warp to reach the barrier
ar = np.arange(R)
print("critical warp (last at barrier): A %.0f  B %.0f  C %.0f  | barrier release->pick1 done %.0f | loop %.0f" % (
    (aA - start)[ar, crit].mean(), (aB - aA)[ar, crit].mean(), (aC - aB)[ar, crit].mean(),
    (aP1.max(1) - aC.max(1)).mean(), (aL.max(1) - aP1.max(1)).mean()))
print("round span (max end - min start) %.0f ; B per update on critical warp %.0f ; loop cycles per extra pick %.0f" % (
    (aL.max(1) - start.min(1)).mean(), ((aB - aA)[ar, crit].sum() / max(nupd[ar, crit].sum(), 1)),
    (aL - aP1).mean(1).sum() / max((K - 1).sum(), 1)))
print("A per sample %.0f" % ((aA - start).mean(1).sum() / np.concatenate([[1], K[:-1]]).sum()))
