"""Where does a tile of the persistent tcgen05 SA kernel spend its time?  Re-times SA2 (batch 16) with parts of the
kernel switched off through PDM_SA_TC3_DEBUG (results are wrong in those runs; timings only)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pdm_ssd_b200 import pointnet2_modules as M, pointnet2_utils as pu, synthetic
dev = "cuda:0"
B, N, Mc, C = 16, 4096, 1024, 32
torch.manual_seed(0)
sa = M.PointnetSAModuleMSG(npoint=Mc, radii=[1.6], nsamples=[32], mlps=[[C, 64, 64, 128]]).to(dev).eval()
xyz = torch.from_numpy(synthetic.kitti_batch(B, N)[..., :3].copy()).to(dev)
feat = torch.randn(B, C, N, device=dev)
with torch.no_grad():
    new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), pu.farthest_point_sample(xyz, Mc)).transpose(1, 2).contiguous()
    def timed(fn, it=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(it): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / it
    tq = timed(lambda: pu.ball_query(1.6, 32, xyz, new_xyz))
    for d in (0, 1, 2, 4, 8, 6, 7, 14, 15):
        os.environ["PDM_SA_TC3_DEBUG"] = str(d)
        print("debug=%2d  %.3f ms" % (d, timed(lambda: sa(xyz, feat, new_xyz)[1]) - tq))
