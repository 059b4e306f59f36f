"""Fused SA scale: tcgen05 tensor-core kernel vs CUDA-core kernel vs the unfused torch path."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pdm_ssd_b200 import pointnet2_modules as M, pointnet2_utils as pu, synthetic
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
dev = "cuda:0"
B = 16
def timed(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
for name, N, Mc, C, mlp in (("SA1", 16384, 4096, 1, [1, 16, 16, 32]), ("SA2", 4096, 1024, 64, [64, 64, 64, 128])):
    torch.manual_seed(0)
    sa = M.PointnetSAModuleMSG(npoint=Mc, radii=[0.8 if N > 5000 else 1.6], nsamples=[32], mlps=[list(mlp)]).to(dev).eval()
    xyz = torch.from_numpy(synthetic.kitti_batch(B, N)[..., :3].copy()).to(dev)
    feat = torch.randn(B, C, N, device=dev)
    with torch.no_grad():
        new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), pu.farthest_point_sample(xyz, Mc)).transpose(1, 2).contiguous()
        def run():
            return sa(xyz, feat, new_xyz)[1]
        # ball query alone (common to all variants)
        tq = timed(lambda: pu.ball_query(sa.groupers[0].radius, 32, xyz, new_xyz))
        os.environ["PDM_SA_TC"] = "1"; t_tc = timed(run); a = run()
        os.environ["PDM_SA_TC"] = "0"; t_cc = timed(run); b = run()
        M.ENABLE_FUSED_SA = False; t_un = timed(run, 5); c = run(); M.ENABLE_FUSED_SA = True
    rel = lambda u, v: float((u - v).abs().max() / v.abs().max())
    print("%s  ball query %.3f ms | after it: tcgen05 %.3f ms, CUDA cores %.3f ms, unfused torch %.3f ms | rel err tc %.1e cc %.1e"
          % (name, tq, t_tc - tq, t_cc - tq, t_un - tq, rel(a, c), rel(b, c)))
