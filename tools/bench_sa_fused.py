"""Fused SA scale, every kernel variant next to the unfused torch path (time after the ball query, batch 16):
   rows     thread-per-row kernel (narrow MLPs, csrc/sa_rows.cu)
   tc3      persistent warp-specialised tcgen05 kernel, bf16 hi/lo operands, two tiles in flight (csrc/sa_tc.cu)
   tc_r1    round 1's tcgen05 kernel (tf32 hi/lo, one tile per CTA)
   cuda     CUDA-core kernel (sa_fused_kernel)
"""
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
VARIANTS = (("rows", dict(PDM_SA_ROWS="1", PDM_SA_TC3="1", PDM_SA_TC="1")), ("tc3", dict(PDM_SA_ROWS="0", PDM_SA_TC3="1", PDM_SA_TC="1")),
            ("tc_r1", dict(PDM_SA_ROWS="0", PDM_SA_TC3="0", PDM_SA_TC="1")), ("cuda", dict(PDM_SA_ROWS="0", PDM_SA_TC3="0", PDM_SA_TC="0")))
for name, N, Mc, C, mlp in (("SA1", 16384, 4096, 1, [1, 16, 16, 32]), ("SA2 (model: 32 feature channels)", 4096, 1024, 32, [32, 64, 64, 128]),
                            ("SA2 (configs[1]: 64 feature channels)", 4096, 1024, 64, [64, 64, 64, 128])):
    torch.manual_seed(0)
    sa = M.PointnetSAModuleMSG(npoint=Mc, radii=[0.8 if N > 5000 else 1.6], nsamples=[32], mlps=[list(mlp)]).to(dev).eval()
    xyz = torch.from_numpy(synthetic.kitti_batch(B, N)[..., :3].copy()).to(dev)
    feat = torch.randn(B, C, N, device=dev)
    feat_pm = feat.clone()
    feat_pm._pdm_point_major = feat.permute(0, 2, 1).contiguous()
    with torch.no_grad():
        new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), pu.farthest_point_sample(xyz, Mc)).transpose(1, 2).contiguous()
        tq = timed(lambda: pu.ball_query(sa.groupers[0].radius, 32, xyz, new_xyz))
        M.ENABLE_FUSED_SA = False; t_un = timed(lambda: sa(xyz, feat, new_xyz)[1], 5); ref = sa(xyz, feat, new_xyz)[1]; M.ENABLE_FUSED_SA = True
        line = "%s  ball query %.3f ms, unfused torch %.3f ms |" % (name, tq, t_un - tq)
        for vn, env in VARIANTS:
            os.environ.update(env)
            for label, f in ((vn, feat), (vn + "+pm", feat_pm)):
                if label.endswith("+pm") and vn != "tc3":
                    continue
                t = timed(lambda: sa(xyz, f, new_xyz)[1])
                out = sa(xyz, f, new_xyz)[1]
                err = float((out - ref).abs().max() / ref.abs().max())
                pm = out._pdm_point_major
                pm_ok = bool(torch.equal(pm.permute(0, 2, 1), out))
                line += " %s %.3f ms (err %.1e%s)" % (label, t - tq, err, "" if pm_ok else ", POINT-MAJOR COPY DIFFERS")
    print(line)
