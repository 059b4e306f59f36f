"""three_nn / three_interpolate (rows a7-a9 of SURVEY 8a: the feature-propagation ops) next to the reference's own kernels
recompiled for sm_100 (oracle/_ref): B = 16, 16384 unknown points, 4096 known points, 64 channels.  One JSON line."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from pdm_ssd_b200 import pointnet2_batch_cuda as ours, synthetic
import build_ref
ref = build_ref.load_ref()
dev = torch.device("cuda:0")
B, n, m, C = 16, 16384, 4096, 64
xyz = torch.from_numpy(synthetic.kitti_batch(B, n)[..., :3].copy()).to(dev)
known = xyz[:, ::4].contiguous()
feat = torch.randn(B, C, m, device=dev)


def timed(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


out = {"workload": "three_nn + three_interpolate, B=%d, n=%d unknown, m=%d known, C=%d" % (B, n, m, C), "ops": {}}
res = {}
for name, ext in (("ours", ours), ("reference_cuda", ref)):
    if ext is None: continue
    d2 = torch.empty(B, n, 3, device=dev); idx = torch.empty(B, n, 3, dtype=torch.int32, device=dev)
    t_nn = timed(lambda: ext.three_nn_wrapper(B, n, m, xyz, known, d2, idx))
    w = 1.0 / (d2.sqrt() + 1e-8); w = (w / w.sum(2, keepdim=True)).contiguous()
    o = torch.empty(B, C, n, device=dev)
    t_int = timed(lambda: ext.three_interpolate_wrapper(B, C, m, n, feat, idx, w, o))
    out["ops"][name] = {"three_nn_ms": t_nn, "three_interpolate_ms": t_int}
    res[name] = (idx.clone(), d2.clone(), o.clone())
if len(res) == 2:
    out["bit_identical"] = all(bool(torch.equal(a, b)) for a, b in zip(res["ours"], res["reference_cuda"]))
    out["speedup_vs_reference_cuda"] = {k: out["ops"]["reference_cuda"][k] / out["ops"]["ours"][k] for k in out["ops"]["ours"]}
ib = B * (n * 3 * 8 + C * m * 4 + C * n * 4)
out["three_interpolate_gbs"] = ib / (out["ops"]["ours"]["three_interpolate_ms"] * 1e-3) / 1e9
print(json.dumps(out))
