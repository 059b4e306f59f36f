"""Per-op timing of the SA op chain (BASELINE config 2) -- ours vs the reference extension.

    python tools/bench_ops.py [--batch 16] [--iters 20] [--out gpurun_out/ops.json]

CUDA-event timing on the current stream, warm-up first; inputs are KITTI-shaped synthetic
frames.  Not the contract bench (that is bench.py); this is the per-kernel breakdown.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from pdm_ssd_b200 import pointnet2_batch_cuda as ours, synthetic  # noqa: E402


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
    return float(np.median(ts)), float(np.min(ts))


def chain(be, B, dev, iters):
    fr = torch.from_numpy(synthetic.kitti_batch(B)).to(dev)
    xyz = fr[..., :3].contiguous()
    feat1 = fr[..., 3:].transpose(1, 2).contiguous()          # (B,1,N)
    feat2 = torch.randn(B, 64, 4096, device=dev)
    res = {}
    layers = [("sa1", 16384, 4096, 0.8, 32, feat1), ("sa2", 4096, 1024, 1.6, 32, feat2)]
    cur = xyz
    for name, N, M, r, S, feat in layers:
        C = feat.shape[1]
        temp = torch.empty(B, N, device=dev)
        idx = torch.empty(B, M, dtype=torch.int32, device=dev)

        def fps():
            temp.fill_(1e10)
            be.farthest_point_sampling_wrapper(B, N, M, cur, temp, idx)
        res[name + "_fps"] = timed(fps, iters)
        xyz_t = cur.transpose(1, 2).contiguous()
        new_t = torch.empty(B, 3, M, device=dev)
        res[name + "_gather"] = timed(lambda: be.gather_points_wrapper(B, 3, N, M, xyz_t, idx, new_t), iters)
        new_xyz = new_t.transpose(1, 2).contiguous()
        bidx = torch.zeros(B, M, S, dtype=torch.int32, device=dev)

        def bq():
            bidx.zero_()
            be.ball_query_wrapper(B, N, M, r, S, new_xyz, cur, bidx)
        res[name + "_ball_query"] = timed(bq, iters)
        gx = torch.empty(B, 3, M, S, device=dev)
        res[name + "_group_xyz"] = timed(lambda: be.group_points_wrapper(B, 3, N, M, S, xyz_t, bidx, gx), iters)
        gf = torch.empty(B, C, M, S, device=dev)
        res[name + "_group_feat"] = timed(lambda: be.group_points_wrapper(B, C, N, M, S, feat, bidx, gf), iters)
        cur = new_xyz
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ops.json"))
    ap.add_argument("--no-ref", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    out = {"batch": a.batch, "gpu": torch.cuda.get_device_name(0)}
    out["ours_ms(median,min)"] = chain(ours, a.batch, dev, a.iters)
    if not a.no_ref:
        import build_ref
        ref = build_ref.load_ref()
        if ref is not None:
            out["reference_ms(median,min)"] = chain(ref, a.batch, dev, max(3, a.iters // 4))
    for k in ("ours_ms(median,min)", "reference_ms(median,min)"):
        if k in out:
            out[k]["total_median"] = sum(v[0] for v in out[k].values())
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
