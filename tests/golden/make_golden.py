"""Generate the golden vectors in tests/golden/*.npz by running the REFERENCE's own CUDA
kernels (oracle/_ref/pointnet2_batch_cuda_ref.so, compiled unmodified from /root/reference by
oracle/build_ref.py) on a B200.

    gpurun -- python tests/golden/make_golden.py gpurun_out/golden      # on the GPU box
    cp gpurun_out/golden/*.npz tests/golden/                            # back here

Every file stores the inputs next to the reference outputs, so the fixtures do not depend on
the synthetic generator staying bit-stable.  The reference has no tests or golden vectors of
its own, hence these are "outputs of the reference itself run here".
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from pdm_ssd_b200 import synthetic  # noqa: E402


def fps_cases():
    rng = np.random.default_rng(11)
    cases = {}
    cases["uniform_n1000_m128"] = (rng.uniform(-5, 5, (2, 1000, 3)).astype(np.float32), 128)  # bs=512, n not pow2
    x = rng.uniform(-5, 5, (2, 1024, 3)).astype(np.float32)
    x[:, 512:] = x[:, :512]  # every point duplicated once -> ties every round
    cases["dups_n1024_m600"] = (x, 600)
    cases["tiny_n37_m20"] = (rng.normal(0, 1, (3, 37, 3)).astype(np.float32), 20)
    cases["same_point_n300_m16"] = (np.ones((1, 300, 3), np.float32) * 1.5, 16)
    g = np.stack(np.meshgrid(np.arange(16), np.arange(16), np.arange(4), indexing="ij"), -1).reshape(-1, 3)
    cases["lattice_n1024_m512"] = (rng.permutation(g.astype(np.float32))[None].copy(), 512)  # many exact ties
    cases["kitti_n4096_m1024"] = (synthetic.kitti_batch(2, 4096)[..., :3].copy(), 1024)
    cases["kitti_n16384_m4096"] = (synthetic.kitti_batch(2, 16384, first_frame=6)[..., :3].copy(), 4096)
    cases["uniform_n2500_m700"] = (rng.uniform(0, 50, (2, 2500, 3)).astype(np.float32), 700)
    cases["n20000_m64"] = (rng.uniform(0, 50, (1, 20000, 3)).astype(np.float32), 64)  # > on-chip capacity
    return cases


def run(out_dir):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref
    ref = build_ref.load_ref()
    assert ref is not None, "oracle/_ref/pointnet2_batch_cuda_ref.so missing"
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(5)

    for name, (xyz, m) in fps_cases().items():
        x = torch.from_numpy(xyz).to(dev)
        B, N, _ = x.shape
        temp = torch.full((B, N), 1e10, device=dev)
        idx = torch.zeros((B, m), dtype=torch.int32, device=dev)
        ref.farthest_point_sampling_wrapper(B, N, m, x, temp, idx)
        torch.cuda.synchronize()
        np.savez_compressed(os.path.join(out_dir, "fps_%s.npz" % name), xyz=xyz, m=m,
                            idx=idx.cpu().numpy(), temp=temp.cpu().numpy())

    # ball query: KITTI-like with FPS centres, plus boundary / empty-ball constructions
    def bq(name, xyz, new_xyz, radius, nsample):
        x = torch.from_numpy(xyz).to(dev)
        q = torch.from_numpy(new_xyz).to(dev)
        B, N, _ = x.shape
        M = q.shape[1]
        idx = torch.zeros((B, M, nsample), dtype=torch.int32, device=dev)
        ref.ball_query_wrapper(B, N, M, radius, nsample, q, x, idx)
        torch.cuda.synchronize()
        np.savez_compressed(os.path.join(out_dir, "bq_%s.npz" % name), xyz=xyz, new_xyz=new_xyz,
                            radius=np.float32(radius), nsample=nsample, idx=idx.cpu().numpy())

    fr = synthetic.kitti_batch(2, 4096, first_frame=3)[..., :3].copy()
    x = torch.from_numpy(fr).to(dev)
    temp = torch.full((2, 4096), 1e10, device=dev)
    fi = torch.zeros((2, 1024), dtype=torch.int32, device=dev)
    ref.farthest_point_sampling_wrapper(2, 4096, 1024, x, temp, fi)
    centres = torch.gather(x, 1, fi.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous().cpu().numpy()
    bq("kitti_n4096_m1024_r0p8_s32", fr, centres, 0.8, 32)
    bq("kitti_n4096_m1024_r1p6_s16", fr, centres, 1.6, 16)
    # exact-boundary: points at distance exactly r (excluded by the strict <) and one ulp inside
    r = np.float32(0.5)
    pts = np.zeros((1, 64, 3), np.float32)
    pts[0, :16, 0] = r
    pts[0, 16:32, 0] = np.nextafter(r, np.float32(0))
    pts[0, 32:48, 1] = np.nextafter(r, np.float32(1))
    pts[0, 48:, :] = rng.uniform(-0.4, 0.4, (16, 3)).astype(np.float32)
    pts = pts[:, rng.permutation(64)]
    bq("boundary_n64", pts, np.zeros((1, 3, 3), np.float32), float(r), 48)
    # empty balls: centres far away from every point -> rows stay zero
    far = rng.uniform(0, 1, (2, 200, 3)).astype(np.float32)
    q = np.concatenate([far[:, :5], far[:, :5] + 100.0], 1).astype(np.float32)
    bq("empty_rows_n200", far, q, 0.3, 8)

    # three_nn / three_interpolate
    unk = rng.uniform(0, 10, (2, 500, 3)).astype(np.float32)
    kn = rng.uniform(0, 10, (2, 77, 3)).astype(np.float32)
    kn[:, 40:] = kn[:, :37]  # duplicate known points -> equal distances
    feats = rng.normal(0, 1, (2, 5, 77)).astype(np.float32)
    d2 = torch.zeros((2, 500, 3), device=dev)
    ni = torch.zeros((2, 500, 3), dtype=torch.int32, device=dev)
    ref.three_nn_wrapper(2, 500, 77, torch.from_numpy(unk).to(dev), torch.from_numpy(kn).to(dev), d2, ni)
    dist = torch.sqrt(d2)
    inv = 1.0 / (dist + 1e-8)
    wgt = (inv / inv.sum(2, keepdim=True)).contiguous()
    out = torch.zeros((2, 5, 500), device=dev)
    ref.three_interpolate_wrapper(2, 5, 77, 500, torch.from_numpy(feats).to(dev), ni, wgt, out)
    torch.cuda.synchronize()
    np.savez_compressed(os.path.join(out_dir, "interp_n500_m77.npz"), unknown=unk, known=kn, feats=feats,
                        dist2=d2.cpu().numpy(), idx=ni.cpu().numpy(), weight=wgt.cpu().numpy(),
                        out=out.cpu().numpy())
    # two known points only: third neighbour stays at its 1e40 -> inf initial value
    d2b = torch.zeros((1, 4, 3), device=dev)
    nib = torch.zeros((1, 4, 3), dtype=torch.int32, device=dev)
    ref.three_nn_wrapper(1, 4, 2, torch.from_numpy(unk[:1, :4].copy()).to(dev),
                         torch.from_numpy(kn[:1, :2].copy()).to(dev), d2b, nib)
    np.savez_compressed(os.path.join(out_dir, "three_nn_m2.npz"), unknown=unk[:1, :4], known=kn[:1, :2],
                        dist2=d2b.cpu().numpy(), idx=nib.cpu().numpy())

    # group / gather (pure copies; small)
    pts = rng.normal(0, 1, (2, 6, 300)).astype(np.float32)
    gi = rng.integers(0, 300, (2, 50, 7)).astype(np.int32)
    go = torch.zeros((2, 6, 50, 7), device=dev)
    ref.group_points_wrapper(2, 6, 300, 50, 7, torch.from_numpy(pts).to(dev), torch.from_numpy(gi).to(dev), go)
    ga = torch.zeros((2, 6, 50), device=dev)
    ref.gather_points_wrapper(2, 6, 300, 50, torch.from_numpy(pts).to(dev),
                              torch.from_numpy(gi[:, :, 0].copy()).to(dev), ga)
    torch.cuda.synchronize()
    np.savez_compressed(os.path.join(out_dir, "group_gather.npz"), points=pts, idx=gi,
                        grouped=go.cpu().numpy(), gathered=ga.cpu().numpy())
    print("golden vectors written to", out_dir, sorted(os.listdir(out_dir)))


if __name__ == "__main__":
    run(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
