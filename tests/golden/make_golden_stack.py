"""Golden vectors of the stacked operator family (tests/golden/stack_*.npz), produced by running the REFERENCE's own
pointnet2_stack CUDA kernels (oracle/_ref/pointnet2_stack_cuda_ref.so, compiled unmodified from /root/reference by
oracle/build_ref.py) on a B200.

    gpurun -- python tests/golden/make_golden_stack.py gpurun_out/golden_stack     # on the GPU box
    cp gpurun_out/golden_stack/*.npz tests/golden/                                 # back here

Every file stores its inputs next to the reference outputs.  Outputs whose ORDER the reference leaves to atomicAdd
(the stacked neighbour list, grouped_idxs) are stored as produced; the tests compare them as sets / through start_len.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from pdm_ssd_b200 import synthetic  # noqa: E402


def ragged_cloud(counts, first_frame=0, dup_frame=None):
    """frames of different sizes cut from KITTI-shaped synthetic frames; returns xyz (N,3), feat (N,C=8), cnt"""
    rng = np.random.default_rng(100 + first_frame)
    xs = []
    for i, n in enumerate(counts):
        fr = synthetic.kitti_batch(1, 4096, first_frame=first_frame + i)[0, :n, :3].copy()
        if dup_frame == i:
            fr[n // 2: 2 * (n // 2)] = fr[: n // 2]      # duplicated points -> sampling ties
        xs.append(fr)
    xyz = np.concatenate(xs, 0).astype(np.float32)
    feat = rng.normal(0, 1, (xyz.shape[0], 8)).astype(np.float32)
    return xyz, feat, np.asarray(counts, np.int32)


def voxelize(xyz, cnt, voxel=(0.8, 0.8, 0.8)):
    """point_indices (B,Z,Y,X) = global row of one point per voxel (the last in row order) or -1; coords per point"""
    lo = xyz.min(0)
    c = np.floor((xyz - lo) / np.asarray(voxel, np.float32)).astype(np.int32)      # x, y, z
    X, Y, Z = (c.max(0) + 1).tolist()
    B = len(cnt)
    pi = np.full((B, Z, Y, X), -1, np.int32)
    b = np.repeat(np.arange(B), cnt)
    pi[b, c[:, 2], c[:, 1], c[:, 0]] = np.arange(xyz.shape[0], dtype=np.int32)
    coords = np.stack([b, c[:, 2], c[:, 1], c[:, 0]], 1).astype(np.int32)            # [batch, z, y, x]
    return pi, coords


def run(out_dir):
    import build_ref
    ref = build_ref.load_ref_stack()
    assert ref is not None, "oracle/_ref/pointnet2_stack_cuda_ref.so missing"
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)   # noqa: E731

    for tag, counts, mcounts, dup in (("a", [1500, 700, 2300], [128, 64, 200], 1), ("b", [37, 1024, 5, 3000], [20, 300, 5, 17], None)):
        xyz, feat, cnt = ragged_cloud(counts, first_frame=7 if tag == "a" else 11, dup_frame=dup)
        mcnt = np.asarray(mcounts, np.int32)
        N, M = int(cnt.sum()), int(mcnt.sum())
        x, f, xc, mc = T(xyz), T(feat), T(cnt), T(mcnt)

        # farthest point sampling
        temp = torch.full((N,), 1e10, device=dev)
        fidx = torch.zeros((M,), dtype=torch.int32, device=dev)
        ref.stack_farthest_point_sampling_wrapper(x, temp, xc, fidx, mc)
        new_xyz = x[fidx.long()].contiguous()
        # a few centres far away from everything: empty balls
        new_xyz_e = new_xyz.clone()
        new_xyz_e[::17] += 500.0

        # ball query
        bq = {}
        for r, ns in ((1.0, 16), (2.5, 32)):
            idx = torch.zeros((M, ns), dtype=torch.int32, device=dev)
            ref.ball_query_wrapper(len(cnt), M, r, ns, new_xyz_e, mc, x, xc, idx)
            bq["bq_r%g_s%d" % (r, ns)] = idx.cpu().numpy()

        # grouping (+ grad) on the r=1.0 query with the empty balls redirected to 0 as the reference does
        gi = torch.from_numpy(bq["bq_r1_s16"]).to(dev).clone()
        gi[gi[:, 0] == -1] = 0
        out = torch.zeros((M, 8, 16), device=dev)
        ref.group_points_wrapper(len(cnt), M, 8, 16, f, xc, gi, mc, out)
        g_out = torch.from_numpy(np.random.default_rng(3).normal(0, 1, (M, 8, 16)).astype(np.float32)).to(dev)
        g_feat = torch.zeros((N, 8), device=dev)
        ref.group_points_grad_wrapper(len(cnt), M, 8, N, 16, g_out, gi, mc, xc, g_feat)

        # three_nn / interpolate: unknown = all points, known = sampled centres
        d2 = torch.zeros((N, 3), device=dev)
        nn = torch.zeros((N, 3), dtype=torch.int32, device=dev)
        ref.three_nn_wrapper(x, xc, new_xyz, mc, d2, nn)
        w = 1.0 / (torch.sqrt(d2) + 1e-8)
        w = (w / w.sum(1, keepdim=True)).contiguous()
        w[~torch.isfinite(w)] = 0.25
        kf = T(np.random.default_rng(4).normal(0, 1, (M, 8)).astype(np.float32))
        interp = torch.zeros((N, 8), device=dev)
        ref.three_interpolate_wrapper(kf, nn, w, interp)
        g_i = T(np.random.default_rng(5).normal(0, 1, (N, 8)).astype(np.float32))
        g_kf = torch.zeros((M, 8), device=dev)
        ref.three_interpolate_grad_wrapper(g_i, nn, w, g_kf)

        # voxel query
        pi, coords = voxelize(xyz, cnt)
        ncoords = coords[fidx.cpu().numpy()]
        vq = torch.zeros((M, 16), dtype=torch.int32, device=dev)
        ref.voxel_query_wrapper(M, pi.shape[1], pi.shape[2], pi.shape[3], 16, 1.6, 2, 2, 2, new_xyz_e, x, T(ncoords), T(pi), vq)

        torch.cuda.synchronize()
        np.savez_compressed(os.path.join(out_dir, "stack_%s_core.npz" % tag), xyz=xyz, feat=feat, cnt=cnt, mcnt=mcnt,
                            fps_idx=fidx.cpu().numpy(), fps_temp=temp.cpu().numpy(), new_xyz=new_xyz_e.cpu().numpy(),
                            group_out=out.cpu().numpy(), grad_out=g_out.cpu().numpy(), grad_feat=g_feat.cpu().numpy(),
                            nn_dist2=d2.cpu().numpy(), nn_idx=nn.cpu().numpy(), weight=w.cpu().numpy(), known_feat=kf.cpu().numpy(),
                            interp=interp.cpu().numpy(), grad_interp=g_i.cpu().numpy(), grad_known=g_kf.cpu().numpy(),
                            point_indices=pi, new_coords=ncoords, vq_idx=vq.cpu().numpy(), **bq)

        # vector-pool family
        vp = {}
        for name, (grids, dmax, ceg, ns, ntype, ptype) in {"cube_avg": ((3, 3, 3), 1.2, 4, -1, 0, 0), "ball_avg_ns": ((2, 2, 2), 1.5, 8, 24, 1, 0),
                                                         "cube_first": ((3, 3, 2), 1.0, 2, -1, 0, 1)}.items():
            g = grids[0] * grids[1] * grids[2]
            c_out = ceg * g
            mean = 100
            while True:
                nf = torch.zeros((M, c_out), device=dev)
                nl = torch.zeros((M, 3 * g), device=dev)
                pc = torch.zeros((M, g), dtype=torch.int32, device=dev)
                cap = mean * M
                grp = torch.zeros((cap, 3), dtype=torch.int32, device=dev)
                cum = ref.vector_pool_wrapper(x, xc, f, new_xyz, mc, nf, nl, pc, grp, grids[0], grids[1], grids[2], dmax, 1, cap, ns, ntype, ptype)
                if cum <= cap:
                    break
                mean = cum // M + 1
            grp = grp[:cum].contiguous()
            g_nf = T(np.random.default_rng(6).normal(0, 1, (M, c_out)).astype(np.float32))
            g_sf = torch.zeros((N, 8), device=dev)
            if cum > 0:
                ref.vector_pool_grad_wrapper(g_nf, pc, grp, g_sf)
            vp.update({name + "_cfg": np.asarray(list(grids) + [ceg, ns, ntype, ptype], np.int32), name + "_dmax": np.float32(dmax),
                       name + "_nf": nf.cpu().numpy(), name + "_nl": nl.cpu().numpy(), name + "_pc": pc.cpu().numpy(),
                       name + "_grp": grp.cpu().numpy(), name + "_gnf": g_nf.cpu().numpy(), name + "_gsf": g_sf.cpu().numpy()})
        # two-step three-nn of the local-interpolate path
        for name, (dmax, ns, ntype) in {"ln_cube": (1.2, -1, 0), "ln_ball_ns": (2.0, 12, 1)}.items():
            avg = 200
            while True:
                lst = torch.zeros((avg * M,), dtype=torch.int32, device=dev)
                sl = torch.zeros((M, 2), dtype=torch.int32, device=dev)
                cs = torch.zeros((1,), dtype=torch.int32, device=dev)
                ref.query_stacked_local_neighbor_idxs_wrapper_stack(x, xc, new_xyz, mc, lst, sl, cs, avg, dmax, ns, ntype)
                tot = int(cs.item())
                if tot <= avg * M:
                    break
                avg = tot // M + 1
            centers = (new_xyz[:, None, :] + T(np.random.default_rng(8).uniform(-1, 1, (1, 6, 3)).astype(np.float32))).contiguous()
            gidx = torch.full((M, 6, 3), -1, dtype=torch.int32, device=dev)
            gd2 = torch.zeros((M, 6, 3), device=dev)
            ref.query_three_nn_by_stacked_local_idxs_wrapper_stack(x, new_xyz, centers, gidx, gd2, lst[:tot].contiguous(), sl, M, 6)
            vp.update({name + "_cfg": np.asarray([ns, ntype, avg, tot], np.int32), name + "_dmax": np.float32(dmax),
                       name + "_list": lst[:tot].cpu().numpy(), name + "_start_len": sl.cpu().numpy(),
                       name + "_centers": centers.cpu().numpy(), name + "_gidx": gidx.cpu().numpy(), name + "_gd2": gd2.cpu().numpy()})
        torch.cuda.synchronize()
        np.savez_compressed(os.path.join(out_dir, "stack_%s_vpool.npz" % tag), xyz=xyz, feat=feat, cnt=cnt, mcnt=mcnt,
                            new_xyz=new_xyz.cpu().numpy(), **vp)
    print("wrote", sorted(os.listdir(out_dir)))


if __name__ == "__main__":
    run(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden_stack"))
