"""GPU parity tests of the stacked (ragged-batch) operator family: our CUDA path (ctypes -> C ABI through
pdm_ssd_b200.pointnet2_stack_cuda, the drop-in for the reference's `pointnet2_stack_cuda`) against
  * the golden vectors the reference's own kernels produced on a B200 (tests/golden/stack_*.npz),
  * the C restatement oracle/pdm_stack_oracle.c on further seeded cases,
  * the live reference extension (oracle/_ref) and the reference's OWN Python layers mounted on both.
Bit-exact for every index / copy / in-order-sum output; gradients (atomicAdd order in the reference) to 1e-5.
"""
import glob
import os

import numpy as np
import pytest
import torch

from pdm_ssd_b200 import pointnet2_stack_cuda as ours

import stack_oracle as so
from golden.make_golden_stack import ragged_cloud, voxelize

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.fixture(scope="module")
def ref_stack():
    import build_ref
    return build_ref.load_ref_stack()


def _fps(ext, xyz, cnt, mcnt):
    x, xc, mc = T(xyz), T(cnt), T(mcnt)
    temp = torch.full((xyz.shape[0],), 1e10, device=DEV)
    idx = torch.zeros((int(mcnt.sum()),), dtype=torch.int32, device=DEV)
    ext.stack_farthest_point_sampling_wrapper(x, temp, xc, idx, mc)
    return idx.cpu().numpy(), temp.cpu().numpy()


def _bq(ext, r, ns, xyz, cnt, new_xyz, mcnt):
    idx = torch.zeros((new_xyz.shape[0], ns), dtype=torch.int32, device=DEV)
    ext.ball_query_wrapper(len(cnt), new_xyz.shape[0], r, ns, T(new_xyz), T(mcnt), T(xyz), T(cnt), idx)
    return idx.cpu().numpy()


def _rows_sorted(a):
    a = np.asarray(a)
    return a[np.lexsort(a.T[::-1])]


def _vector_pool(ext, x, xc, f, q, qc, grids, dmax, ceg, ns, ntype, ptype):
    g = grids[0] * grids[1] * grids[2]
    m = q.shape[0]
    mean = 100
    while True:
        nf = torch.zeros((m, ceg * g), device=DEV)
        nl = torch.zeros((m, 3 * g), device=DEV)
        pc = torch.zeros((m, g), dtype=torch.int32, device=DEV)
        cap = mean * m
        grp = torch.zeros((cap, 3), dtype=torch.int32, device=DEV)
        cum = ext.vector_pool_wrapper(x, xc, f, q, qc, nf, nl, pc, grp, grids[0], grids[1], grids[2], dmax, 1, cap, ns, ntype, ptype)
        if cum <= cap:
            return nf, nl, pc, grp[:cum].contiguous(), cum
        mean = cum // m + 1


def _local_list(ext, x, xc, q, qc, dmax, ns, ntype, avg=200):
    m = q.shape[0]
    while True:
        lst = torch.zeros((avg * m,), dtype=torch.int32, device=DEV)
        sl = torch.zeros((m, 2), dtype=torch.int32, device=DEV)
        cs = torch.zeros((1,), dtype=torch.int32, device=DEV)
        ext.query_stacked_local_neighbor_idxs_wrapper_stack(x, xc, q, qc, lst, sl, cs, avg, dmax, ns, ntype)
        tot = int(cs.item())
        if tot <= avg * m:
            return lst[:tot].contiguous(), sl, tot
        avg = tot // m + 1


def _lists_per_centre(lst, sl):
    lst, sl = np.asarray(lst), np.asarray(sl)
    return [lst[s:s + n].tolist() for s, n in sl]


# ------------------------------------------------------------------------------------------------- goldens
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "stack_*_core.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_core_ops_vs_reference_golden(path):
    g = np.load(path)
    xyz, feat, cnt, mcnt = g["xyz"], g["feat"], g["cnt"], g["mcnt"]
    N, M, B = xyz.shape[0], int(mcnt.sum()), len(cnt)
    idx, temp = _fps(ours, xyz, cnt, mcnt)
    assert np.array_equal(idx, g["fps_idx"]) and np.array_equal(temp, g["fps_temp"])
    for key in ("bq_r1_s16", "bq_r2.5_s32"):
        r, ns = (1.0, 16) if key == "bq_r1_s16" else (2.5, 32)
        assert np.array_equal(_bq(ours, r, ns, xyz, cnt, g["new_xyz"], mcnt), g[key]), key
    assert (g["bq_r1_s16"][:, 0] == -1).sum() > 0                      # the fixture does contain empty balls
    gi = g["bq_r1_s16"].copy()
    gi[gi[:, 0] == -1] = 0
    out = torch.zeros((M, 8, 16), device=DEV)
    ours.group_points_wrapper(B, M, 8, 16, T(feat), T(cnt), T(gi), T(mcnt), out)
    assert np.array_equal(out.cpu().numpy(), g["group_out"])
    for det in (False, True):
        torch.use_deterministic_algorithms(det)
        try:
            gf = torch.zeros((N, 8), device=DEV)
            ours.group_points_grad_wrapper(B, M, 8, N, 16, T(g["grad_out"]), T(gi), T(mcnt), T(cnt), gf)
            gk = torch.zeros((M, 8), device=DEV)
            ours.three_interpolate_grad_wrapper(T(g["grad_interp"]), T(g["nn_idx"]), T(g["weight"]), gk)
        finally:
            torch.use_deterministic_algorithms(False)
        np.testing.assert_allclose(gf.cpu().numpy(), g["grad_feat"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(gk.cpu().numpy(), g["grad_known"], rtol=1e-5, atol=2e-5)
    new_xyz = xyz[g["fps_idx"]]
    d2 = torch.zeros((N, 3), device=DEV)
    nn = torch.zeros((N, 3), dtype=torch.int32, device=DEV)
    ours.three_nn_wrapper(T(xyz), T(cnt), T(new_xyz), T(mcnt), d2, nn)
    assert np.array_equal(nn.cpu().numpy(), g["nn_idx"]) and np.array_equal(d2.cpu().numpy(), g["nn_dist2"])
    interp = torch.zeros((N, 8), device=DEV)
    ours.three_interpolate_wrapper(T(g["known_feat"]), T(g["nn_idx"]), T(g["weight"]), interp)
    assert np.array_equal(interp.cpu().numpy(), g["interp"])
    pi = g["point_indices"]
    vq = torch.zeros((M, 16), dtype=torch.int32, device=DEV)
    ours.voxel_query_wrapper(M, pi.shape[1], pi.shape[2], pi.shape[3], 16, 1.6, 2, 2, 2, T(g["new_xyz"]), T(xyz), T(g["new_coords"]), T(pi), vq)
    assert np.array_equal(vq.cpu().numpy(), g["vq_idx"])


GOLD_VP = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "stack_*_vpool.npz")))


@pytest.mark.parametrize("path", GOLD_VP, ids=[os.path.basename(p) for p in GOLD_VP])
def test_vector_pool_family_vs_reference_golden(path):
    g = np.load(path)
    x, f, xc, mc, q = T(g["xyz"]), T(g["feat"]), T(g["cnt"]), T(g["mcnt"]), T(g["new_xyz"])
    N = g["xyz"].shape[0]
    for name in ("cube_avg", "ball_avg_ns", "cube_first"):
        gx, gy, gz, ceg, ns, ntype, ptype = g[name + "_cfg"].tolist()
        nf, nl, pc, grp, cum = _vector_pool(ours, x, xc, f, q, mc, (gx, gy, gz), float(g[name + "_dmax"]), ceg, ns, ntype, ptype)
        assert cum == g[name + "_grp"].shape[0]
        assert np.array_equal(pc.cpu().numpy(), g[name + "_pc"]), name
        assert np.array_equal(nf.cpu().numpy(), g[name + "_nf"]), name          # in-order sums: bit-exact
        assert np.array_equal(nl.cpu().numpy(), g[name + "_nl"]), name
        assert np.array_equal(_rows_sorted(grp.cpu().numpy()), _rows_sorted(g[name + "_grp"])), name
        gs = torch.zeros((N, 8), device=DEV)
        if cum:
            ours.vector_pool_grad_wrapper(T(g[name + "_gnf"]), pc, grp, gs)
        np.testing.assert_allclose(gs.cpu().numpy(), g[name + "_gsf"], rtol=1e-5, atol=1e-5)
    for name in ("ln_cube", "ln_ball_ns"):
        ns, ntype, _, tot = g[name + "_cfg"].tolist()
        lst, sl, total = _local_list(ours, x, xc, q, mc, float(g[name + "_dmax"]), ns, ntype)
        assert total == tot
        assert _lists_per_centre(lst.cpu().numpy(), sl.cpu().numpy()) == _lists_per_centre(g[name + "_list"], g[name + "_start_len"])
        M = q.shape[0]
        gidx = torch.full((M, 6, 3), -1, dtype=torch.int32, device=DEV)
        gd2 = torch.zeros((M, 6, 3), device=DEV)
        ours.query_three_nn_by_stacked_local_idxs_wrapper_stack(x, q, T(g[name + "_centers"]), gidx, gd2, lst, sl, M, 6)
        assert np.array_equal(gidx.cpu().numpy(), g[name + "_gidx"]) and np.array_equal(gd2.cpu().numpy(), g[name + "_gd2"])


# ------------------------------------------------------------------------------------------------- oracle, seeded
@pytest.mark.parametrize("counts,mcounts", [([900, 2100, 64], [100, 256, 64]), ([4096, 4096], [1024, 512]), ([1, 300], [1, 7]),
                                            ([16384, 700], [512, 100]), ([30000, 9000], [300, 100])])   # last: register top-k lookup
def test_fps_ball_query_group_vs_oracle(counts, mcounts):
    xyz, feat, cnt = ragged_cloud(counts, first_frame=31, dup_frame=0 if counts[0] > 1 and counts[0] < 5000 else None) \
        if max(counts) <= 4096 else (None, None, None)
    if xyz is None:          # frames larger than the synthetic generator's 4096: uniform clouds
        rng = np.random.default_rng(5)
        cnt = np.asarray(counts, np.int32)
        xyz = rng.uniform(0, 40, (int(cnt.sum()), 3)).astype(np.float32)
        feat = rng.normal(0, 1, (xyz.shape[0], 8)).astype(np.float32)
    mcnt = np.asarray(mcounts, np.int32)
    idx, temp = _fps(ours, xyz, cnt, mcnt)
    oidx, otemp = so.fps(xyz, cnt, mcnt, return_temp=True)
    assert np.array_equal(idx, oidx) and np.array_equal(temp, otemp)
    new_xyz = xyz[idx].copy()
    new_xyz[::13] -= 300.0
    for r, ns in ((0.6, 8), (1.7, 32), (4.0, 64)):
        mine = _bq(ours, r, ns, xyz, cnt, new_xyz, mcnt)
        assert np.array_equal(mine, so.ball_query(r, ns, xyz, cnt, new_xyz, mcnt)), (r, ns)
    gi = mine.copy()
    gi[gi[:, 0] == -1] = 0
    for c in (8, 5):
        ft = np.ascontiguousarray(feat[:, :c])
        out = torch.zeros((gi.shape[0], c, gi.shape[1]), device=DEV)
        ours.group_points_wrapper(len(cnt), gi.shape[0], c, gi.shape[1], T(ft), T(cnt), T(gi), T(mcnt), out)
        assert np.array_equal(out.cpu().numpy(), so.group_points(ft, cnt, gi, mcnt))


def test_fps_frame_larger_than_one_sm_and_forced_generic(monkeypatch):
    rng = np.random.default_rng(9)
    cnt = np.asarray([20000, 1500], np.int32)
    mcnt = np.asarray([64, 200], np.int32)
    xyz = rng.uniform(0, 60, (int(cnt.sum()), 3)).astype(np.float32)
    oidx, otemp = so.fps(xyz, cnt, mcnt, return_temp=True)
    idx, temp = _fps(ours, xyz, cnt, mcnt)
    assert np.array_equal(idx, oidx) and np.array_equal(temp, otemp)
    monkeypatch.setenv("PDM_FPS_KERNEL", "generic")
    idx, temp = _fps(ours, xyz, cnt, mcnt)
    assert np.array_equal(idx, oidx) and np.array_equal(temp, otemp)


def test_three_nn_interpolate_voxel_query_vs_oracle():
    xyz, feat, cnt = ragged_cloud([700, 1300, 90], first_frame=40)
    mcnt = np.asarray([40, 2, 30], np.int32)        # frame 1 has fewer than three known points
    idx, _ = _fps(ours, xyz, cnt, mcnt)
    known = xyz[idx]
    d2 = torch.zeros((xyz.shape[0], 3), device=DEV)
    nn = torch.zeros((xyz.shape[0], 3), dtype=torch.int32, device=DEV)
    ours.three_nn_wrapper(T(xyz), T(cnt), T(known), T(mcnt), d2, nn)
    od2, onn = so.three_nn(xyz, cnt, known, mcnt)
    assert np.array_equal(nn.cpu().numpy(), onn) and np.array_equal(d2.cpu().numpy(), od2)
    w = np.random.default_rng(2).uniform(0, 1, (xyz.shape[0], 3)).astype(np.float32)
    for c in (8, 3):
        kf = np.random.default_rng(3).normal(0, 1, (known.shape[0], c)).astype(np.float32)
        out = torch.zeros((xyz.shape[0], c), device=DEV)
        ours.three_interpolate_wrapper(T(kf), T(onn), T(w), out)
        assert np.array_equal(out.cpu().numpy(), so.three_interpolate(kf, onn, w))
    pi, coords = voxelize(xyz, cnt, voxel=(0.5, 0.5, 0.4))
    for rng_, r, ns in (((1, 1, 1), 0.9, 8), ((2, 3, 4), 2.0, 32)):
        vq = torch.zeros((known.shape[0], ns), dtype=torch.int32, device=DEV)
        ours.voxel_query_wrapper(known.shape[0], pi.shape[1], pi.shape[2], pi.shape[3], ns, r, rng_[0], rng_[1], rng_[2],
                                 T(known), T(xyz), T(coords[idx]), T(pi), vq)
        assert np.array_equal(vq.cpu().numpy(), so.voxel_query(rng_, r, ns, xyz, known, coords[idx], pi))


def test_vector_pool_family_vs_oracle():
    xyz, feat, cnt = ragged_cloud([1200, 500], first_frame=50)
    mcnt = np.asarray([90, 60], np.int32)
    idx, _ = _fps(ours, xyz, cnt, mcnt)
    q = xyz[idx]
    x, f, xc, mc, qt = T(xyz), T(feat), T(cnt), T(mcnt), T(q)
    for grids, dmax, ceg, ns, ntype, ptype in (((3, 3, 3), 1.1, 8, -1, 0, 0), ((2, 3, 2), 1.6, 4, 16, 1, 0), ((2, 2, 2), 0.9, 2, -1, 1, 1),
                                               ((4, 4, 2), 2.0, 1, 5, 0, 1)):
        nf, nl, pc, grp, cum = _vector_pool(ours, x, xc, f, qt, mc, grids, dmax, ceg, ns, ntype, ptype)
        onf, onl, opc, ogrp, ocum = so.vector_pool(xyz, cnt, feat, q, mcnt, grids, dmax, ceg, True, max(cum, 1) + 5, ns, ntype, ptype)
        assert cum == ocum
        assert np.array_equal(pc.cpu().numpy(), opc) and np.array_equal(nf.cpu().numpy(), onf) and np.array_equal(nl.cpu().numpy(), onl)
        assert np.array_equal(_rows_sorted(grp.cpu().numpy()), _rows_sorted(ogrp[:ocum]))
        gnf = np.random.default_rng(1).normal(0, 1, onf.shape).astype(np.float32)
        gs = torch.zeros((xyz.shape[0], 8), device=DEV)
        if cum:
            ours.vector_pool_grad_wrapper(T(gnf), pc, grp, gs)
        np.testing.assert_allclose(gs.cpu().numpy(), so.vector_pool_grad(gnf, opc, ogrp[:ocum], xyz.shape[0], 8), rtol=1e-5, atol=1e-5)
    for dmax, ns, ntype in ((1.0, -1, 0), (2.2, 9, 1), (30.0, -1, 1)):       # the last one hits the 1000-entry cap
        lst, sl, tot = _local_list(ours, x, xc, qt, mc, dmax, ns, ntype)
        olst, osl, otot = so.local_neighbor_idxs(xyz, cnt, q, mcnt, max(tot // q.shape[0] + 1, 1), dmax, ns, ntype)
        assert tot == otot
        assert _lists_per_centre(lst.cpu().numpy(), sl.cpu().numpy()) == _lists_per_centre(olst, osl)
        centers = (q[:, None, :] + np.random.default_rng(4).uniform(-1, 1, (1, 5, 3))).astype(np.float32)
        gidx = torch.full((q.shape[0], 5, 3), -1, dtype=torch.int32, device=DEV)
        gd2 = torch.zeros((q.shape[0], 5, 3), device=DEV)
        ours.query_three_nn_by_stacked_local_idxs_wrapper_stack(x, qt, T(centers), gidx, gd2, lst, sl, q.shape[0], 5)
        od2, oidx = so.three_nn_local(xyz, centers, lst.cpu().numpy(), sl.cpu().numpy())
        assert np.array_equal(gidx.cpu().numpy(), oidx) and np.array_equal(gd2.cpu().numpy(), od2)


def test_deterministic_stack_grads_repeat_bitwise():
    xyz, feat, cnt = ragged_cloud([2000, 1000], first_frame=60)
    mcnt = np.asarray([300, 200], np.int32)
    idx, _ = _fps(ours, xyz, cnt, mcnt)
    gi = _bq(ours, 3.0, 32, xyz, cnt, xyz[idx], mcnt)
    gi[gi[:, 0] == -1] = 0
    go = T(np.random.default_rng(0).normal(0, 1, (gi.shape[0], 8, 32)).astype(np.float32))
    torch.use_deterministic_algorithms(True)
    try:
        runs = []
        for _ in range(3):
            gf = torch.zeros((xyz.shape[0], 8), device=DEV)
            ours.group_points_grad_wrapper(len(cnt), gi.shape[0], 8, xyz.shape[0], 32, go, T(gi), T(mcnt), T(cnt), gf)
            runs.append(gf.cpu().numpy())
    finally:
        torch.use_deterministic_algorithms(False)
    assert np.array_equal(runs[0], runs[1]) and np.array_equal(runs[0], runs[2])
    np.testing.assert_allclose(runs[0], so.group_points_grad(go.cpu().numpy(), gi, mcnt, cnt, xyz.shape[0]), rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------------------- live reference
def test_every_entry_vs_live_reference_extension(ref_stack):
    if ref_stack is None:
        pytest.skip("oracle/_ref/pointnet2_stack_cuda_ref.so not built")
    xyz, feat, cnt = ragged_cloud([3000, 1800, 2500], first_frame=70, dup_frame=2)
    mcnt = np.asarray([400, 256, 300], np.int32)
    a, ta = _fps(ours, xyz, cnt, mcnt)
    b, tb = _fps(ref_stack, xyz, cnt, mcnt)
    assert np.array_equal(a, b) and np.array_equal(ta, tb)
    q = xyz[a].copy()
    q[::11] += 400.0
    for r, ns in ((0.8, 16), (3.0, 64)):
        assert np.array_equal(_bq(ours, r, ns, xyz, cnt, q, mcnt), _bq(ref_stack, r, ns, xyz, cnt, q, mcnt))
    x, f, xc, mc, qt = T(xyz), T(feat), T(cnt), T(mcnt), T(xyz[a])
    for grids, dmax, ceg, ns, ntype, ptype in (((3, 3, 3), 1.3, 8, -1, 0, 0), ((2, 2, 2), 1.0, 4, 10, 1, 1)):
        ra = _vector_pool(ours, x, xc, f, qt, mc, grids, dmax, ceg, ns, ntype, ptype)
        rb = _vector_pool(ref_stack, x, xc, f, qt, mc, grids, dmax, ceg, ns, ntype, ptype)
        assert ra[4] == rb[4]
        for u, v in zip(ra[:3], rb[:3]):
            assert torch.equal(u, v)
        assert np.array_equal(_rows_sorted(ra[3].cpu().numpy()), _rows_sorted(rb[3].cpu().numpy()))


def test_reference_python_layers_run_on_our_stack_extension(ref_stack):
    """INTEGRATION option A for the stacked family: the reference's own pointnet2_stack/*.py, unchanged, on our module."""
    import build_ref
    from pdm_ssd_b200 import iou3d_nms_cuda as our_nms, pointnet2_batch_cuda as our_pn2
    mine = build_ref.load_reference_tree("refpy_stack_on_ours", our_pn2, our_nms, stack_ext=ours)
    if mine is None or not hasattr(mine, "stack_modules"):
        pytest.skip("reference python files not vendored (oracle/_ref/py)")
    theirs = None
    if ref_stack is not None:
        theirs = build_ref.load_reference_tree("refpy_stack_on_ref", build_ref.load_ref(), build_ref.load_ref_nms(), stack_ext=ref_stack)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    xyz, feat, cnt = ragged_cloud([2048, 1500], first_frame=80)
    x, f, xc = T(xyz), T(feat), T(cnt)
    outs = []
    for ns_ in (mine, theirs):
        if ns_ is None:
            continue
        su, sm = ns_.stack_utils, ns_.stack_modules
        idx = su.stack_farthest_point_sample(x, xc, [256, 128])
        mc = torch.tensor([256, 128], dtype=torch.int32, device=DEV)
        q = x[idx.long()].contiguous()
        torch.manual_seed(0)
        sa = sm.StackSAModuleMSG(radii=[0.8, 1.6], nsamples=[16, 32], mlps=[[8, 16, 16], [8, 16, 32]], use_xyz=True, pool_method="max_pool").to(DEV).eval()
        fp = sm.StackPointnetFPModule(mlp=[48 + 8, 32]).to(DEV).eval()
        vp = sm.VectorPoolAggregationModule(input_channels=8, num_local_voxel=(3, 3, 3), post_mlps=(32,), max_neighbor_distance=1.2,
                                            neighbor_nsample=-1, local_aggregation_type="voxel_avg_pool", num_reduced_channels=8,
                                            num_channels_of_local_aggregation=16).to(DEV).eval()
        vi = sm.VectorPoolAggregationModule(input_channels=8, num_local_voxel=(2, 2, 2), post_mlps=(16,), max_neighbor_distance=1.5,
                                            neighbor_nsample=-1, local_aggregation_type="local_interpolation", num_reduced_channels=8,
                                            num_channels_of_local_aggregation=8, neighbor_type=1).to(DEV).eval()
        with torch.no_grad():
            _, nf = sa(x, xc, q, mc, features=f)
            up = fp(x, xc, q, mc, unknown_feats=f, known_feats=nf)
            _, vf = vp(x, xc, q, mc, f)
            _, vif = vi(x, xc, q, mc, f)
        vf = torch.cat([vf, vif], 1)
        ff = f.clone().requires_grad_(True)
        _, nf2 = sa(x, xc, q, mc, features=ff)
        nf2.sum().backward()
        outs.append((idx, nf, up, vf, ff.grad))
    assert outs[0][0].dtype == torch.int32 and outs[0][1].shape == (384, 48)
    if len(outs) == 2:
        for k in range(4):
            assert torch.equal(outs[0][k], outs[1][k]), k
        assert torch.allclose(outs[0][4], outs[1][4], rtol=1e-4, atol=1e-5)
