"""GPU parity tests: our CUDA path (through the C ABI, via the reference-shaped Python API)
against (1) the CPU oracle, (2) the committed golden vectors produced by the reference's
kernels, (3) the reference's own compiled extension when oracle/_ref is present.
Integer outputs must be bit-exact; float outputs of pure copies bit-exact too."""
import glob
import os

import numpy as np
import pytest
import torch

import oracle
from pdm_ssd_b200 import _lib, pointnet2_utils as pu, pointnet2_batch_cuda as ours, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def our_fps(xyz, m, return_temp=False):
    x = _t(xyz)
    B, N, _ = x.shape
    temp = torch.full((B, N), 1e10, device=DEV)
    idx = torch.full((B, m), -7, dtype=torch.int32, device=DEV)
    ours.farthest_point_sampling_wrapper(B, N, m, x, temp, idx)
    torch.cuda.synchronize()
    return (idx.cpu().numpy(), temp.cpu().numpy()) if return_temp else idx.cpu().numpy()


def test_library_is_native():
    assert os.path.exists(_lib.SO_PATH)
    assert _lib.load().pdm_abi_version() == 1


# ------------------------------------------------------------------ FPS
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "fps_*.npz"))))
def test_fps_golden(path):
    g = np.load(path)
    idx, temp = our_fps(g["xyz"], int(g["m"]), return_temp=True)
    assert np.array_equal(idx, g["idx"]), os.path.basename(path)
    assert np.array_equal(temp, g["temp"]), "running-min scratch differs from the reference's"


@pytest.mark.parametrize("n,m", [(16384, 4096), (4096, 1024), (5000, 700), (512, 512), (1000, 64)])
def test_fps_throughput_mode_is_bit_identical(n, m):
    """pdm_set_fps_mode(THROUGHPUT) selects fps_prepare + fps_l2_kernel (coordinates in L2, several
    frames per SM): same indices, same final scratch contents as the on-chip kernel and the oracle."""
    xyz = synthetic.kitti_batch(3, n, first_frame=n % 7)[..., :3].copy()
    xyz[1, n // 2:] = xyz[1, :n - n // 2]                    # duplicates: tie-breaks on every path
    want, want_t = our_fps(xyz, m, return_temp=True)
    _lib.set_fps_mode(_lib.FPS_MODE_THROUGHPUT)
    try:
        got, got_t = our_fps(xyz, m, return_temp=True)
    finally:
        _lib.set_fps_mode(_lib.FPS_MODE_AUTO)
    assert np.array_equal(got, want) and np.array_equal(got_t, want_t)
    if n <= 5000:
        assert np.array_equal(got, oracle.fps(xyz, m))
    with pytest.raises(_lib.PdmOpsError):
        _lib.set_fps_mode(7)


def test_fps_golden_present(golden_dir):
    assert len(glob.glob(os.path.join(golden_dir, "fps_*.npz"))) >= 8


@pytest.mark.parametrize("n,m,kind", [
    (512, 100, "uniform"), (700, 256, "uniform"), (1024, 1024, "uniform"), (1500, 300, "dups"),
    (2048, 512, "kitti"), (3000, 1000, "uniform"), (4096, 1024, "kitti"), (5000, 50, "dups"),
    (8192, 2048, "kitti"), (16384, 1024, "uniform"), (100, 100, "uniform"), (33, 7, "uniform"),
    (1, 1, "uniform"), (2, 2, "uniform"), (511, 64, "dups"), (16385, 40, "uniform"),
])
def test_fps_vs_oracle(n, m, kind):
    rng = np.random.default_rng(n * 7 + m)
    if kind == "kitti":
        xyz = synthetic.kitti_batch(3, n, first_frame=n % 11)[..., :3].copy()
    else:
        xyz = rng.uniform(-20, 20, (3, n, 3)).astype(np.float32)
        if kind == "dups":
            h = n // 3
            xyz[:, -h:] = xyz[:, :h]
    want, want_t = oracle.fps(xyz, m, return_temp=True)
    got, got_t = our_fps(xyz, m, return_temp=True)
    assert np.array_equal(got, want)
    assert np.array_equal(got_t, want_t)


def test_fps_full_size_properties():
    """BASELINE size (16 x 16384 -> 4096): oracle on 2 frames, structural properties on all."""
    frames = synthetic.kitti_batch(16)[..., :3].copy()
    idx = our_fps(frames, 4096)
    assert idx.shape == (16, 4096) and (idx[:, 0] == 0).all()
    assert idx.min() >= 0 and idx.max() < 16384
    want = oracle.fps(frames[:2], 4096)
    assert np.array_equal(idx[:2], want)
    for b in range(16):
        # distinct coordinates are never sampled twice before all distinct points are used
        pts = frames[b][idx[b]]
        nuniq = len(np.unique(frames[b], axis=0))
        assert len(np.unique(pts, axis=0)) == min(4096, nuniq)
    # batch independence: frame 5 alone gives the same answer
    assert np.array_equal(our_fps(frames[5:6], 4096)[0], idx[5])


def test_chain_full_size_properties():
    """BASELINE configs[1] at full size (16 x 16384 -> 4096 -> 1024, r = 0.8 / 1.6, 32 samples): properties
    that do not need the oracle -- ball-query rows hold the SMALLEST hit indices in ascending order and are
    padded with the first one, every listed neighbour is inside the ball and (away from the boundary) no
    in-ball point with a smaller index is missing; grouping is an exact gather; a second run is identical;
    the oracle itself on two frames."""
    from pdm_ssd_b200.sa_chain import SAChain
    frames = _t(synthetic.kitti_batch(16))
    xyz = frames[..., :3].contiguous()
    f1 = frames[..., 3:].transpose(1, 2).contiguous()
    f2 = torch.randn(16, 64, 4096, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3))
    chain = SAChain(16, 16384, device=DEV)
    ws = chain.run(xyz, (f1, f2))
    first = [{k: v.clone() for k, v in w.items() if torch.is_tensor(v)} for w in ws]
    cur = xyz
    for L, w, feat in zip(chain.layers, first, (f1, f2)):
        idx, new_xyz, bq, grouped = w["fps_idx"].long(), w["new_xyz"], w["ball_idx"].long(), w["grouped"]
        B, M, S = bq.shape
        assert torch.equal(new_xyz, torch.gather(cur, 1, idx[..., None].expand(-1, -1, 3)))
        # rows: strictly ascending up to the hit count, then copies of the first hit
        rising = bq[:, :, 1:] > bq[:, :, :-1]
        cnt = 1 + rising.long().cumprod(-1).sum(-1)                              # hits listed per row
        pos = torch.arange(S, device=DEV)[None, None, :]
        assert torch.equal(torch.where(pos < cnt[..., None], bq, bq[:, :, :1].expand(-1, -1, S)), bq)
        # every listed neighbour is inside the ball (the centre itself is a point of the cloud: no empty rows)
        nb = torch.gather(cur, 1, bq.reshape(B, -1)[..., None].expand(-1, -1, 3)).view(B, M, S, 3)
        d2 = ((nb.double() - new_xyz[:, :, None, :].double()) ** 2).sum(-1)
        r2 = float(np.float32(L.radius) * np.float32(L.radius))
        assert bool((d2 < r2 * (1 + 1e-5)).all())
        # completeness on a sample of centres: the in-ball points with the smallest indices are the row
        for b in (0, 7, 15):
            sel = torch.arange(0, M, 97, device=DEV)
            dd = ((cur[b][None, :, :].double() - new_xyz[b, sel][:, None, :].double()) ** 2).sum(-1)   # (sel, N)
            clear_in = dd < r2 * (1 - 1e-5)
            for row, c in enumerate(sel.tolist()):
                want = torch.nonzero(clear_in[row]).flatten()[:S]
                got = bq[b, c, : int(cnt[b, c])]
                # every clearly-inside point below the largest listed index is listed
                lim = int(got[-1]) if int(cnt[b, c]) == S else cur.shape[1]
                assert set(want[want <= lim].tolist()) <= set(got.tolist())
        # grouping = exact gather of [xyz - centre, features]
        C = feat.shape[1]
        gx = nb.permute(0, 3, 1, 2) - new_xyz.transpose(1, 2)[..., None]
        assert torch.equal(grouped[:, :3], gx)
        gf = torch.gather(feat, 2, bq.reshape(B, 1, -1).expand(-1, C, -1)).view(B, C, M, S)
        assert torch.equal(grouped[:, 3:], gf)
        cur = new_xyz
    # determinism: a second run (and one in throughput mode) gives identical tensors
    ws2 = chain.run(xyz, (f1, f2))
    for a, b_ in zip(first, ws2):
        for k in ("fps_idx", "new_xyz", "ball_idx", "grouped"):
            assert torch.equal(a[k], b_[k]), k
    _lib.set_fps_mode(_lib.FPS_MODE_THROUGHPUT)
    try:
        ws3 = chain.run(xyz, (f1, f2))
        for a, b_ in zip(first, ws3):
            assert torch.equal(a["fps_idx"], b_["fps_idx"]) and torch.equal(a["ball_idx"], b_["ball_idx"])
    finally:
        _lib.set_fps_mode(_lib.FPS_MODE_AUTO)
    # and the oracle on two frames
    fr = xyz[:2].cpu().numpy()
    fi = oracle.fps(fr, 4096)
    assert np.array_equal(first[0]["fps_idx"][:2].cpu().numpy(), fi)
    nx = np.take_along_axis(fr, fi[..., None].astype(np.int64).repeat(3, -1), 1)
    assert np.array_equal(first[0]["ball_idx"][:2].cpu().numpy(), oracle.ball_query(0.8, 32, fr, nx))


def test_fps_generic_kernel_matches(monkeypatch):
    xyz = synthetic.kitti_batch(2, 4096)[..., :3].copy()
    a = our_fps(xyz, 512)
    monkeypatch.setenv("PDM_FPS_KERNEL", "generic")
    b = our_fps(xyz, 512)
    assert np.array_equal(a, b)


def test_fps_vs_reference_extension(ref_ext):
    if ref_ext is None:
        pytest.skip("oracle/_ref not built")
    for n, m in [(16384, 4096), (4096, 1024), (1000, 333)]:
        xyz = _t(synthetic.kitti_batch(4, n, first_frame=20)[..., :3].copy())
        t1 = torch.full((4, n), 1e10, device=DEV); i1 = torch.zeros((4, m), dtype=torch.int32, device=DEV)
        t2 = torch.full((4, n), 1e10, device=DEV); i2 = torch.zeros((4, m), dtype=torch.int32, device=DEV)
        ref_ext.farthest_point_sampling_wrapper(4, n, m, xyz, t1, i1)
        ours.farthest_point_sampling_wrapper(4, n, m, xyz, t2, i2)
        torch.cuda.synchronize()
        assert torch.equal(i1, i2)
        assert torch.equal(t1, t2)


# ------------------------------------------------------------------ ball query
def our_bq(radius, nsample, xyz, new_xyz):
    return pu.ball_query(float(radius), int(nsample), _t(xyz), _t(new_xyz)).cpu().numpy()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "bq_*.npz"))))
def test_ball_query_golden(path):
    g = np.load(path)
    got = our_bq(g["radius"], g["nsample"], g["xyz"], g["new_xyz"])
    assert np.array_equal(got, g["idx"]), os.path.basename(path)


def test_ball_query_alternative_kernels_agree():
    """Every lookup kernel forced on every shape (PDM_BQ_KERNEL=bitmap: shared-memory bitmaps, the default up to 32768
    points and for nsample > 64; =topk: the register list, the default beyond; =tiled: reference-shaped scan) gives the
    oracle's rows.  Each variant runs in its own interpreter."""
    import subprocess, sys
    code = (
        "import numpy as np, torch, sys; sys.path.insert(0, %r); sys.path.insert(0, %r);"
        "import oracle; from pdm_ssd_b200 import pointnet2_utils as pu, synthetic;"
        "ok = True\n"
        "for n, m, r, s in [(16384, 4096, 0.8, 32), (4096, 1024, 1.6, 16), (1000, 37, 2.5, 64), (4096, 512, 3.0, 100)]:\n"
        "    fr = synthetic.kitti_batch(2, n, first_frame=3)[..., :3].copy()\n"
        "    c = oracle.fps(fr, m)\n"
        "    q = np.take_along_axis(fr, c[..., None].astype(np.int64).repeat(3, -1), 1)\n"
        "    got = pu.ball_query(r, s, torch.from_numpy(fr).cuda(), torch.from_numpy(q).cuda()).cpu().numpy()\n"
        "    ok = ok and np.array_equal(got, oracle.ball_query(r, s, fr, q))\n"
        "print('AGREE' if ok else 'DIFFER')"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    for variant in ("bitmap", "topk", "tiled"):
        env = dict(os.environ, PDM_BQ_KERNEL=variant)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        assert "AGREE" in out.stdout, (variant, out.stdout[-500:])


@pytest.mark.parametrize("n,m,r,s", [(4096, 1024, 0.8, 32), (4096, 1024, 1.6, 16), (1000, 37, 2.5, 64),
                                     (16384, 4096, 0.8, 32), (300, 300, 0.05, 4), (129, 1, 100.0, 200), (4096, 777, 2.4, 48),
                                     (16384, 1024, 3.5, 64), (2048, 2048, 1.0, 33), (5000, 999, 1.2, 1), (4096, 300, 4.0, 128)])
def test_ball_query_vs_oracle(n, m, r, s):
    fr = synthetic.kitti_batch(2, n, first_frame=n % 5)[..., :3].copy()
    cidx = oracle.fps(fr, m)
    new_xyz = np.take_along_axis(fr, cidx[..., None].astype(np.int64).repeat(3, -1), 1)
    assert np.array_equal(our_bq(r, s, fr, new_xyz), oracle.ball_query(r, s, fr, new_xyz))


def test_ball_query_leaves_rows_without_hits_untouched():
    xyz = _t(np.random.default_rng(0).uniform(0, 1, (1, 50, 3)).astype(np.float32))
    q = _t(np.full((1, 4, 3), 50.0, np.float32))
    idx = torch.full((1, 4, 6), 1234, dtype=torch.int32, device=DEV)
    ours.ball_query_wrapper(1, 50, 4, 0.5, 6, q, xyz, idx)
    assert (idx == 1234).all()


# ------------------------------------------------------------------ group / gather
@pytest.mark.parametrize("B,C,N,M,S", [(2, 3, 4096, 1024, 32), (2, 64, 4096, 1024, 32), (3, 1, 16384, 512, 32),
                                       (1, 5, 100, 7, 3), (2, 17, 333, 41, 5), (1, 8, 64, 1, 1)])
def test_group_and_gather(B, C, N, M, S):
    rng = np.random.default_rng(B * 100 + C)
    pts = rng.normal(0, 1, (B, C, N)).astype(np.float32)
    idx = rng.integers(0, N, (B, M, S)).astype(np.int32)
    got = pu.grouping_operation(_t(pts), _t(idx)).cpu().numpy()
    assert np.array_equal(got, oracle.group_points(pts, idx))
    g1 = pu.gather_operation(_t(pts), _t(idx[:, :, 0].copy())).cpu().numpy()
    assert np.array_equal(g1, oracle.gather_points(pts, idx[:, :, 0].copy()))


def test_group_gather_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "group_gather.npz"))
    assert np.array_equal(pu.grouping_operation(_t(g["points"]), _t(g["idx"])).cpu().numpy(), g["grouped"])
    assert np.array_equal(pu.gather_operation(_t(g["points"]), _t(g["idx"][:, :, 0].copy())).cpu().numpy(),
                          g["gathered"])


def test_group_gather_grads_match_oracle():
    rng = np.random.default_rng(3)
    pts = torch.tensor(rng.normal(0, 1, (2, 4, 200)).astype(np.float32), device=DEV, requires_grad=True)
    idx = rng.integers(0, 200, (2, 30, 6)).astype(np.int32)
    go = rng.normal(0, 1, (2, 4, 30, 6)).astype(np.float32)
    out = pu.grouping_operation(pts, _t(idx))
    out.backward(_t(go))
    np.testing.assert_allclose(pts.grad.cpu().numpy(), oracle.group_points_grad(go, idx, 200), rtol=1e-5, atol=1e-5)
    pts.grad = None
    out = pu.gather_operation(pts, _t(idx[:, :, 0].copy()))
    out.backward(_t(go[:, :, :, 0].copy()))
    np.testing.assert_allclose(pts.grad.cpu().numpy(),
                               oracle.gather_points_grad(go[:, :, :, 0].copy(), idx[:, :, 0].copy(), 200),
                               rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------ three_nn / interpolate
def test_interp_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "interp_n500_m77.npz"))
    B, n, _ = g["unknown"].shape
    m = g["known"].shape[1]
    d2 = torch.zeros((B, n, 3), device=DEV); ni = torch.zeros((B, n, 3), dtype=torch.int32, device=DEV)
    ours.three_nn_wrapper(B, n, m, _t(g["unknown"]), _t(g["known"]), d2, ni)
    assert np.array_equal(ni.cpu().numpy(), g["idx"])
    assert np.array_equal(d2.cpu().numpy(), g["dist2"])
    out = pu.three_interpolate(_t(g["feats"]), _t(g["idx"]), _t(g["weight"])).cpu().numpy()
    assert np.array_equal(out, g["out"])
    g2 = np.load(os.path.join(golden_dir, "three_nn_m2.npz"))
    d2 = torch.zeros((1, 4, 3), device=DEV); ni = torch.zeros((1, 4, 3), dtype=torch.int32, device=DEV)
    ours.three_nn_wrapper(1, 4, 2, _t(g2["unknown"]), _t(g2["known"]), d2, ni)
    assert np.array_equal(ni.cpu().numpy(), g2["idx"]) and np.array_equal(d2.cpu().numpy(), g2["dist2"])


@pytest.mark.parametrize("B,n,m,C", [(2, 1024, 256, 16), (1, 3000, 1500, 3), (2, 77, 5, 9), (1, 10, 1, 2)])
def test_three_nn_interpolate_vs_oracle(B, n, m, C):
    rng = np.random.default_rng(n + m)
    unk = rng.uniform(0, 10, (B, n, 3)).astype(np.float32)
    kn = rng.uniform(0, 10, (B, m, 3)).astype(np.float32)
    feats = rng.normal(0, 1, (B, C, m)).astype(np.float32)
    want_d2, want_i = oracle.three_nn(unk, kn)
    dist, idx = pu.three_nn(_t(unk), _t(kn))
    assert np.array_equal(idx.cpu().numpy(), want_i)
    assert np.array_equal(dist.cpu().numpy(), np.sqrt(want_d2))
    w = rng.uniform(0, 1, (B, n, 3)).astype(np.float32)
    got = pu.three_interpolate(_t(feats), _t(want_i), _t(w)).cpu().numpy()
    assert np.array_equal(got, oracle.three_interpolate(feats, want_i, w))
    f = torch.tensor(feats, device=DEV, requires_grad=True)
    go = rng.normal(0, 1, (B, C, n)).astype(np.float32)
    pu.three_interpolate(f, _t(want_i), _t(w)).backward(_t(go))
    np.testing.assert_allclose(f.grad.cpu().numpy(), oracle.three_interpolate_grad(go, want_i, w, m),
                               rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------ whole-op reference extension
def test_all_ops_vs_reference_extension(ref_ext):
    if ref_ext is None:
        pytest.skip("oracle/_ref not built")
    fr = _t(synthetic.kitti_batch(2, 4096, first_frame=9)[..., :3].copy())
    feats = torch.randn(2, 32, 4096, device=DEV)
    res = []
    for be in (ours, ref_ext):
        with pu.use_backend(be):
            fi = pu.farthest_point_sample(fr, 1024)
            new_xyz = pu.gather_operation(fr.transpose(1, 2).contiguous(), fi).transpose(1, 2).contiguous()
            grp = pu.QueryAndGroup(1.6, 32)(fr, new_xyz, feats)
            dist, ni = pu.three_nn(fr, new_xyz)
            w = 1.0 / (dist + 1e-8)
            w = (w / w.sum(2, keepdim=True)).contiguous()
            itp = pu.three_interpolate(grp[:, :, :, 0].contiguous(), ni, w)
            res.append((fi, new_xyz, grp, dist, ni, itp))
    for a, b in zip(*res):
        assert torch.equal(a, b)


# ------------------------------------------------------------------ error behaviour
def test_errors_raise_instead_of_exit():
    x = torch.zeros((1, 8, 3), device=DEV)
    with pytest.raises(RuntimeError):
        ours.farthest_point_sampling_wrapper(1, 8, 4, x.cpu(), torch.zeros(1, 8), torch.zeros(1, 4, dtype=torch.int32))
    with pytest.raises(RuntimeError):
        ours.farthest_point_sampling_wrapper(1, 8, 4, x, torch.zeros((1, 8), device=DEV),
                                             torch.zeros((1, 4), dtype=torch.int64, device=DEV))
    with pytest.raises(RuntimeError):
        ours.ball_query_wrapper(1, 8, 2, 0.5, 4, x[:, :2].transpose(1, 2), x, torch.zeros((1, 2, 4), dtype=torch.int32, device=DEV))
    with pytest.raises(_lib.PdmOpsError):
        ours.farthest_point_sampling_wrapper(1, 0, 4, x[:, :0].contiguous(), torch.zeros((1, 0), device=DEV),
                                             torch.zeros((1, 4), dtype=torch.int32, device=DEV))


# ------------------------------------------------------------------ streams, devices, big frames
def test_ops_follow_the_current_stream():
    """Launches go to torch's current stream (the reference uses the legacy default stream)."""
    xyz = _t(synthetic.kitti_batch(2, 4096, first_frame=30)[..., :3].copy())
    want = oracle.fps(xyz.cpu().numpy(), 512)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        big = torch.empty(64 << 20, device=DEV).normal_()       # keeps the side stream busy first
        idx = pu.farthest_point_sample(xyz, 512)
        new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
        bq = pu.ball_query(0.8, 16, xyz, new_xyz)
    side.synchronize()
    assert np.array_equal(idx.cpu().numpy(), want)
    assert np.array_equal(bq.cpu().numpy(), oracle.ball_query(0.8, 16, xyz.cpu().numpy(), new_xyz.cpu().numpy()))
    del big


def test_cuda_graph_capture_of_the_chain():
    """The whole chain (kernels, scratch cudaMallocAsync/FreeAsync) is capturable and replayable."""
    from pdm_ssd_b200.sa_chain import SAChain, SALayerCfg
    layers = (SALayerCfg(512, 0.8, 16, 1), SALayerCfg(128, 1.6, 16, 4))
    fr = synthetic.kitti_batch(2, 2048, first_frame=40)
    xyz = _t(fr[..., :3].copy())
    f1 = _t(fr[..., 3:].transpose(0, 2, 1).copy())
    f2 = torch.randn(2, 4, 512, device=DEV)
    chain = SAChain(2, 2048, layers, DEV)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        chain.run(xyz, (f1, f2))
    st.synchronize()
    eager = [w["ball_idx"].clone() for w in chain.ws] + [chain.ws[1]["grouped_feat"].clone()]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        chain.run(xyz, (f1, f2))
    for w in chain.ws:
        w["ball_idx"].fill_(-1)
    chain.ws[1]["grouped_feat"].zero_()
    g.replay()
    torch.cuda.synchronize()
    got = [w["ball_idx"] for w in chain.ws] + [chain.ws[1]["grouped_feat"]]
    for a, b in zip(eager, got):
        assert torch.equal(a, b)


def test_waymo_scale_frame():
    """BASELINE configs[4] shape: 163840 points.  FPS takes the any-size kernel, ball query the grid."""
    rng = np.random.default_rng(9)
    xyz = (rng.uniform(0, 1, (1, 163840, 3)) * np.array([150.4, 150.4, 6.0]) - np.array([75.2, 75.2, 2.0])).astype(np.float32)
    xyz[0, 100000:] = xyz[0, :63840]                       # heavy duplication -> ties
    m = 1024
    want, want_t = oracle.fps(xyz, m, return_temp=True)
    got, got_t = our_fps(xyz, m, return_temp=True)
    assert np.array_equal(got, want) and np.array_equal(got_t, want_t)
    new_xyz = np.take_along_axis(xyz, want[..., None].astype(np.int64).repeat(3, -1), 1)
    assert np.array_equal(our_bq(0.8, 32, xyz, new_xyz), oracle.ball_query(0.8, 32, xyz, new_xyz))
    assert np.array_equal(our_bq(40.0, 64, xyz, new_xyz[:, :64]), oracle.ball_query(40.0, 64, xyz, new_xyz[:, :64]))


@pytest.mark.parametrize("b,n,m", [(3, 50000, 300), (2, 16385, 200), (1, 196608, 64)])
def test_fps_cluster_kernel_vs_oracle(b, n, m):
    """Frames larger than one SM take the thread-block-cluster kernel (fps_cluster.cu): same samples and
    same final scratch as the oracle, duplicates (= exact ties) included; batch > 1, chunk tails, and the
    largest supported frame."""
    rng = np.random.default_rng(n)
    xyz = (rng.uniform(0, 1, (b, n, 3)) * np.array([150.4, 150.4, 6.0]) - np.array([75.2, 75.2, 2.0])).astype(np.float32)
    xyz[:, n // 2: n // 2 + n // 4] = xyz[:, : n // 4]          # a quarter of the points are duplicates
    want, want_t = oracle.fps(xyz, m, return_temp=True)
    got, got_t = our_fps(xyz, m, return_temp=True)
    assert np.array_equal(got, want) and np.array_equal(got_t, want_t)


@pytest.mark.parametrize("b,n,m,cl,kern", [(2, 40000, 3000, 0, ""), (1, 30000, 500, 4, ""), (2, 20000, 257, 16, ""), (2, 33000, 300, 0, "cluster"),
                                           (9, 163840, 96, 0, ""), (2, 40000, 2000, 0, "map=global"), (1, 30000, 400, 2, "map=global"),
                                           (2, 50000, 300, 13, "map=smem"), (8, 163840, 64, 0, "map=smem"), (1, 230000, 40, 0, "")])
def test_fps_cluster_bucket_kernel_variants(monkeypatch, b, n, m, cl, kern):
    """The bucket-pruned cluster kernel (fps_cluster_bucket.cu) with many samples (multi-sample rounds), forced cluster
    sizes (chunks with long padded tails, 16 small chunks), more frames than resident clusters (two waves), and the
    full-sweep cluster kernel it replaced (PDM_FPS_KERNEL=cluster), all against the oracle incl. the final scratch."""
    if cl:
        monkeypatch.setenv("PDM_FPS_CLUSTER", str(cl))
    if kern.startswith("map="):
        monkeypatch.setenv("PDM_FPS_CLUSTER_MAP", kern[4:])
    elif kern:
        monkeypatch.setenv("PDM_FPS_KERNEL", kern)
    rng = np.random.default_rng(n + m)
    xyz = (rng.uniform(0, 1, (b, n, 3)) * np.array([150.4, 150.4, 6.0]) - np.array([75.2, 75.2, 2.0])).astype(np.float32)
    xyz[:, n // 2: n // 2 + n // 8] = xyz[:, : n // 8]          # duplicates: exact ties
    want, want_t = oracle.fps(xyz, m, return_temp=True)
    got, got_t = our_fps(xyz, m, return_temp=True)
    assert np.array_equal(got, want) and np.array_equal(got_t, want_t)


def test_ball_query_degenerate_inputs():
    rng = np.random.default_rng(4)
    # all points identical; radius larger than the scene; radius tiny; nsample > n; non-finite coordinates
    same = np.ones((1, 40, 3), np.float32)
    assert np.array_equal(our_bq(0.1, 8, same, same[:, :3]), oracle.ball_query(0.1, 8, same, same[:, :3]))
    pts = rng.uniform(-1, 1, (2, 300, 3)).astype(np.float32)
    q = pts[:, :17].copy()
    for r, s in ((1000.0, 16), (1e-6, 4), (0.5, 400)):
        assert np.array_equal(our_bq(r, s, pts, q), oracle.ball_query(r, s, pts, q))
    bad = pts.copy()
    bad[0, 5] = np.nan
    bad[1, 7, 0] = np.inf
    assert np.array_equal(our_bq(0.5, 16, bad, q), oracle.ball_query(0.5, 16, bad, q))


def test_empty_and_trivial_sizes():
    z = torch.zeros((0, 8, 3), device=DEV)
    ours.farthest_point_sampling_wrapper(0, 8, 4, z, torch.zeros((0, 8), device=DEV), torch.zeros((0, 4), dtype=torch.int32, device=DEV))
    x = torch.rand((2, 50, 3), device=DEV)
    idx = torch.full((2, 0), 7, dtype=torch.int32, device=DEV)
    ours.farthest_point_sampling_wrapper(2, 50, 0, x, torch.full((2, 50), 1e10, device=DEV), idx)   # m = 0: no-op
    one = our_fps(x.cpu().numpy(), 1)
    assert (one == 0).all()
    out = torch.empty((2, 3, 0), device=DEV)
    ours.gather_points_wrapper(2, 3, 50, 0, x.transpose(1, 2).contiguous(), torch.zeros((2, 0), dtype=torch.int32, device=DEV), out)
    bq = torch.zeros((2, 0, 4), dtype=torch.int32, device=DEV)
    ours.ball_query_wrapper(2, 50, 0, 0.5, 4, torch.zeros((2, 0, 3), device=DEV), x, bq)


@pytest.mark.parametrize("C,use_xyz", [(0, True), (1, True), (64, True), (17, False)])
def test_query_and_group_fused_is_bit_identical(C, use_xyz, ref_ext):
    """a5: one-pass QueryAndGroup against the reference-shaped op sequence (and the reference kernels)."""
    fr = _t(synthetic.kitti_batch(2, 4096, first_frame=12)[..., :3].copy())
    feats = torch.randn(2, C, 4096, device=DEV) if C else None
    with torch.no_grad():
        fi = pu.farthest_point_sample(fr, 512)
        new_xyz = pu.gather_operation(fr.transpose(1, 2).contiguous(), fi).transpose(1, 2).contiguous()
        fused = pu.QueryAndGroup(1.2, 16, use_xyz=use_xyz)(fr, new_xyz, feats)
    with torch.enable_grad():                      # autograd on -> reference-shaped sequence over our kernels
        plain = pu.QueryAndGroup(1.2, 16, use_xyz=use_xyz)(fr, new_xyz, feats)
    assert fused.shape == plain.shape == (2, (3 if use_xyz else 0) + C, 512, 16)
    assert torch.equal(fused, plain)
    if ref_ext is not None:
        with pu.use_backend(ref_ext), torch.no_grad():
            assert torch.equal(pu.QueryAndGroup(1.2, 16, use_xyz=use_xyz)(fr, new_xyz, feats), fused)


# ---- deterministic backward (csrc/det_backward.cu) -----------------------------------------------
@pytest.fixture
def deterministic():
    prev = torch.are_deterministic_algorithms_enabled()
    torch.use_deterministic_algorithms(True, warn_only=True)
    yield
    torch.use_deterministic_algorithms(prev)


def test_deterministic_grads_match_oracle_and_repeat_bit_for_bit(deterministic):
    """With torch.use_deterministic_algorithms(True) the gather / group / three_interpolate gradients are summed in
    ascending source order: equal to the CPU oracle's serial sums to fp32 rounding, and bit-identical across runs even
    with heavy collisions (every index hit ~100 times), where atomicAdd results wobble in the last bits."""
    import oracle
    from pdm_ssd_b200 import pointnet2_batch_cuda as be
    rng = np.random.default_rng(11)
    B, C, N, M, S = 3, 13, 257, 400, 32
    idx = rng.integers(0, 40, (B, M, S)).astype(np.int32)            # 40 hot targets: long segments
    go = rng.standard_normal((B, C, M, S)).astype(np.float32)
    outs = []
    for _ in range(3):
        gp = torch.zeros((B, C, N), dtype=torch.float32, device=DEV)
        be.group_points_grad_wrapper(B, C, N, M, S, torch.from_numpy(go).to(DEV), torch.from_numpy(idx).to(DEV), gp)
        outs.append(gp.cpu().numpy())
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    np.testing.assert_allclose(outs[0], oracle.group_points_grad(go, idx, N), rtol=2e-5, atol=2e-5)
    # gather
    gi = rng.integers(0, 9, (B, M)).astype(np.int32)
    gg = rng.standard_normal((B, C, M)).astype(np.float32)
    a = torch.zeros((B, C, N), device=DEV)
    b2 = torch.zeros((B, C, N), device=DEV)
    be.gather_points_grad_wrapper(B, C, N, M, torch.from_numpy(gg).to(DEV), torch.from_numpy(gi).to(DEV), a)
    be.gather_points_grad_wrapper(B, C, N, M, torch.from_numpy(gg).to(DEV), torch.from_numpy(gi).to(DEV), b2)
    assert torch.equal(a, b2)
    np.testing.assert_allclose(a.cpu().numpy(), oracle.gather_points_grad(gg, gi, N), rtol=2e-5, atol=2e-5)
    # three_interpolate: (B, C, n) gradients to m known points through (idx, weight)
    n, m = 500, 37
    ti = rng.integers(0, m, (B, n, 3)).astype(np.int32)
    tw = rng.uniform(0, 1, (B, n, 3)).astype(np.float32)
    tg = rng.standard_normal((B, C, n)).astype(np.float32)
    r = []
    for _ in range(2):
        gp = torch.zeros((B, C, m), device=DEV)
        be.three_interpolate_grad_wrapper(B, C, n, m, torch.from_numpy(tg).to(DEV), torch.from_numpy(ti).to(DEV), torch.from_numpy(tw).to(DEV), gp)
        r.append(gp.cpu().numpy())
    assert np.array_equal(r[0], r[1])
    np.testing.assert_allclose(r[0], oracle.three_interpolate_grad(tg, ti, tw, m), rtol=2e-5, atol=2e-5)
    # "+=" contract: a second call accumulates on top of the first
    be.gather_points_grad_wrapper(B, C, N, M, torch.from_numpy(gg).to(DEV), torch.from_numpy(gi).to(DEV), a)
    np.testing.assert_allclose(a.cpu().numpy(), 2 * oracle.gather_points_grad(gg, gi, N), rtol=2e-5, atol=2e-5)
