"""Seeded random-shape sweeps of the index kernels against the C oracle (bit-exact): ball query over every lookup kernel
(bitmap / register top-k, one or two list registers, nsample above 64) and both grid builds (cluster / one CTA per frame),
the stacked variants on ragged frames, and farthest point sampling across the on-chip / cluster-bucket / any-size kernels.
Shapes are drawn from a fixed seed; each case is sized so that the oracle finishes in well under a second."""
import numpy as np
import pytest
import torch

from pdm_ssd_b200 import _lib, pointnet2_batch_cuda as ours, pointnet2_stack_cuda as stack

import oracle
import stack_oracle as so

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _cloud(rng, b, n, kind):
    if kind == 0:      # planar LiDAR-like
        xyz = rng.uniform(0, 1, (b, n, 3)) * np.array([70.0, 80.0, 4.0]) - np.array([0.0, 40.0, 3.0])
    elif kind == 1:    # volumetric
        xyz = rng.normal(0, 3.0, (b, n, 3))
    else:              # clustered with duplicates (ties, crowded balls)
        c = rng.uniform(-20, 20, (b, 8, 3))
        xyz = np.stack([c[i][rng.integers(0, 8, n)] for i in range(b)]) + rng.normal(0, 0.7, (b, n, 3))
        xyz[:, n // 2:] = xyz[:, : n - n // 2]
    return xyz.astype(np.float32)


def _bq_cases():
    rng = np.random.default_rng(20261019)
    out = []
    for i in range(24):
        b = int(rng.integers(1, 4))
        n = int(rng.choice([7, 300, 4096, 9000, 33000, 40000]))
        m = int(min(n, rng.choice([1, 33, 257, 700])))
        out.append((i, b, n, m, float(rng.choice([0.05, 0.6, 1.7, 6.0])), int(rng.choice([1, 5, 16, 32, 48, 64, 100])), int(rng.integers(0, 3))))
    return out


@pytest.mark.parametrize("seed,b,n,m,r,ns,kind", _bq_cases())
def test_ball_query_random_shapes(seed, b, n, m, r, ns, kind):
    rng = np.random.default_rng(1000 + seed)
    xyz = _cloud(rng, b, n, kind)
    new_xyz = xyz[:, rng.permutation(n)[:m]].copy()
    new_xyz[:, ::7] += rng.normal(0, r, new_xyz[:, ::7].shape).astype(np.float32)      # centres off the points; some empty balls
    want = oracle.ball_query(r, ns, xyz, new_xyz)
    for mode in (_lib.FPS_MODE_LATENCY, _lib.FPS_MODE_THROUGHPUT):                      # cluster build / one CTA per frame
        idx = torch.zeros((b, m, ns), dtype=torch.int32, device=DEV)
        with _lib.fps_mode(mode):
            ours.ball_query_wrapper(b, n, m, r, ns, T(new_xyz), T(xyz), idx)
        assert np.array_equal(idx.cpu().numpy(), want), (b, n, m, r, ns, kind, mode)
    # the same frames stacked with ragged tails
    cnt = np.asarray([n - 3 * i for i in range(b)], np.int32).clip(1)
    mcnt = np.asarray([max(1, m - i) for i in range(b)], np.int32)
    sx = np.concatenate([xyz[i, :cnt[i]] for i in range(b)])
    sq = np.concatenate([new_xyz[i, :mcnt[i]] for i in range(b)])
    sidx = torch.zeros((sq.shape[0], ns), dtype=torch.int32, device=DEV)
    stack.ball_query_wrapper(b, sq.shape[0], r, ns, T(sq), T(mcnt), T(sx), T(cnt), sidx)
    assert np.array_equal(sidx.cpu().numpy(), so.ball_query(r, ns, sx, cnt, sq, mcnt))


def _fps_cases():
    rng = np.random.default_rng(77)
    out = []
    for i in range(14):
        n = int(rng.choice([3, 600, 1024, 5000, 16384, 16385, 23000, 52000]))
        m = int(min(n, rng.choice([1, 2, 64, 300]) if n > 20000 else min(n, rng.choice([1, 17, 256, 1200]))))
        out.append((i, int(rng.integers(1, 4)), n, m, int(rng.integers(0, 3))))
    return out


@pytest.mark.parametrize("seed,b,n,m,kind", _fps_cases())
def test_fps_random_shapes(seed, b, n, m, kind):
    rng = np.random.default_rng(500 + seed)
    xyz = _cloud(rng, b, n, kind)
    want, want_t = oracle.fps(xyz, m, return_temp=True)
    for mode in (_lib.FPS_MODE_LATENCY, _lib.FPS_MODE_THROUGHPUT):
        temp = torch.full((b, n), 1e10, device=DEV)
        idx = torch.zeros((b, m), dtype=torch.int32, device=DEV)
        with _lib.fps_mode(mode):
            ours.farthest_point_sampling_wrapper(b, n, m, T(xyz), temp, idx)
        assert np.array_equal(idx.cpu().numpy(), want) and np.array_equal(temp.cpu().numpy(), want_t), (b, n, m, kind, mode)
    # stacked: ragged frames, the 1024-thread tie-break whatever the frame size
    cnt = np.asarray([max(1, n - 5 * i) for i in range(b)], np.int32)
    mcnt = np.asarray([min(int(cnt[i]), max(1, m - i)) for i in range(b)], np.int32)
    sx = np.concatenate([xyz[i, :cnt[i]] for i in range(b)])
    temp = torch.full((sx.shape[0],), 1e10, device=DEV)
    idx = torch.zeros((int(mcnt.sum()),), dtype=torch.int32, device=DEV)
    stack.stack_farthest_point_sampling_wrapper(T(sx), temp, T(cnt), idx, T(mcnt))
    swant, swant_t = so.fps(sx, cnt, mcnt, return_temp=True)
    assert np.array_equal(idx.cpu().numpy(), swant) and np.array_equal(temp.cpu().numpy(), swant_t)
