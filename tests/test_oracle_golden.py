"""CPU tests: the C oracle against the golden vectors produced by the REFERENCE's own CUDA
kernels on a B200 (tests/golden/make_golden.py).  This is what pins the oracle."""
import glob
import os

import numpy as np
import pytest

import oracle

G = os.path.join(os.path.dirname(__file__), "golden")


def test_golden_files_present():
    assert len(glob.glob(os.path.join(G, "fps_*.npz"))) >= 9
    assert len(glob.glob(os.path.join(G, "bq_*.npz"))) >= 4
    for f in ("interp_n500_m77.npz", "three_nn_m2.npz", "group_gather.npz"):
        assert os.path.exists(os.path.join(G, f))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(G, "fps_*.npz"))), ids=os.path.basename)
def test_fps_matches_reference_kernels(path):
    g = np.load(path)
    idx, temp = oracle.fps(g["xyz"], int(g["m"]), return_temp=True)
    assert np.array_equal(idx, g["idx"])
    assert np.array_equal(temp, g["temp"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(G, "bq_*.npz"))), ids=os.path.basename)
def test_ball_query_matches_reference_kernels(path):
    g = np.load(path)
    got = oracle.ball_query(float(g["radius"]), int(g["nsample"]), g["xyz"], g["new_xyz"])
    assert np.array_equal(got, g["idx"])


def test_boundary_case_is_meaningful():
    g = np.load(os.path.join(G, "bq_boundary_n64.npz"))
    xyz, r = g["xyz"][0], np.float32(g["radius"])
    hits = set(np.unique(g["idx"][0, 0]).tolist())
    on_sphere = np.flatnonzero((xyz[:, 0] == r) & (xyz[:, 1] == 0) & (xyz[:, 2] == 0))
    one_ulp_in = np.flatnonzero((xyz[:, 0] == np.nextafter(r, np.float32(0))) & (xyz[:, 1] == 0))
    one_ulp_out = np.flatnonzero(xyz[:, 1] == np.nextafter(r, np.float32(1)))
    assert len(on_sphere) == 16 and len(one_ulp_in) == 16 and len(one_ulp_out) == 16
    assert not hits & set(on_sphere.tolist())      # strict <: distance exactly r is outside
    assert not hits & set(one_ulp_out.tolist())
    assert set(one_ulp_in.tolist()) <= hits        # one ulp inside is inside
    e = np.load(os.path.join(G, "bq_empty_rows_n200.npz"))
    assert (e["idx"][:, 5:] == 0).all() and (e["idx"][:, :5] >= 0).all()


def test_interpolation_matches_reference_kernels():
    g = np.load(os.path.join(G, "interp_n500_m77.npz"))
    d2, idx = oracle.three_nn(g["unknown"], g["known"])
    assert np.array_equal(idx, g["idx"]) and np.array_equal(d2, g["dist2"])
    assert np.array_equal(oracle.three_interpolate(g["feats"], g["idx"], g["weight"]), g["out"])
    g2 = np.load(os.path.join(G, "three_nn_m2.npz"))
    d2, idx = oracle.three_nn(g2["unknown"], g2["known"])
    assert np.array_equal(idx, g2["idx"]) and np.array_equal(d2, g2["dist2"])
    assert np.isinf(d2[..., 2]).all()  # 1e40 narrowed to +inf, as in the kernel


def test_group_gather_matches_reference_kernels():
    g = np.load(os.path.join(G, "group_gather.npz"))
    assert np.array_equal(oracle.group_points(g["points"], g["idx"]), g["grouped"])
    assert np.array_equal(oracle.gather_points(g["points"], g["idx"][:, :, 0].copy()), g["gathered"])


def test_thread_count_does_not_change_results():
    rng = np.random.default_rng(1)
    xyz = rng.uniform(0, 10, (4, 700, 3)).astype(np.float32)
    oracle.set_threads(1)
    a = oracle.fps(xyz, 100)
    qa = oracle.ball_query(1.0, 16, xyz, xyz[:, :50].copy())
    oracle.set_threads(4)
    assert np.array_equal(a, oracle.fps(xyz, 100))
    assert np.array_equal(qa, oracle.ball_query(1.0, 16, xyz, xyz[:, :50].copy()))
    oracle.set_threads(1)


def test_opt_n_threads_matches_reference_host_formula():
    # cuda_utils.h:10-14
    for n, want in [(1, 1), (2, 2), (3, 2), (37, 32), (300, 256), (1000, 512), (1024, 1024), (4096, 1024), (16384, 1024)]:
        assert oracle.opt_n_threads(n) == want


def test_fps_tiebreak_theory():
    """The decomposition-free key our CUDA kernels use (DESIGN.md 'FPS tie-break'):
    argmax == smallest (bitreverse_p(k mod bs), k div bs) among equal maxima.  Checked against
    the literal tournament simulation in the oracle on inputs made of duplicates."""
    rng = np.random.default_rng(3)
    for n in (64, 300, 1024, 2500):
        base = rng.integers(0, 6, (1, n, 3)).astype(np.float32)  # tiny lattice: ties everywhere
        m = min(n, 200)
        want = oracle.fps(base, m)[0]
        bs = oracle.opt_n_threads(n)
        p = bs.bit_length() - 1
        temp = np.full(n, 1e10, np.float32)
        k = np.arange(n)
        rev = np.array([int(format(int(v), "0%db" % p)[::-1], 2) if p else 0 for v in (k % bs)])
        tiekey = rev.astype(np.int64) * (1 << 32) + k // bs
        got = [0]
        for _ in range(1, m):
            c = base[0, got[-1]]
            d = base[0] - c
            dist = (np.float32(d[:, 1] * d[:, 1]) + d[:, 0] * d[:, 0] + d[:, 2] * d[:, 2]).astype(np.float32)  # exact on small ints
            temp = np.minimum(temp, dist)
            cand = np.flatnonzero(temp == temp.max())
            got.append(int(cand[np.argmin(tiekey[cand])]))
        assert np.array_equal(np.array(got), want)
