"""CPU tests of the stacked (pointnet2_stack) family: the C restatement oracle/pdm_stack_oracle.c against the golden
vectors the reference's own kernels produced on a B200 (tests/golden/stack_*.npz, generator make_golden_stack.py), the
shim's names / arity against pointnet2_stack/src/pointnet2_api.cpp:12-31, and the no-CPU-fallback rule."""
import glob
import inspect
import os

import numpy as np
import pytest
import torch

import stack_oracle as so

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = sorted(glob.glob(os.path.join(HERE, "golden", "stack_*_core.npz")))
GOLD_VP = sorted(glob.glob(os.path.join(HERE, "golden", "stack_*_vpool.npz")))


def _rows_sorted(a):
    a = np.asarray(a)
    return a[np.lexsort(a.T[::-1])]


def _lists(lst, sl):
    return [np.asarray(lst)[s:s + n].tolist() for s, n in np.asarray(sl)]


def test_golden_fixtures_are_committed():
    assert len(GOLD) >= 2 and len(GOLD_VP) >= 2


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_oracle_core_ops_vs_reference_golden(path):
    g = np.load(path)
    xyz, feat, cnt, mcnt = g["xyz"], g["feat"], g["cnt"], g["mcnt"]
    idx, temp = so.fps(xyz, cnt, mcnt, return_temp=True)
    assert np.array_equal(idx, g["fps_idx"]) and np.array_equal(temp, g["fps_temp"])
    assert np.array_equal(so.ball_query(1.0, 16, xyz, cnt, g["new_xyz"], mcnt), g["bq_r1_s16"])
    assert np.array_equal(so.ball_query(2.5, 32, xyz, cnt, g["new_xyz"], mcnt), g["bq_r2.5_s32"])
    gi = g["bq_r1_s16"].copy()
    gi[gi[:, 0] == -1] = 0
    assert np.array_equal(so.group_points(feat, cnt, gi, mcnt), g["group_out"])
    np.testing.assert_allclose(so.group_points_grad(g["grad_out"], gi, mcnt, cnt, xyz.shape[0]), g["grad_feat"], rtol=1e-5, atol=1e-5)
    d2, nn = so.three_nn(xyz, cnt, xyz[g["fps_idx"]], mcnt)
    assert np.array_equal(nn, g["nn_idx"]) and np.array_equal(d2, g["nn_dist2"])
    assert np.array_equal(so.three_interpolate(g["known_feat"], g["nn_idx"], g["weight"]), g["interp"])
    np.testing.assert_allclose(so.three_interpolate_grad(g["grad_interp"], g["nn_idx"], g["weight"], int(mcnt.sum())),
                               g["grad_known"], rtol=1e-5, atol=2e-5)
    assert np.array_equal(so.voxel_query((2, 2, 2), 1.6, 16, xyz, g["new_xyz"], g["new_coords"], g["point_indices"]), g["vq_idx"])


@pytest.mark.parametrize("path", GOLD_VP, ids=[os.path.basename(p) for p in GOLD_VP])
def test_oracle_vector_pool_family_vs_reference_golden(path):
    g = np.load(path)
    xyz, feat, cnt, mcnt, q = g["xyz"], g["feat"], g["cnt"], g["mcnt"], g["new_xyz"]
    for name in ("cube_avg", "ball_avg_ns", "cube_first"):
        gx, gy, gz, ceg, ns, ntype, ptype = g[name + "_cfg"].tolist()
        n_ent = g[name + "_grp"].shape[0]
        nf, nl, pc, grp, cum = so.vector_pool(xyz, cnt, feat, q, mcnt, (gx, gy, gz), float(g[name + "_dmax"]), ceg, True,
                                              n_ent + 3, ns, ntype, ptype)
        assert cum == n_ent
        assert np.array_equal(pc, g[name + "_pc"]) and np.array_equal(nf, g[name + "_nf"]) and np.array_equal(nl, g[name + "_nl"])
        assert np.array_equal(_rows_sorted(grp[:cum]), _rows_sorted(g[name + "_grp"]))
        np.testing.assert_allclose(so.vector_pool_grad(g[name + "_gnf"], pc, grp[:cum], xyz.shape[0], 8), g[name + "_gsf"],
                                   rtol=1e-5, atol=1e-5)
    for name in ("ln_cube", "ln_ball_ns"):
        ns, ntype, avg, tot = g[name + "_cfg"].tolist()
        lst, sl, total = so.local_neighbor_idxs(xyz, cnt, q, mcnt, avg, float(g[name + "_dmax"]), ns, ntype)
        assert total == tot
        assert _lists(lst, sl) == _lists(g[name + "_list"], g[name + "_start_len"])
        d2, gi = so.three_nn_local(xyz, g[name + "_centers"], g[name + "_list"], g[name + "_start_len"])
        assert np.array_equal(gi, g[name + "_gidx"]) and np.array_equal(d2, g[name + "_gd2"])


def test_oracle_stack_fps_equals_batch_fps_on_uniform_1024_frames():
    """with n a multiple of 1024 >= 1024 both families use the 1024-thread tournament: same picks, global rows"""
    import oracle
    rng = np.random.default_rng(3)
    xyz = rng.uniform(0, 20, (2, 2048, 3)).astype(np.float32)
    xyz[1, 1024:] = xyz[1, :1024]      # duplicates: ties
    b = oracle.fps(xyz, 300)
    s = so.fps(xyz.reshape(-1, 3), [2048, 2048], [300, 300])
    assert np.array_equal(s.reshape(2, 300), b + np.asarray([[0], [2048]], np.int32))


def test_shim_has_reference_names_and_arity():
    """pointnet2_stack/src/pointnet2_api.cpp:12-31 names; positional arity of the C++ wrappers"""
    from pdm_ssd_b200 import pointnet2_stack_cuda as m
    want = {"ball_query_wrapper": 9, "voxel_query_wrapper": 14, "farthest_point_sampling_wrapper": 6,
            "stack_farthest_point_sampling_wrapper": 5, "group_points_wrapper": 9, "group_points_grad_wrapper": 10,
            "three_nn_wrapper": 6, "three_interpolate_wrapper": 4, "three_interpolate_grad_wrapper": 4,
            "query_stacked_local_neighbor_idxs_wrapper_stack": 11, "query_three_nn_by_stacked_local_idxs_wrapper_stack": 9,
            "vector_pool_wrapper": 18, "vector_pool_grad_wrapper": 4}
    for name, arity in want.items():
        assert len(inspect.signature(getattr(m, name)).parameters) == arity, name


def test_no_cpu_fallback_cpu_tensors_raise():
    from pdm_ssd_b200 import pointnet2_stack_cuda as m
    x = torch.zeros(8, 3)
    cnt = torch.tensor([8], dtype=torch.int32)
    with pytest.raises(RuntimeError):
        m.ball_query_wrapper(1, 2, 0.5, 4, x[:2], torch.tensor([2], dtype=torch.int32), x, cnt, torch.zeros(2, 4, dtype=torch.int32))
    with pytest.raises(RuntimeError):
        m.stack_farthest_point_sampling_wrapper(x, torch.zeros(8), cnt, torch.zeros(4, dtype=torch.int32), torch.tensor([4], dtype=torch.int32))
    with pytest.raises(RuntimeError):
        m.three_interpolate_wrapper(torch.zeros(4, 8), torch.zeros(8, 3, dtype=torch.int32), torch.zeros(8, 3), torch.zeros(8, 8))
