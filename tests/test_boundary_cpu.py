"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/*.h declares; the Python shim mirrors the reference's nine entry points and raises
(never exits, never falls back to a CPU path); host modules keep the reference's plugin API."""
import ctypes
import glob
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names += re.findall(r"\b(pdm_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_the_nine_ops():
    syms = _declared_symbols()
    for op in ("farthest_point_sampling", "gather_points", "gather_points_grad", "ball_query", "group_points",
               "group_points_grad", "three_nn", "three_interpolate", "three_interpolate_grad"):
        assert "pdm_" + op in syms


def test_library_exports_every_declared_symbol():
    from pdm_ssd_b200 import _lib
    assert os.path.exists(_lib.SO_PATH), "build with python -m pdm_ssd_b200.build"
    lib = ctypes.CDLL(_lib.SO_PATH)
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert _lib.load().pdm_abi_version() == 1
    assert set(_lib.SIGNATURES) <= set(_declared_symbols())


def test_shim_has_reference_names_and_arity():
    """pointnet2_api.cpp:10-24 names; positional arity of the C++ wrappers."""
    from pdm_ssd_b200 import pointnet2_batch_cuda as m
    want = {"ball_query_wrapper": 8, "group_points_wrapper": 8, "group_points_grad_wrapper": 8,
            "gather_points_wrapper": 7, "gather_points_grad_wrapper": 7, "farthest_point_sampling_wrapper": 6,
            "three_nn_wrapper": 7, "three_interpolate_wrapper": 8, "three_interpolate_grad_wrapper": 8}
    for name, arity in want.items():
        assert len(inspect.signature(getattr(m, name)).parameters) == arity, name


def test_no_cpu_fallback_cpu_tensors_raise():
    from pdm_ssd_b200 import pointnet2_batch_cuda as m, pointnet2_utils as pu
    x = torch.zeros(1, 8, 3)
    with pytest.raises(RuntimeError):
        m.farthest_point_sampling_wrapper(1, 8, 4, x, torch.zeros(1, 8), torch.zeros(1, 4, dtype=torch.int32))
    with pytest.raises(RuntimeError):
        m.ball_query_wrapper(1, 8, 2, 0.5, 4, x[:, :2], x, torch.zeros(1, 2, 4, dtype=torch.int32))
    with pytest.raises(Exception):
        pu.farthest_point_sample(x, 4)  # allocating on a CPU tensor's device cannot reach a kernel


def test_product_package_never_touches_the_oracle():
    for path in glob.glob(os.path.join(ROOT, "pdm_ssd_b200", "**", "*.*"), recursive=True):
        if path.endswith((".py", ".cu", ".cuh", ".h")):
            src = open(path).read()
            assert "import oracle" not in src and "pdm_oracle" not in src and "build_ref" not in src, path


def test_sa_module_plugin_api_and_state_dict_keys():
    from pdm_ssd_b200 import pointnet2_modules as M
    mlps = [[1, 16, 16, 32], [1, 32, 32, 64]]
    sa = M.PointnetSAModuleMSG(npoint=64, radii=[0.4, 0.8], nsamples=[8, 16], mlps=mlps, use_xyz=True)
    assert mlps[0][0] == 4  # in-place +3 like the reference (pointnet2_modules.py:87-88)
    keys = set(sa.state_dict().keys())
    for i in (0, 1):
        for k in (0, 3, 6):
            assert "mlps.%d.%d.weight" % (i, k) in keys
            assert "mlps.%d.%d.running_mean" % (i, k + 1) in keys
    assert sa.npoint == 64 and len(sa.groupers) == 2 and sa.pool_method == "max_pool"
    assert sa.groupers[0].radius == 0.4 and sa.groupers[1].nsample == 16
    single = M.PointnetSAModule(mlp=[3, 8], npoint=4, radius=1.0, nsample=2)
    assert isinstance(single, M.PointnetSAModuleMSG)
    fp = M.PointnetFPModule(mlp=[10, 20, 30])
    assert "mlp.0.weight" in fp.state_dict() and fp.mlp[0].weight.shape == (20, 10, 1, 1)
    ga = M.PointnetSAModuleMSG(npoint=None, radii=[None], nsamples=[None], mlps=[[2, 4]])
    xyz, f = torch.randn(2, 5, 3), torch.randn(2, 2, 5)
    new_xyz, out = ga(xyz, f)  # GroupAll path is pure torch: runs on CPU
    assert new_xyz is None and out.shape == (2, 4, 1)


def test_fp_module_without_known_points_runs_on_cpu():
    from pdm_ssd_b200 import pointnet2_modules as M
    fp = M.PointnetFPModule(mlp=[6, 8]).eval()
    out = fp(torch.randn(2, 7, 3), None, torch.randn(2, 2, 7), torch.randn(2, 4, 1))
    assert out.shape == (2, 8, 7)


def test_synthetic_frames_are_deterministic_and_kitti_shaped():
    import numpy as np
    from pdm_ssd_b200 import synthetic as S
    a, b = S.kitti_frame(1003), S.kitti_frame(1003)
    assert a.shape == (16384, 4) and a.dtype == np.float32 and np.array_equal(a, b)
    lo, hi = S.KITTI_RANGE[:3], S.KITTI_RANGE[3:]
    assert (a[:, :3] >= lo).all() and (a[:, :3] <= hi).all()
    assert len(np.unique(S.kitti_frame(1007)[:, :3], axis=0)) < 16384  # padded by duplication
    pts = S.to_pcdet_points(S.kitti_batch(2, 256))
    assert pts.shape == (512, 5) and (pts[:256, 0] == 0).all() and (pts[256:, 0] == 1).all()


def test_algorithmic_bytes_match_survey():
    from pdm_ssd_b200.sa_chain import algorithmic_bytes_per_frame
    ab = algorithmic_bytes_per_frame(16384)
    assert ab["sa1_fps"] == 212992 and ab["sa2_group_feat"] == 9568256
    assert ab["total"] == 14921728
