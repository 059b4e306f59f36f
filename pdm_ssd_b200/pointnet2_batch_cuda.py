"""Drop-in for the reference extension module `pointnet2_batch_cuda`.

Same nine function names and positional signatures as the pybind module the reference
builds from pcdet/ops/pointnet2/pointnet2_batch/src/pointnet2_api.cpp:10-24, so the
reference's pointnet2_utils.py runs on top of it unchanged (`from . import
pointnet2_batch_cuda as pointnet2`).  Each call validates its tensors, then forwards raw
device pointers and torch's CURRENT stream to the C ABI in libpdmops.so.

Differences by design: launches go to the current stream instead of the
legacy default stream, and failures raise RuntimeError instead of exit(-1).
"""
import torch

from . import _lib

_F32, _I32 = torch.float32, torch.int32


def _chk(t, name, dtype):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if t.dtype != dtype:
        raise RuntimeError("%s must have dtype %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)
    return t.data_ptr()


def _need(t, name, numel):
    if t.numel() < numel:
        raise RuntimeError("%s has %d elements, the call needs %d" % (name, t.numel(), numel))


def _stream(ref):
    return torch.cuda.current_stream(ref.device).cuda_stream


def _deterministic():
    """The three gradient entries switch to the in-order (sorted, atomics-free) kernels of csrc/det_backward.cu when the
    user asked torch for deterministic algorithms; the default matches the reference (atomicAdd, order unspecified)."""
    return torch.are_deterministic_algorithms_enabled()


def farthest_point_sampling_wrapper(b, n, m, points_tensor, temp_tensor, idx_tensor):
    lib = _lib.load()
    p = _chk(points_tensor, "points", _F32)
    t = _chk(temp_tensor, "temp", _F32)
    i = _chk(idx_tensor, "idx", _I32)
    _need(points_tensor, "points", b * n * 3); _need(temp_tensor, "temp", b * n); _need(idx_tensor, "idx", b * m)
    mode = _lib.current_fps_mode()          # per-thread scheduling hint (`with _lib.fps_mode(...)`), passed per call
    with torch.cuda.device(points_tensor.device):
        if mode is None:
            rc = lib.pdm_farthest_point_sampling(b, n, m, p, t, i, _stream(points_tensor))
        else:
            rc = lib.pdm_farthest_point_sampling_ex(b, n, m, p, t, i, mode, _stream(points_tensor))
        _lib.check(rc, "farthest_point_sampling")
    return 1


def gather_points_wrapper(b, c, n, npoints, points_tensor, idx_tensor, out_tensor):
    lib = _lib.load()
    p = _chk(points_tensor, "points", _F32)
    i = _chk(idx_tensor, "idx", _I32)
    o = _chk(out_tensor, "out", _F32)
    _need(points_tensor, "points", b * c * n); _need(idx_tensor, "idx", b * npoints); _need(out_tensor, "out", b * c * npoints)
    with torch.cuda.device(points_tensor.device):
        _lib.check(lib.pdm_gather_points(b, c, n, npoints, p, i, o, _stream(points_tensor)), "gather_points")
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out_tensor, idx_tensor, grad_points_tensor):
    lib = _lib.load()
    g = _chk(grad_out_tensor, "grad_out", _F32)
    i = _chk(idx_tensor, "idx", _I32)
    o = _chk(grad_points_tensor, "grad_points", _F32)
    _need(grad_out_tensor, "grad_out", b * c * npoints); _need(idx_tensor, "idx", b * npoints); _need(grad_points_tensor, "grad_points", b * c * n)
    fn = lib.pdm_gather_points_grad_det if _deterministic() else lib.pdm_gather_points_grad
    with torch.cuda.device(grad_out_tensor.device):
        _lib.check(fn(b, c, n, npoints, g, i, o, _stream(grad_out_tensor)), "gather_points_grad")
    return 1


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz_tensor, xyz_tensor, idx_tensor):
    lib = _lib.load()
    q = _chk(new_xyz_tensor, "new_xyz", _F32)
    x = _chk(xyz_tensor, "xyz", _F32)
    i = _chk(idx_tensor, "idx", _I32)
    _need(new_xyz_tensor, "new_xyz", b * m * 3); _need(xyz_tensor, "xyz", b * n * 3); _need(idx_tensor, "idx", b * m * nsample)
    mode = _lib.current_fps_mode()          # per-thread scheduling hint (`with _lib.fps_mode(...)`), passed per call
    with torch.cuda.device(xyz_tensor.device):
        if mode is None:
            rc = lib.pdm_ball_query(b, n, m, float(radius), nsample, q, x, i, _stream(xyz_tensor))
        else:
            rc = lib.pdm_ball_query_ex(b, n, m, float(radius), nsample, q, x, i, mode, _stream(xyz_tensor))
        _lib.check(rc, "ball_query")
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points_tensor, idx_tensor, out_tensor):
    lib = _lib.load()
    p = _chk(points_tensor, "points", _F32)
    i = _chk(idx_tensor, "idx", _I32)
    o = _chk(out_tensor, "out", _F32)
    _need(points_tensor, "points", b * c * n); _need(idx_tensor, "idx", b * npoints * nsample)
    _need(out_tensor, "out", b * c * npoints * nsample)
    with torch.cuda.device(points_tensor.device):
        _lib.check(lib.pdm_group_points(b, c, n, npoints, nsample, p, i, o, _stream(points_tensor)), "group_points")
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out_tensor, idx_tensor, grad_points_tensor):
    lib = _lib.load()
    g = _chk(grad_out_tensor, "grad_out", _F32)
    i = _chk(idx_tensor, "idx", _I32)
    o = _chk(grad_points_tensor, "grad_points", _F32)
    _need(grad_out_tensor, "grad_out", b * c * npoints * nsample); _need(idx_tensor, "idx", b * npoints * nsample)
    _need(grad_points_tensor, "grad_points", b * c * n)
    fn = lib.pdm_group_points_grad_det if _deterministic() else lib.pdm_group_points_grad
    with torch.cuda.device(grad_out_tensor.device):
        _lib.check(fn(b, c, n, npoints, nsample, g, i, o, _stream(grad_out_tensor)), "group_points_grad")
    return 1


def three_nn_wrapper(b, n, m, unknown_tensor, known_tensor, dist2_tensor, idx_tensor):
    lib = _lib.load()
    u = _chk(unknown_tensor, "unknown", _F32)
    k = _chk(known_tensor, "known", _F32)
    d = _chk(dist2_tensor, "dist2", _F32)
    i = _chk(idx_tensor, "idx", _I32)
    _need(unknown_tensor, "unknown", b * n * 3); _need(known_tensor, "known", b * m * 3)
    _need(dist2_tensor, "dist2", b * n * 3); _need(idx_tensor, "idx", b * n * 3)
    with torch.cuda.device(unknown_tensor.device):
        _lib.check(lib.pdm_three_nn(b, n, m, u, k, d, i, _stream(unknown_tensor)), "three_nn")


def three_interpolate_wrapper(b, c, m, n, points_tensor, idx_tensor, weight_tensor, out_tensor):
    lib = _lib.load()
    p = _chk(points_tensor, "points", _F32)
    i = _chk(idx_tensor, "idx", _I32)
    w = _chk(weight_tensor, "weight", _F32)
    o = _chk(out_tensor, "out", _F32)
    _need(points_tensor, "points", b * c * m); _need(idx_tensor, "idx", b * n * 3)
    _need(weight_tensor, "weight", b * n * 3); _need(out_tensor, "out", b * c * n)
    with torch.cuda.device(points_tensor.device):
        _lib.check(lib.pdm_three_interpolate(b, c, m, n, p, i, w, o, _stream(points_tensor)), "three_interpolate")


def three_interpolate_grad_wrapper(b, c, n, m, grad_out_tensor, idx_tensor, weight_tensor, grad_points_tensor):
    lib = _lib.load()
    g = _chk(grad_out_tensor, "grad_out", _F32)
    i = _chk(idx_tensor, "idx", _I32)
    w = _chk(weight_tensor, "weight", _F32)
    o = _chk(grad_points_tensor, "grad_points", _F32)
    _need(grad_out_tensor, "grad_out", b * c * n); _need(idx_tensor, "idx", b * n * 3)
    _need(weight_tensor, "weight", b * n * 3); _need(grad_points_tensor, "grad_points", b * c * m)
    fn = lib.pdm_three_interpolate_grad_det if _deterministic() else lib.pdm_three_interpolate_grad
    with torch.cuda.device(grad_out_tensor.device):
        _lib.check(fn(b, c, n, m, g, i, w, o, _stream(grad_out_tensor)), "three_interpolate_grad")
