"""ctypes front-end of oracle/libpdm_oracle.so (numpy in / numpy out).

TEST INFRASTRUCTURE ONLY -- see the header of pdm_oracle.c.  The product package
never imports this module; tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs do.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpdm_oracle.so")
_lib = None

_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    srcs = [os.path.join(_HERE, "pdm_oracle.c"), os.path.join(_HERE, "pdm_stack_oracle.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpdm_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_opt_n_threads.restype = ctypes.c_int
        _lib.oracle_max_threads.restype = ctypes.c_int
    return _lib


def set_threads(n):
    lib().oracle_set_threads(ctypes.c_int(int(n)))


def _fp(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f)


def _ip(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_i)


def opt_n_threads(n):
    return int(lib().oracle_opt_n_threads(ctypes.c_int(n)))


def fps(xyz, npoint, return_temp=False):
    """xyz (B,N,3) f32 -> idx (B,npoint) i32 [, temp (B,N) f32]"""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    B, N, _ = xyz.shape
    temp = np.full((B, N), 1e10, dtype=np.float32)
    idx = np.zeros((B, npoint), dtype=np.int32)
    lib().oracle_fps(B, N, npoint, _fp(xyz), _fp(temp), _ip(idx))
    return (idx, temp) if return_temp else idx


def gather_points(points, idx):
    points = np.ascontiguousarray(points, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    B, C, N = points.shape
    M = idx.shape[1]
    out = np.empty((B, C, M), dtype=np.float32)
    lib().oracle_gather_points(B, C, N, M, _fp(points), _ip(idx), _fp(out))
    return out


def gather_points_grad(grad_out, idx, N):
    grad_out = np.ascontiguousarray(grad_out, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    B, C, M = grad_out.shape
    gp = np.zeros((B, C, N), dtype=np.float32)
    lib().oracle_gather_points_grad(B, C, N, M, _fp(grad_out), _ip(idx), _fp(gp))
    return gp


def ball_query(radius, nsample, xyz, new_xyz):
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    new_xyz = np.ascontiguousarray(new_xyz, dtype=np.float32)
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = np.zeros((B, M, nsample), dtype=np.int32)
    lib().oracle_ball_query(B, N, M, ctypes.c_float(radius), nsample, _fp(new_xyz), _fp(xyz), _ip(idx))
    return idx


def group_points(points, idx):
    points = np.ascontiguousarray(points, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    B, C, N = points.shape
    _, M, S = idx.shape
    out = np.empty((B, C, M, S), dtype=np.float32)
    lib().oracle_group_points(B, C, N, M, S, _fp(points), _ip(idx), _fp(out))
    return out


def group_points_grad(grad_out, idx, N):
    grad_out = np.ascontiguousarray(grad_out, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    B, C, M, S = grad_out.shape
    gp = np.zeros((B, C, N), dtype=np.float32)
    lib().oracle_group_points_grad(B, C, N, M, S, _fp(grad_out), _ip(idx), _fp(gp))
    return gp


def three_nn(unknown, known):
    """returns (dist2, idx): the SQUARED distances the kernel stores (the Python wrapper
    of the reference takes sqrt afterwards, pointnet2_utils.py:97)."""
    unknown = np.ascontiguousarray(unknown, dtype=np.float32)
    known = np.ascontiguousarray(known, dtype=np.float32)
    B, N, _ = unknown.shape
    M = known.shape[1]
    d2 = np.empty((B, N, 3), dtype=np.float32)
    idx = np.empty((B, N, 3), dtype=np.int32)
    lib().oracle_three_nn(B, N, M, _fp(unknown), _fp(known), _fp(d2), _ip(idx))
    return d2, idx


def three_interpolate(points, idx, weight):
    points = np.ascontiguousarray(points, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    weight = np.ascontiguousarray(weight, dtype=np.float32)
    B, C, M = points.shape
    N = idx.shape[1]
    out = np.empty((B, C, N), dtype=np.float32)
    lib().oracle_three_interpolate(B, C, M, N, _fp(points), _ip(idx), _fp(weight), _fp(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, M):
    grad_out = np.ascontiguousarray(grad_out, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    weight = np.ascontiguousarray(weight, dtype=np.float32)
    B, C, N = grad_out.shape
    gp = np.zeros((B, C, M), dtype=np.float32)
    lib().oracle_three_interpolate_grad(B, C, N, M, _fp(grad_out), _ip(idx), _fp(weight), _fp(gp))
    return gp


def boxes_iou_bev(boxes_a, boxes_b):
    """(Na,7), (Nb,7) [x,y,z,dx,dy,dz,heading] -> (Na,Nb) rotated BEV IoU"""
    a = np.ascontiguousarray(boxes_a, dtype=np.float32)
    b = np.ascontiguousarray(boxes_b, dtype=np.float32)
    out = np.empty((len(a), len(b)), dtype=np.float32)
    lib().oracle_boxes_iou_bev(len(a), _fp(a), len(b), _fp(b), _fp(out))
    return out


def nms_bev(boxes_sorted, thresh):
    """boxes (N,7) already sorted by descending score -> kept positions (ascending)"""
    b = np.ascontiguousarray(boxes_sorted, dtype=np.float32)
    keep = np.empty(len(b), dtype=np.int32)
    fn = lib().oracle_nms_bev
    fn.restype = ctypes.c_int
    n = fn(len(b), _fp(b), ctypes.c_float(thresh), _ip(keep))
    return keep[:n].copy()
