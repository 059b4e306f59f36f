"""ctypes binding of libpdmops.so (the C ABI declared in include/pdm_ops.h).

The product path has no CPU fallback: if the shared library is missing or does not
export a symbol, importing/using the ops raises.  (`python -m pdm_ssd_b200.build`
or `__graft_entry__.build()` produces the library; it is built in-tree.)
"""
import contextlib
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libpdmops.so")

_vp = ctypes.c_void_p
_i = ctypes.c_int
_f = ctypes.c_float

# name -> argtypes, mirrors include/pdm_ops.h one to one
SIGNATURES = {
    "pdm_set_fps_mode": [_i],
    "pdm_farthest_point_sampling": [_i, _i, _i, _vp, _vp, _vp, _vp],
    "pdm_farthest_point_sampling_ex": [_i, _i, _i, _vp, _vp, _vp, _i, _vp],
    "pdm_gather_points": [_i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "pdm_gather_points_grad": [_i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "pdm_ball_query": [_i, _i, _i, _f, _i, _vp, _vp, _vp, _vp],
    "pdm_ball_query_ex": [_i, _i, _i, _f, _i, _vp, _vp, _vp, _i, _vp],
    "pdm_group_points": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "pdm_group_points_grad": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "pdm_three_nn": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "pdm_three_interpolate": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "pdm_three_interpolate_grad": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "pdm_gather_points_grad_det": [_i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "pdm_group_points_grad_det": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "pdm_three_interpolate_grad_det": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "pdm_query_and_group": [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "pdm_sa_fused_forward": [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, ctypes.POINTER(_i), _vp, _vp, _vp, _vp],
    "pdm_sa_fused_forward_v2": [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, ctypes.POINTER(_i), _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "pdm_neck_forward": [_i, _i, _i, _vp, _vp, _vp, ctypes.POINTER(_f), ctypes.POINTER(_f),
                         ctypes.POINTER(_i), ctypes.POINTER(_i), _i, _f, _f, _vp, _vp, _vp, _vp],
    "pdm_boxes_iou_bev": [_i, _vp, _i, _vp, _vp, _vp],
    "pdm_boxes_overlap_bev": [_i, _vp, _i, _vp, _vp, _vp],
    "pdm_nms_bev_batched": [_i, _i, _vp, _vp, _f, _vp, _vp, _vp],
    "pdm_nms_normal_batched": [_i, _i, _vp, _vp, _f, _vp, _vp, _vp],
    "pdm_boxes_overlap_bev_paired": [_i, _vp, _vp, _vp, _vp],
    "pdm_neck_forward_split": [_i, _i, _i, _vp, _vp, _vp, ctypes.POINTER(_f), ctypes.POINTER(_f),
                               ctypes.POINTER(_i), ctypes.POINTER(_i), _i, _f, _f, _vp, _vp, _vp],
    "pdm_linear_rows": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "pdm_point_head_forward": [_i, _i, _i, _i, _i, _i, _i, _i, _i, ctypes.POINTER(_f), ctypes.POINTER(_f),
                               _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "pdm_sample_points": [_i, _i, _i, _i, ctypes.c_uint, _vp, _vp, _vp, _vp, _vp],
    "pdm_act_split_bytes": [_i, _i, _i, _i, ctypes.POINTER(ctypes.c_longlong)],
    "pdm_act_split_from_nchw": [_i, _i, _i, _i, _vp, _vp, _vp],
    "pdm_act_split_to_nchw": [_i, _i, _i, _i, _vp, _vp, _vp],
    "pdm_conv_tc_forward": [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp],
    # stacked family (pointnet2_stack)
    "pdm_stack_ball_query": [_i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "pdm_stack_voxel_query": [_i, _i, _i, _i, _i, _f, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "pdm_stack_farthest_point_sampling": [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "pdm_stack_group_points": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "pdm_stack_group_points_grad": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp],
    "pdm_stack_three_nn": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "pdm_stack_three_interpolate": [_i, _i, _vp, _vp, _vp, _vp, _vp],
    "pdm_stack_three_interpolate_grad": [_i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp],
    "pdm_stack_query_local_neighbor_idxs": [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _f, _i, _i, _vp],
    "pdm_stack_query_three_nn_by_local_idxs": [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "pdm_stack_vector_pool": [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _i,
                              _i, _i, _i, _i, _vp, _vp],
    "pdm_stack_vector_pool_grad": [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
}

_lib = None


class PdmOpsError(RuntimeError):
    pass


def load():
    """Load libpdmops.so once; raise loudly when it is absent or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise PdmOpsError(
            "libpdmops.so not found at %s -- build it with `python -m pdm_ssd_b200.build`; "
            "there is deliberately no CPU fallback" % SO_PATH)
    lib = ctypes.CDLL(SO_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.argtypes = args
        fn.restype = _i
    lib.pdm_abi_version.restype = _i
    lib.pdm_last_error.restype = ctypes.c_char_p
    lib.pdm_launch_count.restype = ctypes.c_longlong
    lib.pdm_reset_launch_count.restype = None
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().pdm_last_error().decode("utf-8", "replace")
        raise PdmOpsError("%s failed (code %d): %s" % (what, rc, msg))


FPS_MODE_AUTO, FPS_MODE_LATENCY, FPS_MODE_THROUGHPUT = 0, 1, 2


def set_fps_mode(mode):
    """Process-wide default scheduling hint for farthest point sampling (include/pdm_ops.h); results do not
    depend on it.  Prefer `fps_mode(...)`, which is per thread and passed per call."""
    check(load().pdm_set_fps_mode(int(mode)), "set_fps_mode")


_tls = threading.local()


def current_fps_mode():
    """Mode the calling thread asked for with `fps_mode(...)`, or None (= the process-wide default)."""
    return getattr(_tls, "fps_mode", None)


@contextlib.contextmanager
def fps_mode(mode):
    """`with _lib.fps_mode(_lib.FPS_MODE_THROUGHPUT): ...` -- every farthest-point-sampling call this thread
    makes inside the block (also while a CUDA graph is being captured) carries the mode as an argument
    (pdm_farthest_point_sampling_ex): no process-wide state, other threads / pipelines are unaffected."""
    prev = getattr(_tls, "fps_mode", None)
    _tls.fps_mode = int(mode)
    try:
        yield
    finally:
        _tls.fps_mode = prev


def launch_count():
    return int(load().pdm_launch_count())


def reset_launch_count():
    load().pdm_reset_launch_count()
