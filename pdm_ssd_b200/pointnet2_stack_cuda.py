"""Drop-in for the reference extension module `pointnet2_stack_cuda` (the stacked / ragged-batch operator family).

Same fifteen function names and positional signatures as the pybind module the reference builds from
pcdet/ops/pointnet2/pointnet2_stack/src/pointnet2_api.cpp:12-31, so the reference's own
pointnet2_stack/{pointnet2_utils,pointnet2_modules,voxel_query_utils,voxel_pool_modules}.py run on top of it
unchanged (`from . import pointnet2_stack_cuda as pointnet2`, pointnet2_utils.py:5).  Each call validates its
tensors, then forwards raw device pointers and torch's CURRENT stream to the C ABI in libpdmops.so.

Differences by design: launches go to the current stream instead of the legacy default stream,
failures raise RuntimeError instead of exit(-1), and the two gradient entries use the in-order (atomics-free)
kernels when torch.use_deterministic_algorithms(True) is set.
"""
import torch

from . import _lib
from .pointnet2_batch_cuda import _F32, _I32, _chk, _deterministic, _need, _stream
from .pointnet2_batch_cuda import farthest_point_sampling_wrapper  # noqa: F401  (pointnet2_api.cpp:16: the batch kernel)


def ball_query_wrapper(B, M, radius, nsample, new_xyz_tensor, new_xyz_batch_cnt_tensor, xyz_tensor, xyz_batch_cnt_tensor,
                       idx_tensor):
    lib = _lib.load()
    q = _chk(new_xyz_tensor, "new_xyz", _F32)
    qc = _chk(new_xyz_batch_cnt_tensor, "new_xyz_batch_cnt", _I32)
    x = _chk(xyz_tensor, "xyz", _F32)
    xc = _chk(xyz_batch_cnt_tensor, "xyz_batch_cnt", _I32)
    i = _chk(idx_tensor, "idx", _I32)
    n_total = xyz_tensor.shape[0]
    _need(new_xyz_tensor, "new_xyz", M * 3); _need(idx_tensor, "idx", M * nsample)
    _need(new_xyz_batch_cnt_tensor, "new_xyz_batch_cnt", B); _need(xyz_batch_cnt_tensor, "xyz_batch_cnt", B)
    with torch.cuda.device(xyz_tensor.device):
        _lib.check(lib.pdm_stack_ball_query(B, M, n_total, float(radius), nsample, q, qc, x, xc, i, _stream(xyz_tensor)),
                   "stack_ball_query")
    return 1


def voxel_query_wrapper(M, R1, R2, R3, nsample, radius, z_range, y_range, x_range, new_xyz_tensor, xyz_tensor,
                        new_coords_tensor, point_indices_tensor, idx_tensor):
    lib = _lib.load()
    q = _chk(new_xyz_tensor, "new_xyz", _F32)
    x = _chk(xyz_tensor, "xyz", _F32)
    co = _chk(new_coords_tensor, "new_coords", _I32)
    pi = _chk(point_indices_tensor, "point_indices", _I32)
    i = _chk(idx_tensor, "idx", _I32)
    _need(new_xyz_tensor, "new_xyz", M * 3); _need(new_coords_tensor, "new_coords", M * 4); _need(idx_tensor, "idx", M * nsample)
    if point_indices_tensor.numel() % max(R1 * R2 * R3, 1) != 0:
        raise RuntimeError("point_indices does not have the shape (B, %d, %d, %d)" % (R1, R2, R3))
    with torch.cuda.device(xyz_tensor.device):
        _lib.check(lib.pdm_stack_voxel_query(M, R1, R2, R3, nsample, float(radius), z_range, y_range, x_range, q, x, co, pi, i,
                                             _stream(xyz_tensor)), "stack_voxel_query")
    return 1


def stack_farthest_point_sampling_wrapper(points_tensor, temp_tensor, xyz_batch_cnt_tensor, idx_tensor,
                                          num_sampled_points_tensor):
    lib = _lib.load()
    p = _chk(points_tensor, "points", _F32)
    t = _chk(temp_tensor, "temp", _F32)
    xc = _chk(xyz_batch_cnt_tensor, "xyz_batch_cnt", _I32)
    i = _chk(idx_tensor, "idx", _I32)
    ns = _chk(num_sampled_points_tensor, "num_sampled_points", _I32)
    batch, n = xyz_batch_cnt_tensor.shape[0], points_tensor.shape[0]
    _need(temp_tensor, "temp", n); _need(num_sampled_points_tensor, "num_sampled_points", batch)
    with torch.cuda.device(points_tensor.device):
        _lib.check(lib.pdm_stack_farthest_point_sampling(n, batch, p, t, xc, i, ns, _stream(points_tensor)),
                   "stack_farthest_point_sampling")
    return 1


def group_points_wrapper(B, M, C, nsample, features_tensor, features_batch_cnt_tensor, idx_tensor, idx_batch_cnt_tensor,
                         out_tensor):
    lib = _lib.load()
    f = _chk(features_tensor, "features", _F32)
    fc = _chk(features_batch_cnt_tensor, "features_batch_cnt", _I32)
    i = _chk(idx_tensor, "idx", _I32)
    ic = _chk(idx_batch_cnt_tensor, "idx_batch_cnt", _I32)
    o = _chk(out_tensor, "out", _F32)
    _need(idx_tensor, "idx", M * nsample); _need(out_tensor, "out", M * C * nsample)
    _need(features_batch_cnt_tensor, "features_batch_cnt", B); _need(idx_batch_cnt_tensor, "idx_batch_cnt", B)
    with torch.cuda.device(features_tensor.device):
        _lib.check(lib.pdm_stack_group_points(B, M, C, nsample, f, fc, i, ic, o, _stream(features_tensor)), "stack_group_points")
    return 1


def group_points_grad_wrapper(B, M, C, N, nsample, grad_out_tensor, idx_tensor, idx_batch_cnt_tensor,
                              features_batch_cnt_tensor, grad_features_tensor):
    lib = _lib.load()
    g = _chk(grad_out_tensor, "grad_out", _F32)
    i = _chk(idx_tensor, "idx", _I32)
    ic = _chk(idx_batch_cnt_tensor, "idx_batch_cnt", _I32)
    fc = _chk(features_batch_cnt_tensor, "features_batch_cnt", _I32)
    o = _chk(grad_features_tensor, "grad_features", _F32)
    _need(grad_out_tensor, "grad_out", M * C * nsample); _need(idx_tensor, "idx", M * nsample)
    _need(grad_features_tensor, "grad_features", N * C)
    with torch.cuda.device(grad_out_tensor.device):
        _lib.check(lib.pdm_stack_group_points_grad(B, M, C, N, nsample, g, i, ic, fc, o, int(_deterministic()),
                                                   _stream(grad_out_tensor)), "stack_group_points_grad")
    return 1


def three_nn_wrapper(unknown_tensor, unknown_batch_cnt_tensor, known_tensor, known_batch_cnt_tensor, dist2_tensor, idx_tensor):
    lib = _lib.load()
    u = _chk(unknown_tensor, "unknown", _F32)
    uc = _chk(unknown_batch_cnt_tensor, "unknown_batch_cnt", _I32)
    k = _chk(known_tensor, "known", _F32)
    kc = _chk(known_batch_cnt_tensor, "known_batch_cnt", _I32)
    d = _chk(dist2_tensor, "dist2", _F32)
    i = _chk(idx_tensor, "idx", _I32)
    b, n, m = unknown_batch_cnt_tensor.shape[0], unknown_tensor.shape[0], known_tensor.shape[0]
    _need(known_batch_cnt_tensor, "known_batch_cnt", b); _need(dist2_tensor, "dist2", n * 3); _need(idx_tensor, "idx", n * 3)
    with torch.cuda.device(unknown_tensor.device):
        _lib.check(lib.pdm_stack_three_nn(b, n, m, u, uc, k, kc, d, i, _stream(unknown_tensor)), "stack_three_nn")


def three_interpolate_wrapper(features_tensor, idx_tensor, weight_tensor, out_tensor):
    lib = _lib.load()
    f = _chk(features_tensor, "features", _F32)
    i = _chk(idx_tensor, "idx", _I32)
    w = _chk(weight_tensor, "weight", _F32)
    o = _chk(out_tensor, "out", _F32)
    n, c = idx_tensor.shape[0], features_tensor.shape[1]
    _need(weight_tensor, "weight", n * 3); _need(out_tensor, "out", n * c)
    with torch.cuda.device(features_tensor.device):
        _lib.check(lib.pdm_stack_three_interpolate(n, c, f, i, w, o, _stream(features_tensor)), "stack_three_interpolate")


def three_interpolate_grad_wrapper(grad_out_tensor, idx_tensor, weight_tensor, grad_features_tensor):
    lib = _lib.load()
    g = _chk(grad_out_tensor, "grad_out", _F32)
    i = _chk(idx_tensor, "idx", _I32)
    w = _chk(weight_tensor, "weight", _F32)
    o = _chk(grad_features_tensor, "grad_features", _F32)
    n, c = grad_out_tensor.shape[0], grad_out_tensor.shape[1]
    m = grad_features_tensor.shape[0]
    _need(idx_tensor, "idx", n * 3); _need(weight_tensor, "weight", n * 3); _need(grad_features_tensor, "grad_features", m * c)
    with torch.cuda.device(grad_out_tensor.device):
        _lib.check(lib.pdm_stack_three_interpolate_grad(n, c, m, g, i, w, o, int(_deterministic()), _stream(grad_out_tensor)),
                   "stack_three_interpolate_grad")


def query_stacked_local_neighbor_idxs_wrapper_stack(support_xyz_tensor, xyz_batch_cnt_tensor, new_xyz_tensor,
                                                    new_xyz_batch_cnt_tensor, stack_neighbor_idxs_tensor, start_len_tensor,
                                                    cumsum_tensor, avg_length_of_neighbor_idxs, max_neighbour_distance,
                                                    nsample, neighbor_type):
    lib = _lib.load()
    x = _chk(support_xyz_tensor, "support_xyz", _F32)
    xc = _chk(xyz_batch_cnt_tensor, "xyz_batch_cnt", _I32)
    q = _chk(new_xyz_tensor, "new_xyz", _F32)
    qc = _chk(new_xyz_batch_cnt_tensor, "new_xyz_batch_cnt", _I32)
    sn = _chk(stack_neighbor_idxs_tensor, "stack_neighbor_idxs", _I32)
    sl = _chk(start_len_tensor, "start_len", _I32)
    cs = _chk(cumsum_tensor, "cumsum", _I32)
    b, m = xyz_batch_cnt_tensor.shape[0], new_xyz_tensor.shape[0]
    _need(stack_neighbor_idxs_tensor, "stack_neighbor_idxs", int(avg_length_of_neighbor_idxs) * m)
    _need(start_len_tensor, "start_len", 2 * m); _need(cumsum_tensor, "cumsum", 1)
    with torch.cuda.device(support_xyz_tensor.device):
        _lib.check(lib.pdm_stack_query_local_neighbor_idxs(b, m, x, xc, q, qc, sn, sl, cs, int(avg_length_of_neighbor_idxs),
                                                           float(max_neighbour_distance), int(nsample), int(neighbor_type),
                                                           _stream(support_xyz_tensor)), "stack_query_local_neighbor_idxs")
    return 0


def query_three_nn_by_stacked_local_idxs_wrapper_stack(support_xyz_tensor, new_xyz_tensor, new_xyz_grid_centers_tensor,
                                                       new_xyz_grid_idxs_tensor, new_xyz_grid_dist2_tensor,
                                                       stack_neighbor_idxs_tensor, start_len_tensor, M, num_total_grids):
    lib = _lib.load()
    x = _chk(support_xyz_tensor, "support_xyz", _F32)
    _chk(new_xyz_tensor, "new_xyz", _F32)
    gc = _chk(new_xyz_grid_centers_tensor, "new_xyz_grid_centers", _F32)
    gi = _chk(new_xyz_grid_idxs_tensor, "new_xyz_grid_idxs", _I32)
    gd = _chk(new_xyz_grid_dist2_tensor, "new_xyz_grid_dist2", _F32)
    sn = _chk(stack_neighbor_idxs_tensor, "stack_neighbor_idxs", _I32)
    sl = _chk(start_len_tensor, "start_len", _I32)
    _need(new_xyz_grid_centers_tensor, "new_xyz_grid_centers", M * num_total_grids * 3)
    _need(new_xyz_grid_idxs_tensor, "new_xyz_grid_idxs", M * num_total_grids * 3)
    _need(new_xyz_grid_dist2_tensor, "new_xyz_grid_dist2", M * num_total_grids * 3); _need(start_len_tensor, "start_len", 2 * M)
    with torch.cuda.device(support_xyz_tensor.device):
        _lib.check(lib.pdm_stack_query_three_nn_by_local_idxs(M, num_total_grids, x, gc, gi, gd, sn, sl, _stream(support_xyz_tensor)),
                   "stack_query_three_nn_by_local_idxs")
    return 0


def vector_pool_wrapper(support_xyz_tensor, xyz_batch_cnt_tensor, support_features_tensor, new_xyz_tensor,
                        new_xyz_batch_cnt_tensor, new_features_tensor, new_local_xyz_tensor, point_cnt_of_grid_tensor,
                        grouped_idxs_tensor, num_grid_x, num_grid_y, num_grid_z, max_neighbour_distance, use_xyz,
                        num_max_sum_points, nsample, neighbor_type, pooling_type):
    """Returns the number of (support, centre, cell) entries the call produced, like the reference (which reads the
    counter back with a blocking cudaMemcpy, vector_pool_gpu.cu:360; here: one .item())."""
    lib = _lib.load()
    x = _chk(support_xyz_tensor, "support_xyz", _F32)
    xc = _chk(xyz_batch_cnt_tensor, "xyz_batch_cnt", _I32)
    sf = _chk(support_features_tensor, "support_features", _F32)
    q = _chk(new_xyz_tensor, "new_xyz", _F32)
    qc = _chk(new_xyz_batch_cnt_tensor, "new_xyz_batch_cnt", _I32)
    nf = _chk(new_features_tensor, "new_features", _F32)
    nl = _chk(new_local_xyz_tensor, "new_local_xyz", _F32)
    pc = _chk(point_cnt_of_grid_tensor, "point_cnt_of_grid", _I32)
    gi = _chk(grouped_idxs_tensor, "grouped_idxs", _I32)
    n, b, m = support_xyz_tensor.shape[0], xyz_batch_cnt_tensor.shape[0], new_xyz_tensor.shape[0]
    c_out, c_in, grids = new_features_tensor.shape[1], support_features_tensor.shape[1], point_cnt_of_grid_tensor.shape[1]
    _need(new_local_xyz_tensor, "new_local_xyz", m * 3 * grids); _need(point_cnt_of_grid_tensor, "point_cnt_of_grid", m * grids)
    _need(grouped_idxs_tensor, "grouped_idxs", int(num_max_sum_points) * 3)
    counter = torch.zeros(1, dtype=_I32, device=support_xyz_tensor.device)
    with torch.cuda.device(support_xyz_tensor.device):
        _lib.check(lib.pdm_stack_vector_pool(b, n, m, c_in, c_out, grids, x, xc, sf, q, qc, nf, nl, pc, gi, int(num_grid_x),
                                             int(num_grid_y), int(num_grid_z), float(max_neighbour_distance), int(bool(use_xyz)),
                                             int(num_max_sum_points), int(nsample), int(neighbor_type), int(pooling_type),
                                             counter.data_ptr(), _stream(support_xyz_tensor)), "stack_vector_pool")
    return int(counter.item())


def vector_pool_grad_wrapper(grad_new_features_tensor, point_cnt_of_grid_tensor, grouped_idxs_tensor,
                             grad_support_features_tensor):
    lib = _lib.load()
    g = _chk(grad_new_features_tensor, "grad_new_features", _F32)
    pc = _chk(point_cnt_of_grid_tensor, "point_cnt_of_grid", _I32)
    gi = _chk(grouped_idxs_tensor, "grouped_idxs", _I32)
    o = _chk(grad_support_features_tensor, "grad_support_features", _F32)
    m, c_out = grad_new_features_tensor.shape
    n, c_in = grad_support_features_tensor.shape
    grids, entries = point_cnt_of_grid_tensor.shape[1], grouped_idxs_tensor.shape[0]
    with torch.cuda.device(grad_new_features_tensor.device):
        _lib.check(lib.pdm_stack_vector_pool_grad(m, c_out, n, c_in, grids, entries, g, pc, gi, o, _stream(grad_new_features_tensor)),
                   "stack_vector_pool_grad")
    return 1
