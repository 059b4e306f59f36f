"""numpy front-end of the stacked-op restatements in oracle/pdm_stack_oracle.c (pointnet2_stack family).

TEST INFRASTRUCTURE ONLY -- see the header of pdm_stack_oracle.c.  Argument order follows the reference's Python
Functions (pcdet/ops/pointnet2/pointnet2_stack/pointnet2_utils.py, voxel_query_utils.py).
"""
import ctypes

import numpy as np

from oracle import _fp, _ip, lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def ball_query(radius, nsample, xyz, xyz_cnt, new_xyz, new_cnt):
    """-> idx (M, nsample) raw kernel output (idx[:,0] == -1 flags an empty ball; pointnet2_utils.py:31-37)"""
    xyz, new_xyz, xyz_cnt, new_cnt = _f32(xyz), _f32(new_xyz), _i32(xyz_cnt), _i32(new_cnt)
    m = new_xyz.shape[0]
    idx = np.zeros((m, nsample), np.int32)
    lib().oracle_stack_ball_query(len(xyz_cnt), m, ctypes.c_float(radius), nsample, _fp(new_xyz), _ip(new_cnt), _fp(xyz),
                                  _ip(xyz_cnt), _ip(idx))
    return idx


def fps(xyz, xyz_cnt, m_cnt, return_temp=False):
    xyz, xyz_cnt, m_cnt = _f32(xyz), _i32(xyz_cnt), _i32(m_cnt)
    temp = np.full((xyz.shape[0],), 1e10, np.float32)
    idx = np.zeros((int(m_cnt.sum()),), np.int32)
    lib().oracle_stack_fps(len(xyz_cnt), _fp(xyz), _fp(temp), _ip(xyz_cnt), _ip(idx), _ip(m_cnt))
    return (idx, temp) if return_temp else idx


def group_points(features, features_cnt, idx, idx_cnt):
    features, features_cnt, idx, idx_cnt = _f32(features), _i32(features_cnt), _i32(idx), _i32(idx_cnt)
    m, ns = idx.shape
    c = features.shape[1]
    out = np.empty((m, c, ns), np.float32)
    lib().oracle_stack_group_points(len(idx_cnt), m, c, ns, _fp(features), _ip(features_cnt), _ip(idx), _ip(idx_cnt), _fp(out))
    return out


def group_points_grad(grad_out, idx, idx_cnt, features_cnt, n):
    grad_out, idx, idx_cnt, features_cnt = _f32(grad_out), _i32(idx), _i32(idx_cnt), _i32(features_cnt)
    m, c, ns = grad_out.shape
    gf = np.zeros((n, c), np.float32)
    lib().oracle_stack_group_points_grad(len(idx_cnt), m, c, ns, _fp(grad_out), _ip(idx), _ip(idx_cnt), _ip(features_cnt), _fp(gf))
    return gf


def three_nn(unknown, unknown_cnt, known, known_cnt):
    """-> (dist2, idx): squared distances as the kernel stores them"""
    unknown, known, unknown_cnt, known_cnt = _f32(unknown), _f32(known), _i32(unknown_cnt), _i32(known_cnt)
    n = unknown.shape[0]
    d2 = np.zeros((n, 3), np.float32)
    idx = np.zeros((n, 3), np.int32)
    lib().oracle_stack_three_nn(len(unknown_cnt), n, _fp(unknown), _ip(unknown_cnt), _fp(known), _ip(known_cnt), _fp(d2), _ip(idx))
    return d2, idx


def three_interpolate(features, idx, weight):
    features, idx, weight = _f32(features), _i32(idx), _f32(weight)
    n, c = idx.shape[0], features.shape[1]
    out = np.empty((n, c), np.float32)
    lib().oracle_stack_three_interpolate(n, c, _fp(features), _ip(idx), _fp(weight), _fp(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, idx, weight = _f32(grad_out), _i32(idx), _f32(weight)
    n, c = grad_out.shape
    gf = np.zeros((m, c), np.float32)
    lib().oracle_stack_three_interpolate_grad(n, c, _fp(grad_out), _ip(idx), _fp(weight), _fp(gf))
    return gf


def voxel_query(max_range, radius, nsample, xyz, new_xyz, new_coords, point_indices):
    xyz, new_xyz, new_coords, point_indices = _f32(xyz), _f32(new_xyz), _i32(new_coords), _i32(point_indices)
    m = new_coords.shape[0]
    _, r1, r2, r3 = point_indices.shape
    idx = np.zeros((m, nsample), np.int32)
    zr, yr, xr = max_range
    lib().oracle_stack_voxel_query(m, r1, r2, r3, nsample, ctypes.c_float(radius), zr, yr, xr, _fp(new_xyz), _fp(xyz),
                                   _ip(new_coords), _ip(point_indices), _ip(idx))
    return idx


def local_neighbor_idxs(support_xyz, xyz_cnt, new_xyz, new_cnt, avg_length, dmax, nsample, neighbor_type):
    """-> (stack_neighbor_idxs (avg_length * M), start_len (M,2), cumsum)"""
    support_xyz, new_xyz, xyz_cnt, new_cnt = _f32(support_xyz), _f32(new_xyz), _i32(xyz_cnt), _i32(new_cnt)
    m = new_xyz.shape[0]
    lst = np.zeros((avg_length * m,), np.int32)
    sl = np.zeros((m, 2), np.int32)
    lib().oracle_stack_local_neighbor_idxs.restype = ctypes.c_int
    total = lib().oracle_stack_local_neighbor_idxs(len(xyz_cnt), m, _fp(support_xyz), _ip(xyz_cnt), _fp(new_xyz), _ip(new_cnt),
                                                   _ip(lst), _ip(sl), avg_length, ctypes.c_float(dmax), nsample, neighbor_type)
    return lst, sl, int(total)


def three_nn_local(support_xyz, grid_centers, stack_neighbor_idxs, start_len):
    support_xyz, grid_centers = _f32(support_xyz), _f32(grid_centers)
    stack_neighbor_idxs, start_len = _i32(stack_neighbor_idxs), _i32(start_len)
    m, g, _ = grid_centers.shape
    idxs = np.full((m, g, 3), -1, np.int32)
    d2 = np.zeros((m, g, 3), np.float32)
    lib().oracle_stack_three_nn_local(m, g, _fp(support_xyz), _fp(grid_centers), _ip(idxs), _fp(d2), _ip(stack_neighbor_idxs),
                                      _ip(start_len))
    return d2, idxs


def vector_pool(support_xyz, xyz_cnt, support_features, new_xyz, new_cnt, grids, dmax, c_out_each_grid, use_xyz,
                num_max_sum_points, nsample, neighbor_type, pooling_type):
    """-> (new_features SUMS (M, c_out), new_local_xyz sums (M, 3G), point_cnt_of_grid (M,G), grouped_idxs, cum_sum)"""
    support_xyz, support_features, new_xyz = _f32(support_xyz), _f32(support_features), _f32(new_xyz)
    xyz_cnt, new_cnt = _i32(xyz_cnt), _i32(new_cnt)
    ngx, ngy, ngz = grids
    g = ngx * ngy * ngz
    m, c_in = new_xyz.shape[0], support_features.shape[1]
    c_out = c_out_each_grid * g
    nf = np.zeros((m, c_out), np.float32)
    nl = np.zeros((m, 3 * g), np.float32)
    pc = np.zeros((m, g), np.int32)
    gi = np.zeros((num_max_sum_points, 3), np.int32)
    lib().oracle_stack_vector_pool.restype = ctypes.c_int
    cum = lib().oracle_stack_vector_pool(len(xyz_cnt), m, c_in, c_out, g, _fp(support_xyz), _ip(xyz_cnt), _fp(support_features),
                                         _fp(new_xyz), _ip(new_cnt), _fp(nf), _fp(nl), _ip(pc), _ip(gi), ngx, ngy, ngz,
                                         ctypes.c_float(dmax), int(use_xyz), num_max_sum_points, nsample, neighbor_type,
                                         pooling_type)
    return nf, nl, pc, gi, int(cum)


def vector_pool_grad(grad_new_features, point_cnt_of_grid, grouped_idxs, n, c_in):
    grad_new_features, point_cnt_of_grid, grouped_idxs = _f32(grad_new_features), _i32(point_cnt_of_grid), _i32(grouped_idxs)
    c_out, g = grad_new_features.shape[1], point_cnt_of_grid.shape[1]
    gs = np.zeros((n, c_in), np.float32)
    lib().oracle_stack_vector_pool_grad(c_out, c_in, g, grouped_idxs.shape[0], _fp(grad_new_features), _ip(point_cnt_of_grid),
                                        _ip(grouped_idxs), _fp(gs))
    return gs
