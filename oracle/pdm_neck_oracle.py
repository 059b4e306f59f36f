"""Pure-torch CPU restatement of SPEC_PDM.md (the PDM neck: point dilation, SH x Gaussian feature
filling, multi-centre fusion, height compression).

TEST INFRASTRUCTURE ONLY (also BASELINE.json configs[0]: "PDM neck alone ... pure-torch on CPU").
PARITY UNPINNED: the reference tree has no PDM neck code to check this against (SURVEY.md
section 0.1); this file IS the specification in executable form.  In-tree conventions it follows:
point->cell  floor((p - range_min) / voxel)   pcdet/models/backbones_3d/vfe/dynamic_voxel_vfe.py:60-71
voxel centre (cell + 0.5) * v + min            dynamic_voxel_vfe.py:42-44
sum over colliding cells after dropping z      pcdet/models/backbones_3d/spconv_backbone_voxelnext.py:149-164
dense output (B, C, Y, X)                      pcdet/models/backbones_2d/map_to_bev/height_compression.py:22-25
Gaussian of squared distance                   pcdet/models/model_utils/centernet_utils.py:89-96
in-order segmented sums                        pcdet/ops/bev_pool/src/bev_pool_cuda.cu:37-41
"""
import torch

SH_C = (0.28209479177387814, 0.4886025119029199, 1.0925484305920792, 0.31539156525252005, 0.5462742152960396)


def grid_size(point_cloud_range, voxel_size):
    r = torch.as_tensor(point_cloud_range, dtype=torch.float64)
    v = torch.as_tensor(voxel_size, dtype=torch.float64)
    return [int(x) for x in torch.round((r[3:] - r[:3]) / v).tolist()]


def offsets(dilation):
    kx, ky, kz = dilation
    return torch.tensor([(ox, oy, oz) for ox in range(-kx, kx + 1) for oy in range(-ky, ky + 1)
                         for oz in range(-kz, kz + 1)], dtype=torch.int32)


def sh_basis(u, degree):
    """u (...,3) unit directions -> (..., (degree+1)^2), products rounded one by one (no fma)."""
    ux, uy, uz = u[..., 0], u[..., 1], u[..., 2]
    f = u.new_tensor
    ys = [torch.full_like(ux, SH_C[0])]
    if degree >= 1:
        ys += [f(SH_C[1]) * uy, f(SH_C[1]) * uz, f(SH_C[1]) * ux]
    if degree >= 2:
        ys += [f(SH_C[2]) * (ux * uy), f(SH_C[2]) * (uy * uz), f(SH_C[3]) * (f(3.0) * (uz * uz) - f(1.0)),
               f(SH_C[2]) * (ux * uz), f(SH_C[4]) * (ux * ux - uy * uy)]
    return torch.stack(ys, -1)


def dilate(point_coords, point_cloud_range, voxel_size, dilation, grid):
    """n1.  Returns cells (P,K,3) int32, valid (P,K) bool, key3 (P,K) int64."""
    rmin = torch.tensor(point_cloud_range[:3], dtype=torch.float32)
    v = torch.tensor(voxel_size, dtype=torch.float32)
    xyz = point_coords[:, 1:4].float()
    b = point_coords[:, 0].long()
    c0 = torch.floor((xyz - rmin) / v).int()
    g = torch.tensor(grid, dtype=torch.int32)
    keep = ((c0 >= 0) & (c0 < g)).all(1)
    cells = c0[:, None, :] + offsets(dilation)[None]
    valid = ((cells >= 0) & (cells < g)).all(2) & keep[:, None]
    X, Y, Z = grid
    c = cells.long()
    key3 = ((b[:, None] * X + c[..., 0]) * Y + c[..., 1]) * Z + c[..., 2]
    return cells, valid, key3


def fill_weights(point_coords, coef, cells, point_cloud_range, voxel_size, degree, sigma, eps):
    """n2.  coef (P,n_sh) -> w (P,K) f32."""
    rmin = torch.tensor(point_cloud_range[:3], dtype=torch.float32)
    v = torch.tensor(voxel_size, dtype=torch.float32)
    xyz = point_coords[:, 1:4].float()
    ctr = (cells.float() + 0.5) * v + rmin
    d = ctr - xyz[:, None, :]
    r2 = d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2]
    u = d / torch.sqrt(r2 + eps)[..., None]
    ys = sh_basis(u, degree)                     # (P,K,n_sh)
    acc = coef[:, None, 0] * ys[..., 0]
    for i in range(1, ys.shape[-1]):
        acc = acc + coef[:, None, i] * ys[..., i]
    two_s2 = torch.tensor(2.0 * sigma * sigma, dtype=torch.float32)
    return acc * torch.exp(-r2 / two_s2)


def _inorder_segment_sum(values, seg, nseg, rank):
    """values (E,...) summed per segment in ascending entry order (serial fp32 sums)."""
    out = torch.zeros((nseg,) + tuple(values.shape[1:]), dtype=values.dtype)
    for r in range(int(rank.max().item()) + 1 if rank.numel() else 0):
        sel = rank == r
        out[seg[sel]] = out[seg[sel]] + values[sel]
    return out


def neck_forward(point_coords, point_features, coef, batch_size, point_cloud_range, voxel_size,
                 dilation=(1, 1, 1), degree=2, sigma=0.8, eps=1e-6, return_debug=False):
    """Full neck.  point_coords (P,4) [b,x,y,z], point_features (P,C), coef (P,(degree+1)^2).
    Returns spatial_features (B,C,Y,X) (and the intermediate keys when return_debug)."""
    grid = grid_size(point_cloud_range, voxel_size)
    X, Y, Z = grid
    P, C = point_features.shape
    cells, valid, key3 = dilate(point_coords, point_cloud_range, voxel_size, dilation, grid)
    w = fill_weights(point_coords, coef, cells, point_cloud_range, voxel_size, degree, sigma, eps)
    K = cells.shape[1]
    eid = torch.arange(P * K).view(P, K)
    vk, ve, vw = key3[valid], eid[valid], w[valid]
    order = torch.sort(vk, stable=True).indices      # ties keep ascending entry id
    sk, se, sw = vk[order], ve[order], vw[order]
    sp = se // K
    # per 3-D cell
    ukey, seg = torch.unique_consecutive(sk, return_inverse=True)
    first = torch.ones_like(sk, dtype=torch.bool)
    first[1:] = sk[1:] != sk[:-1]
    start = torch.nonzero(first).flatten()
    rank = torch.arange(len(sk)) - start[seg]
    num = _inorder_segment_sum(sw[:, None] * point_features[sp], seg, len(ukey), rank)
    den = _inorder_segment_sum(sw.abs(), seg, len(ukey), rank)
    F = num / (den + eps)[:, None]
    # height compression: sum the cells of a pillar in ascending z
    pkey = ukey // Z
    upil, pseg = torch.unique_consecutive(pkey, return_inverse=True)
    pfirst = torch.ones_like(pkey, dtype=torch.bool)
    pfirst[1:] = pkey[1:] != pkey[:-1]
    pstart = torch.nonzero(pfirst).flatten()
    prank = torch.arange(len(pkey)) - pstart[pseg]
    bev_rows = _inorder_segment_sum(F, pseg, len(upil), prank)
    bev = torch.zeros((batch_size, C, Y, X), dtype=torch.float32)
    pb, pxy = upil // (X * Y), upil % (X * Y)
    px, py = pxy // Y, pxy % Y
    bev[pb, :, py, px] = bev_rows
    if return_debug:
        return bev, dict(cells=cells, valid=valid, key3=key3, w=w, sorted_keys=sk, sorted_entries=se,
                         cell_keys=ukey, cell_features=F)
    return bev
