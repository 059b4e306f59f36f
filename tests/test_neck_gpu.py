"""PDM neck: CUDA path (C ABI via the plugin) against the torch oracle of SPEC_PDM.md.
Cell keys bit-exact; weights and BEV features within 1e-3 relative (tolerance stated by
BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

import pdm_neck_oracle as O
from pdm_ssd_b200 import pdm_neck, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RANGE = [0.0, -40.0, -3.0, 70.4, 40.0, 1.0]


def _case(batch, m, c, seed, voxel=(0.4, 0.4, 0.4), out_of_range=False):
    rng = np.random.default_rng(seed)
    rows = []
    for b in range(batch):
        fr = synthetic.kitti_frame(1000 + seed * 10 + b)[:, :3]
        sel = rng.choice(len(fr), m, replace=False)
        xyz = fr[sel].copy()
        if out_of_range:
            xyz[: m // 8] += np.array([100.0, 0, 0], np.float32)   # dropped by the in-range mask
            xyz[m // 8: m // 4, 2] = 0.9999                        # top voxel layer: offsets leave the grid
        rows.append(np.concatenate([np.full((m, 1), b, np.float32), xyz], 1))
    coords = torch.from_numpy(np.concatenate(rows, 0))
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(batch * m, c, generator=g)
    return coords, feats


def _rel(a, b):
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-12))


@pytest.mark.parametrize("batch,m,c,degree,dil,oor", [
    (1, 4096, 256, 2, (1, 1, 1), False),     # BASELINE configs[0]
    (2, 1024, 64, 2, (1, 1, 1), True),
    (3, 333, 40, 1, (2, 1, 0), True),
    (1, 512, 300, 0, (1, 1, 1), False),      # C > 256: second channel pass
    (2, 64, 8, 2, (0, 0, 0), False),         # no dilation
])
def test_neck_matches_oracle(batch, m, c, degree, dil, oor):
    coords, feats = _case(batch, m, c, seed=batch * 7 + c, out_of_range=oor)
    g = torch.Generator().manual_seed(1)
    coef = torch.randn(batch * m, (degree + 1) ** 2, generator=g) * 0.5
    voxel = [0.4, 0.4, 0.4]
    grid = O.grid_size(RANGE, voxel)
    want, dbg = O.neck_forward(coords, feats, coef, batch, RANGE, voxel, dil, degree, 0.8, 1e-6, return_debug=True)
    got, keys, wts = pdm_neck.neck_forward(coords.to(DEV), feats.to(DEV), coef.to(DEV), batch, RANGE, voxel, grid,
                                           dil, degree, 0.8, 1e-6, return_debug=True)
    torch.cuda.synchronize()
    want_keys = torch.where(dbg["valid"], dbg["key3"], torch.full_like(dbg["key3"], -1)).int()
    assert torch.equal(keys.cpu(), want_keys), "dilation-grid keys must be bit-exact"
    wv = dbg["w"][dbg["valid"]]
    werr = _rel(wts.cpu()[dbg["valid"]], wv)
    assert werr < 1e-5, "entry weights differ: %g" % werr
    assert got.shape == want.shape
    ferr = _rel(got.cpu(), want)
    assert ferr < 1e-3, "BEV features differ: %g" % ferr
    # empty pillars are exactly zero and occupancy matches
    occ_got, occ_want = (got.cpu() != 0).any(1), (want != 0).any(1)
    assert torch.equal(occ_got, occ_want), "occupancy differs in %d pillars" % int((occ_got != occ_want).sum())


def test_neck_is_deterministic():
    coords, feats = _case(2, 2048, 128, seed=5)
    coef = torch.randn(4096, 9) * 0.5
    voxel = [0.4, 0.4, 0.4]
    grid = O.grid_size(RANGE, voxel)
    args = (coords.to(DEV), feats.to(DEV), coef.to(DEV), 2, RANGE, voxel, grid)
    a = pdm_neck.neck_forward(*args)
    b = pdm_neck.neck_forward(*args)
    assert torch.equal(a, b)  # in-order segmented sums: no atomics on float data


def test_neck_plugin_module_contract():
    cfg = dict(NUM_BEV_FEATURES=32, VOXEL_SIZE=[0.4, 0.4, 0.4], POINT_CLOUD_RANGE=RANGE, DILATION=[1, 1, 1],
               SH_DEGREE=2, SIGMA=0.8)
    torch.manual_seed(0)
    neck = pdm_neck.PDMNeck(cfg, grid_size=np.array([176, 200, 10])).to(DEV).eval()
    assert neck.num_bev_features == 32 and set(neck.state_dict()) == {"coef.weight", "coef.bias"}
    coords, feats = _case(2, 256, 32, seed=9)
    bd = {"batch_size": 2, "point_coords": coords.to(DEV), "point_features": feats.to(DEV)}
    with torch.no_grad():
        out = neck(bd)
    assert out["spatial_features"].shape == (2, 32, 200, 176)
    coef = torch.nn.functional.linear(feats, neck.coef.weight.cpu(), neck.coef.bias.cpu())
    want = O.neck_forward(coords, feats, coef.detach(), 2, RANGE, [0.4, 0.4, 0.4])
    assert _rel(out["spatial_features"].cpu(), want) < 1e-3


def test_neck_crowded_cells_take_the_selection_path():
    """> 32 entries in one cell (many centres inside one voxel) exercises the min-selection branch."""
    rng = np.random.default_rng(3)
    m = 200
    xyz = np.array([10.2, 0.2, -1.0], np.float32) + rng.uniform(-0.15, 0.15, (m, 3)).astype(np.float32)
    coords = torch.from_numpy(np.concatenate([np.zeros((m, 1), np.float32), xyz], 1))
    feats = torch.randn(m, 16, generator=torch.Generator().manual_seed(2))
    coef = torch.randn(m, 9, generator=torch.Generator().manual_seed(3))
    voxel = [0.4, 0.4, 0.4]
    want = O.neck_forward(coords, feats, coef, 1, RANGE, voxel)
    got = pdm_neck.neck_forward(coords.to(DEV), feats.to(DEV), coef.to(DEV), 1, RANGE, voxel, O.grid_size(RANGE, voxel))
    assert _rel(got.cpu(), want) < 1e-3
