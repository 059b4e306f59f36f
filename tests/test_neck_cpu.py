"""CPU tests of the PDM-neck oracle (SPEC_PDM.md): structural properties the spec promises."""
import numpy as np
import torch

import pdm_neck_oracle as O

RANGE = [0.0, -40.0, -3.0, 70.4, 40.0, 1.0]
VOX = [0.4, 0.4, 0.4]


def test_grid_and_offsets():
    assert O.grid_size(RANGE, VOX) == [176, 200, 10]
    off = O.offsets((1, 1, 1))
    assert off.shape == (27, 3) and off[0].tolist() == [-1, -1, -1] and off[1].tolist() == [-1, -1, 0]
    assert O.offsets((0, 0, 0)).tolist() == [[0, 0, 0]]


def test_cell_of_a_point_follows_pcdet_convention():
    pc = torch.tensor([[0, 0.39, -39.99, -2.99], [0, 0.41, 0.0, 0.9999], [1, 70.39, 39.99, -3.0], [0, 70.4, 0, 0]])
    cells, valid, key3 = O.dilate(pc, RANGE, VOX, (0, 0, 0), O.grid_size(RANGE, VOX))
    assert cells[:, 0].tolist() == [[0, 0, 0], [1, 100, 9], [175, 199, 0], [176, 100, 7]]
    assert valid[:, 0].tolist() == [True, True, True, False]   # x == range max is outside
    assert key3[2, 0].item() == ((1 * 176 + 175) * 200 + 199) * 10 + 0


def test_sh_basis_values():
    u = torch.tensor([[0.0, 0.0, 1.0], [1.0, 0.0, 0.0]])
    y = O.sh_basis(u, 2)
    assert y.shape == (2, 9)
    np.testing.assert_allclose(y[0].numpy(), [0.2820948, 0, 0.4886025, 0, 0, 0, 0.6307831, 0, 0], atol=1e-6)
    np.testing.assert_allclose(y[1].numpy(), [0.2820948, 0, 0, 0.4886025, 0, 0, -0.3153916, 0, 0.5462742], atol=1e-6)


def test_single_centre_degree0_gives_normalised_feature():
    """One centre, L = 0: every dilated cell holds f * w/|w| = +-f; a pillar sums its z-cells."""
    pc = torch.tensor([[0, 10.2, 0.2, -1.0]])
    f = torch.tensor([[1.0, -2.0, 3.0]])
    coef = torch.tensor([[2.0]])
    bev = O.neck_forward(pc, f, coef, 1, RANGE, VOX, (1, 1, 1), 0, 0.8, 1e-6)
    cx, cy = int(np.floor(10.2 / 0.4)), int(np.floor(40.2 / 0.4))
    assert (bev.abs().sum(1) > 0).sum().item() == 9
    np.testing.assert_allclose(bev[0, :, cy, cx].numpy(), 3 * f[0].numpy(), rtol=1e-4)  # 3 z-cells, each ~ f
    assert bev[0, :, cy + 2, cx].abs().sum() == 0


def test_linearity_in_features_and_batch_independence():
    torch.manual_seed(0)
    pc = torch.cat([torch.zeros(200, 1), torch.rand(200, 3) * torch.tensor([70.0, 79.0, 3.9]) + torch.tensor([0.0, -39.5, -2.95])], 1)
    f1, f2, coef = torch.randn(200, 6), torch.randn(200, 6), torch.randn(200, 9)
    a = O.neck_forward(pc, f1, coef, 1, RANGE, VOX)
    b = O.neck_forward(pc, f2, coef, 1, RANGE, VOX)
    ab = O.neck_forward(pc, f1 + f2, coef, 1, RANGE, VOX)
    assert float((a + b - ab).abs().max()) < 1e-4 * float(ab.abs().max())
    pc2 = pc.clone(); pc2[:, 0] = 1
    two = O.neck_forward(torch.cat([pc, pc2]), torch.cat([f1, f2]), torch.cat([coef, coef]), 2, RANGE, VOX)
    assert torch.equal(two[0], a[0]) and torch.equal(two[1], b[0])
