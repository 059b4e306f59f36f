"""Golden vectors for the box decode of the hybrid head, produced by the REFERENCE's own
PointResidualCoder.decode_torch (pcdet/utils/box_coder_utils.py:189-222) imported from /root/reference
(CPU; its __init__ calls .cuda(), so the instance is assembled by hand -- decode_torch itself is untouched).

    python tests/golden/make_golden_head.py        # writes tests/golden/head_decode.npz
"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PDM_REFERENCE_ROOT", "/root/reference")
MEAN_SIZE = [[3.9, 1.6, 1.56], [0.8, 0.6, 1.73], [1.76, 0.6, 1.73]]


def reference_coder():
    spec = importlib.util.spec_from_file_location("ref_box_coder_utils", os.path.join(REF, "pcdet/utils/box_coder_utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    coder = object.__new__(mod.PointResidualCoder)
    coder.code_size, coder.use_mean_size = 8, True
    coder.mean_size = torch.tensor(MEAN_SIZE, dtype=torch.float32)
    return coder


if __name__ == "__main__":
    g = torch.Generator().manual_seed(20261018)
    n = 4096
    enc = torch.randn(n, 8, generator=g) * torch.tensor([1.5, 1.5, 1.5, 0.4, 0.4, 0.4, 1.0, 1.0])
    enc[:8, 6:] = torch.tensor([[1, 0], [-1, 0], [0, 1], [0, -1], [0, 0], [-1, -0.0], [1e-30, -1e-30], [3, 4]], dtype=torch.float32)
    pts = torch.rand(n, 3, generator=g) * torch.tensor([70.4, 80.0, 4.0]) + torch.tensor([0.0, -40.0, -3.0])
    cls = torch.randint(1, 4, (n,), generator=g)
    out = reference_coder().decode_torch(enc, pts, cls)
    np.savez_compressed(os.path.join(HERE, "head_decode.npz"), enc=enc.numpy(), points=pts.numpy(),
                        pred_classes=cls.numpy().astype(np.int64), boxes=out.numpy())
    print("wrote head_decode.npz", tuple(out.shape))
