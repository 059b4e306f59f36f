"""On-GPU sample_points + collate (csrc/sample_points.cu) against its numpy restatement (oracle/sample_points_oracle.py):
the chosen rows and the collated tensor are identical, for ragged batches covering every branch of
pcdet/datasets/processor/data_processor.py:182-212."""
import numpy as np
import pytest
import torch

import sample_points_oracle as so
from pdm_ssd_b200.data_processor import sample_points
from test_sample_points_cpu import _frame

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(frames, n, seed):
    pts = torch.from_numpy(np.concatenate(frames, 0) if sum(len(f) for f in frames) else np.zeros((0, 4), np.float32)).to(DEV)
    counts = torch.tensor([len(f) for f in frames], dtype=torch.int32, device=DEV)
    out, choice = sample_points(pts, counts, n, seed=seed, return_choice=True)
    return out.cpu().numpy(), choice.cpu().numpy()


def test_matches_oracle_on_every_branch():
    frames = [_frame(50000, 0.1, 1), _frame(50000, 0.6, 2), _frame(10000, 0.1, 3), _frame(16384, 0.2, 4), _frame(100, 0.0, 5),
              np.zeros((0, 4), np.float32), _frame(123457, 0.15, 6), _frame(1, 0.0, 7)]
    out, choice = _run(frames, 16384, seed=20261018)
    want_out, want_choice = so.sample_points(frames, 16384, seed=20261018)
    assert np.array_equal(choice, want_choice)
    assert np.array_equal(out, want_out)
    # different seed: different choice, same invariants
    out2, choice2 = _run(frames, 16384, seed=1)
    assert not np.array_equal(choice[0], choice2[0])
    assert len(np.unique(choice2[0])) == 16384


def test_feeds_the_detector_and_captures_in_a_graph():
    from pdm_ssd_b200.backbone import _split_points
    frames = [_frame(30000 + 1000 * i, 0.1, 10 + i) for i in range(4)]
    pts = torch.from_numpy(np.concatenate(frames, 0)).to(DEV)
    counts = torch.tensor([len(f) for f in frames], dtype=torch.int32, device=DEV)
    out = sample_points(pts, counts, 4096, seed=3)
    bidx, xyz, feats = _split_points(out, 4)
    assert xyz.shape == (4, 4096, 3) and feats.shape == (4, 1, 4096)
    assert torch.equal(bidx.view(4, 4096), torch.arange(4, device=DEV, dtype=torch.float32)[:, None].expand(4, 4096))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        sample_points(pts, counts, 4096, seed=3)              # sizes the library scratch of this stream
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            again = sample_points(pts, counts, 4096, seed=3)
        g.replay()
    s.synchronize()
    assert torch.equal(again, out)


def test_errors():
    with pytest.raises(RuntimeError):
        sample_points(torch.zeros(10, 4), torch.tensor([10], dtype=torch.int32, device=DEV), 8)       # CPU points
    with pytest.raises(RuntimeError):
        sample_points(torch.zeros(10, 2, device=DEV), torch.tensor([10], dtype=torch.int32, device=DEV), 8)   # < 3 channels
