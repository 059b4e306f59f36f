"""The numpy oracle of the on-GPU sample_points keeps the reference's selection rules
(pcdet/datasets/processor/data_processor.py:182-212) -- checked as invariants, since numpy's random stream itself cannot
be reproduced on the device."""
import numpy as np

import sample_points_oracle as so


def _frame(n, far_frac, seed):
    rng = np.random.default_rng(seed)
    r = np.where(rng.uniform(size=n) < far_frac, rng.uniform(41, 70, n), rng.uniform(1, 39, n))
    a = rng.uniform(-0.7, 0.7, n)
    return np.stack([r * np.cos(a), r * np.sin(a), rng.uniform(-2, 1, n), rng.uniform(0, 1, n)], 1).astype(np.float32)


def _far(p):
    return np.linalg.norm(p[:, :3], axis=1) >= 40.0


def test_more_points_than_needed_keeps_all_far_points():
    p = _frame(50000, 0.1, 1)
    ch = so.sample_frame(p, 0, 16384, seed=7)
    assert len(ch) == 16384 and len(np.unique(ch)) == 16384                 # no replacement (data_processor.py:197)
    far_rows = np.where(_far(p))[0]
    assert np.isin(far_rows, ch).all()                                      # every far point kept (:198-199)
    assert not np.array_equal(ch, np.sort(ch))                              # shuffled (:203)


def test_far_points_alone_exceed_the_budget():
    p = _frame(50000, 0.6, 2)
    ch = so.sample_frame(p, 3, 16384, seed=7)
    assert len(np.unique(ch)) == 16384                                      # uniform subset of everything, no replacement (:200-202)
    assert 0.5 < _far(p[ch]).mean() < 0.7


def test_fewer_points_are_padded_without_losing_any():
    p = _frame(10000, 0.1, 3)
    ch = so.sample_frame(p, 1, 16384, seed=7)
    counts = np.bincount(ch, minlength=10000)
    assert counts.min() == 1 and counts.max() == 2 and (counts == 2).sum() == 6384   # all points + choice without replacement (:205-209)
    assert so.sample_frame(p, 1, 10000, seed=7).tolist() != list(range(10000))       # n == N: a pure shuffle
    assert sorted(so.sample_frame(p, 1, 10000, seed=7).tolist()) == list(range(10000))


def test_seed_and_frame_change_the_choice_and_collate_layout():
    p = _frame(30000, 0.1, 4)
    a, b, c = so.sample_frame(p, 0, 4096, 1), so.sample_frame(p, 0, 4096, 2), so.sample_frame(p, 1, 4096, 1)
    assert not np.array_equal(a, b) and not np.array_equal(a, c)
    assert np.array_equal(a, so.sample_frame(p, 0, 4096, 1))                # deterministic
    out, choice = so.sample_points([p, _frame(100, 0.0, 5), np.zeros((0, 4), np.float32)], 2048, seed=9)
    assert out.shape == (3 * 2048, 5)
    assert (out[:2048, 0] == 0).all() and (out[2048:4096, 0] == 1).all() and (out[4096:, 0] == 2).all()    # dataset.py:240-243
    assert np.array_equal(out[:2048, 1:], p[choice[0]]) and (choice[2] == -1).all() and (out[4096:, 1:] == 0).all()
    assert np.bincount(choice[1], minlength=100).min() >= 20                # 100 points padded cyclically to 2048
