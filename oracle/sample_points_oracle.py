"""numpy restatement of csrc/sample_points.cu (DataProcessor.sample_points + collate for a batch, hash-defined randomness).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED against the reference's random stream (numpy's Mersenne Twister cannot be
reproduced on the device); the SELECTION RULES follow pcdet/datasets/processor/data_processor.py:182-212 and are checked as
invariants by tests/test_sample_points_*.py:
  n > N: every far point (depth >= 40) is kept when they fit, the rest is a subset of the near points without replacement
         (or a subset of all points without replacement when the far points alone exceed N); n <= N: every point appears, the
         padding is a choice without replacement; the result is a permutation of the selection.
"""
import numpy as np

M32 = np.uint64(0xffffffff)


def fmix32(h):
    h = h.astype(np.uint64) & M32
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85ebca6b)) & M32
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xc2b2ae35)) & M32
    h ^= h >> np.uint64(16)
    return h.astype(np.uint32)


def key(seed, stream, frame, i):
    i = np.asarray(i, dtype=np.uint64)
    v = (np.uint64(seed & 0xffffffff) ^ ((np.uint64(frame + 1) * np.uint64(0x9E3779B9)) & M32) ^ (((i + np.uint64(1)) * np.uint64(0x85EBCA6B)) & M32)
         ^ ((np.uint64(stream) * np.uint64(0xC2B2AE35)) & M32))
    return fmix32(v)


def sample_frame(points, frame, num_points, seed):
    """points (n, C) float32 -> choice (num_points,) int32 rows of `points` (-1 if n == 0)."""
    n = len(points)
    if n == 0:
        return np.full(num_points, -1, np.int32)
    x, y, z = points[:, 0].astype(np.float32), points[:, 1].astype(np.float32), points[:, 2].astype(np.float32)
    d2 = (x * x + y * y).astype(np.float32) + (z * z).astype(np.float32)
    far = d2 >= np.float32(1600.0)
    k1 = key(seed, 1, frame, np.arange(n)).astype(np.uint64)
    if n > num_points and num_points > int(far.sum()):
        k1 = k1 | (np.where(far, 0, 1).astype(np.uint64) << np.uint64(32))
    order = np.argsort(k1, kind="stable")
    if n > num_points:
        sel = order[:num_points]
    else:
        extra = order[np.arange(num_points - n) % n]
        sel = np.concatenate([np.arange(n), extra])
    k2 = key(seed, 2, frame, np.arange(num_points))
    return sel[np.argsort(k2, kind="stable")].astype(np.int32)


def sample_points(frames, num_points, seed=0):
    """frames: list of (n_i, C) float32 -> (B * num_points, 1 + C) float32 collated points, choice (B, num_points)."""
    C = frames[0].shape[1]
    out = np.zeros((len(frames) * num_points, 1 + C), np.float32)
    choice = np.zeros((len(frames), num_points), np.int32)
    for f, pts in enumerate(frames):
        ch = sample_frame(pts, f, num_points, seed)
        choice[f] = ch
        rows = out[f * num_points:(f + 1) * num_points]
        rows[:, 0] = f
        if len(pts):
            rows[:, 1:] = pts[ch]
    return out, choice
