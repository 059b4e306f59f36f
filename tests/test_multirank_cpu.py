"""world_size-2 gloo tests of the N > 1 path (SURVEY section 4 iv): frames are sharded by rank with no data-path
collective except the all-gather of the fixed-shape detections (`detector.gather_detections`, replacing the
reference's pickle-file merge, pcdet/utils/common_utils.py:229-250); the timing reduction is a MAX all-reduce.

The detections a rank contributes are produced here by the product's own head (torch path on CPU, post_cfg=None:
device-side NMS is GPU-only) on that rank's frames; the test asserts that what every rank holds after the gather is
exactly the un-sharded, one-process result, in frame order."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _detections(frames):
    """Per-frame detections (F, K, 9) of the product's BEV context + hybrid head on a deterministic synthetic BEV map
    and point set per frame id (CPU, torch path).  Frame-independent, so any sharding must reproduce it."""
    from pdm_ssd_b200.detector import BEVContext, HybridHead, default_cfg
    cfg = default_cfg()
    torch.manual_seed(0)
    ctx = BEVContext(cfg.BACKBONE_2D, 128).eval()
    head = HybridHead(cfg.DENSE_HEAD, 128, 128, 3, cfg.POINT_CLOUD_RANGE, cfg.VOXEL_SIZE, post_cfg=None).eval()
    out = []
    with torch.no_grad():
        for f in frames:
            g = torch.Generator().manual_seed(1000 + f)
            sf = torch.randn(1, 128, 24, 32, generator=g) * (torch.rand(1, 1, 24, 32, generator=g) < 0.3)
            P = 256
            coords = torch.cat([torch.zeros(P, 1), torch.rand(P, 3, generator=g) * torch.tensor([12.0, 9.0, 4.0]) + torch.tensor([0.0, -40.0, -3.0])], 1)
            bd = ctx({"spatial_features": sf, "batch_size": 1})
            bd.update(point_coords=coords, point_features=torch.randn(P, 128, generator=g))
            out.append(head(bd)["detections"][0])
    return torch.stack(out)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from pdm_ssd_b200.detector import gather_detections
    frames_per_rank = 3
    mine = list(range(rank * frames_per_rank, (rank + 1) * frames_per_rank))        # contiguous blocks of frames per rank
    det = _detections(mine)
    gathered = gather_detections(det)
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, float(t.item()), gathered.numpy(), det.numpy()))
    dist.destroy_process_group()


def test_two_rank_sharded_detections_equal_single_rank_result():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(world)), key=lambda r: r[0])
    for p in ps:
        p.join(timeout=60)
    assert all(r[1] == 11.0 for r in res)                                    # timing reduction: MAX over ranks
    single = _detections(list(range(6))).numpy()                             # the one-process, un-sharded result
    assert single.shape == (6, 100, 9)
    for r in res:
        assert r[2].shape == single.shape
        assert (r[2] == single).all()                                        # every rank holds all detections, frame order kept
    assert (res[0][3] == single[:3]).all() and (res[1][3] == single[3:]).all()
    assert (res[0][3] != res[1][3]).any()                                    # ranks really worked on different frames


def test_gather_is_identity_without_process_group():
    from pdm_ssd_b200.detector import gather_detections
    d = torch.zeros(2, 5, 9)
    assert gather_detections(d) is d
