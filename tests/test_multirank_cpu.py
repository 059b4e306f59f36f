"""world_size-2 gloo test of the N>1 harness logic: frames are sharded by rank with no
data-path collective; the timing reduction is a MAX all-reduce; rank-disjoint inputs."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "oracle"))
    import bench
    import oracle
    frames, feat2 = bench.make_host_batches(rank, pool=1, batch=2)[0]
    idx = oracle.fps(np.ascontiguousarray(frames[:, :512, :3]), 16)
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [torch.zeros(2, 16, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(idx))
    q.put((rank, float(t.item()), float(frames.sum()), [g.numpy().tolist() for g in gathered]))
    dist.destroy_process_group()


def test_two_rank_sharding_and_max_timing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
    assert all(r[1] == 11.0 for r in res)          # MAX over ranks
    assert res[0][2] != res[1][2]                   # ranks work on different frames
    assert res[0][3] == res[1][3]                   # gathered results identical on every rank
    assert res[0][3][0] != res[0][3][1]             # and per-rank shards differ
