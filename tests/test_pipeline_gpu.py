"""pipeline.PipelinedDetector: the CUDA-graph / multi-stream / pinned-host form of the detector returns exactly what
an eager forward returns, slot by slot, replay after replay (what bench.py times)."""
import numpy as np
import pytest
import torch

from pdm_ssd_b200 import _lib, synthetic
from pdm_ssd_b200.detector import PDMSSD, default_cfg
from pdm_ssd_b200.pipeline import PipelinedDetector

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_pipelined_graphs_match_eager_forward_device_and_host():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    B, N, S = 2, 4096, 3
    model = PDMSSD(default_cfg(N)).to(DEV).eval()
    host = [synthetic.to_pcdet_points(synthetic.kitti_batch(B, N, first_frame=10 * s)) for s in range(S)]
    # eager references: default (latency) sampling kernel -- the pipelines run the throughput-mode kernel, same indices
    want = [model({"batch_size": B, "points": torch.from_numpy(h).to(DEV)})["detections"].clone() for h in host]
    assert not torch.equal(want[0], want[1])

    pipe = PipelinedDetector(model, B, N, S, DEV, host=False)
    pipe.capture([torch.from_numpy(h).to(DEV) for h in host])
    assert pipe.launches_per_step >= 20
    for rounds in range(2):                                   # replays are repeatable
        pipe.begin()
        for _ in range(2 * S):
            pipe.submit()
        pipe.end()
        torch.cuda.synchronize()
        for k in range(S):
            assert torch.equal(pipe.det[k], want[k]), "slot %d differs from the eager forward" % k

    pinned = [torch.from_numpy(h).pin_memory() for h in host]
    hpipe = PipelinedDetector(model, B, N, S, DEV, host=True)
    hpipe.capture(pinned)
    assert hpipe.h2d_bytes == B * N * 5 * 4 and hpipe.d2h_bytes == B * 100 * 9 * 4
    hpipe.begin()
    for _ in range(S):
        hpipe.submit()
    hpipe.end()
    torch.cuda.synchronize()
    for k in range(S):
        assert torch.equal(hpipe.h_det[k], want[k].cpu())
    # the pinned input is re-read by every replay: refill slot 0 with slot 1's frames -> slot 0 now returns slot 1's result
    pinned[0].copy_(pinned[1])
    hpipe.begin()
    hpipe.submit()
    hpipe.end()
    torch.cuda.synchronize()
    assert torch.equal(hpipe.h_det[0], want[1].cpu())


def test_fps_mode_is_per_thread_and_per_call():
    """`_lib.fps_mode` replaces the process-wide switch of round 1 (ADVICE / VERDICT r1 #15): it is thread-local and
    restored on exit, and both kernels return the same indices."""
    import threading
    from pdm_ssd_b200 import pointnet2_utils as pu
    xyz = torch.from_numpy(synthetic.kitti_batch(2, 8192)[..., :3].copy()).to(DEV)
    base = pu.farthest_point_sample(xyz, 512)
    seen = {}

    def other():
        seen["mode"] = _lib.current_fps_mode()
    with _lib.fps_mode(_lib.FPS_MODE_THROUGHPUT):
        assert _lib.current_fps_mode() == _lib.FPS_MODE_THROUGHPUT
        t = threading.Thread(target=other)
        t.start()
        t.join()
        assert torch.equal(pu.farthest_point_sample(xyz, 512), base)
    assert seen["mode"] is None and _lib.current_fps_mode() is None
