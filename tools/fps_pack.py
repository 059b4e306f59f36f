"""How well do one-CTA-per-frame FPS launches pack onto the 148 SMs?
(1) one launch with B frames, B = 16..296: duration vs B shows the per-SM independence;
(2) S streams x R launches of 16 frames: ms per launch vs the ideal 16 * t_frame / 148."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pdm_ssd_b200 import pointnet2_batch_cuda as ours, synthetic  # noqa: E402

dev = torch.device("cuda:0")
N, M = 16384, 4096
base = torch.from_numpy(synthetic.kitti_batch(16, N)[..., :3].copy()).to(dev)


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for B in (16, 74, 148, 149, 296):
    xyz = base.repeat((B + 15) // 16, 1, 1)[:B].contiguous()
    temp = torch.full((B, N), 1e10, device=dev)
    idx = torch.empty(B, M, dtype=torch.int32, device=dev)
    ours.farthest_point_sampling_wrapper(B, N, M, xyz, temp, idx)
    temp.fill_(1e10)
    ms = timed(lambda: ours.farthest_point_sampling_wrapper(B, N, M, xyz, temp, idx))
    print("one launch, %3d frames: %.3f ms  (%.1f us/frame)" % (B, ms, 1e3 * ms / B), flush=True)

R = 8
for S in (1, 4, 9, 12, 18, 27):
    streams = [torch.cuda.Stream() for _ in range(S)]
    temps = [torch.full((16, N), 1e10, device=dev) for _ in range(S)]
    idxs = [torch.empty(16, M, dtype=torch.int32, device=dev) for _ in range(S)]

    def go():
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event(); ev.record(cur)
        for r in range(R):
            for k, st in enumerate(streams):
                with torch.cuda.stream(st):
                    if r == 0:
                        st.wait_event(ev)
                    ours.farthest_point_sampling_wrapper(16, N, M, base, temps[k], idxs[k])
        for st in streams:
            e = torch.cuda.Event(); e.record(st); cur.wait_event(e)
    go()
    ms = timed(go)
    print("%2d streams x %d launches of 16 frames: %.3f ms per launch (%.1f us/frame)" % (S, R, ms / (S * R), 1e3 * ms / (S * R * 16)), flush=True)
