"""Throughput form of the detector: several batches in flight, one CUDA graph of the whole forward per slot.

A batch alone keeps ~16 of the 148 SMs busy while it samples (farthest point sampling is one CTA per frame), so
`PipelinedDetector` sends consecutive batches round-robin to `n_streams` CUDA streams; each slot owns its static
input buffer, its captured graph (every kernel launch of `PDMSSD.forward`, ~40 of them, becomes one graph launch)
and its output buffers.  With `host=True` a step starts from a PINNED HOST point cloud and ends with the
detections in pinned host memory, both copies inside the slot's graph -- the serving loop's real boundary
(pcdet: `load_data_to_gpu`, pcdet/models/__init__.py:23-36, and `generate_prediction_dicts` reading the boxes
back, tools/eval_utils/eval_utils.py:58-79).  With `gather=True` (torch.distributed initialised) every step ends
with the one collective of the sharded pipeline: an NCCL all-gather of the fixed-shape detections
(`detector.gather_detections`, replacing the reference's pickle-file merge, pcdet/utils/common_utils.py:229-250),
issued on the slot's stream right after the graph replay.  With more than ~6 slots set CUDA_DEVICE_MAX_CONNECTIONS (hardware
work queues, default 8) to at least n_streams + 2 before CUDA initialises: streams that alias one queue serialise, and the
tiny all-gather then waits behind whole forward graphs of other slots (bench.py does this).
"""
import torch

from . import _lib
from .detector import gather_detections


class PipelinedDetector:
    def __init__(self, model, batch, n_points, n_streams=6, device="cuda:0", host=False, gather=False,
                 fps_mode=_lib.FPS_MODE_THROUGHPUT, point_channels=5):
        self.model, self.B, self.N = model, int(batch), int(n_points)
        self.dev = torch.device(device)
        self.host, self.gather, self.fps_mode = bool(host), bool(gather), fps_mode
        self.streams = [torch.cuda.Stream(device=self.dev) for _ in range(n_streams)]
        rows = self.B * self.N
        self.d_points = [torch.empty((rows, point_channels), dtype=torch.float32, device=self.dev) for _ in range(n_streams)]
        self.h_points = [None] * n_streams
        self.h_det = [None] * n_streams       # this rank's detections, pinned host (host=True)
        self.det = [None] * n_streams         # device detections written by the graph
        self.gathered = [None] * n_streams    # all ranks' detections (gather=True)
        self.h_gathered = [None] * n_streams
        self.graphs = None
        self.launches_per_step = None
        self.i = 0
        self._start = torch.cuda.Event()

    def _step(self, k):
        if self.host:
            self.d_points[k].copy_(self.h_points[k], non_blocking=True)
        out = self.model({"batch_size": self.B, "points": self.d_points[k]})
        det = out["detections"]
        if self.host:
            if self.h_det[k] is None:
                self.h_det[k] = torch.empty(det.shape, dtype=det.dtype).pin_memory()
            self.h_det[k].copy_(det, non_blocking=True)
        return det

    def capture(self, slot_inputs):
        """slot_inputs[k]: the (B*N, C) points slot k will always run on -- a pinned host tensor when host=True
        (re-read by every replay, so the caller may refill it between steps), else a device tensor (copied once)."""
        assert len(slot_inputs) == len(self.streams)
        torch.cuda.synchronize(self.dev)
        self.graphs = []
        with torch.no_grad(), _lib.fps_mode(self.fps_mode):
            for k, (st, x) in enumerate(zip(self.streams, slot_inputs)):
                if self.host:
                    if not x.is_pinned():
                        raise RuntimeError("host=True needs pinned host inputs")
                    self.h_points[k] = x
                else:
                    self.d_points[k].copy_(x)
                st.wait_stream(torch.cuda.current_stream(self.dev))
                with torch.cuda.stream(st):
                    self._step(k)                     # warm-up: allocator pools, kernel attributes, library scratch
                    before = _lib.launch_count()
                    self._step(k)
                    self.launches_per_step = _lib.launch_count() - before
                st.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=st):
                    self.det[k] = self._step(k)
                self.graphs.append(g)
                if self.gather:
                    import torch.distributed as dist
                    world = dist.get_world_size()
                    self.gathered[k] = self.det[k].new_empty((world * self.det[k].shape[0],) + tuple(self.det[k].shape[1:]))
                    if self.host and dist.get_rank() == 0:
                        self.h_gathered[k] = torch.empty(self.gathered[k].shape, dtype=self.gathered[k].dtype).pin_memory()
        torch.cuda.synchronize(self.dev)

    def begin(self):
        """Fork: every slot stream waits for what is already queued on the current stream."""
        self._start.record(torch.cuda.current_stream(self.dev))
        for st in self.streams:
            st.wait_event(self._start)
        self.i = 0

    def submit(self):
        """Next step on the next slot (round-robin); returns the slot index."""
        k = self.i % len(self.streams)
        self.i += 1
        with torch.cuda.stream(self.streams[k]):
            self.graphs[k].replay()
            if self.gather:
                import torch.distributed as dist
                dist.all_gather_into_tensor(self.gathered[k], self.det[k])
                if self.h_gathered[k] is not None:
                    self.h_gathered[k].copy_(self.gathered[k], non_blocking=True)
        return k

    def end(self):
        """Join: the current stream waits for every slot stream."""
        cur = torch.cuda.current_stream(self.dev)
        for st in self.streams:
            ev = torch.cuda.Event()
            ev.record(st)
            cur.wait_event(ev)

    @property
    def h2d_bytes(self):
        return self.d_points[0].numel() * 4 if self.host else 0

    @property
    def d2h_bytes(self):
        if not self.host or self.h_det[0] is None:
            return 0
        extra = self.h_gathered[0].numel() * 4 if self.h_gathered[0] is not None else 0
        return self.h_det[0].numel() * 4 + extra


__all__ = ["PipelinedDetector", "gather_detections"]
