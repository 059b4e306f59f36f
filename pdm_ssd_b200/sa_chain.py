"""The set-abstraction operator chain of BASELINE config 2, as one callable.

    SA1: FPS 16384 -> 4096, gather centres, ball query r=0.8 / 32, group xyz, group features (C=1)
    SA2: FPS  4096 -> 1024, gather centres, ball query r=1.6 / 32, group xyz, group features (C=64)

It produces what `_PointnetSAModuleBase.forward` computes up to the shared MLP in the reference
(pointnet2_modules.py:19-55): new_xyz and the QueryAndGroup tensor (B, 3+C, npoint, nsample) of
each layer.  With the reference extension as backend it issues exactly the reference's operator
sequence (pointnet2_utils.py:241-264: ball query, xyz^T, group, subtract, group, cat) through the
nine-function API; with our backend the grouping half is the one-pass `pdm_query_and_group`
(bit-identical output).  Outputs are pre-allocated (steady-state serving: no allocator traffic
inside a step with our backend).
"""
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import _lib, pointnet2_batch_cuda as _ours


@dataclass
class SALayerCfg:
    npoint: int
    radius: float
    nsample: int
    channels: int  # feature channels grouped at this layer


KITTI_CHAIN = (SALayerCfg(4096, 0.8, 32, 1), SALayerCfg(1024, 1.6, 32, 64))


def algorithmic_bytes_per_frame(n_points: int, layers=KITTI_CHAIN) -> Dict[str, int]:
    """Compulsory HBM bytes per frame, op by op (SURVEY.md section 8d / BASELINE.md 2.4):
    every API-visible input read once, every API-visible output written once, fp32/int32."""
    out = {}
    n = n_points
    for li, L in enumerate(layers, 1):
        m, s, c = L.npoint, L.nsample, L.channels
        out["sa%d_fps" % li] = 12 * n + 4 * m
        out["sa%d_gather" % li] = 4 * m + 24 * m
        out["sa%d_ball_query" % li] = 12 * n + 12 * m + 4 * m * s
        out["sa%d_group_xyz" % li] = 4 * m * s + 12 * n + 12 * m * s
        out["sa%d_group_feat" % li] = 4 * m * s + 4 * c * n + 4 * c * m * s
        n = m
    out["total"] = sum(out.values())
    return out


class SAChain:
    def __init__(self, batch: int, n_points: int = 16384, layers=KITTI_CHAIN, device="cuda:0", backend=None):
        self.B, self.N, self.layers = batch, n_points, tuple(layers)
        self.dev = torch.device(device)
        self.be = backend if backend is not None else _ours
        self.ws = []
        n = n_points
        for L in self.layers:
            m, s, c = L.npoint, L.nsample, L.channels
            f32 = dict(dtype=torch.float32, device=self.dev)
            i32 = dict(dtype=torch.int32, device=self.dev)
            grouped = torch.empty((batch, 3 + c, m, s), **f32)   # QueryAndGroup output: [xyz - centre, features]
            self.ws.append(dict(
                temp=torch.empty((batch, n), **f32), fps_idx=torch.empty((batch, m), **i32),
                xyz_t=torch.empty((batch, 3, n), **f32), new_t=torch.empty((batch, 3, m), **f32),
                new_xyz=torch.empty((batch, m, 3), **f32), ball_idx=torch.empty((batch, m, s), **i32),
                grouped=grouped, grouped_xyz=grouped[:, :3], grouped_feat=grouped[:, 3:]))
            n = m
        self.fused_group = self.be is _ours   # one-pass QueryAndGroup (pdm_query_and_group)

    def run(self, xyz: torch.Tensor, feats) -> Dict[str, torch.Tensor]:
        """xyz (B,N,3) device tensor; feats[i] (B,C_i,N_i) device tensor per layer.
        Returns the per-layer workspaces (device tensors, overwritten by the next call)."""
        be, B = self.be, self.B
        cur, n = xyz, self.N
        for L, ws, feat in zip(self.layers, self.ws, feats):
            m, s, c = L.npoint, L.nsample, L.channels
            ws["temp"].fill_(1e10)                                      # pointnet2_utils.py:26
            be.farthest_point_sampling_wrapper(B, n, m, cur, ws["temp"], ws["fps_idx"])
            ws["xyz_t"].copy_(cur.transpose(1, 2))                      # pointnet2_modules.py:30
            be.gather_points_wrapper(B, 3, n, m, ws["xyz_t"], ws["fps_idx"], ws["new_t"])
            ws["new_xyz"].copy_(ws["new_t"].transpose(1, 2))            # pointnet2_modules.py:32-35
            ws["ball_idx"].zero_()                                      # pointnet2_utils.py:218
            be.ball_query_wrapper(B, n, m, L.radius, s, ws["new_xyz"], cur, ws["ball_idx"])
            if self.fused_group:
                _lib.check(_lib.load().pdm_query_and_group(
                    B, c, n, m, s, 1, cur.data_ptr(), ws["new_xyz"].data_ptr(), feat.data_ptr(),
                    ws["ball_idx"].data_ptr(), ws["grouped"].data_ptr(),
                    torch.cuda.current_stream(self.dev).cuda_stream), "pdm_query_and_group")
            else:                                                       # the reference's sequence
                gx = torch.empty((B, 3, m, s), dtype=torch.float32, device=self.dev)
                be.group_points_wrapper(B, 3, n, m, s, ws["xyz_t"], ws["ball_idx"], gx)
                gx.sub_(ws["new_t"].unsqueeze(-1))                      # pointnet2_utils.py:252
                gf = torch.empty((B, c, m, s), dtype=torch.float32, device=self.dev)
                be.group_points_wrapper(B, c, n, m, s, feat, ws["ball_idx"], gf)
                torch.cat([gx, gf], dim=1, out=ws["grouped"])          # pointnet2_utils.py:257
            cur, n = ws["new_xyz"], m
        return self.ws


class HostSAChain:
    """End-to-end form: host (pinned) buffers in, host results out; copies inside the call.

    `io="points"` (default) is what a serving loop moves per batch: the raw point clouds (B,N,4)
    [x,y,z,intensity] go in -- in the detector they are the only host input (`load_data_to_gpu`,
    pcdet/models/__init__.py:23-36); the SA2 feature tensor stands for the output of SA1's shared MLP,
    which lives on the device, so it is uploaded once -- and the chain's products come out: sampled
    indices and centres of both layers and the final layer's ball-query indices (SA1's ball-query
    indices are an intermediate that grouping consumes on the device).
    `io="full"` moves everything every step: points + the SA2 feature tensor (B,64,4096) in, sampled
    indices, centres and ball-query indices of both layers out."""

    def __init__(self, batch, n_points=16384, layers=KITTI_CHAIN, device="cuda:0", backend=None, io="points"):
        assert io in ("points", "full")
        self.io = io
        self.chain = SAChain(batch, n_points, layers, device, backend)
        dev = self.chain.dev
        self.d_points = torch.empty((batch, n_points, 4), dtype=torch.float32, device=dev)
        self.d_xyz = torch.empty((batch, n_points, 3), dtype=torch.float32, device=dev)
        self.d_feat1 = torch.empty((batch, 1, n_points), dtype=torch.float32, device=dev)
        self.d_feat2 = torch.empty((batch, layers[1].channels, layers[0].npoint), dtype=torch.float32, device=dev)
        self._feat2_from = None
        self.h_out = []
        last = len(self.chain.ws) - 1
        for li, ws in enumerate(self.chain.ws):
            keys = ("fps_idx", "new_xyz", "ball_idx") if (io == "full" or li == last) else ("fps_idx", "new_xyz")
            self.h_out.append({k: torch.empty(ws[k].shape, dtype=ws[k].dtype).pin_memory() for k in keys})
        self.h2d_bytes = self.d_points.numel() * 4 + (self.d_feat2.numel() * 4 if io == "full" else 0)
        self.d2h_bytes = sum(t.numel() * t.element_size() for o in self.h_out for t in o.values())

    def run(self, h_points: torch.Tensor, h_feat2: torch.Tensor):
        self.d_points.copy_(h_points, non_blocking=True)
        if self.io == "full":
            self.d_feat2.copy_(h_feat2, non_blocking=True)
        elif self._feat2_from is not h_feat2:      # device-resident stand-in for SA1's MLP output: uploaded once
            self.d_feat2.copy_(h_feat2, non_blocking=True)
            self._feat2_from = h_feat2
        self.d_xyz.copy_(self.d_points[..., :3])
        self.d_feat1.copy_(self.d_points[..., 3:].transpose(1, 2))
        ws = self.chain.run(self.d_xyz, (self.d_feat1, self.d_feat2))
        for o, w in zip(self.h_out, ws):
            for k, t in o.items():
                t.copy_(w[k], non_blocking=True)
        return self.h_out


class PipelinedSAChain:
    """Throughput form of the chain: consecutive batches go round-robin to `n_streams` CUDA streams,
    each with its own workspaces, so independent batches overlap on the GPU.  One batch keeps only
    ~16 SMs busy for most of its time (FPS runs one CTA per frame), so several batches in flight
    fill the other SMs; per-batch latency is unchanged.  `host=True` adds the pinned-host H2D /
    D2H copies of HostSAChain to every step (on the step's stream, overlapping other steps).

    `capture(slot_args)` records each slot's step (all kernel launches and copies; the library's scratch
    buffer of the slot's stream becomes private to the graph, csrc/capi.cu) into a CUDA graph bound to that slot's static input buffers; `submit()`
    then replays graphs, which removes the ~0.8 ms of Python/launch overhead per step that otherwise
    bounds the pipeline once four or more batches are in flight.

    `fps_mode`: scheduling hint for farthest point sampling while the steps are captured / run
    (`_lib.FPS_MODE_*`, include/pdm_ops.h; passed per call, no process-wide state).  With many batches in flight the THROUGHPUT kernel (several
    frames per SM) gives more frames/s at a longer per-batch latency; results are bit-identical."""

    def __init__(self, batch, n_streams=4, n_points=16384, layers=KITTI_CHAIN, device="cuda:0", host=False, backend=None,
                 fps_mode=None, host_io="points"):
        self.dev = torch.device(device)
        self.host = host
        self.fps_mode = fps_mode
        self.streams = [torch.cuda.Stream(device=self.dev) for _ in range(n_streams)]
        if host:
            self.chains = [HostSAChain(batch, n_points, layers, self.dev, backend, io=host_io) for _ in range(n_streams)]
        else:
            self.chains = [SAChain(batch, n_points, layers, self.dev, backend) for _ in range(n_streams)]
        self.graphs = None
        self.slot_args = None
        self.launches_per_step = None
        self.i = 0
        self._start = torch.cuda.Event()

    def capture(self, slot_args):
        """slot_args[k] = the (static) argument tuple slot k will always run on."""
        from . import _lib
        assert len(slot_args) == len(self.streams)
        self.slot_args, self.graphs = list(slot_args), []
        torch.cuda.synchronize(self.dev)
        if self.fps_mode is not None:
            with _lib.fps_mode(self.fps_mode):    # per call, per thread: the captured graphs keep the kernels chosen now
                self._capture_all(_lib)
        else:
            self._capture_all(_lib)
        torch.cuda.synchronize(self.dev)

    def _capture_all(self, _lib):
        for st, chain, args in zip(self.streams, self.chains, self.slot_args):
            with torch.cuda.stream(st):
                chain.run(*args)                      # warm-up: allocator pools, kernel attributes
                before = _lib.launch_count()
                chain.run(*args)
                self.launches_per_step = _lib.launch_count() - before
            st.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                chain.run(*args)
            self.graphs.append(g)

    def begin(self):
        """Fork: every worker stream waits for what is already queued on the current stream."""
        self._start.record(torch.cuda.current_stream(self.dev))
        for st in self.streams:
            st.wait_event(self._start)
        self.i = 0

    def submit(self, *args):
        """Next step.  With captured graphs the step runs on its slot's bound inputs (args ignored)."""
        k = self.i % len(self.streams)
        self.i += 1
        with torch.cuda.stream(self.streams[k]):
            if self.graphs is not None:
                self.graphs[k].replay()
                return self.chains[k].h_out if self.host else self.chains[k].ws
            return self.chains[k].run(*args)

    def end(self):
        """Join: the current stream waits for every worker stream."""
        cur = torch.cuda.current_stream(self.dev)
        for st in self.streams:
            ev = torch.cuda.Event()
            ev.record(st)
            cur.wait_event(ev)

    @property
    def h2d_bytes(self):
        return self.chains[0].h2d_bytes

    @property
    def d2h_bytes(self):
        return self.chains[0].d2h_bytes
