#!/usr/bin/env python
"""bench.py -- frames/s of the set-abstraction operator chain (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the chain over one batch of 16 synthetic KITTI-shaped frames per GPU
(16384 points; FPS 16384->4096->1024, ball query r=0.8/1.6 x32, xyz + feature grouping).
Frames are independent, so ranks just take different frames ("weak" scaling, no data-path
collective); the only collective is the max-reduction of the timing.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (CUDA events on the launch
stream) with STREAMS batches in flight: the steps go round-robin to STREAMS CUDA streams, each
replaying a CUDA graph of one step (one batch keeps only ~16 of 148 SMs busy while it samples, so
independent batches overlap; nothing is skipped, every step is the full chain on its own batch).
`latency` is the same step on a single stream, eager launches.  `e2e` is the pipelined throughput
through the host API (`HostSAChain`): pinned point clouds in, the chain's products out, the H2D/D2H
copies inside every step (`e2e_full_io`: the same with every tensor of the chain crossing PCIe);
`roofline` is for the dominant kernel (farthest point sampling, SA1); `cpu_baseline` is the CPU
oracle (oracle/, a port of the reference kernels' semantics) on this box's host cores.

`--impl reference` times that CPU path alone with all host threads (the reference has no CPU
implementation of these ops -- its ops are CUDA-only -- so the oracle port stands in, as
BASELINE.json's north_star prescribes).  The reference's own CUDA kernels, recompiled for sm_100
(oracle/_ref), are timed next to ours and reported under "reference_cuda" when present.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BATCH = 16
N_POINTS = 16384
METRIC = "frames/s (SA op chain, 16384-pt frames, batch 16 per GPU)"     # BASELINE.json metric, both arms
WORKLOAD = ("configs[1]: pointnet2 SA op chain (FPS 16384->4096->1024, ball query r=0.8/1.6 x32, "
            "xyz+feature grouping C=1/64), KITTI-shaped synthetic frames")
STREAMS = int(os.environ.get("PDM_BENCH_STREAMS", "24"))  # batches in flight (one CUDA stream + graph + input batch + workspace each)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons while the timed regions run (NVML, 5 ms period; nvidia-smi fallback)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.max_sm, self.mask, self._stop_evt = index, [], None, 0, threading.Event()

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while not self._stop_evt.is_set():
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self._stop_evt.wait(0.005)
        except Exception:
            q = "clocks.sm,clocks.max.sm"
            while not self._stop_evt.is_set():
                try:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    a, b = [float(c) for c in out.strip().split(",")]
                    self.sm.append(a)
                    self.max_sm = b
                except Exception:
                    pass
                self._stop_evt.wait(0.1)

    def summary(self):
        self._stop_evt.set()
        self.join(timeout=6)
        reasons = sorted(n for bit, n in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                "reasons": reasons, "samples": len(self.sm)}


def make_host_batches(rank, pool=STREAMS, batch=BATCH):
    from pdm_ssd_b200 import synthetic
    rng = np.random.default_rng(77 + rank)
    out = []
    for p in range(pool):
        frames = synthetic.kitti_batch(batch, N_POINTS, first_frame=(rank * pool + p) * batch)
        feat2 = rng.standard_normal((batch, 64, 4096), dtype=np.float32)
        out.append((frames, feat2))
    return out


def cpu_chain(frames, feat2, threads):
    """The chain on the CPU oracle (numpy in/out).  frames (b,N,4)."""
    import oracle
    oracle.set_threads(threads)
    xyz = np.ascontiguousarray(frames[..., :3])
    feats = [np.ascontiguousarray(frames[..., 3:].transpose(0, 2, 1)), feat2]
    cur = xyz
    for (m, r, s), feat in zip(((4096, 0.8, 32), (1024, 1.6, 32)), feats):
        fi = oracle.fps(cur, m)
        cur_t = np.ascontiguousarray(cur.transpose(0, 2, 1))
        new_t = oracle.gather_points(cur_t, fi)
        new_xyz = np.ascontiguousarray(new_t.transpose(0, 2, 1))
        bi = oracle.ball_query(r, s, cur, new_xyz)
        gx = oracle.group_points(cur_t, bi)
        gx -= new_t[..., None]
        oracle.group_points(np.ascontiguousarray(feat[:len(cur)]), bi)
        cur = new_xyz
    return cur


def run_reference_arm(args, rank, world):
    """CPU arm: the oracle port with all host threads, bounded sample per step."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    cores = os.cpu_count() or 1
    frames_per_step = max(1, min(BATCH, 480 // max(1, args.steps + args.warmup)))
    frames, feat2 = make_host_batches(0, pool=1, batch=frames_per_step)[0]
    for _ in range(min(args.warmup, 1)):
        cpu_chain(frames[:1], feat2[:1], cores)
    for _ in range(max(0, args.warmup - 1)):
        cpu_chain(frames, feat2, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_chain(frames, feat2, cores)
    dt = time.perf_counter() - t0
    value = frames_per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "points_per_frame": N_POINTS,
                   "frames_per_step": frames_per_step},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": "%d frame(s) per step x %d steps, all ops of the chain on the CPU oracle" % (frames_per_step, args.steps)},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=192)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from pdm_ssd_b200 import _lib
    from pdm_ssd_b200.sa_chain import SAChain, HostSAChain, algorithmic_bytes_per_frame

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    _lib.load()  # fail loudly if libpdmops.so is missing
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["NCCL_DEBUG"] = os.environ.get("PDM_NCCL_DEBUG", "WARN")  # keep NCCL's banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)

    from pdm_ssd_b200.sa_chain import PipelinedSAChain
    host = make_host_batches(rank, pool=STREAMS)
    dev_batches = []
    for frames, feat2 in host:
        pts = torch.from_numpy(frames).to(dev)
        dev_batches.append((pts[..., :3].contiguous(), (pts[..., 3:].transpose(1, 2).contiguous(),
                                                        torch.from_numpy(feat2).to(dev))))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput: STREAMS batches in flight, one CUDA graph per slot ----------
    # FPS in THROUGHPUT mode (fps_l2_kernel: 2 frames per SM); the latency pass below uses the default kernel
    pipe = PipelinedSAChain(BATCH, STREAMS, N_POINTS, device=dev, fps_mode=_lib.FPS_MODE_THROUGHPUT)
    pipe.capture(dev_batches)

    def pipelined(p, nsteps):
        p.begin()
        for i in range(nsteps):
            p.submit()
        p.end()

    pipelined(pipe, max(args.warmup, STREAMS))
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    tw0 = time.perf_counter()
    e0.record()
    pipelined(pipe, args.steps)
    e1.record()
    tw1 = time.perf_counter()
    barrier()
    ms_rank = e0.elapsed_time(e1)
    if os.environ.get("PDM_BENCH_DEBUG"):
        print("rank %d: pipelined %.3f ms (events) for %d steps; submit loop %.3f ms wall; total %.3f ms wall"
              % (rank, ms_rank, args.steps, (tw1 - tw0) * 1e3, (time.perf_counter() - tw0) * 1e3), file=sys.stderr, flush=True)
    ms_max = reduce_max(ms_rank)
    value = world * BATCH * args.steps / (ms_max * 1e-3)
    launches = pipe.launches_per_step * args.steps   # our kernels inside the replayed graphs

    # ---- latency: the same step on ONE stream, eager launches; times the dominant kernel ------------
    chain = SAChain(BATCH, N_POINTS, device=dev)
    fps_ev = []
    orig_be = chain.be
    orig_fps = orig_be.farthest_point_sampling_wrapper

    class _Timed:  # thin proxy: CUDA events around SA1's FPS launch on the launch stream
        def __getattr__(self, name):
            return getattr(orig_be, name)

        def farthest_point_sampling_wrapper(self, b, n, m, *a):
            if n != N_POINTS:
                return orig_fps(b, n, m, *a)
            x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x0.record()
            r = orig_fps(b, n, m, *a)
            x1.record()
            fps_ev.append((x0, x1))
            return r
    for i in range(3):
        chain.run(*dev_batches[i % STREAMS])
    chain.be = _Timed()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    l0.record()
    nlat = max(5, min(args.steps, 20))
    for i in range(nlat):
        chain.run(*dev_batches[i % STREAMS])
    l1.record()
    barrier()
    chain.be = orig_be
    latency_ms = reduce_max(l0.elapsed_time(l1)) / nlat
    fps_lat_kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in fps_ev]))
    # the kernel the pipelined region runs (throughput mode: fps_prepare + fps_l2), timed alone the same way
    fps_ev.clear()
    _lib.set_fps_mode(_lib.FPS_MODE_THROUGHPUT)
    tfps = _Timed()
    ws0 = chain.ws[0]
    for i in range(8):
        ws0["temp"].fill_(1e10)
        tfps.farthest_point_sampling_wrapper(BATCH, N_POINTS, ws0["fps_idx"].shape[1], dev_batches[i % STREAMS][0], ws0["temp"], ws0["fps_idx"])
    _lib.set_fps_mode(_lib.FPS_MODE_AUTO)
    torch.cuda.synchronize()
    fps_kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in fps_ev[2:]]))

    # ---- end to end: host pinned buffers in, host results out, copies inside every step ---------------
    # "points": what a serving loop moves per batch (point clouds in, the chain's products out);
    # "full": every tensor of the chain through the host every step (see HostSAChain)
    pinned = [(torch.from_numpy(f).pin_memory(), torch.from_numpy(g).pin_memory()) for f, g in host]
    e2e = {}
    for io in ("points", "full"):
        hpipe = PipelinedSAChain(BATCH, STREAMS, N_POINTS, device=dev, host=True, fps_mode=_lib.FPS_MODE_THROUGHPUT, host_io=io)
        hpipe.capture(pinned)
        pipelined(hpipe, STREAMS)
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        pipelined(hpipe, args.steps)
        h1.record()
        barrier()
        checksum = int(sum(int(c.h_out[1]["fps_idx"].sum().item()) for c in hpipe.chains))  # reads the host copies
        e2e[io] = {"value": world * BATCH * args.steps / (reduce_max(h0.elapsed_time(h1)) * 1e-3), "unit": "frames/s",
                   "h2d_bytes_per_step": hpipe.h2d_bytes, "d2h_bytes_per_step": hpipe.d2h_bytes, "checksum": checksum,
                   "streams": STREAMS}
        del hpipe
        torch.cuda.empty_cache()
    e2e["points"]["what"] = ("pinned host point clouds (B,N,4) in, sampled indices + centres of both layers and the final layer's "
                             "ball-query indices out, copies inside every step; the SA2 feature tensor stands for SA1's MLP output "
                             "and stays on the device")
    e2e["full"]["what"] = "every tensor of the chain through the host every step: points + SA2 feature tensor in, indices/centres/ball indices of both layers out"
    clocks = sampler.summary() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel + chain-level bytes --------------------------------
    peak, peak_src = _peaks()
    ab = algorithmic_bytes_per_frame(N_POINTS)
    fps_bytes = ab["sa1_fps"] * BATCH
    achieved = fps_bytes / (fps_kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "fps_sa1_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    chain_gbs = ab["total"] * BATCH * args.steps / (ms_max * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "batch_per_gpu": BATCH, "points_per_frame": N_POINTS,
                   "streams": STREAMS, "cuda_graphs": True,
                   "fps_mode": "throughput (fps_l2_kernel, 2 frames/SM) in the pipelined and e2e regions; latency pass: on-chip fps_bucket_kernel",
                   "l2": "step working set %.0f MB > 126 MB L2; %d distinct input batches (one per stream slot)" % (ab["total"] * BATCH / 1e6, STREAMS)},
        "roofline": {"kernel": "fps_prepare_kernel + fps_l2_kernel (SA1 farthest point sampling, throughput mode)", "bound": "hbm", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": fps_bytes, "kernel_ms": fps_kernel_ms,
                     "note": "timed alone on one stream after the latency pass; latency/issue-bound by design: 4095 dependent "
                             "argmax rounds per frame, one CTA per frame -> see rounds_per_s; throughput comes from co-resident frames "
                             "and overlapping batches",
                     "rounds_per_s": 4095.0 / (fps_kernel_ms * 1e-3), "latency_mode_kernel_ms": fps_lat_kernel_ms},
        "chain_hbm": {"algorithmic_bytes_per_frame": ab["total"], "achieved_gbs": chain_gbs, "frac": chain_gbs / peak},
        "latency": {"ms_per_step_single_stream": latency_ms, "frames_per_s_single_stream": world * BATCH / (latency_ms * 1e-3),
                    "what": "same step, one stream, eager launches (no graphs, no overlap between batches)"},
        "e2e": e2e["points"], "e2e_full_io": e2e["full"],
        "gpu_launches": int(launches), "clocks": clocks,
    }

    # ---- reference CUDA kernels on the same GPU (informational; the >=10x denominator) -------
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import build_ref
        ref = build_ref.load_ref()
        if ref is not None:
            rchain = SAChain(BATCH, N_POINTS, device=dev, backend=ref)
            for i in range(2):
                rchain.run(*dev_batches[i % STREAMS])
            torch.cuda.synchronize()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            nref = max(3, min(10, args.steps))
            r0.record()
            for i in range(nref):
                rchain.run(*dev_batches[i % STREAMS])
            r1.record()
            torch.cuda.synchronize()
            rms = r0.elapsed_time(r1) / nref
            line["reference_cuda"] = {"value": BATCH / (rms * 1e-3), "unit": "frames/s", "ms_per_step": rms,
                                      "what": "reference pointnet2_batch kernels recompiled for sm_100 (oracle/_ref), same chain, 1 GPU; "
                                              "they launch on the legacy default stream, so batches cannot overlap"}
    except Exception as ex:  # informational only
        line["reference_cuda"] = {"unavailable": str(ex)[:120]}

    # ---- CPU baseline (oracle port) on this box's host cores, bounded sample ------------------
    if world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        cores = os.cpu_count() or 1
        frames, feat2 = host[0]
        cpu_chain(frames[:1], feat2[:1], cores)
        t0 = time.perf_counter()
        cpu_chain(frames, feat2, cores)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": BATCH / dt, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": "one batch of %d frames, whole chain, oracle/pdm_oracle.c with %d threads" % (BATCH, cores)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
