#!/usr/bin/env python
"""bench.py -- frames/s of full PDM-SSD inference (BASELINE.json configs[2]; configs[3] when N > 1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one forward of the whole detector over one batch of 16 synthetic KITTI-shaped frames per GPU (16384 points
each): SA backbone 16384 -> 4096 -> 1024 (FPS, ball query, fused grouping + shared MLP + max-pool), PDM neck
(dilation, SH x Gaussian filling, fusion, height compression), BEV context convolutions, hybrid head (heatmap branch,
fused per-point FC stacks + score calibration + box decode), batched rotated NMS -> fixed-shape detections (16,100,9).
Frames are independent, so ranks take different frames ("weak" scaling: 16 frames per GPU per step at every N); with
N > 1 every step ends with the path's one collective, the NCCL all-gather of the detections
(`detector.gather_detections`, replacing the reference's pickle-file merge pcdet/utils/common_utils.py:229-250), inside
the timed region.

Prints ONE JSON line (rank 0, last line of stdout).
  value      device-resident throughput: STREAMS batches in flight, one CUDA graph of the whole forward per slot
             (`pipeline.PipelinedDetector`), CUDA events on the launch stream, max over ranks
  e2e        the same through the host API: every step starts from a PINNED HOST point cloud (H2D inside the step's
             graph) and ends with the detections in pinned host memory (D2H inside the step)
  latency    one batch alone: ONE CUDA graph of the forward replayed on one stream, latency-mode sampling (the eager-launch
             figure rides along)
  roofline   the kernel with the largest SM-time share of a step (tcgen05 convolution 128->128 of the BEV context
             block), timed alone with CUDA events: algorithmic fp32-conv FLOPs / time against the measured bf16 peak
  stage_ms / kernels / sa_chain / neck_config0 / waymo_config4 / stack_family   sub-records (per-stage times, per-kernel
             achieved GB/s or TFLOP/s, the set-abstraction op chain of configs[1] -- round 1's headline --, the neck alone at
             configs[0], the Waymo-scale chain of configs[4] -- at N > 1 on all GPUs at once --, the stacked operator family
             next to the reference's own kernels)
  cpu_baseline   the same detector on this box's host cores (oracle/pdm_model_cpu.py: torch CPU modules + the C oracle
             for the CUDA-only ops), bounded sample
`--impl reference` times that CPU arm alone with all host threads (BASELINE.json north_star prescribes it: the
reference's ops are CUDA-only, its host path is torch).
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
METRIC = "frames/s (full PDM-SSD inference, 16384-pt frames, batch 16 per GPU)"     # BASELINE.json metric, both arms
WORKLOAD = ("configs[2]: full PDM-SSD KITTI 3-class inference (SA backbone 16384->4096->1024 + PDM neck + BEV context + "
            "hybrid heatmap/point head + rotated NMS), random-init weights, batch 16 per GPU; with N > 1 = configs[3]: frames "
            "sharded by rank, detections all-gathered over NCCL every step")
STREAMS = int(os.environ.get("PDM_BENCH_STREAMS", "16"))       # batches in flight for the model (measured: 6 -> 7.80k, 8 -> 8.18k,
                                                                # 12 -> 8.54k, 16 -> 8.66k, 24 -> 8.74k frames/s on one B200)
SA_STREAMS = int(os.environ.get("PDM_BENCH_SA_STREAMS", "24"))  # for the SA-chain sub-record


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d["bf16_tflops"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, 1400.0, "fallback (B200_PROFILING.md)"


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


def host_points(rank, slot, batch=BATCH):
    """(batch*N, 5) float32 pcdet `points` of `batch` synthetic frames, distinct per (rank, slot)."""
    from pdm_ssd_b200 import synthetic
    return synthetic.to_pcdet_points(synthetic.kitti_batch(batch, N_POINTS, first_frame=(rank * 64 + slot) * batch))


def make_host_batches(rank, pool, batch=BATCH):
    """SA-chain inputs (configs[1] sub-record; also used by tests/test_multirank_cpu.py)."""
    from pdm_ssd_b200 import synthetic
    rng = np.random.default_rng(77 + rank)
    out = []
    for p in range(pool):
        frames = synthetic.kitti_batch(batch, N_POINTS, first_frame=(rank * pool + p) * batch)
        out.append((frames, rng.standard_normal((batch, 64, 4096), dtype=np.float32)))
    return out


def build_model(device):
    import torch
    from pdm_ssd_b200.detector import PDMSSD, default_cfg
    torch.manual_seed(0)                                   # random-init weights, torch.manual_seed(0), eval mode
    return PDMSSD(default_cfg(N_POINTS)).to(device).eval()


def cpu_model_arm(frames_per_step, steps, warmup, threads):
    """The detector on the host CPU (oracle/pdm_model_cpu.py).  Returns (frames/s, seconds per step)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pdm_model_cpu
    model = build_model("cpu")
    pts = torch.from_numpy(host_points(0, 63, batch=frames_per_step))
    for _ in range(warmup):
        pdm_model_cpu.cpu_forward(model, pts, frames_per_step, threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        pdm_model_cpu.cpu_forward(model, pts, frames_per_step, threads=threads)
    dt = time.perf_counter() - t0
    return frames_per_step * steps / dt, dt / steps


def run_reference_arm(args, rank, world):
    """`--impl reference`: the CPU arm with all host threads, bounded sample per step; rank 0 alone works."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    fps_step = 2
    steps = max(1, min(args.steps, 60))                    # ~0.5 s per step: the whole run stays within a few minutes
    value, sec = cpu_model_arm(fps_step, steps, min(args.warmup, 2), cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 2), "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "points_per_frame": N_POINTS, "frames_per_step": fps_step},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": "%d frames per step x %d steps: the detector's torch modules on CPU + the C oracle for the "
                                   "CUDA-only pointnet2 / NMS ops + the torch neck oracle (oracle/pdm_model_cpu.py)" % (fps_step, steps)},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=96)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-subrecords", action="store_true", help="skip the SA-chain / neck / per-kernel sub-records")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    # 16 slot streams + NCCL's stream on the default 8 hardware work queues alias each other: a 35 KB all-gather then waits
    # behind whole forward graphs of unrelated slots (measured at 2 GPUs: 16.1k frames/s with 8 queues, 17.3k with 32; without
    # any gather 17.5k).  Must be set before the CUDA context exists.
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    import torch
    import torch.distributed as dist
    from pdm_ssd_b200 import _lib
    from pdm_ssd_b200.pipeline import PipelinedDetector

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    _lib.load()  # fail loudly if libpdmops.so is missing
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's verbosity (NCCL_DEBUG ...) is left exactly as the launcher set it; its log only moves from stdout to
        # stderr, because NCCL also prints at communicator teardown and at process exit -- after the JSON line, which must
        # stay the LAST line of stdout
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def pipelined(p, nsteps):
        p.begin()
        for _ in range(nsteps):
            p.submit()
        p.end()

    def timed(p, nsteps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        pipelined(p, nsteps)
        e1.record()
        barrier()
        return reduce_max(e0.elapsed_time(e1))

    model = build_model(dev)
    host = [host_points(rank, s) for s in range(STREAMS)]

    # ---- device-resident throughput: STREAMS batches in flight, inputs resident in HBM ----------------------------
    pipe = PipelinedDetector(model, BATCH, N_POINTS, STREAMS, dev, host=False, gather=(world > 1 and not os.environ.get("PDM_BENCH_NO_GATHER")))
    pipe.capture([torch.from_numpy(h).to(dev) for h in host])
    pipelined(pipe, max(args.warmup, STREAMS))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_max = timed(pipe, args.steps)
    value = world * BATCH * args.steps / (ms_max * 1e-3)
    launches = pipe.launches_per_step * args.steps       # our kernels inside the replayed graphs
    det_shape = list(pipe.det[0].shape)
    gathered_shape = list(pipe.gathered[0].shape) if (world > 1 and pipe.gathered[0] is not None) else None

    # ---- end to end: pinned host points in, detections in pinned host memory out, copies inside every step --------
    pinned = [torch.from_numpy(h).pin_memory() for h in host]
    hpipe = PipelinedDetector(model, BATCH, N_POINTS, STREAMS, dev, host=True, gather=(world > 1 and not os.environ.get("PDM_BENCH_NO_GATHER")))
    hpipe.capture(pinned)
    pipelined(hpipe, max(args.warmup, STREAMS))
    e2e_ms = timed(hpipe, args.steps)
    checksum = float(sum(float(h[..., 7].sum()) for h in hpipe.h_det))       # reads the host copies of the detections
    e2e = {"value": world * BATCH * args.steps / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": hpipe.h2d_bytes,
           "d2h_bytes_per_step": hpipe.d2h_bytes, "ms_per_step": e2e_ms / args.steps, "streams": STREAMS, "checksum": checksum,
           "what": "PDMSSD.forward through pipeline.PipelinedDetector(host=True): pinned host (B*N,5) points -> H2D -> full forward "
                   "-> NMS -> detections (B,100,9) -> D2H to pinned host, every step; with N > 1 the all-gathered detections of all "
                   "ranks are copied to rank 0's host as well"}
    del hpipe

    # ---- latency: one batch alone, one stream, eager launches; per-stage CUDA events ----------------------------------
    dev_pts = [torch.from_numpy(h).to(dev) for h in host[:2]]
    stage_ev = {}
    names = ("backbone_3d", "pdm_neck", "bev_context", "hybrid_head")
    with torch.no_grad():
        for i in range(3):
            model({"batch_size": BATCH, "points": dev_pts[i % 2]})
        nlat = max(5, min(args.steps, 20))
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        l0.record()
        for i in range(nlat):
            model({"batch_size": BATCH, "points": dev_pts[i % 2]})
        l1.record()
        barrier()
        eager_ms = reduce_max(l0.elapsed_time(l1)) / nlat
    # the same as ONE CUDA graph replayed on one stream (latency-mode sampling): what a batch alone costs on the GPU, without
    # the host's ~70 eager launches per forward in the way (on a busy host the eager figure doubles, the graph one does not)
    lpipe = PipelinedDetector(model, BATCH, N_POINTS, 1, dev, host=False, gather=False, fps_mode=_lib.FPS_MODE_LATENCY)
    lpipe.capture([dev_pts[0]])
    pipelined(lpipe, 3)
    latency_ms = timed(lpipe, nlat) / nlat
    del lpipe
    with torch.no_grad():
        for i in range(5):
            bd = {"batch_size": BATCH, "points": dev_pts[i % 2], "pdm_fused_dense": True}
            for name, m in zip(names, model.module_list):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                bd = m(bd)
                b.record()
                stage_ev.setdefault(name, []).append((a, b))
        torch.cuda.synchronize()
    stage_ms = {k: float(np.mean([a.elapsed_time(b) for a, b in v[1:]])) for k, v in stage_ev.items()}
    clocks = sampler.summary() if rank == 0 else None

    # configs[4] at N GPUs: every rank runs the Waymo-scale chain on its own 8 frames at the same time (a child process per
    # GPU, the same tool as the 1-GPU sub-record); rank 0 aggregates frames over the slowest rank's time
    waymo_multi = None
    if world > 1 and not args.no_subrecords:
        dist.barrier()
        vis = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v]
        mine = tool_record("tools/bench_waymo.py", ["--batch", "8", "--iters", "2"],
                           env=dict(os.environ, CUDA_VISIBLE_DEVICES=vis[local_rank] if local_rank < len(vis) else str(local_rank)))
        recs = [None] * world
        dist.all_gather_object(recs, mine)
        if rank == 0:
            ok = [r for r in recs if isinstance(r, dict) and "chain_ms_per_batch" in r]
            if len(ok) == world:
                slow = max(r["chain_ms_per_batch"] for r in ok)
                waymo_multi = {"workload": ok[0]["workload"] + ", per GPU, %d GPUs at once" % world,
                               "chain_ms_per_batch_max_over_ranks": slow, "chain_frames_per_s": world * 8 / (slow * 1e-3),
                               "per_rank_chain_ms": [r["chain_ms_per_batch"] for r in ok], "ops_ms_rank0": ok[0]["ops_ms"],
                               "neck_ms_rank0": ok[0]["neck_ms"]}
            else:
                waymo_multi = {"error": [r for r in recs if r not in ok][:1]}

    # every collective is behind us: all ranks leave the process group together, rank 0 goes on alone
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
    if rank != 0:
        return

    # ================= rank 0 only from here (single-GPU measurements and the JSON line) ===========================
    hbm_peak, tf_peak, tf_sustained, peak_src = _peaks()
    from pdm_ssd_b200.conv_tc import SplitAct

    def time_alone(fn, iters=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    # ---- roofline of the dominant kernel: conv_tc_kernel, BEV context 128 -> 128 on (16,128,200,176) ----------------
    with torch.no_grad():
        bd = model.map_to_bev_module(model.backbone_3d({"batch_size": BATCH, "points": dev_pts[0], "pdm_fused_dense": True}))
        bev_split = bd["spatial_features_split"]
        ctx_layers = model.backbone_2d._packed(dev)
        conv_ms = time_alone(lambda: ctx_layers[0](bev_split, want_split=True))
    Yb, Xb, Cb = bev_split.Y, bev_split.X, bev_split.C
    conv_flops = 2.0 * BATCH * Yb * Xb * Cb * Cb * 9
    achieved_tf = conv_flops / (conv_ms * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "conv_tc_r2_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    dense_ms = 2 * conv_ms
    roofline = {"kernel": "conv_tc_kernel (BEV context 3x3 conv 128->128 + BN + ReLU on (16,128,200,176), tcgen05 implicit GEMM)",
                "bound": "tensor", "achieved": achieved_tf, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved_tf / tf_peak,
                "traffic": traffic, "peak_source": peak_src + ", burst figure (kernel timed alone)",
                "algorithmic_flops_per_launch": conv_flops, "kernel_ms": conv_ms,
                "executed_tflops": 3 * achieved_tf, "executed_frac": 3 * achieved_tf / tf_peak,
                "note": "algorithmic = fp32 convolution FLOPs (2*B*Y*X*Cout*Cin*9).  To stay inside the 1e-3 fp32 budget every product is "
                        "formed from bf16 hi/lo halves as hi*hi + lo*hi + hi*lo, so the tensor cores EXECUTE 3x the algorithmic FLOPs: "
                        "frac is bounded by 1/3 by construction, executed_frac is the utilisation of the pipe",
                "share_of_step": "2 launches x %.3f ms of the %.2f ms single-stream step" % (conv_ms, latency_ms)}

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "points_per_frame": N_POINTS, "streams": STREAMS, "cuda_graphs": True,
                   "weights": "random init, torch.manual_seed(0), eval mode", "classes": 3,
                   "dense_layers": "tcgen05 implicit-GEMM convolutions, fp32 carried as bf16 hi/lo pairs (3 MMAs per product)",
                   "fps_mode": "throughput (fps_l2_kernel) in the pipelined and e2e regions; latency pass: on-chip fps_bucket_kernel",
                   "collective": ("all_gather_into_tensor of detections %s -> %s every step (NCCL)" % (det_shape, gathered_shape)) if world > 1 else None,
                   "l2": "per-step activations (BEV maps 4 x 288 MB) exceed the 126 MB L2; %d distinct input batches (one per slot)" % STREAMS},
        "roofline": roofline,
        "latency": {"ms_per_step_single_stream": latency_ms, "frames_per_s_single_stream": world * BATCH / (latency_ms * 1e-3),
                    "eager_ms_per_step": eager_ms,
                    "what": "same forward, one batch alone: one CUDA graph replayed on one stream, latency-mode sampling "
                            "(eager_ms_per_step: the same with eager launches, host-dependent)"},
        "stage_ms": stage_ms,
        "e2e": e2e, "gpu_launches": int(launches), "gpu_launches_per_step": int(pipe.launches_per_step), "clocks": clocks,
    }

    if world == 1 and not args.no_subrecords:
        try:
            line["kernels"] = kernel_records(model, dev, dev_pts[0], hbm_peak, tf_peak)
        except Exception as ex:  # sub-records are informational
            line["kernels"] = {"error": repr(ex)[:200]}
        try:
            line["sa_chain"] = sa_chain_record(dev, rank, hbm_peak)
        except Exception as ex:
            line["sa_chain"] = {"error": repr(ex)[:200]}
        torch.cuda.empty_cache()
        # the other BASELINE configs, each by its own tool in a child process (bounded; informational sub-records)
        line["neck_config0"] = tool_record("tools/bench_neck.py", [])                               # configs[0]
        line["waymo_config4"] = tool_record("tools/bench_waymo.py", ["--batch", "8", "--iters", "2"])   # configs[4], one GPU's share
        line["stack_family"] = tool_record("tools/bench_stack.py", [])      # pointnet2_stack ops next to the reference's kernels
    if waymo_multi is not None:
        line["waymo_config4"] = waymo_multi

    # ---- CPU baseline: the same detector on this box's host cores, bounded sample ---------------------------------------
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, sec = cpu_model_arm(2, 6, 1, cores)
        line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": "6 steps x 2 frames of the full detector on CPU (oracle/pdm_model_cpu.py: torch modules + C oracle "
                                          "for the CUDA-only ops), %d threads, %.2f s per step" % (cores, sec)}
    print(json.dumps(line), flush=True)


def tool_record(script, argv, timeout=240, env=None):
    """Run one of the per-config tools and return the JSON line it prints (or the error)."""
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, script)] + argv, capture_output=True, text=True, timeout=timeout, env=env)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": (out.stderr or out.stdout)[-300:]}
    except Exception as ex:
        return {"error": repr(ex)[:200]}


def kernel_records(model, dev, pts, hbm_peak, tf_peak):
    """Per-kernel achieved bandwidth / throughput on the model's own tensors (batch 16), each timed alone with CUDA
    events: the 'op HBM GB/s vs peak' half of BASELINE.json's metric.  Algorithmic bytes per SURVEY 8d."""
    import torch
    from pdm_ssd_b200 import pointnet2_utils as pu
    from pdm_ssd_b200.pdm_neck import neck_forward
    out = {}

    def t(fn, iters=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    def rec(name, ms, nbytes=None, flops=None):
        r = {"ms": ms}
        if nbytes is not None:
            r.update(algorithmic_bytes=int(nbytes), achieved_gbs=nbytes / ms / 1e6, frac_of_hbm_peak=nbytes / ms / 1e6 / hbm_peak)
        if flops is not None:
            r.update(algorithmic_flops=float(flops), achieved_tflops=flops / ms / 1e9, frac_of_bf16_peak=flops / ms / 1e9 / tf_peak)
        out[name] = r

    B = BATCH
    with torch.no_grad():
        xyz = pts[:, 1:4].contiguous().view(B, -1, 3)
        feat = pts[:, 4:].contiguous().view(B, -1, 1).permute(0, 2, 1).contiguous()
        sa1, sa2 = model.backbone_3d.SA_modules
        n, m1, m2 = xyz.shape[1], sa1.npoint, sa2.npoint
        rec("fps_sa1 (latency mode)", t(lambda: pu.farthest_point_sample(xyz, m1)), B * (12 * n + 4 * m1))
        idx1 = pu.farthest_point_sample(xyz, m1)
        new1 = pu.gather_operation(xyz.transpose(1, 2).contiguous(), idx1).transpose(1, 2).contiguous()
        g1 = sa1.groupers[0]
        rec("ball_query_sa1", t(lambda: pu.ball_query(g1.radius, g1.nsample, xyz, new1)), B * (12 * n + 12 * m1 + 4 * m1 * g1.nsample))
        rec("sa1 (fps + gather + ball query + fused group/MLP/max-pool)", t(lambda: sa1(xyz, feat)))
        x1, f1 = sa1(xyz, feat)
        rec("fps_sa2 (latency mode)", t(lambda: pu.farthest_point_sample(x1, m2)), B * (12 * m1 + 4 * m2))
        rec("sa2 (fps + gather + ball query + fused group/MLP/max-pool)", t(lambda: sa2(x1, f1)))
        bd = model.backbone_3d({"batch_size": B, "points": pts})
        neck = model.map_to_bev_module
        feats, coords = bd["point_features"].contiguous(), bd["point_coords"].contiguous()
        coef = neck.coef(feats).contiguous()
        args = (coords, feats, coef, B, neck.point_cloud_range, neck.voxel_size, neck.grid_size, neck.dilation, neck.sh_degree, neck.sigma, neck.eps)
        X, Y = neck.grid_size[0], neck.grid_size[1]
        C = feats.shape[1]
        nb = feats.numel() * 4 + coords.numel() * 4 + B * C * Y * X * 4
        rec("pdm_neck (split NHWC8 output)", t(lambda: neck_forward(*args, output="split")), nb)
        rec("pdm_neck (fp32 NCHW output)", t(lambda: neck_forward(*args)), nb)
        split = neck_forward(*args, output="split")
        ctx = model.backbone_2d._packed(dev)
        head = model.dense_head._packed(dev)
        px = B * Y * X
        rec("conv_tc 128->128 (bev context)", t(lambda: ctx[0](split, want_split=True)), 2 * px * C * 4, 2.0 * px * C * C * 9)
        s2, _ = ctx[1](ctx[0](split, want_split=True)[0], want_split=True)
        rec("conv_tc 128->64 (shared conv)", t(lambda: head.shared(s2, want_split=True)), px * (C + 64) * 4, 2.0 * px * 64 * C * 9)
        xs, _ = head.shared(s2, want_split=True)
        rec("conv_tc 64->64 (heatmap conv 1)", t(lambda: head.hm1(xs, want_split=True)), px * 128 * 4, 2.0 * px * 64 * 64 * 9)
        hs, _ = head.hm1(xs, want_split=True)
        rec("conv_tc 64->3 (heatmap conv 2, sigmoid)", t(lambda: head.hm2(hs, want_split=False, want_nchw=True)), px * (64 + 3) * 4, 2.0 * px * 3 * 64 * 9)
        _, hm = head.hm2(hs, want_split=False, want_nchw=True)
        P = feats.shape[0]
        rec("point_head (fused gather + FC stacks + score + decode)", t(lambda: model.dense_head._point_head_fused(head, coords, feats, xs, hm)),
            P * (C + 64 + 4 + 3 + 3 + 7 + 2) * 4, 2.0 * P * ((C + 64) * (head.hc + head.hb) + head.hc * 3 + head.hb * 8))
    return out


def sa_chain_record(dev, rank, hbm_peak):
    """BASELINE configs[1] (round 1's headline): the set-abstraction op chain alone, pipelined with CUDA graphs, next to
    the reference's own CUDA kernels recompiled for sm_100 (oracle/_ref) through the same host code."""
    import torch
    from pdm_ssd_b200 import _lib
    from pdm_ssd_b200.sa_chain import PipelinedSAChain, SAChain, algorithmic_bytes_per_frame
    host = make_host_batches(rank, pool=SA_STREAMS)
    dev_batches = []
    for frames, feat2 in host:
        pts = torch.from_numpy(frames).to(dev)
        dev_batches.append((pts[..., :3].contiguous(), (pts[..., 3:].transpose(1, 2).contiguous(), torch.from_numpy(feat2).to(dev))))
    pipe = PipelinedSAChain(BATCH, SA_STREAMS, N_POINTS, device=dev, fps_mode=_lib.FPS_MODE_THROUGHPUT)
    pipe.capture(dev_batches)

    def run(n):
        pipe.begin()
        for _ in range(n):
            pipe.submit()
        pipe.end()
    run(SA_STREAMS)
    torch.cuda.synchronize()
    steps = 4 * SA_STREAMS
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run(steps)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    ab = algorithmic_bytes_per_frame(N_POINTS)
    rec = {"workload": "configs[1]: FPS 16384->4096->1024, ball query r=0.8/1.6 x32, xyz + feature grouping (C=1/64), batch 16",
           "frames_per_s": BATCH / (ms * 1e-3), "ms_per_step": ms, "streams": SA_STREAMS,
           "algorithmic_bytes_per_frame": ab["total"], "achieved_gbs": ab["total"] * BATCH / ms / 1e6,
           "frac_of_hbm_peak": ab["total"] * BATCH / ms / 1e6 / hbm_peak}
    chain = SAChain(BATCH, N_POINTS, device=dev)
    for i in range(3):
        chain.run(*dev_batches[i])
    torch.cuda.synchronize()
    a.record()
    for i in range(10):
        chain.run(*dev_batches[i % SA_STREAMS])
    b.record()
    torch.cuda.synchronize()
    rec["single_stream_ms_per_step"] = a.elapsed_time(b) / 10
    del pipe
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import build_ref
        ref = build_ref.load_ref()
        if ref is not None:
            rchain = SAChain(BATCH, N_POINTS, device=dev, backend=ref)
            for i in range(2):
                rchain.run(*dev_batches[i])
            torch.cuda.synchronize()
            a.record()
            for i in range(5):
                rchain.run(*dev_batches[i])
            b.record()
            torch.cuda.synchronize()
            rms = a.elapsed_time(b) / 5
            rec["reference_cuda"] = {"frames_per_s": BATCH / (rms * 1e-3), "ms_per_step": rms,
                                     "what": "the reference's pointnet2_batch kernels recompiled for sm_100 (oracle/_ref), same chain, same host "
                                             "code; they launch on the legacy default stream, so batches cannot overlap"}
            rec["speedup_vs_reference_cuda"] = {"pipelined": rms / ms, "single_stream": rms / rec["single_stream_ms_per_step"]}
    except Exception as ex:
        rec["reference_cuda"] = {"unavailable": repr(ex)[:160]}
    return rec


if __name__ == "__main__":
    main()
