"""BASELINE configs[2]/[3]: full PDM-SSD inference (backbone + PDM neck + BEV context + hybrid head),
random-init weights, batch 16 per GPU, frames sharded across ranks, detections gathered with NCCL.

    python tools/bench_model.py [--steps 20]          /  torchrun --nproc-per-node N tools/bench_model.py
"""
import argparse, json, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pdm_ssd_b200 import synthetic, _lib
from pdm_ssd_b200.detector import PDMSSD, default_cfg, gather_detections

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=5)
ap.add_argument("--batch", type=int, default=16)
a = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
model = PDMSSD(default_cfg()).to(dev).eval()
batches = [torch.from_numpy(synthetic.to_pcdet_points(synthetic.kitti_batch(a.batch, first_frame=(rank * 4 + p) * a.batch))).to(dev) for p in range(4)]
stage_ms = {}


def run(i, timed=False):
    bd = {"batch_size": a.batch, "points": batches[i % 4]}
    if not timed:
        bd = model(bd)
    else:
        for name, m in zip(("backbone_3d", "pdm_neck", "bev_context", "hybrid_head"), model.module_list):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); bd = m(bd); e1.record()
            stage_ms.setdefault(name, []).append((e0, e1))
    return gather_detections(bd["detections"])


with torch.no_grad():
    for i in range(a.warmup):
        run(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    _lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        det = run(i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches = _lib.launch_count()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    for i in range(5):
        run(i, timed=True)
    torch.cuda.synchronize()
if rank == 0:
    ms = float(ms.item())
    print(json.dumps({"workload": "configs[2/3]: full PDM-SSD KITTI 3-class inference, random-init, batch %d/GPU" % a.batch,
                      "n_gpus": world, "ms_per_step": ms / a.steps, "frames_per_s": world * a.batch * a.steps / (ms * 1e-3),
                      "detections_gathered": list(det.shape), "our_kernel_launches_per_step": launches / a.steps,
                      "stage_ms": {k: float(np.mean([x.elapsed_time(y) for x, y in v])) for k, v in stage_ms.items()}}))
if world > 1:
    dist.destroy_process_group()
