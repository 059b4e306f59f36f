"""Throughput of the SA chain with 1..N batches in flight, eager launches vs CUDA graphs
(device-resident and host end-to-end).  SWEEP_STREAMS="12,16,24" / SWEEP_GRAPHS_ONLY=1 narrow the
sweep; CUDA_DEVICE_MAX_CONNECTIONS (hardware work queues, default 8) is passed through."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from pdm_ssd_b200.sa_chain import PipelinedSAChain
dev = torch.device("cuda:0")
steps = 96
SS = tuple(int(x) for x in os.environ.get("SWEEP_STREAMS", "1,2,4,6,8,12").split(","))
GG = (True,) if os.environ.get("SWEEP_GRAPHS_ONLY") else (False, True)
print("CUDA_DEVICE_MAX_CONNECTIONS =", os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"))
for hostmode in (False, True):
    for graphs in GG:
        for S in SS:
            host = bench.make_host_batches(0, pool=S)
            if hostmode:
                args = [(torch.from_numpy(f).pin_memory(), torch.from_numpy(g).pin_memory()) for f, g in host]
            else:
                args = []
                for frames, feat2 in host:
                    pts = torch.from_numpy(frames).to(dev)
                    args.append((pts[..., :3].contiguous(), (pts[..., 3:].transpose(1, 2).contiguous(), torch.from_numpy(feat2).to(dev))))
            pipe = PipelinedSAChain(16, S, device=dev, host=hostmode)
            if graphs:
                pipe.capture(args)
            for rep in range(2):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                pipe.begin()
                for i in range(steps):
                    pipe.submit(*args[i % S])
                pipe.end()
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print("host=%s graphs=%s streams=%2d  %.3f ms/step  %.0f frames/s" % (hostmode, graphs, S, ms / steps, 16 * steps / (ms * 1e-3)), flush=True)
            del pipe, args
            torch.cuda.empty_cache()
