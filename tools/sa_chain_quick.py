"""SA-chain pipelined throughput (configs[1]) alone: quick A/B harness (same code as bench.py's sa_chain sub-record).
    python tools/sa_chain_quick.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
rec = bench.sa_chain_record(dev, 0, 6530.3)
print(json.dumps({"frames_per_s": rec["frames_per_s"], "ms_per_step": rec["ms_per_step"], "single": rec.get("single_stream_ms_per_step"),
                  "bq_cl": os.environ.get("PDM_BQ_BUILD_CL", "auto")}))
