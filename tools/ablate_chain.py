"""Which part of the chain bounds the pipelined throughput?  Replays the bench's pipelined region with
parts of the step left out (their outputs are kept from a full warm-up run, so the remaining ops see
the same data): full / no FPS / no ball query / no grouping / FPS only.
    python tools/ablate_chain.py [steps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from pdm_ssd_b200 import _lib, sa_chain
from pdm_ssd_b200.sa_chain import PipelinedSAChain, SAChain

dev = torch.device("cuda:0")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 96
S = bench.STREAMS
SKIP = set()


class AblatedChain(SAChain):
    def run(self, xyz, feats):
        be, B = self.be, self.B
        cur, n = xyz, self.N
        for L, ws, feat in zip(self.layers, self.ws, feats):
            m, s, c = L.npoint, L.nsample, L.channels
            if "fps" not in SKIP:
                ws["temp"].fill_(1e10)
                be.farthest_point_sampling_wrapper(B, n, m, cur, ws["temp"], ws["fps_idx"])
            if "glue" not in SKIP:
                ws["xyz_t"].copy_(cur.transpose(1, 2))
                be.gather_points_wrapper(B, 3, n, m, ws["xyz_t"], ws["fps_idx"], ws["new_t"])
                ws["new_xyz"].copy_(ws["new_t"].transpose(1, 2))
            if "bq" not in SKIP:
                ws["ball_idx"].zero_()
                be.ball_query_wrapper(B, n, m, L.radius, s, ws["new_xyz"], cur, ws["ball_idx"])
            if "group" not in SKIP:
                _lib.check(_lib.load().pdm_query_and_group(
                    B, c, n, m, s, 1, cur.data_ptr(), ws["new_xyz"].data_ptr(), feat.data_ptr(),
                    ws["ball_idx"].data_ptr(), ws["grouped"].data_ptr(),
                    torch.cuda.current_stream(self.dev).cuda_stream), "pdm_query_and_group")
            cur, n = ws["new_xyz"], m
        return self.ws


sa_chain.SAChain = AblatedChain
host = bench.make_host_batches(0, pool=S)
args = []
for frames, feat2 in host:
    pts = torch.from_numpy(frames).to(dev)
    args.append((pts[..., :3].contiguous(), (pts[..., 3:].transpose(1, 2).contiguous(), torch.from_numpy(feat2).to(dev))))
for name, skip in [("full", ()), ("no fps", ("fps",)), ("no ball query", ("bq",)), ("no grouping", ("group",)),
                   ("no bq, no grouping", ("bq", "group")), ("fps only", ("bq", "group", "glue")), ("bq only", ("fps", "group", "glue")),
                   ("grouping only", ("fps", "bq", "glue"))]:
    SKIP.clear()
    pipe = PipelinedSAChain(16, S, device=dev, fps_mode=_lib.FPS_MODE_THROUGHPUT)
    for ch, a in zip(pipe.chains, args):      # full run first: every workspace holds valid indices
        ch.run(*a)
    torch.cuda.synchronize()
    SKIP.update(skip)
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
    ms = e0.elapsed_time(e1) / steps
    print("%-22s %.3f ms/step  %.0f frames/s" % (name, ms, 16e3 / ms), flush=True)
    del pipe
    torch.cuda.empty_cache()
