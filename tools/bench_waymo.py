"""BASELINE configs[4]: Waymo-scale synthetic frames (163840 points, 150 m range) through the SA chain
(163840 -> 16384 -> 4096) and the PDM neck on the larger dilation grid (376 x 376 x 15), batch 8 per GPU.
Per-op CUDA-event times on one stream; one JSON line.
    python tools/bench_waymo.py [--batch 8] [--iters 3]"""
import argparse, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pdm_ssd_b200 import pdm_neck, pointnet2_batch_cuda as ours, synthetic, _lib
from pdm_ssd_b200.sa_chain import SAChain, SALayerCfg, algorithmic_bytes_per_frame

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda:0")
B, N = a.batch, 163840
LAYERS = (SALayerCfg(16384, 0.8, 32, 2), SALayerCfg(4096, 1.6, 32, 64))
frames = torch.from_numpy(synthetic.waymo_batch(B, N)).to(dev)
xyz = frames[..., :3].contiguous()
feat1 = frames[..., 3:].transpose(1, 2).contiguous()
feat2 = torch.randn(B, 64, 16384, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
chain = SAChain(B, N, LAYERS, dev)


def timed(fn, iters=a.iters):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ops = {}
cur, n = xyz, N
for li, (L, ws, feat) in enumerate(zip(LAYERS, chain.ws, (feat1, feat2)), 1):
    m, s, c = L.npoint, L.nsample, L.channels

    def fps():
        ws["temp"].fill_(1e10)
        ours.farthest_point_sampling_wrapper(B, n, m, cur, ws["temp"], ws["fps_idx"])
    ops["sa%d_fps_%d_to_%d" % (li, n, m)] = timed(fps)
    ws["xyz_t"].copy_(cur.transpose(1, 2))
    ours.gather_points_wrapper(B, 3, n, m, ws["xyz_t"], ws["fps_idx"], ws["new_t"])
    ws["new_xyz"].copy_(ws["new_t"].transpose(1, 2))

    def bq():
        ws["ball_idx"].zero_()
        ours.ball_query_wrapper(B, n, m, L.radius, s, ws["new_xyz"], cur, ws["ball_idx"])
    ops["sa%d_ball_query" % li] = timed(bq)

    def grp():
        _lib.check(_lib.load().pdm_query_and_group(B, c, n, m, s, 1, cur.data_ptr(), ws["new_xyz"].data_ptr(), feat.data_ptr(),
                                                   ws["ball_idx"].data_ptr(), ws["grouped"].data_ptr(),
                                                   torch.cuda.current_stream(dev).cuda_stream), "qg")
    ops["sa%d_query_and_group" % li] = timed(grp)
    cur, n = ws["new_xyz"], m
total = timed(lambda: chain.run(xyz, (feat1, feat2)))
# neck: the 4096 SA2 centres of every frame, 256 channels, 150 m grid
RANGE, VOX = [-75.2, -75.2, -2.0, 75.2, 75.2, 4.0], [0.4, 0.4, 0.4]
grid = [376, 376, 15]
centres = chain.ws[1]["new_xyz"]
coords = torch.cat([torch.arange(B, device=dev, dtype=torch.float32).repeat_interleave(4096)[:, None], centres.reshape(-1, 3)], 1).contiguous()
feats = torch.randn(B * 4096, 256, device=dev)
coef = torch.randn(B * 4096, 9, device=dev) * 0.5
neck_ms = timed(lambda: pdm_neck.neck_forward(coords, feats, coef, B, RANGE, VOX, grid), 10)
ab = algorithmic_bytes_per_frame(N, LAYERS)["total"]
print(json.dumps({"workload": "configs[4]: Waymo-scale synthetic frames, %d points, batch %d, SA %d->16384->4096 + PDM neck on %s" % (N, B, N, grid),
                  "ops_ms": ops, "chain_ms_per_batch": total, "chain_frames_per_s": B / (total * 1e-3),
                  "chain_algorithmic_bytes_per_frame": ab, "chain_hbm_gbs": ab * B / (total * 1e-3) / 1e9,
                  "neck_ms": neck_ms, "neck_out_bytes": B * 256 * 376 * 376 * 4, "neck_gbs": B * 256 * 376 * 376 * 4 / (neck_ms * 1e-3) / 1e9}))
