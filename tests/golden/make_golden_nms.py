"""Golden vectors for rotated BEV IoU / NMS from the REFERENCE's own CUDA kernels
(oracle/_ref/iou3d_nms_cuda_ref.so = pcdet/ops/iou3d_nms compiled unmodified by oracle/build_ref.py).

    gpurun -- python tests/golden/make_golden_nms.py gpurun_out/golden   # on the GPU box
    cp gpurun_out/golden/nms_*.npz tests/golden/

Inputs are stored next to the reference outputs (IoU matrix of boxes_iou_bev_gpu, keep list of
nms_gpu on score-sorted boxes)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def box_cases():
    """name -> (boxes (N,7) in score order, thresh)"""
    from pdm_ssd_b200 import synthetic
    cases = {}
    cases["scene_n300_t0p1"] = (synthetic.random_boxes(300, seed=1), 0.1)
    cases["scene_n1000_t0p7"] = (synthetic.random_boxes(1000, seed=2, clusters=40), 0.7)
    cases["dense_n512_t0p01"] = (synthetic.random_boxes(512, seed=3, extent=12.0), 0.01)
    cases["clusters_n777_t0p25"] = (synthetic.random_boxes(777, seed=4, clusters=25), 0.25)
    b = synthetic.random_boxes(128, seed=5)
    b[64:] = b[:64]                      # exact duplicates: IoU == 1 pairs
    b[:, 6] = np.round(b[:, 6] / (np.pi / 2)) * (np.pi / 2)   # axis-aligned: parallel / collinear edges
    cases["aligned_dups_n128_t0p5"] = (b.astype(np.float32), 0.5)
    cases["tiny_n3_t0p1"] = (synthetic.random_boxes(3, seed=6, extent=2.0), 0.1)
    cases["one_n1_t0p1"] = (synthetic.random_boxes(1, seed=7), 0.1)
    cases["big_n2500_t0p1"] = (synthetic.random_boxes(2500, seed=8, clusters=60), 0.1)
    return cases


def run(out_dir):
    import build_ref
    ref = build_ref.load_ref_nms()
    assert ref is not None, "oracle/_ref/iou3d_nms_cuda_ref.so missing"
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")
    for name, (boxes, thresh) in box_cases().items():
        b = torch.from_numpy(boxes).to(dev)
        n = b.shape[0]
        keep = torch.zeros(n, dtype=torch.int64)
        num = ref.nms_gpu(b, keep, float(thresh))
        m = min(n, 400)                                   # IoU matrix of the first 400 boxes
        iou = torch.zeros((m, m), device=dev)
        ref.boxes_iou_bev_gpu(b[:m].contiguous(), b[:m].contiguous(), iou)
        ovl = torch.zeros((m, m), device=dev)
        ref.boxes_overlap_bev_gpu(b[:m].contiguous(), b[:m].contiguous(), ovl)
        torch.cuda.synchronize()
        np.savez_compressed(os.path.join(out_dir, "nms_%s.npz" % name), boxes=boxes, thresh=np.float32(thresh),
                            keep=keep[:num].numpy().astype(np.int32), iou=iou.cpu().numpy(), overlap=ovl.cpu().numpy())
        print(name, n, "kept", num)


if __name__ == "__main__":
    run(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
