"""Drop-in proof (INTEGRATION.md option A): the REFERENCE's own Python layers -- pointnet2_utils.py,
pointnet2_modules.py, iou3d_nms_utils.py, model_nms_utils.py, vendored unchanged under oracle/_ref/py by
oracle/build_ref.py -- are imported twice: once with the extension-module slots `pointnet2_batch_cuda` /
`iou3d_nms_cuda` (pointnet2_utils.py:7, iou3d_nms_utils.py:9) filled by OUR shims, once filled by the reference's
compiled kernels (oracle/_ref/*.so).  Same inputs, same weights -> bit-identical outputs."""
import pytest
import torch

from pdm_ssd_b200 import iou3d_nms_cuda as our_nms, pointnet2_batch_cuda as our_pn2, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def trees():
    import build_ref
    ref_pn2, ref_nms = build_ref.load_ref(), build_ref.load_ref_nms()
    ours = build_ref.load_reference_tree("refpy_on_ours", our_pn2, our_nms)
    if ours is None:
        pytest.skip("reference python files not vendored (oracle/_ref/py)")
    ref = build_ref.load_reference_tree("refpy_on_ref", ref_pn2, ref_nms) if ref_pn2 is not None and ref_nms is not None else None
    return ours, ref


@pytest.fixture(autouse=True)
def _fp32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def _cloud(B, N, first=0):
    pts = torch.from_numpy(synthetic.kitti_batch(B, N, first_frame=first)).to(DEV)
    return pts[..., :3].contiguous(), pts[..., 3:].transpose(1, 2).contiguous()


def test_reference_sa_and_fp_modules_run_on_our_extension(trees):
    ours, ref = trees
    if ref is None:
        pytest.skip("oracle/_ref kernels not built")
    xyz, feat = _cloud(2, 4096, first=21)
    outs = []
    for ns in (ours, ref):
        torch.manual_seed(0)
        sa = ns.pointnet2_modules.PointnetSAModuleMSG(npoint=512, radii=[0.8, 1.6], nsamples=[16, 32],
                                                      mlps=[[1, 16, 32], [1, 16, 32]]).to(DEV).eval()
        fp = ns.pointnet2_modules.PointnetFPModule(mlp=[64 + 1, 32]).to(DEV).eval()
        with torch.no_grad():
            new_xyz, new_feat = sa(xyz, feat)                      # FPS, gather, ball query, grouping through the slot
            up = fp(xyz, new_xyz, feat, new_feat)                  # three_nn, three_interpolate
        outs.append((new_xyz, new_feat, up))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_reference_autograd_functions_on_our_extension(trees):
    """The reference's Function classes allocate with torch.cuda.IntTensor / FloatTensor (pointnet2_utils.py:25-26):
    our argument checks accept exactly those, and the backward kernels run through the same slot."""
    ours, ref = trees
    xyz, feat = _cloud(2, 2048, first=3)
    pu = ours.pointnet2_utils
    idx = pu.farthest_point_sample(xyz, 128)
    assert idx.dtype == torch.int32 and idx.shape == (2, 128) and (idx[:, 0] == 0).all()
    f = feat.clone().requires_grad_(True)
    g = pu.gather_operation(f, idx)
    g.sum().backward()
    assert f.grad is not None and float(f.grad.sum()) == 2 * 128
    bq = pu.ball_query(1.0, 16, xyz, xyz[:, :64].contiguous())
    assert bq.shape == (2, 64, 16)
    if ref is not None:
        assert torch.equal(idx, ref.pointnet2_utils.farthest_point_sample(xyz, 128))
        assert torch.equal(bq, ref.pointnet2_utils.ball_query(1.0, 16, xyz, xyz[:, :64].contiguous()))


def test_reference_class_agnostic_nms_on_our_extension(trees):
    ours, ref = trees
    if ref is None:
        pytest.skip("oracle/_ref kernels not built")

    class Cfg(dict):
        __getattr__ = dict.__getitem__
    cfg = Cfg(NMS_TYPE="nms_gpu", NMS_THRESH=0.1, NMS_PRE_MAXSIZE=1024, NMS_POST_MAXSIZE=100)
    g = torch.Generator().manual_seed(4)
    n = 3000
    boxes = torch.cat([torch.rand(n, 2, generator=g) * 40, torch.zeros(n, 1), torch.rand(n, 3, generator=g) * 3 + 1,
                       torch.rand(n, 1, generator=g) * 6.28], 1).to(DEV)
    scores = torch.rand(n, generator=g).to(DEV)
    sel_o, sc_o = ours.model_nms_utils.class_agnostic_nms(scores, boxes, cfg, score_thresh=0.2)
    sel_r, sc_r = ref.model_nms_utils.class_agnostic_nms(scores, boxes, cfg, score_thresh=0.2)
    assert torch.equal(sel_o, sel_r) and torch.equal(sc_o, sc_r) and len(sel_o) > 10
    # IoU values: the overlap construction is ill-conditioned for nearly coincident boxes (tests/test_nms_gpu.py
    # test_iou_vs_oracle_and_reference), so values are compared tightly but not bit for bit; keep lists above are exact
    iou_o = ours.iou3d_nms_utils.boxes_iou3d_gpu(boxes[:300], boxes[300:500])
    iou_r = ref.iou3d_nms_utils.boxes_iou3d_gpu(boxes[:300], boxes[300:500])
    assert float((iou_o - iou_r).abs().max()) < 1e-5
    bev_o = ours.iou3d_nms_utils.boxes_iou_bev(boxes[:200], boxes[100:300])
    bev_r = ref.iou3d_nms_utils.boxes_iou_bev(boxes[:200], boxes[100:300])
    assert float((bev_o - bev_r).abs().max()) < 1e-5 and float(bev_o.diagonal(-100).min()) > 0.999
