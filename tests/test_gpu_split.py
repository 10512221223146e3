"""GPU tests of the fp32-grade tensor-core mode (YB_MODE_FP32_TC, precision='fp32').

The reference computes the convolution stack in fp32 (darknet.py:27-53, :118).  This mode reproduces it on the tcgen05
tensor pipe: every activation and weight is an fp16 pair hi + lo (22 mantissa bits), three partial products per k-step
accumulated in fp32 in TMEM (csrc/conv_tc.cu, SPLIT instantiations).

Layer level: every layer kind through yb_run_layer (fp32 NHWC at the boundary) against a float64 evaluation of the same
layer.  Tolerance: 2e-6 * max|ref| + 2e-6 * |ref| -- an fp32 implementation of a K <= 9216 dot product is expected at
1e-7 .. 1e-6 relative; a dropped partial product or a wrong hi/lo pairing shows up at 5e-4.
Network level (the north-star bar, BASELINE.json): head logits within 1e-4 * max|logit| of the fp32 oracle at 608x608,
batch 4; decoded scores allclose(atol=1e-4, rtol=1e-6); coordinates at the fp32 noise floor of two 75-layer stacks.
"""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from yolo_v3_b200 import _lib, synth, topology

pytestmark = pytest.mark.gpu


def vp(t):
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.fixture(scope="module")
def sd():
    return synth.make_state_dict(seed=1234, recipe="calibrated")


@pytest.fixture(scope="module")
def split_ctx(sd):
    lib = _lib.load()
    ctx = _lib.create_ctx(0, 80, None)
    for k, v in sd.items():
        if "num_batches" in k:
            continue
        v = v.contiguous()
        _lib.check(lib.yb_set_tensor(ctx, k.encode(), vp(v), v.numel(), 1), ctx)
    _lib.check(lib.yb_finalize(ctx, _lib.YB_MODE_FP32_TC), ctx)
    yield lib, ctx
    lib.yb_destroy(ctx)


def ref_layer64(sd, spec, x_nhwc, res_nhwc):
    """float64 conv + the fused epilogue with the engine's fp32 BN folding, NHWC out."""
    k = spec["key"]
    x = x_nhwc.double().permute(0, 3, 1, 2)
    if spec["bn"]:
        y = F.conv2d(x, sd[k + ".conv.weight"].double(), None, spec["stride"], (spec["ks"] - 1) // 2)
        invstd = 1.0 / torch.sqrt(sd[k + ".bn.running_var"] + 1e-5)
        alpha = invstd * sd[k + ".bn.weight"]
        beta = sd[k + ".bn.bias"] - sd[k + ".bn.running_mean"] * alpha
        y = y * alpha.double().view(1, -1, 1, 1) + beta.double().view(1, -1, 1, 1)
        y = F.leaky_relu(y, 0.1)
    else:
        y = F.conv2d(x, sd[k + ".weight"].double(), None) + sd[k + ".bias"].double().view(1, -1, 1, 1)
    y = y.permute(0, 2, 3, 1).contiguous()
    if res_nhwc is not None:
        y = y + res_nhwc.double()
    return y


# (layer index in cfg order, B, H, W, with_residual)
CASES = [
    (2, 2, 24, 40, False),     # 64->32 1x1: 32-column sub-tiles
    (3, 2, 24, 40, True),      # 32->64 3x3 s1, Cin=32 (64B swizzle), residual
    (1, 2, 48, 40, False),     # 32->64 3x3 s2, Cin=32
    (4, 3, 38, 38, False),     # 64->128 3x3 s2
    (6, 3, 19, 19, True),      # 64->128 3x3 s1 + residual, M=1083 (tail, tiles straddle images)
    (5, 1, 19, 19, False),     # 128->64 1x1
    (11, 2, 20, 12, True),     # 128->256 3x3 + residual
    (43, 2, 38, 38, False),    # 512->1024 3x3 s2 (K = 3 * 4608)
    (44, 3, 19, 19, False),    # 1024->512 1x1
    (45, 3, 19, 19, True),     # 512->1024 3x3 + residual (deepest, N=1024, CTA pairs)
    (58, 3, 19, 19, False),    # head 1024->255 1x1, fp32 out, padded N
    (60, 1, 38, 38, False),    # 768->256 1x1 (concat input width)
    (74, 1, 76, 76, False),    # head 256->255 @ /8
    (4, 4, 304, 304, False),   # 64->128 s2: enough tiles for the BN = 128 pair mode
    (6, 2, 152, 152, True),    # the real stage-1 map + residual
    (10, 2, 76, 76, False),    # 256->128 1x1 at 76^2: resident weights
]


@pytest.mark.parametrize("li,B,H,W,with_res", CASES)
def test_split_layer_vs_float64(split_ctx, sd, li, B, H, W, with_res):
    lib, ctx = split_ctx
    spec = topology.layer_specs(80)[li]
    rs = np.random.RandomState(300 + li)
    x = torch.from_numpy(rs.standard_normal((B, H, W, spec["cin"])).astype(np.float32))
    Ho, Wo = H // spec["stride"], W // spec["stride"]
    res = torch.from_numpy(rs.standard_normal((B, Ho, Wo, spec["cout"])).astype(np.float32)) if with_res else None
    head = not spec["bn"]
    cout_store = (spec["cout"] + 15) // 16 * 16 if head else spec["cout"]
    out = torch.full((B, Ho, Wo, cout_store), float("nan"), device="cuda", dtype=torch.float32)
    xd = x.cuda()
    rd = res.cuda() if with_res else None
    _lib.check(lib.yb_run_layer(ctx, li, vp(xd), B, H, W, vp(rd) if with_res else None, vp(out), stream()), ctx)
    torch.cuda.synchronize()
    y = out.cpu()[..., :spec["cout"]].double()
    ref = ref_layer64(sd, spec, x, res)
    err = (y - ref).abs()
    tol = 2e-6 * ref.abs().max() + 2e-6 * ref.abs()
    bad = (err > tol) | torch.isnan(y)
    # signed diagnostics: a truncating accumulator shows up as a negative slope of the error on the value (shrink toward zero)
    slope = float(((y - ref) * ref).sum() / (ref * ref).sum())
    print(f"layer {li} {spec['key']}: max err {float(err.max()):.3e} = {float(err.max() / ref.abs().max()):.2e} of max|ref|, "
          f"rms {float(err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()):.2e} of rms|ref|, slope {slope:+.2e}")
    assert not bad.any(), (f"layer {li} {spec['key']}: {int(bad.sum())} of {bad.numel()} outside tolerance, "
                           f"max err {float(err.max()):.4g} (ref max {float(ref.abs().max()):.4g}); first bad index "
                           f"{tuple(int(v) for v in bad.nonzero()[0])}")


@pytest.mark.parametrize("B,H,W", [(2, 40, 56), (3, 17, 23), (3, 17, 24), (1, 64, 608), (2, 13, 152), (1, 3, 40)])
def test_split_stem_vs_float64(split_ctx, sd, B, H, W):
    lib, ctx = split_ctx
    spec = topology.layer_specs(80)[0]
    x = torch.from_numpy(np.random.RandomState(7).rand(B, 3, H, W).astype(np.float32))
    out = torch.full((B, H, W, 32), float("nan"), device="cuda", dtype=torch.float32)
    xd = x.cuda()
    _lib.check(lib.yb_run_layer(ctx, 0, vp(xd), B, H, W, None, vp(out), stream()), ctx)
    torch.cuda.synchronize()
    ref = ref_layer64(sd, spec, x.permute(0, 2, 3, 1), None)
    err = (out.cpu().double() - ref).abs()
    assert float(err.max()) <= 2e-6 * float(ref.abs().max()), float(err.max())


def _net(sd, precision, hw):
    from yolo_v3_b200 import YoloNet
    net = YoloNet(hw, precision=precision)
    net.load_state_dict(sd)
    return net.cuda().eval()


def test_split_net_608_meets_north_star_tolerance(sd):
    """The north-star bar at the benched image size: 608x608, batch 4, against the fp32 CPU oracle."""
    from oracle import yolo_oracle as O
    x = synth.make_images(4, 608, 608, seed=11)
    net = _net(sd, "fp32", (608, 608))
    ls = [l.cpu() for l in net.head_logits(x.cuda())]
    ref_ls = O.head_logits(sd, x)
    mx = max(float(r.abs().max()) for r in ref_ls)
    worst = max(float((a - b).abs().max()) for a, b in zip(ls, ref_ls))
    print(f"split mode 608 b4: head logits max|d| = {worst:.3e} = {worst / mx:.2e} of max|logit| {mx:.2f}")
    assert worst <= 1e-4 * mx
    det = torch.cat(net(x.cuda(), None), 1).cpu()
    ref = torch.cat(O.forward(sd, x), 1)
    assert det.shape == (4, 22743, 85)
    d = (det - ref).abs()
    print(f"split mode 608 b4: max|d| xy {float(d[..., :2].max()):.2e} px, wh rel {float((d[..., 2:4] / ref[..., 2:4]).max()):.2e}, "
          f"obj/cls {float(d[..., 4:].max()):.2e}")
    np.testing.assert_allclose(det[..., 4:].numpy(), ref[..., 4:].numpy(), rtol=1e-6, atol=1e-4)       # scores: the stated bar
    # coordinates: sigmoid' * stride amplifies the logit noise; the floor between two fp32 75-layer stacks is ~5e-4 px at 608
    np.testing.assert_allclose(det[..., :2].numpy(), ref[..., :2].numpy(), rtol=0, atol=1e-3)
    np.testing.assert_allclose(det[..., 2:4].numpy(), ref[..., 2:4].numpy(), rtol=5e-4, atol=1e-4)


def test_split_detections_equal_oracle_detections(sd):
    """Survivors of the fp32-grade mode vs the oracle's own end-to-end result at conf 0.5 (the bench thresholds): the two
    candidate sets may differ only where a score sits within 1e-4 of the threshold."""
    from oracle import yolo_oracle as O
    x = synth.make_images(4, 608, 608, seed=12)
    net = _net(sd, "fp32", (608, 608))
    got = net.detect(x.cuda(), 0.5, 0.4)
    ref = O.postprocessing(torch.cat(O.forward(sd, x), 1), 80, 0.5, 0.4)
    assert len(got) == len(ref)
    for g, r in zip(got, ref):
        g = g.cpu()
        assert abs(len(g) - len(r)) <= 1
        if len(g) == len(r) and len(g):
            assert torch.equal(g[:, 6], r[:, 6])                                        # same classes in the same order
            # corners of boxes thousands of pixels wide (random weights): w/h carry the logit noise as a RELATIVE error
            wh = (r[:, 2:4] - r[:, 0:2]).abs().repeat(1, 2).numpy()
            assert (np.abs(g[:, :4].numpy() - r[:, :4].numpy()) <= 2e-3 + 5e-4 * wh).all()
            np.testing.assert_allclose(g[:, 4:6].numpy(), r[:, 4:6].numpy(), rtol=0, atol=1e-4)


def test_split_matches_cuda_core_fp32_path(sd):
    """The tensor-core fp32-grade mode and the CUDA-core fp32 debugging path agree to fp32 noise on a small non-square batch."""
    x = synth.make_images(3, 160, 224, seed=5).cuda()
    a = torch.cat(_net(sd, "fp32", (224, 160))(x, None), 1)
    b = torch.cat(_net(sd, "fp32_simt", (224, 160))(x, None), 1)
    np.testing.assert_allclose(a[..., 4:].cpu().numpy(), b[..., 4:].cpu().numpy(), rtol=1e-6, atol=1e-4)
    np.testing.assert_allclose(a[..., :2].cpu().numpy(), b[..., :2].cpu().numpy(), rtol=0, atol=1e-3)


def test_split_backbone_and_other_class_count(sd):
    """Backbone-only output (hi + lo recombined at the boundary) and a 20-class head (75 channels padded to 80)."""
    from oracle import yolo_oracle as O
    x = synth.make_images(2, 96, 128, seed=3)
    net = _net(sd, "fp32", (128, 96))
    bb = net.backbone(x.cuda()).cpu()
    with torch.no_grad():
        ref = O.backbone(sd, x)[0]
    assert float((bb - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    sd20 = synth.make_state_dict(seed=77, num_classes=20, recipe="analytic")
    from yolo_v3_b200 import YoloNet
    n20 = YoloNet((128, 96), numClass=20, precision="fp32")
    n20.load_state_dict(sd20)
    det = torch.cat(n20.cuda().eval()(x.cuda(), None), 1).cpu()
    ref20 = torch.cat(O.forward(sd20, x, num_classes=20), 1)
    np.testing.assert_allclose(det[..., 4:].numpy(), ref20[..., 4:].numpy(), rtol=1e-6, atol=1e-4)
