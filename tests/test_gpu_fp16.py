"""GPU tests of the tcgen05/TMA convolution path (YB_MODE_FP16).

Layer level: every kind of layer the network has (1x1, 3x3 stride 1, 3x3 stride 2, Cin=32 64B-swizzle
variant, residual, N=255 head with fp32 output, M tails, tiles that straddle image boundaries) is run
through yb_run_layer and compared with a torch fp32 convolution of the SAME fp16-rounded operands, so
the only differences are fp32 summation order and the final fp16 rounding:
    |y - ref| <= 3e-3 * max|ref| + 2e-3 * |ref|.
Network level: the deviation of the fp16 path from the fp32 oracle is measured and reported; it is
bounded loosely (SURVEY.md 7.2 predicts ~0.1-0.2 on logits of std 1.25), never asserted to 1e-4.
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
def fp16_ctx(sd):
    lib = _lib.load()
    ctx = _lib.create_ctx(0, 80, None)
    for k, v in sd.items():
        if "num_batches" in k:
            continue
        v = v.contiguous()
        _lib.check(lib.yb_set_tensor(ctx, k.encode(), vp(v), v.numel(), 1), ctx)
    _lib.check(lib.yb_finalize(ctx, _lib.YB_MODE_FP16), ctx)
    yield lib, ctx
    lib.yb_destroy(ctx)


def ref_layer(sd, spec, x_nhwc16, res_nhwc16):
    """torch fp32 conv on fp16-rounded operands + the fused epilogue, NHWC out."""
    k = spec["key"]
    x = x_nhwc16.float().permute(0, 3, 1, 2)
    if spec["bn"]:
        w = sd[k + ".conv.weight"].half().float()
        y = F.conv2d(x, w, None, spec["stride"], (spec["ks"] - 1) // 2)
        invstd = 1.0 / torch.sqrt(sd[k + ".bn.running_var"] + 1e-5)
        alpha = invstd * sd[k + ".bn.weight"]
        beta = sd[k + ".bn.bias"] - sd[k + ".bn.running_mean"] * alpha
        y = y * alpha.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)
        y = F.leaky_relu(y, 0.1)
    else:
        w = sd[k + ".weight"].half().float()
        y = F.conv2d(x, w, None) + sd[k + ".bias"].view(1, -1, 1, 1)
    y = y.permute(0, 2, 3, 1).contiguous()
    if res_nhwc16 is not None:
        y = y + res_nhwc16.float()
    return y


# (layer index in cfg order, B, H, W, with_residual)
CASES = [
    (2, 2, 24, 40, False),     # 64->32 1x1 @ stage 0
    (3, 2, 24, 40, True),      # 32->64 3x3 s1, Cin=32 (64B swizzle), residual
    (1, 2, 48, 40, False),     # 32->64 3x3 s2, Cin=32
    (4, 3, 38, 38, False),     # 64->128 3x3 s2
    (6, 3, 19, 19, True),      # 64->128 3x3 s1 + residual, M=1083 (tail, tiles straddle images)
    (5, 1, 19, 19, False),     # 128->64 1x1
    (9, 2, 40, 24, False),     # 128->256 3x3 s2
    (11, 2, 20, 12, True),     # 128->256 3x3 + residual
    (43, 2, 38, 38, False),    # 512->1024 3x3 s2 (K=4608)
    (44, 3, 19, 19, False),    # 1024->512 1x1
    (45, 3, 19, 19, True),     # 512->1024 3x3 + residual (deepest, N=1024)
    (58, 3, 19, 19, False),    # head 1024->255 1x1, fp32 out, padded N
    (59, 2, 19, 19, False),    # up1 512->256 1x1
    (60, 1, 38, 38, False),    # 768->256 1x1 (concat input width)
    (74, 1, 76, 76, False),    # head 256->255 @ /8
    # widths that are multiples of 38 take the halo-tile kernel (conv_halo.cu): 3 x 38 tiles, nine shifted views of one patch
    (3, 2, 10, 76, True),      # 32->64, Cin=32 (64-byte pixels), residual, last row strip partial (10 = 3*3 + 1)
    (3, 1, 5, 38, False),      # same without residual (two-deep staging ring), single column tile
    (6, 2, 7, 114, True),      # 64->128, Cin=64, two 64-channel halves per tile, residual
    (6, 1, 152, 152, True),    # the real stage-1 map
    (1, 2, 20, 152, False),    # 32->64 stride 2 through four parity planes (output 10 x 76: partial last strip)
    (1, 1, 14, 76, False),     # same, one column tile, odd number of output rows
    # 64->128 stride 2 with enough tiles to run as CTA pairs with resident half weight slabs (BN = 128 pair mode)
    (4, 4, 304, 304, False),   # M = 92 416 = 361 full 256-row units
    (4, 7, 304, 152, False),   # M = 80 864: odd number of 128-row tiles and a partial last one
    (6, 9, 104, 104, True),    # 64->128 stride 1 at a width the halo kernel does not take: pair mode + residual ring
    # 32-column sub-tiles: the two halves of the epilogue warps alternate over tiles (Cout = 32) / sub-tiles (fp32 heads)
    (2, 4, 152, 152, False),   # 722 tiles: four or five per CTA, both parities
    (74, 8, 76, 76, False),    # head at /8: 181 pair tiles, eight fp32 sub-tiles each
    # halo kernel in pair mode (Cout = 128: two spatial tiles per cta_group::2 UMMA)
    (6, 1, 9, 38, True),       # three tiles: rank 1 of the second pair gets the zero-filled / clipped tile past the batch
    (6, 3, 5, 76, False),      # no residual: the four-slot ring recycled by the stores alone; partial last row strip
    # halo kernel at widths that are not multiples of 38: the last column tile is partial (zero-filled patch, clipped store)
    (3, 2, 10, 104, True),     # 32->64 at the 416 network's stage-0 width / 2: 2.74 column tiles, residual
    (3, 1, 7, 208, False),     # 416 network, 5.47 column tiles
    (6, 2, 7, 64, True),       # pair mode at the 256 network's width: 1.68 column tiles, residual
    (6, 1, 5, 128, False),     # pair mode, 3.37 column tiles
    (1, 2, 20, 208, False),    # 32->64 stride 2 (four parity planes), output 10 x 104
    (1, 1, 14, 256, False),    # stride 2 at the 256 network's input width, output 7 x 128
    # the benched batch: 608x608 batch 32 shapes of the last stage (2.49 waves of pair tiles: the tail wave is partial)
    (45, 32, 19, 19, True),    # 512->1024 3x3 + residual, M = 11 552
    (27, 32, 38, 38, False),   # 512->256 1x1 at 38^2, M = 46 208
]


@pytest.mark.parametrize("li,B,H,W,with_res", CASES)
def test_tc_layer_vs_torch(fp16_ctx, sd, li, B, H, W, with_res):
    lib, ctx = fp16_ctx
    spec = topology.layer_specs(80)[li]
    rs = np.random.RandomState(100 + li)
    x = torch.from_numpy(rs.standard_normal((B, H, W, spec["cin"])).astype(np.float32)).half()
    Ho, Wo = H // spec["stride"], W // spec["stride"]
    res = torch.from_numpy(rs.standard_normal((B, Ho, Wo, spec["cout"])).astype(np.float32)).half() if with_res else None
    head = not spec["bn"]
    cout_store = (spec["cout"] + 15) // 16 * 16 if head else spec["cout"]
    out = torch.full((B, Ho, Wo, cout_store), float("nan"), device="cuda", dtype=torch.float32 if head else torch.float16)
    xd = x.cuda()
    rd = res.cuda() if with_res else None
    _lib.check(lib.yb_run_layer(ctx, li, vp(xd), B, H, W, vp(rd) if with_res else None, vp(out), stream()), ctx)
    torch.cuda.synchronize()
    y = out.float().cpu()[..., :spec["cout"]]
    ref = ref_layer(sd, spec, x, res)
    err = (y - ref).abs()
    tol = 3e-3 * ref.abs().max() + 2e-3 * ref.abs()
    bad = (err > tol) | torch.isnan(y)
    assert not bad.any(), (f"layer {li} {spec['key']}: {int(bad.sum())} of {bad.numel()} outside tolerance, "
                           f"max err {float(err.max()):.4g} (ref max {float(ref.abs().max()):.4g}); first bad index "
                           f"{tuple(int(v) for v in bad.nonzero()[0])}")


@pytest.mark.parametrize("B,H,W", [(2, 40, 56), (1, 64, 32), (3, 17, 24), (1, 64, 608), (2, 10, 76), (1, 3, 40), (3, 7, 116),
                                   (2, 9, 152), (1, 1, 8), (5, 2, 36)])
def test_tc_stem_vs_torch(fp16_ctx, sd, B, H, W):
    """Cin=3 stem on the tensor cores: NCHW fp32 image in, NHWC fp16 out (im2col rows built by producer warps)."""
    lib, ctx = fp16_ctx
    rs = np.random.RandomState(7)
    x = torch.from_numpy(rs.rand(B, 3, H, W).astype(np.float32))
    out = torch.full((B, H, W, 32), float("nan"), device="cuda", dtype=torch.float16)
    xd = x.cuda()
    _lib.check(lib.yb_run_layer(ctx, 0, vp(xd), B, H, W, None, vp(out), stream()), ctx)
    torch.cuda.synchronize()
    spec = topology.layer_specs(80)[0]
    ref = ref_layer(sd, spec, x.half().permute(0, 2, 3, 1).contiguous(), None)
    y = out.float().cpu()
    err = (y - ref).abs()
    tol = 3e-3 * ref.abs().max() + 2e-3 * ref.abs()
    bad = (err > tol) | torch.isnan(y)
    assert not bad.any(), f"stem: {int(bad.sum())} of {bad.numel()} outside tolerance, max err {float(err.max()):.4g}"


@pytest.mark.parametrize("B,H,W", [(2, 40, 64), (1, 64, 76), (3, 18, 128), (1, 64, 608), (2, 10, 152), (1, 2, 64), (2, 6, 304),
                                   (1, 96, 416), (5, 14, 208)])
def test_fused_stem_block_vs_torch(fp16_ctx, sd, B, H, W):
    """Stem + first stride-2 convolution in one kernel (stem_block.cu; reference darknet.py:66-69): NCHW fp32 image in, layer
    1's NHWC fp16 output out, against torch fp32 convolutions of the same fp16-rounded operands with the stem output rounded
    to fp16 in between (what the two separate kernels produce).  Shapes cover partial row / column tiles, images smaller
    than one tile and every image border."""
    lib, ctx = fp16_ctx
    rs = np.random.RandomState(11)
    x = torch.from_numpy(rs.rand(B, 3, H, W).astype(np.float32))
    out = torch.full((B, H // 2, W // 2, 64), float("nan"), device="cuda", dtype=torch.float16)
    xd = x.cuda()
    _lib.check(lib.yb_run_stem_block(ctx, vp(xd), B, H, W, vp(out), stream()), ctx)
    torch.cuda.synchronize()
    specs = topology.layer_specs(80)
    y0 = ref_layer(sd, specs[0], x.half().permute(0, 2, 3, 1).contiguous(), None).half()
    ref = ref_layer(sd, specs[1], y0, None)
    y = out.float().cpu()
    err = (y - ref).abs()
    tol = 6e-3 * ref.abs().max() + 4e-3 * ref.abs()
    bad = (err > tol) | torch.isnan(y)
    assert not bad.any(), f"fused stem block: {int(bad.sum())} of {bad.numel()} outside tolerance, max err {float(err.max()):.4g}"


def test_fused_stem_block_error_codes(fp16_ctx):
    """yb_run_stem_block: YB_E_UNSUPPORTED (-8) for shapes the fused kernel does not take (odd H, widths TMA cannot read, widths
    whose last column tile would be < 80 % full), YB_E_ARG (-1) for null tensors / empty shapes."""
    lib, ctx = fp16_ctx
    x = torch.zeros(1, 3, 64, 160, device="cuda")
    out = torch.zeros(1, 32, 80, 64, device="cuda", dtype=torch.float16)
    assert lib.yb_run_stem_block(ctx, vp(x), 1, 64, 160, vp(out), stream()) == -8      # 80 output columns = 2.1 tiles
    assert lib.yb_run_stem_block(ctx, vp(x), 1, 63, 64, vp(out), stream()) == -8       # odd height
    assert lib.yb_run_stem_block(ctx, vp(x), 1, 64, 70, vp(out), stream()) == -8       # row pitch not a multiple of 16 bytes
    assert lib.yb_run_stem_block(ctx, None, 1, 64, 64, vp(out), stream()) == -1
    assert lib.yb_run_stem_block(ctx, vp(x), 0, 64, 64, vp(out), stream()) == -1
    assert lib.yb_run_stem_block(ctx, vp(x), 1, 64, 64, vp(out), stream()) == 0
    torch.cuda.synchronize()


def test_fused_stem_block_equals_separate_kernels(fp16_ctx, sd):
    """The fused kernel against the two separate kernels (halo stem, then the stride-2 halo convolution) on the same image:
    same fp16 operands and the same fp16 hand-over, so the results agree to the accumulation order of the tensor core."""
    lib, ctx = fp16_ctx
    B, H, W = 2, 96, 152
    x = torch.rand(B, 3, H, W, device="cuda")
    fused = torch.empty((B, H // 2, W // 2, 64), device="cuda", dtype=torch.float16)
    _lib.check(lib.yb_run_stem_block(ctx, vp(x), B, H, W, vp(fused), stream()), ctx)
    y0 = torch.empty((B, H, W, 32), device="cuda", dtype=torch.float16)
    _lib.check(lib.yb_run_layer(ctx, 0, vp(x), B, H, W, None, vp(y0), stream()), ctx)
    sep = torch.empty_like(fused)
    _lib.check(lib.yb_run_layer(ctx, 1, vp(y0), B, H, W, None, vp(sep), stream()), ctx)
    torch.cuda.synchronize()
    d = (fused.float() - sep.float()).abs()
    assert float(d.max()) <= 4e-3 * float(sep.float().abs().max()), f"max diff {float(d.max()):.4g}"


def test_fp16_net_deviation_report(oracle, sd):
    """End to end at 416 (one image) and 608 (two images): the deviation of the fp16 tensor-core path from the fp32 CPU
    oracle is reported and bounded at about twice what is measured (rounds 1-2: logits max 0.07-0.15, xy 0.26-0.47 px,
    conf/cls 0.022-0.035), so that a real regression of the benched path fails here."""
    from yolo_v3_b200 import YoloNet
    net = YoloNet((416, 416), precision="fp16")
    net.load_state_dict(sd)
    net = net.cuda().eval()
    for B, hw, seed in ((1, 416, 1), (2, 608, 0)):
        x = synth.make_images(B, hw, hw, seed=seed)
        ls = net.head_logits(x.cuda())
        ref_ls = oracle.head_logits(sd, x)
        for i, (l, r) in enumerate(zip(ls, ref_ls)):
            d = (l.cpu() - r).abs()
            print(f"[fp16 deviation] {hw}px head{i}: logits max|d|={float(d.max()):.4f} mean|d|={float(d.mean()):.5f} "
                  f"(logit std {float(r.std()):.3f})")
            assert float(d.max()) < 0.3 and float(d.mean()) < 0.02
        det = torch.cat(net(x.cuda(), None), 1).cpu()
        ref = torch.cat(oracle.forward(sd, x), 1)
        dxy = (det[..., :2] - ref[..., :2]).abs().max()
        dconf = (det[..., 4:] - ref[..., 4:]).abs().max()
        print(f"[fp16 deviation] {hw}px boxes: max|d xy|={float(dxy):.3f}px  max|d conf/cls|={float(dconf):.4f}")
        assert float(dxy) < 0.75 and float(dconf) < 0.06


def test_fp16_final_detections_vs_oracle_at_bench_thresholds(oracle, sd):
    """the FINAL detections of the benched fp16 path against the fp32 oracle's own end-to-end result
    (oracle.forward + oracle.postprocessing) on 8 images of the bench workload at conf 0.5 / nms 0.4: boxes are matched
    one to one (same class, IOU > 0.5); the sets may differ only where a score sits at the threshold or an IOU at the
    NMS threshold.  bench.py reports the same figures in its `parity` block."""
    import importlib.util
    import os
    from yolo_v3_b200 import YoloNet
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    net = YoloNet((608, 608), precision="fp16")
    net.load_state_dict(sd)
    net = net.cuda().eval()
    x = synth.make_images(8, 608, 608, seed=100)           # the first resident batch of bench.py, first 8 images
    got = net.detect(x.cuda(), 0.5, 0.4)
    ref = oracle.postprocessing(torch.cat(oracle.forward(sd, x), 1), 80, 0.5, 0.4)
    assert len(got) == len(ref) == 8
    m = g = r = 0
    ious = []
    for a, b in zip(got, ref):
        mm, ng, nr, mi, _ = bench.match_detections(a.cpu(), b)
        m += mm; g += ng; r += nr
        if mm:
            ious.append(mi)
    set_iou = m / max(1, g + r - m)
    print(f"[fp16 vs oracle, conf 0.5] ours {g}, oracle {r}, matched {m}, set IOU {set_iou:.4f}, mean box IOU {sum(ious) / len(ious):.4f}")
    # measured (round 2): 2330 vs 2336 detections, 2255 matched -> set IOU 0.935, mean box IOU of the matches 0.942 -- random
    # weights put many scores at the 0.5 threshold and many box pairs at the NMS threshold; bounds at twice the deviation
    assert r > 500 and set_iou > 0.87 and sum(ious) / len(ious) > 0.88


def test_fp16_detect_matches_own_postprocess(sd):
    """The fused yb_detect call equals forward + postprocessing on the same device tensor."""
    from yolo_v3_b200 import YoloNet, postprocessing
    net = YoloNet((608, 608), precision="fp16")
    net.load_state_dict(sd)
    net = net.cuda().eval()
    x = synth.make_images(2, 608, 608, seed=4).cuda()
    det = torch.cat(net(x, None), 1)
    a = postprocessing(det, 80, 0.1, 0.4)
    b = net.detect(x, 0.1, 0.4)
    assert len(a) == len(b) == 2
    for p, q in zip(a, b):
        assert torch.equal(p, q)


def test_fp16_other_shapes_and_class_counts(oracle):
    """Non-square input, batch 3, and a 20-class head (75 channels -> padded 80: the narrow fp32 staging path)."""
    from yolo_v3_b200 import YoloNet, postprocessing
    for nc, (h, w), B in ((20, (160, 224), 3), (80, (320, 192), 1)):
        sd = synth.make_state_dict(seed=7, num_classes=nc, recipe="calibrated", calib_hw=96)
        x = synth.make_images(B, h, w, seed=9)
        net = YoloNet((w, h), numClass=nc, precision="fp16")
        net.load_state_dict(sd)
        net = net.cuda().eval()
        det = torch.cat(net(x.cuda(), None), 1)
        ref = torch.cat(oracle.forward(sd, x, num_classes=nc), 1)
        assert det.shape == ref.shape == (B, topology.num_boxes(h, w), 5 + nc)
        d = (det.cpu() - ref).abs()
        print(f"[fp16 deviation] {nc} classes {h}x{w}: max|d xy|={float(d[..., :2].max()):.3f}px max|d conf/cls|={float(d[..., 4:].max()):.4f}")
        assert float(d[..., :2].max()) < 1.0 and float(d[..., 4:].max()) < 0.08
        res, idx = postprocessing(det, nc, 0.05, 0.4, return_index=True)
        ref_res, ref_idx = oracle.postprocessing_c(det.cpu(), nc, 0.05, 0.4)
        for r, e, i, ei in zip(res, ref_res, idx, ref_idx):
            assert np.array_equal(i, ei) and torch.equal(r, e)


def test_fp16_net_where_the_first_two_layers_are_not_fused(oracle, sd):
    """Widths whose last 38-column tile of layer 1 would be less than 80 % full (W = 160 -> 80 output columns = 2.1 tiles)
    run the stem (stem_halo.cu) and the first stride-2 convolution as separate kernels: same deviation bounds against the
    fp32 oracle as the fused path, and the post-process stays bit-exact on the same candidates."""
    from yolo_v3_b200 import YoloNet, postprocessing
    lib = _lib.load()
    for (h, w) in ((160, 160), (128, 96)):
        x = synth.make_images(2, h, w, seed=21)
        net = YoloNet((w, h), precision="fp16")
        net.load_state_dict(sd)
        net = net.cuda().eval()
        n0 = None
        det = torch.cat(net(x.cuda(), None), 1)
        n0 = lib.yb_launch_count(net._ctx)
        det = torch.cat(net(x.cuda(), None), 1)
        launches = lib.yb_launch_count(net._ctx) - n0
        assert launches == 78, launches                        # 75 convolution kernels (nothing fused) + 2 upsample copies + decode
        ref = torch.cat(oracle.forward(sd, x), 1)
        d = (det.cpu() - ref).abs()
        print(f"[fp16 deviation] unfused first layers {h}x{w}: max|d xy|={float(d[..., :2].max()):.3f}px max|d conf/cls|={float(d[..., 4:].max()):.4f}")
        assert float(d[..., :2].max()) < 0.75 and float(d[..., 4:].max()) < 0.06
        res, idx = postprocessing(det, 80, 0.05, 0.4, return_index=True)
        ref_res, ref_idx = oracle.postprocessing_c(det.cpu(), 80, 0.05, 0.4)
        for r, e, i, ei in zip(res, ref_res, idx, ref_idx):
            assert np.array_equal(i, ei) and torch.equal(r, e)


def test_fp16_plan_cache_two_shapes(sd):
    """Alternating shapes reuses cached plans and keeps results identical."""
    from yolo_v3_b200 import YoloNet
    net = YoloNet((416, 416), precision="fp16")
    net.load_state_dict(sd)
    net = net.cuda().eval()
    xa = synth.make_images(1, 416, 416, seed=1).cuda()
    xb = synth.make_images(2, 224, 224, seed=2).cuda()
    a1 = torch.cat(net(xa, None), 1).clone()
    b1 = torch.cat(net(xb, None), 1).clone()
    a2 = torch.cat(net(xa, None), 1)
    b2 = torch.cat(net(xb, None), 1)
    assert torch.equal(a1, a2) and torch.equal(b1, b2)


# ---- fp16 input images (yb_set_input_dtype) ---------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W", [(2, 40, 56), (3, 17, 24), (1, 64, 608)])
def test_stem_fp16_input_same_bits_as_fp32(fp16_ctx, B, H, W):
    """The stem rounds every fp32 pixel to fp16 (round to nearest even) before the tensor core sees it, so reading the
    host-rounded fp16 image must give bit-identical output."""
    lib, ctx = fp16_ctx
    rs = np.random.RandomState(11)
    x = torch.from_numpy(rs.rand(B, 3, H, W).astype(np.float32)).cuda()
    xh = x.half()
    out32 = torch.full((B, H, W, 32), float("nan"), device="cuda", dtype=torch.float16)
    out16 = torch.full((B, H, W, 32), float("nan"), device="cuda", dtype=torch.float16)
    _lib.check(lib.yb_run_layer(ctx, 0, vp(x), B, H, W, None, vp(out32), stream()), ctx)
    try:
        _lib.check(lib.yb_set_input_dtype(ctx, _lib.YB_INPUT_F16), ctx)
        _lib.check(lib.yb_run_layer(ctx, 0, vp(xh), B, H, W, None, vp(out16), stream()), ctx)
    finally:
        _lib.check(lib.yb_set_input_dtype(ctx, _lib.YB_INPUT_F32), ctx)
    torch.cuda.synchronize()
    assert not torch.isnan(out32.float()).any()
    assert torch.equal(out32.view(torch.int16), out16.view(torch.int16))


def test_detect_right_after_an_async_h2d_copy_sees_the_new_batch(sd):
    """The first kernel is launched with the programmatic-stream-serialisation attribute (its prologue overlaps the previous
    KERNEL's tail).  A host->device copy enqueued on the same stream right before the call is not a kernel: the launch must
    stay fully ordered behind it.  Alternating batches are copied into ONE device buffer and detected immediately; every
    result must equal the one computed from a synchronised copy of that batch."""
    from yolo_v3_b200 import YoloNet
    net = YoloNet((608, 608), precision="fp16")
    net.load_state_dict(sd)
    net = net.cuda().eval()
    hx = [synth.make_images(8, 608, 608, seed=31 + i).pin_memory() for i in range(2)]
    ref = []
    for h in hx:
        d = h.cuda()
        torch.cuda.synchronize()
        ref.append([r.clone() for r in net.detect(d, 0.1, 0.4)])
        torch.cuda.synchronize()
    assert not all(torch.equal(a, b) for a, b in zip(ref[0], ref[1]) if a.shape == b.shape) or any(a.shape != b.shape for a, b in zip(ref[0], ref[1]))
    dev = torch.empty_like(hx[0], device="cuda")
    for i in range(8):
        dev.copy_(hx[i & 1], non_blocking=True)              # same stream, no synchronisation before the detect call
        out = net.detect(dev, 0.1, 0.4)
        assert len(out) == len(ref[i & 1])
        for a, b in zip(out, ref[i & 1]):
            assert torch.equal(a, b), f"iteration {i}: the detect call did not see the batch copied right before it"


def test_detect_fp16_input_same_detections(sd):
    from yolo_v3_b200 import YoloNet
    net = YoloNet((224, 160), precision="fp16")
    net.load_state_dict(sd)
    net = net.cuda().eval()
    x = synth.make_images(2, 160, 224, seed=5).cuda()
    a = net.detect(x, 0.05, 0.4)
    b = net.detect(x.half(), 0.05, 0.4)          # read by the stem as fp16
    c = net.detect(x, 0.05, 0.4)                 # and back to fp32 input on the same context
    assert len(a) == len(b) == len(c)
    for ra, rb, rc in zip(a, b, c):
        assert torch.equal(ra, rb) and torch.equal(ra, rc)


@pytest.mark.parametrize("precision", ["fp16", "fp32"])
def test_graph_replay_equals_stream_launches(sd, precision):
    """yb_set_graph_mode: the CUDA-graph replay of the launches after the stem gives bit-identical results (same kernels, same
    arguments) on the capture call, on replays, for a second shape, and for the backbone-only entry of the same plan."""
    from yolo_v3_b200 import YoloNet
    net = YoloNet((416, 416), precision=precision)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    xa = synth.make_images(1, 416, 416, seed=1).cuda()
    xb = synth.make_images(2, 96, 160, seed=2).cuda()
    net.set_graph_mode("never")
    ra, rb, bb = torch.cat(net(xa, None), 1).clone(), torch.cat(net(xb, None), 1).clone(), net.backbone(xa).clone()
    da = net.detect(xa, 0.3, 0.4)
    net.set_graph_mode("always")
    for it in range(4):                                     # eager, capture, replay, replay
        assert torch.equal(torch.cat(net(xa, None), 1), ra), it
        assert torch.equal(torch.cat(net(xb, None), 1), rb), it
        assert torch.equal(net.backbone(xa), bb), it
        db = net.detect(xa, 0.3, 0.4)
        assert len(da) == len(db) and all(torch.equal(p, q) for p, q in zip(da, db))
    replays = _lib.load().yb_graph_replays(net._ctx)
    assert replays >= 6, f"the graph was not replayed ({replays}): capture must have failed"     # 3 entries x 2 replay iterations
    net.set_graph_mode("auto")
    assert torch.equal(torch.cat(net(xa, None), 1), ra)
