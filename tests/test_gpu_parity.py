"""GPU parity tests: the CUDA path (through the C ABI / its ctypes binding) against the CPU oracle
and the reference-generated golden fixtures.  Run on the B200 box with `pytest -m gpu`.

Tolerances (stated where used):
  * post-process: bit-exact rows and identical survivor indices (integer/index work);
  * decode given identical fp32 logits: allclose(atol=1e-4, rtol=1e-6)  (1 fp32 ulp at
    608 px is 6.1e-5; exp(tw)*anchor reaches thousands of px, hence the rtol term);
  * fp32-grade convolution stack (precision='fp32' = YB_MODE_FP32_TC: the tensor-core split mode, csrc/conv_tc.cu): head
    logits within 1e-4 * max|logit| (measured 1.1e-5 at 608x608 batch 4 -- the fp32 CPU oracle is itself 5e-6 away from
    a float64 evaluation).  The decode turns a logit error d into: conf/cls 0.25*d, xy 0.25*stride*d (<= 8*d), w/h a
    RELATIVE error d.  End to end: conf/cls allclose(atol=1e-4, rtol=1e-6) -- the north-star bar (measured 3.8e-5) --
    xy atol 1e-3 px (measured 4.6e-4), w/h rtol 5e-4 (measured 1.4e-4): the noise floor between two fp32-grade
    75-layer stacks, not a loosened kernel tolerance (decode alone holds 1e-4, above);
  * fp16 tensor-core stack: every layer against a torch conv on the same fp16-rounded operands
    (atol 3e-3*max|y| + rtol 2e-3 = fp16 output rounding); end-to-end deviation vs the fp32 oracle is
    REPORTED and bounded at twice the measured figures (tests/test_gpu_fp16.py), as BASELINE/SURVEY state (fp16 cannot
    meet 1e-4); the final detections are compared with the oracle's at the bench thresholds.
"""
import ctypes

import numpy as np
import pytest
import torch

from yolo_v3_b200 import _lib, synth, topology

pytestmark = pytest.mark.gpu

ANCHORS = [(10, 13), (16, 30), (33, 23), (30, 61), (62, 45), (59, 119), (116, 90), (156, 198), (373, 326)]
MASKS = ([6, 7, 8], [3, 4, 5], [0, 1, 2])


def vp(t):
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def unpack(g, prefix):
    if bool(g[prefix + "_is_empty_list"]):
        return []
    counts, rows = g[prefix + "_counts"], g[prefix + "_rows"]
    out, o = [], 0
    for c in counts:
        out.append(rows[o:o + c])
        o += c
    return out


@pytest.fixture(scope="module")
def sd_analytic():
    return synth.make_state_dict(seed=1234, recipe="analytic")


@pytest.fixture(scope="module")
def sd_calibrated():
    return synth.make_state_dict(seed=1234, recipe="calibrated")


def make_net(sd, precision, hw=(416, 416)):
    from yolo_v3_b200 import YoloNet
    net = YoloNet(hw, precision=precision)
    net.load_state_dict(sd)
    return net.cuda().eval()


# ---------------------------------------------------------------------------------------------
# L1: decode
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["a", "b"])
def test_decode_vs_reference_golden(golden, tag):
    g = golden("decode_golden.npz")
    h, w = (int(v) for v in g[f"{tag}_img_hw"])
    nc = int(g[f"{tag}_num_classes"])
    lib = _lib.load()
    ctx = _lib.create_ctx(0, nc, None)
    maps = [torch.from_numpy(g[f"{tag}_logits{i}"]).cuda() for i in range(3)]
    B = maps[0].shape[0]
    det = torch.empty(B, topology.num_boxes(h, w), 5 + nc, device="cuda")
    _lib.check(lib.yb_decode(ctx, vp(maps[0]), vp(maps[1]), vp(maps[2]), B, h, w, vp(det), stream()), ctx)
    ref = np.concatenate([g[f"{tag}_det{i}"] for i in range(3)], 1)
    np.testing.assert_allclose(det.cpu().numpy(), ref, atol=1e-4, rtol=1e-6)
    lib.yb_destroy(ctx)


def test_decode_608_vs_oracle(oracle):
    maps = synth.make_head_logits(2, 608, 608, seed=3, obj_mu=-2.0)
    lib = _lib.load()
    ctx = _lib.create_ctx(0, 80, None)
    dm = [m.cuda() for m in maps]
    det = torch.empty(2, 22743, 85, device="cuda")
    _lib.check(lib.yb_decode(ctx, vp(dm[0]), vp(dm[1]), vp(dm[2]), 2, 608, 608, vp(det), stream()), ctx)
    ref = torch.cat([oracle.decode(m, ANCHORS, MASKS[i], (608, 608)) for i, m in enumerate(maps)], 1)
    np.testing.assert_allclose(det.cpu().numpy(), ref.numpy(), atol=1e-4, rtol=1e-6)
    lib.yb_destroy(ctx)


def test_yololayer_module_matches_oracle(oracle):
    from yolo_v3_b200 import YoloLayer
    x = synth.make_head_logits(1, 96, 160, seed=5)[1]
    layer = YoloLayer(ANCHORS, [3, 4, 5], (160, 96), 80)
    y = layer(x.cuda(), (160, 96), None)
    ref = oracle.decode(x, ANCHORS, [3, 4, 5], (160, 96))
    np.testing.assert_allclose(y.cpu().numpy(), ref.numpy(), atol=1e-4, rtol=1e-6)
    with pytest.raises(NotImplementedError):
        layer(x.cuda(), (160, 96), torch.zeros(1, 1, 5))


# ---------------------------------------------------------------------------------------------
# L0: post-process (bit-exact)
# ---------------------------------------------------------------------------------------------
MODES = ["nms", "nms_low", "eval", "raw", "raw_eval", "none", "none_eval"]


@pytest.mark.parametrize("tag", ["c80", "c20"])
@pytest.mark.parametrize("mode", MODES)
def test_postprocess_vs_reference_golden(golden, tag, mode):
    from yolo_v3_b200 import postprocessing
    g = golden("postprocess_golden.npz")
    det = torch.from_numpy(g[f"{tag}_det"]).cuda()
    keep = det.clone()
    nc = int(g[f"{tag}_num_classes"])
    conf, nms, is_eval, use_nms = g[f"{tag}_{mode}_kw"]
    res = postprocessing(det, nc, float(conf), float(nms), bool(is_eval), bool(use_nms))
    assert torch.equal(det, keep), "input must not be modified"
    ref = unpack(g, f"{tag}_{mode}")
    assert isinstance(res, list) and len(res) == len(ref)
    for r, e in zip(res, ref):
        assert not r.is_cuda
        if e.shape[0] == 0 and r.numel() == 0:
            continue
        assert tuple(r.shape) == e.shape
        assert np.array_equal(r.numpy(), e), "rows must be bit-equal to the reference"


def _stress_det(oracle, B, dense, mu, seed=7):
    maps = synth.make_head_logits(B, 608, 608, seed=seed, obj_mu=mu, dense=dense)
    return torch.cat([oracle.decode(m, ANCHORS, MASKS[i], (608, 608)) for i, m in enumerate(maps)], 1)


@pytest.mark.parametrize("dense,mu", [(False, -6.5), (True, -5.0)])
def test_postprocess_stress_608_bit_exact(oracle, dense, mu):
    """BASELINE config 4 shape (608x608, conf 0.001): ~10k candidates per image, wide and dense."""
    from yolo_v3_b200 import postprocessing
    det = _stress_det(oracle, 3, dense, mu)
    ref, ref_idx = oracle.postprocessing_c(det, 80, 0.001, 0.4)
    res, idx = postprocessing(det.cuda(), 80, 0.001, 0.4, return_index=True)
    assert len(res) == len(ref) == 3
    ncand = 0
    for r, e, i, ei in zip(res, ref, idx, ref_idx):
        assert tuple(r.shape) == tuple(e.shape)
        assert np.array_equal(i, ei), "NMS survivor indices must be identical"
        assert torch.equal(r, e)
        ncand += len(e)
    assert ncand > 3000


def test_postprocess_eval_mode_608(oracle):
    from yolo_v3_b200 import postprocessing
    det = _stress_det(oracle, 2, False, -3.0, seed=11)
    ref, ref_idx = oracle.postprocessing_c(det, 80, 0.005, 0.45, True, True)
    res, idx = postprocessing(det.cuda(), 80, 0.005, 0.45, is_eval=True, return_index=True)
    for r, e, i, ei in zip(res, ref, idx, ref_idx):
        assert np.array_equal(i, ei)
        assert torch.equal(r, e)


def test_postprocess_single_class_worst_case(oracle):
    """Every box in one class: one (image, class) segment of N boxes (global-memory NMS path)."""
    from yolo_v3_b200 import postprocessing
    rs = np.random.RandomState(1)
    n = 9000
    d = np.zeros((1, n, 85), np.float32)
    d[0, :, 0:2] = rs.uniform(0, 600, (n, 2))
    d[0, :, 2:4] = rs.uniform(10, 80, (n, 2))
    d[0, :, 4] = rs.uniform(0.5, 1, n)
    d[0, :, 5 + 3] = rs.uniform(0.5, 1, n)
    det = torch.from_numpy(d)
    ref, ref_idx = oracle.postprocessing_c(det, 80, 0.2, 0.4)
    res, idx = postprocessing(det.cuda(), 80, 0.2, 0.4, return_index=True)
    assert np.array_equal(idx[0], ref_idx[0])
    assert torch.equal(res[0], ref[0])


def test_postprocess_idempotent_on_survivors(oracle):
    """Size-independent property at the full BASELINE size: survivors of NMS survive NMS again
    (re-encoded as a detection tensor, same thresholds), for every image of a batch of 8."""
    from yolo_v3_b200 import postprocessing
    det = _stress_det(oracle, 8, True, -5.0, seed=21)
    res = postprocessing(det.cuda(), 80, 0.001, 0.4)
    kmax = max(len(r) for r in res)
    again = torch.zeros(len(res), kmax, 85)
    for b, r in enumerate(res):
        k = len(r)
        again[b, :k, 0] = (r[:, 0] + r[:, 2]) / 2
        again[b, :k, 1] = (r[:, 1] + r[:, 3]) / 2
        again[b, :k, 2] = r[:, 2] - r[:, 0]
        again[b, :k, 3] = r[:, 3] - r[:, 1]
        again[b, :k, 4] = 1.0
        again[b, torch.arange(k), 5 + r[:, 6].long()] = r[:, 5]
    res2 = postprocessing(again.cuda(), 80, 0.001, 0.4 + 1e-3)
    for r, r2 in zip(res, res2):
        assert len(r2) == len(r)
        assert torch.equal(r2[:, 5], r[:, 5]) and torch.equal(r2[:, 6], r[:, 6])


# ---------------------------------------------------------------------------------------------
# L3: loaders through the C ABI
# ---------------------------------------------------------------------------------------------
def test_darknet_blob_roundtrip_c_abi(oracle, sd_analytic):
    lib = _lib.load()
    ctx = _lib.create_ctx(0, 80, None)
    assert lib.yb_num_tensors(ctx) == 438
    keys = []
    for i in range(438):
        n = ctypes.c_size_t()
        k = lib.yb_tensor_key(ctx, i, ctypes.byref(n)).decode()
        keys.append(k)
        assert n.value == sd_analytic[k].numel()
    assert keys == list(sd_analytic.keys())
    blob = oracle.darknet_blob_from_state_dict(sd_analytic)
    consumed = ctypes.c_size_t()
    _lib.check(lib.yb_load_darknet_blob(ctx, ctypes.c_void_p(blob.ctypes.data), len(blob), 0, ctypes.byref(consumed)), ctx)
    assert consumed.value == 62001757
    for k in ("feature.mlist.0.conv.weight", "feature.mlist.14.conv2.bn.running_var", "pre_det2.mlist.6.bias",
              "up2.conv.bn.weight", "pre_det3.mlist.6.weight"):
        out = np.empty(sd_analytic[k].numel(), np.float32)
        _lib.check(lib.yb_get_tensor(ctx, k.encode(), ctypes.c_void_p(out.ctypes.data), out.size), ctx)
        assert np.array_equal(out, sd_analytic[k].numpy().ravel()), k
    back = np.empty(len(blob), np.float32)
    written = ctypes.c_size_t()
    _lib.check(lib.yb_save_darknet_blob(ctx, ctypes.c_void_p(back.ctypes.data), len(back), 0, ctypes.byref(written)), ctx)
    assert written.value == len(blob) and np.array_equal(back, blob)
    _lib.check(lib.yb_load_darknet_blob(ctx, ctypes.c_void_p(blob.ctypes.data), 40620640, 1, ctypes.byref(consumed)), ctx)
    assert consumed.value == 40620640
    rc = lib.yb_load_darknet_blob(ctx, ctypes.c_void_p(blob.ctypes.data), 1000, 0, ctypes.byref(consumed))
    assert rc != 0 and b"ends inside" in lib.yb_last_error(ctx)
    rc = lib.yb_set_tensor(ctx, b"nope.weight", ctypes.c_void_p(blob.ctypes.data), 1, 1)
    assert rc == -2
    lib.yb_destroy(ctx)


def test_darknet_file_python_loader(tmp_path, oracle, sd_analytic):
    from yolo_v3_b200 import YoloNet
    blob = oracle.darknet_blob_from_state_dict(sd_analytic)
    path = tmp_path / "synth.weights"
    with open(path, "wb") as fp:
        np.array([0, 2, 0, 32013312, 0], np.int32).tofile(fp)
        blob.tofile(fp)
    net = YoloNet((64, 64))
    net.loadWeight(str(path), "darknet")
    assert int(net.seen) == 32013312
    for k, v in net.state_dict().items():
        if "num_batches" not in k:
            assert torch.equal(v, sd_analytic[k]), k
    out = tmp_path / "back.weights"
    net.saveWeight(str(out), "darknet")
    assert open(out, "rb").read() == open(path, "rb").read()


# ---------------------------------------------------------------------------------------------
# L2a: fp32 convolution stack, end to end
# ---------------------------------------------------------------------------------------------
def test_fp32_net_vs_reference_golden(golden, sd_analytic):
    g = golden("net_golden.npz")
    h, w = (int(v) for v in g["img_hw"])
    x = synth.make_images(2, h, w, seed=int(g["img_seed"]))
    net = make_net(sd_analytic, "fp32", (h, w))
    ls = net.head_logits(x.cuda())
    for i, l in enumerate(ls):
        ref = g[f"logits{i}"]
        np.testing.assert_allclose(l.cpu().numpy(), ref, rtol=0, atol=1e-4 * np.abs(ref).max())
    dets = net(x.cuda(), None)
    for i, d in enumerate(dets):
        np.testing.assert_allclose(d.cpu().numpy(), g[f"det{i}"], rtol=5e-4, atol=1e-4)
    bb = net.backbone(x.cuda())
    np.testing.assert_allclose(bb.cpu().numpy(), g["backbone"], rtol=0, atol=1e-4 * np.abs(g["backbone"]).max())


def test_fp32_net_416_plumbing_config(oracle, sd_calibrated):
    """BASELINE config 1 shape: one 416x416 image through forward + decode + NMS."""
    from yolo_v3_b200 import postprocessing
    x = synth.make_images(1, 416, 416, seed=1)
    net = make_net(sd_calibrated, "fp32")
    d1, d2, d3 = net(x.cuda(), None)
    assert d1.shape == (1, 507, 85) and d2.shape == (1, 2028, 85) and d3.shape == (1, 8112, 85)   # yolo_detect.ipynb:643
    det = torch.cat((d1, d2, d3), 1)
    ref = torch.cat(oracle.forward(sd_calibrated, x), 1)
    np.testing.assert_allclose(det.cpu().numpy(), ref.numpy(), rtol=5e-4, atol=1e-4)
    np.testing.assert_allclose(det[..., :2].cpu().numpy(), ref[..., :2].numpy(), rtol=0, atol=1e-3)
    np.testing.assert_allclose(det[..., 4:].cpu().numpy(), ref[..., 4:].numpy(), rtol=0, atol=1e-4)
    # identical candidates -> identical survivors: run both post-processes on the SAME tensor
    res, idx = postprocessing(det, 80, 0.1, 0.4, return_index=True)
    ref_res, ref_idx = oracle.postprocessing_c(det.cpu(), 80, 0.1, 0.4)
    assert len(res) == 1 and np.array_equal(idx[0], ref_idx[0]) and torch.equal(res[0], ref_res[0])
    assert len(res[0]) > 10
    fused = net.detect(x.cuda(), 0.1, 0.4)
    assert torch.equal(fused[0], res[0])


def test_fp32_net_608_batch(oracle, sd_calibrated):
    x = synth.make_images(2, 608, 608, seed=0)
    net = make_net(sd_calibrated, "fp32", (608, 608))
    det = torch.cat(net(x.cuda(), None), 1)
    assert det.shape == (2, 22743, 85)
    ref = torch.cat(oracle.forward(sd_calibrated, x), 1)
    np.testing.assert_allclose(det.cpu().numpy(), ref.numpy(), rtol=5e-4, atol=1e-4)
    np.testing.assert_allclose(det[..., :2].cpu().numpy(), ref[..., :2].numpy(), rtol=0, atol=1e-3)
    np.testing.assert_allclose(det[..., 4:].cpu().numpy(), ref[..., 4:].numpy(), rtol=0, atol=1e-4)


def test_standalone_darknet_backbone(oracle, sd_calibrated, tmp_path):
    """Darknet(blkList).forward stand-alone (darknet.py:72-88), the way the reference's notebooks use the backbone for
    classification pre-training: its own parameters, fp32-grade mode, against the oracle's backbone; and loadWeight of a
    backbone-only darknet stream (darknet.py:102-104)."""
    from yolo_v3_b200.darknet import Darknet
    d = Darknet([1, 2, 8, 8, 4])
    d.load_state_dict({k[len("feature."):]: v for k, v in sd_calibrated.items() if k.startswith("feature.")})
    d = d.cuda().eval()
    x = synth.make_images(2, 96, 128, seed=8)
    owner = d._engine_owner()
    owner.precision = "fp32"
    y = d(x.cuda())
    with torch.no_grad():
        ref = oracle.backbone(sd_calibrated, x)[0]
    assert y.shape == ref.shape == (2, 1024, 3, 4)
    assert float((y.cpu() - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    blob = oracle.darknet_blob_from_state_dict(sd_calibrated, 80, backbone_only=True)
    path = tmp_path / "darknet53.conv.74"
    with open(path, "wb") as fp:
        np.array([0, 2, 0, 0, 0], np.int32).tofile(fp)
        (blob * 0.5).astype(np.float32).tofile(fp)
    d.loadWeight(str(path))
    assert torch.equal(d.mlist[0].conv.weight.detach().cpu(), sd_calibrated["feature.mlist.0.conv.weight"] * 0.5)
    y2 = d(x.cuda())
    assert not torch.equal(y, y2)
    with pytest.raises(NotImplementedError):
        Darknet([1, 1, 1, 1, 1]).cuda().eval()(x.cuda())


# ---------------------------------------------------------------------------------------------
# L4: drop-in conventions
# ---------------------------------------------------------------------------------------------
def test_dropin_conventions(sd_calibrated):
    from yolo_v3_b200 import YoloNet, postprocessing
    net = make_net(sd_calibrated, "fp32")
    x = synth.make_images(2, 64, 64, seed=2).cuda()
    with pytest.raises(NotImplementedError):
        net(x, torch.zeros(2, 1, 5))
    with pytest.raises(RuntimeError):
        net(x.cpu(), None)
    with pytest.raises(ValueError):
        net(torch.zeros(1, 3, 60, 64, device="cuda"), None)
    net.train()
    with pytest.raises(RuntimeError):
        net(x, None)
    net.eval()
    dets = net(x, None)
    det = torch.cat(dets, 1)
    assert det.is_cuda and det.dtype == torch.float32
    assert postprocessing(det, 80, 2.0, 0.4) == []                      # nothing passes anywhere -> []
    d = det.clone()
    d[1, :, 4] = 0                                                        # image 1 has no candidate
    res = postprocessing(d, 80, 0.01, 0.4)
    assert len(res) == 2 and res[1].numel() == 0 and res[1].shape == (0,) and res[0].shape[1] == 7
    # weights changed -> engine re-finalises
    before = det.clone()
    with torch.no_grad():
        net.pre_det1.mlist[6].bias.add_(1.0)
    after = torch.cat(net(x, None), 1)
    assert not torch.equal(before[:, :12], after[:, :12]) and torch.equal(before[:, 12:], after[:, 12:])


def test_c_abi_error_codes():
    lib = _lib.load()
    ctx = _lib.create_ctx(0, 80, None)
    x = torch.zeros(1, 3, 64, 64, device="cuda")
    det = torch.zeros(1, 252, 85, device="cuda")
    assert lib.yb_forward(ctx, vp(x), 1, 64, 64, vp(det), stream()) == -3          # not finalized
    _lib.check(lib.yb_finalize(ctx, _lib.YB_MODE_FP32), ctx)
    assert lib.yb_forward(ctx, vp(x), 1, 60, 64, vp(det), stream()) == -1          # H not multiple of 32
    assert lib.yb_forward(ctx, vp(x), 0, 64, 64, vp(det), stream()) == -1
    assert lib.yb_finalize(ctx, 7) == -1
    assert lib.yb_forward(ctx, vp(x), 1, 64, 64, vp(det), stream()) == 0
    assert lib.yb_launch_count(ctx) == 76
    # round-2 entry points: the graph-mode switch validates its argument; the fp32-grade tensor-core mode finalises, runs
    # (77 launches: 75 layers + the two upsample copies, + the decode), and rejects fp16 input images like the other fp32 mode
    assert lib.yb_set_graph_mode(ctx, 3) == -1 and lib.yb_set_graph_mode(ctx, 0) == 0
    assert lib.yb_graph_replays(ctx) == 0
    _lib.check(lib.yb_finalize(ctx, _lib.YB_MODE_FP32_TC), ctx)
    n0 = lib.yb_launch_count(ctx)
    assert lib.yb_forward(ctx, vp(x), 1, 64, 64, vp(det), stream()) == 0
    assert lib.yb_launch_count(ctx) - n0 == 78
    _lib.check(lib.yb_set_input_dtype(ctx, _lib.YB_INPUT_F16), ctx)
    assert lib.yb_forward(ctx, vp(x), 1, 64, 64, vp(det), stream()) == -8          # YB_E_UNSUPPORTED
    _lib.check(lib.yb_set_input_dtype(ctx, _lib.YB_INPUT_F32), ctx)
    torch.cuda.synchronize()
    lib.yb_destroy(ctx)


# ---------------------------------------------------------------------------------------------
# N2: correct_yolo_boxes (the step right after the path)
# ---------------------------------------------------------------------------------------------
def test_correct_yolo_boxes_bit_exact(golden, oracle):
    from yolo_v3_b200.boundingbox import correct_yolo_boxes, correct_yolo_boxes_batch
    g = golden("boxes_golden.npz")
    for i, (ow, oh, iw, ih) in enumerate(g["cases"]):
        b = torch.from_numpy(g[f"in{i}"])
        for mode, flag in (("letterbox", True), ("resize", False)):
            y = correct_yolo_boxes(b, int(ow), int(oh), int(iw), int(ih), flag)            # CPU in -> CPU out
            assert not y.is_cuda and np.array_equal(y.numpy(), g[f"{mode}{i}"]), (i, mode)
    # batch form on device rows7 with counts: two images of different original size
    rows = torch.zeros(2, 64, 7)
    rows[0, :, :4] = torch.from_numpy(g["in0"]); rows[1, :, :4] = torch.from_numpy(g["in1"])
    counts = torch.tensor([64, 40], dtype=torch.int32)
    out = correct_yolo_boxes_batch(rows.cuda(), counts.cuda(), [(602, 452), (602, 452)], 416, 416, True).cpu()
    assert np.array_equal(out[0].numpy(), g["letterbox0"])
    ref1 = oracle.correct_yolo_boxes(torch.from_numpy(g["in1"]), 602, 452, 416, 416, True)
    assert torch.equal(out[1, :40], ref1[:40]) and float(out[1, 40:].abs().sum()) == 0.0


# ---------------------------------------------------------------------------------------------
# N1: letterbox pre-process (the step right before the path)
# ---------------------------------------------------------------------------------------------
def test_letterbox_bit_exact(golden, oracle):
    """yb_letterbox against the canvases the reference's own letterbox_image produced (OpenCV's portable code path:
    bit-exact; the IPP-accelerated build: within one grey level), one launch for the whole ragged batch."""
    from yolo_v3_b200.utils import letterbox_batch, letterbox_image
    g = golden("letterbox_golden.npz")
    by_dim = {}
    for i, (sh, sw, dim, seed) in enumerate(g["cases"]):
        by_dim.setdefault(int(dim), []).append((i, synth.make_photo(int(sh), int(sw), int(seed))))
    for dim, items in by_dim.items():
        canv, trans = letterbox_batch([im for _, im in items], (dim, dim), want_canvas=True)
        x, trans2 = letterbox_batch([im for _, im in items], (dim, dim))
        canv, x = canv.cpu().numpy(), x.cpu()
        assert torch.equal(trans, trans2)
        for k, (i, im) in enumerate(items):
            assert np.array_equal(canv[k], g[f"canvas{i}"]), i                                  # bit-exact (bytes)
            assert np.array_equal(trans[k].numpy(), g[f"trans{i}"]), i
            ref = torch.from_numpy(g[f"canvas{i}"].astype(np.int64)).float().permute(2, 0, 1) / 255   # utils.py:71
            assert torch.equal(x[k], ref), i
            ipp = g[f"canvas{i}"].astype(np.int64) + g[f"ipp_delta{i}"]
            assert np.abs(canv[k].astype(np.int64) - ipp).max() <= 1
    # reference signature: numpy integer canvas + 5-element transform
    img = synth.make_photo(97, 131, 34)
    canvas, t = letterbox_image(img, (160, 160))
    ref_canvas, ref_t = oracle.letterbox_image(img, (160, 160))
    assert canvas.dtype == np.int64 and np.array_equal(canvas, ref_canvas) and torch.equal(t, ref_t)


def test_letterbox_edge_cases(oracle):
    from yolo_v3_b200.utils import letterbox_batch, letterbox_image
    rs = np.random.RandomState(3)
    # ragged, tiny, strongly up- and down-scaled noise images against the oracle (which is pinned to cv2 on the CPU side)
    shapes = [(1, 1), (2, 7), (5, 3), (33, 257), (301, 17), (64, 64), (240, 427)]
    imgs = [rs.randint(0, 256, (h, w, 3)).astype(np.uint8) for h, w in shapes]
    for dim in (32, 96, 224):
        canv, trans = letterbox_batch(imgs, (dim, dim), want_canvas=True)
        for k, im in enumerate(imgs):
            h, w = im.shape[:2]
            if int(w * min(dim / w, dim / h)) == 0 or int(h * min(dim / w, dim / h)) == 0:
                continue
            ref, rt = oracle.letterbox_image(im, (dim, dim))
            assert np.array_equal(canv[k].cpu().numpy(), ref), (dim, k)
            assert torch.equal(trans[k], rt)
    # an image that collapses to an empty box, and a non-square dim whose box does not fit the reference's
    # transposed canvas: the reference raises (cv2 / numpy), so does the mirror
    with pytest.raises(Exception):
        letterbox_batch([np.zeros((1, 400, 3), np.uint8)], (32, 32))
    with pytest.raises(ValueError):
        letterbox_image(np.zeros((100, 300, 3), np.uint8), (320, 96))
    # identity geometry = load_image(mode=None): exact copy, /255, HWC -> CHW
    im = imgs[-1]
    x, _ = letterbox_batch([im], (im.shape[1], im.shape[0]), canvas_hw=im.shape[:2])
    assert torch.equal(x[0].cpu(), torch.from_numpy(im).float().permute(2, 0, 1) / 255)


def test_detect_fused_equals_two_kernel_path(oracle, sd_calibrated):
    """yb_detect (decode + score fused, det never materialised) == yb_forward + yb_postprocess, bit for bit, in fp32
    mode at a low threshold (most anchors take the full path) and a high one (most are skipped on objectness)."""
    from yolo_v3_b200 import YoloNet, postprocessing
    net = YoloNet((96, 160), precision="fp32")
    net.load_state_dict(sd_calibrated)
    net = net.cuda().eval()
    x = synth.make_images(3, 160, 96, seed=11).cuda()
    det = torch.cat(net(x, None), 1)
    ref_det = torch.cat(oracle.forward(sd_calibrated, x.cpu()), 1)
    np.testing.assert_allclose(det.cpu().numpy(), ref_det.numpy(), rtol=1e-3, atol=2e-3)
    for thr in (0.001, 0.05, 0.5, 0.999):
        a = net.detect(x, thr, 0.4)
        b = postprocessing(det, 80, thr, 0.4)
        assert len(a) == len(b)
        assert all(torch.equal(p, q) for p, q in zip(a, b)), thr
        ref, _ = oracle.postprocessing_c(det.cpu(), 80, thr, 0.4)
        assert len(ref) == len(b) and all(torch.equal(p, q) for p, q in zip(ref, b)), thr


# ---------------------------------------------------------------------------------------------
# N4: the notebook's own inline post-process (yolo_detect.ipynb cell 35)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["c80", "c20"])
@pytest.mark.parametrize("mode", ["default", "low", "none"])
def test_notebook_postprocess_vs_reference_golden(golden, tag, mode):
    from yolo_v3_b200.notebook import postprocessing as nb_post
    g = golden("notebook_golden.npz")
    det = torch.from_numpy(g[f"{tag}_det"])
    ct, nt = g[f"{tag}_{mode}_kw"]
    res = nb_post(det, int(g[f"{tag}_num_classes"]), float(ct), float(nt))
    ref = unpack(g, f"{tag}_{mode}")
    assert isinstance(res, list) and len(res) == len(ref)
    for a, b in zip(res, ref):
        assert not a.is_cuda
        assert np.array_equal(a.numpy().reshape(-1, 7), b)          # bit-exact rows in the notebook's order


def test_notebook_postprocess_608_vs_oracle(oracle):
    """A 608x608 head (22 743 boxes, 2 images) through the notebook variant against its oracle: identical rows
    and source indices."""
    from yolo_v3_b200.notebook import postprocessing as nb_post
    logits = synth.make_head_logits(2, 608, 608, 80, seed=7, obj_mu=-3.0)
    det = torch.cat([oracle.decode(l, ANCHORS, MASKS[i], (608, 608), 80) for i, l in enumerate(logits)], 1)
    res, idx = nb_post(det, 80, 0.3, 0.4, return_index=True)
    ref, ridx = oracle.notebook_postprocessing(det, 80, 0.3, 0.4, return_index=True)
    assert sum(len(r) for r in ref) > 200
    for a, b, i, j in zip(res, ref, idx, ridx):
        assert torch.equal(a, b) and np.array_equal(i, j)


def test_resize_mode_bit_exact(golden, oracle, tmp_path):
    """yb_resize (load_image mode 'resize') against the reference's own output, and load_image end to end through a
    PNG file for the three modes."""
    from yolo_v3_b200.utils import load_image, resize_batch
    g = golden("letterbox_golden.npz")
    for i, (sh, sw, dw, dh, seed) in enumerate(g["resize_cases"]):
        img = synth.make_photo(int(sh), int(sw), int(seed))
        u8 = resize_batch([img], (int(dw), int(dh)), want_u8=True)[0].cpu().numpy()
        assert np.array_equal(u8, g[f"resize{i}"]), i
        x = resize_batch([img], (int(dw), int(dh)))[0].cpu()
        assert torch.equal(x, torch.from_numpy(g[f"resize{i}"]).float().permute(2, 0, 1) / 255)
    cv2 = pytest.importorskip("cv2")
    img = synth.make_photo(97, 131, 42)
    path = str(tmp_path / "img.png")
    cv2.imwrite(path, cv2.cvtColor(img, cv2.COLOR_RGB2BGR))
    x, t = load_image(path, "resize", (160, 96))
    assert t is None and x.is_cuda and torch.equal(x.cpu(), oracle.load_image_resize(img, (160, 96)))
    x, t = load_image(path, "letterbox", (160, 160))
    rx, rt = oracle.load_image_letterbox(img, (160, 160))
    assert torch.equal(x.cpu(), rx) and torch.equal(t, rt)
    x, t = load_image(path)
    assert t is None and torch.equal(x.cpu(), torch.from_numpy(img).float().permute(2, 0, 1) / 255)


def test_iaa_letterbox_variant(oracle):
    """The dataset transform's letterbox (transforms.IaaLetterbox): same bicubic resize, (w - bw)//2 offsets, [h,w] canvas."""
    from yolo_v3_b200.utils import letterbox_batch
    imgs = [synth.make_photo(500, 333, 33), synth.make_photo(97, 131, 34), synth.make_photo(240, 427, 5)]
    for dim in ((416, 416), (320, 224)):
        canv, _ = letterbox_batch(imgs, dim, want_canvas=True, iaa=True)
        assert tuple(canv.shape) == (3, dim[1], dim[0], 3)
        for k, im in enumerate(imgs):
            assert np.array_equal(canv[k].cpu().numpy(), oracle.iaa_letterbox(im, dim)), (dim, k)


def test_eval_json_writer(tmp_path, oracle, sd_calibrated):
    """evaluate.py's results writer on the device path: eval-mode detect (conf 0.005 / nms 0.45) + correct_yolo_boxes,
    JSON text identical to the reference's format built from the oracle's rows."""
    import json
    from collections import OrderedDict
    from yolo_v3_b200 import YoloNet
    from yolo_v3_b200.evaluate import open_json_pred_writer, predict_and_process
    net = YoloNet((96, 96), precision="fp32")
    net.load_state_dict(sd_calibrated)
    net = net.cuda().eval()
    x = synth.make_images(2, 96, 96, seed=4)
    sample = {"img": x, "org_img": torch.zeros(2, 3, 120, 160), "img_path": ["a/COCO_val2014_000000000042.jpg", "b/000000000139.jpg"]}
    out = tmp_path / "res.json"
    with open_json_pred_writer(str(out), ["c"] * 80, is_letterbox=True) as w:
        predict_and_process([sample], net, 80, w)
    got = json.loads(out.read_text())
    det = torch.cat(net(x.cuda(), None), 1).cpu()
    ref, _ = oracle.postprocessing_c(det, 80, 0.005, 0.45, is_eval=True)
    exp = []
    for b, (iid, rows) in enumerate(zip((42, 139), ref)):
        boxes = oracle.correct_yolo_boxes(rows[:, :4], 160, 120, 96, 96, True)
        for r, bb in zip(rows.tolist(), boxes.tolist()):
            exp.append(OrderedDict(image_id=iid, category_id=int(r[6]), bbox=bb, score=r[5]))
    assert len(got) == len(exp) > 50
    assert got == json.loads(json.dumps(exp))
    text = out.read_text()
    assert text.startswith('[{\n    "image_id":42,') and text.endswith("}]")
