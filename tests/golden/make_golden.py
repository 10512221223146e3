"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on seeded inputs.

Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

Harness shims, both test-side only:
  * torch.Tensor.cuda -> identity, because YoloLayer.forward hard-codes .cuda() (yololayer.py:98-100)
    and this container has no GPU;
  * torch.Tensor.sort -> stable=True, because torch.sort(descending=True) is unstable on ties and the
    parity contract fixes the tie-break to score-descending, candidate-order-ascending.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

torch.Tensor.cuda = lambda self, *a, **k: self
_orig_sort = torch.Tensor.sort


def _stable_sort(self, *a, **k):
    k.setdefault("stable", True)
    return _orig_sort(self, *a, **k)


torch.Tensor.sort = _stable_sort

import darknet            # noqa: E402  (reference)
import utils as ref_utils  # noqa: E402  (reference)
import yololayer          # noqa: E402  (reference)

from yolo_v3_b200 import synth  # noqa: E402

ANCHORS = [(10, 13), (16, 30), (33, 23), (30, 61), (62, 45), (59, 119), (116, 90), (156, 198), (373, 326)]
MASKS = ([6, 7, 8], [3, 4, 5], [0, 1, 2])


def pack_results(res, prefix, out):
    """list-of-tensors -> flat arrays (npz cannot hold ragged lists)."""
    out[prefix + "_is_empty_list"] = np.array(isinstance(res, list) and len(res) == 0)
    counts, rows = [], []
    for r in res:
        r = r.numpy() if r.numel() else np.zeros((0, 7), np.float32)
        counts.append(len(r))
        rows.append(r.reshape(-1, 7))
    out[prefix + "_counts"] = np.array(counts, np.int64)
    out[prefix + "_rows"] = np.concatenate(rows, 0).astype(np.float32) if rows else np.zeros((0, 7), np.float32)


def make_decode():
    out = {}
    for tag, (hw, nc, b) in {"a": ((64, 64), 80, 2), "b": ((96, 160), 20, 1)}.items():
        h, w = hw
        rs = np.random.RandomState(11)
        out[f"{tag}_img_hw"] = np.array([h, w])
        out[f"{tag}_num_classes"] = np.array(nc)
        for i, s in enumerate((32, 16, 8)):
            x = torch.from_numpy((rs.standard_normal((b, 3 * (5 + nc), h // s, w // s)) * 1.5).astype(np.float32))
            layer = yololayer.YoloLayer(ANCHORS, MASKS[i], (w, h), nc)
            with torch.no_grad():
                y = layer(x, (w, h), None)
            out[f"{tag}_logits{i}"] = x.numpy()
            out[f"{tag}_det{i}"] = y.numpy()
    np.savez_compressed(os.path.join(HERE, "decode_golden.npz"), **out)


def synth_det(rs, b, n, nc, frac_hot=0.25):
    """Decoded-looking rows: clustered boxes so NMS suppresses, a few exact score ties, a zero-area box."""
    d = np.zeros((b, n, 5 + nc), np.float32)
    centers = rs.uniform(40, 360, (b, 12, 2))
    which = rs.randint(0, 12, (b, n))
    for i in range(b):
        d[i, :, 0:2] = centers[i, which[i]] + rs.standard_normal((n, 2)) * 9
    d[..., 2:4] = rs.uniform(20, 120, (b, n, 2))
    d[..., 4] = np.where(rs.rand(b, n) < frac_hot, rs.uniform(0.3, 1.0, (b, n)), rs.uniform(0, 0.05, (b, n)))
    d[..., 5:] = rs.uniform(0, 0.2, (b, n, nc))
    hot = rs.randint(0, min(nc, 6), (b, n))
    for i in range(b):
        d[i, np.arange(n), 5 + hot[i]] = rs.uniform(0.5, 1.0, n)
    # exact ties: copy score-defining entries between a few rows of the same class
    d[0, 10] = d[0, 3]
    d[0, 10, 0] += 200.0
    d[0, 17] = d[0, 3]
    d[0, 17, 1] += 150.0
    # zero-area box with a top score (NaN self-IOU: never kept, never suppresses)
    d[0, 5, 2] = 0.0
    d[0, 5, 4] = 0.99
    d[0, 5, 5:] = 0.0
    d[0, 5, 5] = 0.99
    return d


def make_postprocess():
    out = {}
    rs = np.random.RandomState(5)
    cases = {
        "c80": (synth_det(rs, 2, 300, 80), 80),
        "c20": (synth_det(rs, 3, 200, 20), 20),
    }
    # an image with no candidates in a batch that has some
    cases["c20"][0][1, :, 4] = 0.0
    for tag, (det, nc) in cases.items():
        out[f"{tag}_det"] = det
        out[f"{tag}_num_classes"] = np.array(nc)
        t = torch.from_numpy(det)
        modes = {
            "nms": dict(obj_conf_thr=0.5, nms_thr=0.4),
            "nms_low": dict(obj_conf_thr=0.05, nms_thr=0.45),
            "eval": dict(obj_conf_thr=0.05, nms_thr=0.45, is_eval=True),
            "raw": dict(obj_conf_thr=0.3, nms_thr=0.4, use_nms=False),
            "raw_eval": dict(obj_conf_thr=0.1, nms_thr=0.4, is_eval=True, use_nms=False),
            "none": dict(obj_conf_thr=2.0, nms_thr=0.4),
            "none_eval": dict(obj_conf_thr=2.0, nms_thr=0.4, is_eval=True),
        }
        for m, kw in modes.items():
            res = ref_utils.postprocessing(t.clone(), nc, **kw)
            pack_results(res, f"{tag}_{m}", out)
            out[f"{tag}_{m}_kw"] = np.array([kw.get("obj_conf_thr"), kw.get("nms_thr"),
                                             float(kw.get("is_eval", False)), float(kw.get("use_nms", True))])
    np.savez_compressed(os.path.join(HERE, "postprocess_golden.npz"), **out)


def make_net():
    out = {}
    sd = synth.make_state_dict(seed=1234, recipe="analytic")
    net = darknet.YoloNet((64, 96)).eval()
    net.load_state_dict(sd)
    x = synth.make_images(2, 64, 96, seed=3)
    feats = {}
    net.pre_det1.mlist[6].register_forward_hook(lambda m, i, o: feats.__setitem__("l0", o.detach().numpy()))
    net.pre_det2.mlist[6].register_forward_hook(lambda m, i, o: feats.__setitem__("l1", o.detach().numpy()))
    net.pre_det3.mlist[6].register_forward_hook(lambda m, i, o: feats.__setitem__("l2", o.detach().numpy()))
    net.feature.register_forward_hook(lambda m, i, o: feats.__setitem__("backbone", o.detach().numpy()))
    with torch.no_grad():
        dets = net(x, None)
    out["img_hw"] = np.array([64, 96])
    ref_sd = net.state_dict()
    out["state_dict_keys"] = np.array(list(ref_sd.keys()))
    out["state_dict_shapes"] = np.array([",".join(str(int(d)) for d in v.shape) for v in ref_sd.values()])
    out["seed"] = np.array(1234)
    out["img_seed"] = np.array(3)
    for i, d in enumerate(dets):
        out[f"det{i}"] = d.numpy()
        out[f"logits{i}"] = feats[f"l{i}"]
    out["backbone"] = feats["backbone"]
    res = ref_utils.postprocessing(torch.cat(dets, 1).clone(), 80, obj_conf_thr=0.05, nms_thr=0.4)
    pack_results(res, "post", out)
    # weight-stream order (WeightManager, darknet.py:249-303): write a synthetic darknet file from the
    # state_dict in our claimed order, load it with the reference loader, and record what it consumed.
    from oracle import yolo_oracle as O
    blob = O.darknet_blob_from_state_dict(sd)
    path = "/tmp/_golden_synth.weights"
    with open(path, "wb") as fp:
        np.array([0, 2, 0, 32013312, 0], np.int32).tofile(fp)
        blob.tofile(fp)
    net2 = darknet.YoloNet((64, 96)).eval()
    wm = darknet.WeightManager(net2)
    consumed = wm.loadWeight(path)
    os.remove(path)
    sd2 = net2.state_dict()
    same = all(torch.equal(sd[k], sd2[k]) for k in sd if not k.endswith("num_batches_tracked"))
    out["darknet_consumed"] = np.array(consumed)
    out["darknet_total"] = np.array(len(blob))
    out["darknet_roundtrip_equal"] = np.array(same)
    out["darknet_seen"] = np.array(int(wm.seen))
    # order-sensitive digest of the float stream, cheap to recompute anywhere
    w = np.arange(1, 1 + 4096, dtype=np.float64)
    out["darknet_blob_probe"] = np.array([float(blob[:4096].astype(np.float64) @ w),
                                          float(blob[-4096:].astype(np.float64) @ w),
                                          float(blob[20_000_000:20_004_096].astype(np.float64) @ w)])
    np.savez_compressed(os.path.join(HERE, "net_golden.npz"), **out)
    print("darknet consumed", consumed, "of", len(blob), "roundtrip equal", same)


def make_boxes():
    """correct_yolo_boxes (boundingbox.py:139-149) on network-space boxes, letterbox and plain resize."""
    import boundingbox as ref_bb   # noqa: E402  (reference)
    out = {}
    rs = np.random.RandomState(21)
    cases = [(602, 452, 416, 416), (1280, 720, 608, 608), (333, 500, 416, 416), (640, 640, 608, 608)]
    out["cases"] = np.array(cases)
    for i, (ow, oh, iw, ih) in enumerate(cases):
        k = 64
        x1 = rs.uniform(-20, iw, k); y1 = rs.uniform(-20, ih, k)
        b = np.stack([x1, y1, x1 + rs.uniform(1, 300, k), y1 + rs.uniform(1, 300, k)], 1).astype(np.float32)
        b[5] = 0.0                                     # all-zero row: passed through by the mask
        t = torch.from_numpy(b)
        out[f"in{i}"] = b
        out[f"letterbox{i}"] = ref_bb.correct_yolo_boxes(t.clone(), ow, oh, iw, ih, True).numpy()
        out[f"resize{i}"] = ref_bb.correct_yolo_boxes(t.clone(), ow, oh, iw, ih, False).numpy()
    np.savez_compressed(os.path.join(HERE, "boxes_golden.npz"), **out)


LETTERBOX_CASES = [  # (src_h, src_w, dim, seed): the dog-cycle-car.png shape, VGA, tall, upscale, 720p, square upscale
    (452, 602, 416, 31), (480, 640, 224, 32), (500, 333, 192, 33), (97, 131, 160, 34), (720, 1280, 160, 35), (64, 64, 96, 36)]


RESIZE_CASES = [(480, 640, 224, 224, 41), (97, 131, 160, 96, 42), (720, 1280, 192, 128, 43), (60, 40, 33, 77, 44)]  # sh, sw, dim_w, dim_h, seed


def make_letterbox():
    """utils.letterbox_image / load_image (utils.py:44-72) on seeded uint8 images (synth.make_photo, regenerated by
    the tests from the seed).  Stored per case: the canvas with cv2's own implementation (cv2.ipp.setUseIPP(False)),
    the difference of the IPP-accelerated build to it (int8, |d| <= 1), and `trans`."""
    import cv2
    out = {"cases": np.array(LETTERBOX_CASES)}
    for i, (sh, sw, dim, seed) in enumerate(LETTERBOX_CASES):
        img = synth.make_photo(sh, sw, seed)
        cv2.ipp.setUseIPP(False)
        canvas, trans = ref_utils.letterbox_image(img, (dim, dim))
        cv2.ipp.setUseIPP(True)
        canvas_ipp, _ = ref_utils.letterbox_image(img, (dim, dim))
        assert canvas.min() >= 0 and canvas.max() <= 255
        out[f"canvas{i}"] = canvas.astype(np.uint8)
        out[f"ipp_delta{i}"] = (canvas_ipp - canvas).astype(np.int8)
        out[f"trans{i}"] = trans.numpy()
        # what load_image returns after the decode: float CHW / 255 (utils.py:71) -- digest only
        x = torch.from_numpy(canvas).float().permute(2, 0, 1) / 255
        out[f"sum{i}"] = np.array(float(x.double().sum()))
    # load_image end to end (utils.py:60-72) through a lossless PNG: 'resize' (cv2.resize(img, dim), INTER_LINEAR; IPP
    # does not divert 8-bit linear, so both builds agree), 'letterbox' and mode=None
    import tempfile
    out["resize_cases"] = np.array(RESIZE_CASES)
    with tempfile.TemporaryDirectory() as td:
        for i, (sh, sw, dw, dh, seed) in enumerate(RESIZE_CASES):
            img = synth.make_photo(sh, sw, seed)
            path = os.path.join(td, f"img{i}.png")
            cv2.imwrite(path, cv2.cvtColor(img, cv2.COLOR_RGB2BGR))
            res = {}
            for ipp in (False, True):
                cv2.ipp.setUseIPP(ipp)
                res[ipp], trans = ref_utils.load_image(path, "resize", (dw, dh))
                assert trans is None
            assert torch.equal(res[False], res[True])
            x = res[False]
            assert tuple(x.shape) == (3, dh, dw)
            out[f"resize{i}"] = torch.round(x * 255).permute(1, 2, 0).numpy().astype(np.uint8)
            out[f"resize_sum{i}"] = np.array(float(x.double().sum()))
            plain, _ = ref_utils.load_image(path)                   # mode=None: /255 + CHW only
            out[f"plain_sum{i}"] = np.array(float(plain.double().sum()))
    np.savez_compressed(os.path.join(HERE, "letterbox_golden.npz"), **out)


def make_notebook():
    """The notebook's own post-process: cells 30 (torch_unique), 33 (iou_vectorized, reduce_row_by_column, nms) and
    35 (postprocessing) of /root/reference/yolo_detect.ipynb are executed as they are (Tensor.cuda is the identity
    shim above, Tensor.sort is stable) on the synthetic detections of the post-process fixtures."""
    import json
    nb = json.load(open("/root/reference/yolo_detect.ipynb"))
    ns = {"torch": torch, "np": np, "Tensor": torch.Tensor}
    for i in (30, 33, 35):
        src = "".join(nb["cells"][i]["source"])
        assert any(k in src for k in ("def torch_unique", "def iou_vectorized", "def postprocessing")), i
        exec(compile(src, f"yolo_detect.ipynb#cell{i}", "exec"), ns)
    out = {}
    rs = np.random.RandomState(9)
    cases = {"c80": (synth_det(rs, 2, 300, 80), 80), "c20": (synth_det(rs, 3, 200, 20), 20)}
    cases["c20"][0][1, :, 4] = 0.0                     # an image without detections
    for tag, (det, nc) in cases.items():
        out[f"{tag}_det"] = det
        out[f"{tag}_num_classes"] = np.array(nc)
        for m, (ct, nt) in {"default": (0.5, 0.4), "low": (0.02, 0.45), "none": (2.0, 0.4)}.items():
            res = ns["postprocessing"](torch.from_numpy(det).clone(), nc, obj_conf_thr=ct, nms_thr=nt)
            assert isinstance(res, list) and len(res) == det.shape[0]
            pack_results(res, f"{tag}_{m}", out)
            out[f"{tag}_{m}_kw"] = np.array([ct, nt])
    np.savez_compressed(os.path.join(HERE, "notebook_golden.npz"), **out)


if __name__ == "__main__":
    if "--notebook-only" in sys.argv:
        make_notebook()
        sys.exit(0)
    make_notebook()
    make_letterbox()
    if "--letterbox-only" in sys.argv:
        sys.exit(0)
    make_boxes()
    make_decode()
    make_postprocess()
    make_net()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
