"""CPU oracle for the YOLOv3 inference hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (torch fp32 on the host + numpy) of the algorithm the
reference implements in

    /root/reference/darknet.py      (conv_bn_relu :27-44, res_layer :46-53, Darknet :72-100,
                                     PreDetectionConvGroup :107-150, UpsampleGroup :153-162,
                                     YoloNet.forward :198-231, WeightManager :249-303)
    /root/reference/yololayer.py    (YoloLayer.forward inference branch :31-59, :97-105)
    /root/reference/boundingbox.py  (bbox_cxcywh_to_x1y1x2y2 :25-29; correct_yolo_boxes :139-149 with
                                     letterbox_reverse :95-116, rescale_bbox :119-137)
    /root/reference/utils.py        (iou_vectorized :98-119, get_nms_detections :148-202,
                                     get_raw_detections :204-224, postprocessing :226-258;
                                     letterbox_transforms :34-42, letterbox_image :44-57, load_image :60-72,
                                     whose cv2.resize arithmetic is OpenCV's -- see the pre-process section)

It is the *checker* for the CUDA path.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import it; nothing under
yolo_v3_b200/ does (the product fails loudly without its CUDA library).

Parity pin: the reference has no runnable tests or golden vectors for this path
(SURVEY.md section 4/8c), so the oracle is pinned against outputs of the reference itself,
generated in the build container by tests/golden/make_golden.py (which imports
/root/reference unmodified) and committed as tests/golden/*.npz.  tests/test_oracle_golden.py
checks every function here against those fixtures.

Tie-break convention: torch.sort(descending=True) on the CPU is unstable, so
"bit-exact NMS survivors" is defined under score-descending, candidate-order-ascending
(= stable sort), which is what the fixtures were generated with.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

DEFAULT_ANCHORS = [10, 13, 16, 30, 33, 23, 30, 61, 62, 45, 59, 119, 116, 90, 156, 198, 373, 326]
ANCHOR_MASKS = ([6, 7, 8], [3, 4, 5], [0, 1, 2])     # yolo1, yolo2, yolo3 (darknet.py:184,189,194)
BN_EPS = 1e-5                                        # nn.BatchNorm2d default (darknet.py:39)
LEAKY = 0.1                                          # darknet.py:41
BLOCKS = [1, 2, 8, 8, 4]                             # darknet.py:179


# --------------------------------------------------------------------------------------
# Network topology, restated as a flat table of the 75 convolutions in darknet-cfg order
# (= WeightManager.find_conv_layers order, darknet.py:292-303).
# --------------------------------------------------------------------------------------
def conv_table(num_classes: int = 80) -> List[dict]:
    """Each entry: dict(key, cin, cout, ks, stride, bn).  `key` is the state_dict prefix of
    the conv_bn_relu block (then `.conv.weight`, `.bn.*`) or of the plain nn.Conv2d
    (`.weight`, `.bias`) -- names as registered by darknet.py:76-79,112-118,156."""
    t: List[dict] = []

    def cbr(key, cin, cout, ks, s=1):
        t.append(dict(key=key, cin=cin, cout=cout, ks=ks, stride=s, bn=True))

    # Darknet-53 backbone (darknet.py:72-79, make_res_stack :68-70)
    cbr("feature.mlist.0", 3, 32, 3)
    idx, ch = 1, 32
    for nb in BLOCKS:
        cbr(f"feature.mlist.{idx}", ch, ch * 2, 3, 2)
        idx += 1
        ch *= 2
        for _ in range(nb):
            cbr(f"feature.mlist.{idx}.conv1", ch, ch // 2, 1)
            cbr(f"feature.mlist.{idx}.conv2", ch // 2, ch, 3)
            idx += 1

    def predet(name, nin, nout):                      # darknet.py:107-118
        for i in range(3):
            cbr(f"{name}.mlist.{2 * i}", nin, nout, 1)
            cbr(f"{name}.mlist.{2 * i + 1}", nout, nout * 2, 3)
            nin = nout * 2
        t.append(dict(key=f"{name}.mlist.6", cin=nin, cout=(num_classes + 5) * 3, ks=1, stride=1, bn=False))

    predet("pre_det1", 1024, 512)
    cbr("up1.conv", 512, 256, 1)
    predet("pre_det2", 768, 256)
    cbr("up2.conv", 256, 128, 1)
    predet("pre_det3", 384, 128)
    return t


def _cbr(sd: Dict[str, torch.Tensor], key: str, x: torch.Tensor, ks: int, stride: int = 1) -> torch.Tensor:
    """conv_bn_relu.forward (darknet.py:43-44): LeakyReLU_0.1(BN_eval(Conv2d(x))), pad=(ks-1)//2."""
    y = F.conv2d(x, sd[key + ".conv.weight"], None, stride, (ks - 1) // 2)
    y = F.batch_norm(y, sd[key + ".bn.running_mean"], sd[key + ".bn.running_var"],
                     sd[key + ".bn.weight"], sd[key + ".bn.bias"], False, 0.0, BN_EPS)
    return F.leaky_relu(y, LEAKY)


def backbone(sd, x) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Darknet.forward (darknet.py:83-88).  Returns (out, route36, route61): the outputs of
    mlist[28], mlist[14] (256 ch, /8) and mlist[23] (512 ch, /16) (darknet.py:180-181)."""
    x = _cbr(sd, "feature.mlist.0", x, 3)
    idx = 1
    cached = {}
    for nb in BLOCKS:
        x = _cbr(sd, f"feature.mlist.{idx}", x, 3, 2)
        idx += 1
        for _ in range(nb):
            k = f"feature.mlist.{idx}"
            x = x + _cbr(sd, k + ".conv2", _cbr(sd, k + ".conv1", x, 1), 3)   # res_layer, darknet.py:52-53
            if idx in (14, 23):
                cached[idx] = x
            idx += 1
    return x, cached[14], cached[23]


def _predet(sd, name, x):
    """PreDetectionConvGroup.forward (darknet.py:121-126); returns (head logits, mlist[4] output)."""
    route = None
    for i in range(6):
        x = _cbr(sd, f"{name}.mlist.{i}", x, 1 if i % 2 == 0 else 3)
        if i == 4:
            route = x                                  # addCachedOut(-3) -> index 7-3 (darknet.py:185,146-148)
    logits = F.conv2d(x, sd[f"{name}.mlist.6.weight"], sd[f"{name}.mlist.6.bias"])
    return logits, route


def _up(sd, name, head, tail):
    """UpsampleGroup.forward (darknet.py:159-162): cat(upsample2x(conv1x1(head)), tail)."""
    y = _cbr(sd, f"{name}.conv", head, 1)
    y = F.interpolate(y, scale_factor=2, mode="nearest")
    return torch.cat((y, tail), 1)


def head_logits(sd, x) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """The convolutional part of YoloNet.forward (darknet.py:198-223): three [B,255,h,w] maps."""
    with torch.no_grad():
        out, r36, r61 = backbone(sd, x)
        l1, h1 = _predet(sd, "pre_det1", out)
        l2, h2 = _predet(sd, "pre_det2", _up(sd, "up1", h1, r61))
        l3, _ = _predet(sd, "pre_det3", _up(sd, "up2", h2, r36))
    return l1, l2, l3


def decode(logits: torch.Tensor, anchors_all: Sequence[Tuple[float, float]], mask: Sequence[int],
           img_dim: Tuple[int, int], num_classes: int = 80) -> torch.Tensor:
    """YoloLayer.forward with target=None (yololayer.py:31-59, 97-105).

    channel = a*(5+C) + attr (view at :42); row = (h*W + w)*A + a (permute at :104);
    op order  b_xy = (sigmoid(t_xy) + cell) * stride,  b_wh = (exp(t_wh) * (anchor/stride)) * stride.
    """
    nB, _, nH, nW = logits.shape
    nA = len(mask)
    attrs = 5 + num_classes
    stride = img_dim[1] / nH                                            # python float (:36)
    anc = (torch.tensor(anchors_all, dtype=torch.float32) / stride)[list(mask)]   # (:37-38)
    p = logits.view(nB, nA, attrs, nH, nW).permute(0, 1, 3, 4, 2).contiguous()
    xy = p[..., :2].sigmoid()
    wh = p[..., 2:4]
    conf = p[..., 4].sigmoid()
    cls = p[..., 5:].sigmoid()
    gx = torch.arange(nW, dtype=torch.float32).view(1, 1, 1, nW).expand(1, 1, nH, nW)
    gy = torch.arange(nH, dtype=torch.float32).view(1, 1, nH, 1).expand(1, 1, nH, nW)
    boxes = torch.empty(nB, nA, nH, nW, 4, dtype=torch.float32)
    boxes[..., 0] = xy[..., 0] + gx
    boxes[..., 1] = xy[..., 1] + gy
    boxes[..., 2:4] = wh.exp() * anc.view(1, nA, 1, 1, 2)
    out = torch.cat((boxes * stride, conf.unsqueeze(4), cls), 4)
    return out.permute(0, 2, 3, 1, 4).contiguous().view(nB, nA * nH * nW, attrs)


def forward(sd, x, anchors=DEFAULT_ANCHORS, num_classes: int = 80):
    """YoloNet.forward(x, None) (darknet.py:198-231) -> (det1, det2, det3)."""
    anchors_all = [(anchors[i], anchors[i + 1]) for i in range(0, len(anchors), 2)]   # darknet.py:176
    img_dim = (x.shape[3], x.shape[2])                                                # darknet.py:199
    ls = head_logits(sd, x)
    return tuple(decode(l, anchors_all, m, img_dim, num_classes) for l, m in zip(ls, ANCHOR_MASKS))


# --------------------------------------------------------------------------------------
# Post-process
# --------------------------------------------------------------------------------------
def iou_matrix(b: torch.Tensor) -> torch.Tensor:
    """iou_vectorized (utils.py:98-119): every op rounded separately in fp32, no +1 convention."""
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    ltx = torch.max(x1.unsqueeze(1), x1.unsqueeze(0))
    lty = torch.max(y1.unsqueeze(1), y1.unsqueeze(0))
    rbx = torch.min(x2.unsqueeze(1), x2.unsqueeze(0))
    rby = torch.min(y2.unsqueeze(1), y2.unsqueeze(0))
    inter = torch.clamp(rbx - ltx, min=0) * torch.clamp(rby - lty, min=0)
    area = (x2 - x1) * (y2 - y1)
    union = area.unsqueeze(0) + area.unsqueeze(1) - inter
    return inter / union


def greedy_keep(boxes_sorted: torch.Tensor, nms_thr: float) -> np.ndarray:
    """The row loop of get_nms_detections (utils.py:175-193) on score-sorted boxes: returns a bool
    keep vector.  A row whose own diagonal entry is not `> thr` (zero-area box -> 0/0 = NaN) is
    never kept and never suppresses."""
    over = (iou_matrix(boxes_sorted) > nms_thr).numpy()
    n = over.shape[0]
    alive = over.diagonal().copy()
    for i in range(n):
        if not alive[i]:
            continue
        sup = over[i, i + 1:] & alive[i + 1:]
        alive[i + 1:][sup] = False
    return alive


def candidates(det: torch.Tensor, num_classes: int, obj_conf_thr: float, is_eval: bool):
    """First half of postprocessing (utils.py:227-251).  Returns (det_xyxy_scored, index[K,3]) or
    (det, None) when nothing passes.  Works on a copy (the reference mutates CPU input in place)."""
    det = det.detach().cpu().clone()
    cx, cy, w, h = det[..., 0].clone(), det[..., 1].clone(), det[..., 2].clone(), det[..., 3].clone()
    det[..., 0], det[..., 2] = cx - w / 2, cx + w / 2               # boundingbox.py:25-29
    det[..., 1], det[..., 3] = cy - h / 2, cy + h / 2
    det[..., 5:5 + num_classes] = det[..., 5:5 + num_classes] * det[..., 4].unsqueeze(-1)
    if is_eval:
        index = (det[..., 5:5 + num_classes] > obj_conf_thr).nonzero()
    else:
        score, cls = torch.max(det[..., 5:5 + num_classes], -1)
        m = score > obj_conf_thr
        if not m.any():
            return det, None
        index = torch.cat((m.nonzero(), cls[m].unsqueeze(-1)), -1)
    if len(index) == 0:
        return det, None
    return det, index


def postprocessing(det: torch.Tensor, num_classes: int, obj_conf_thr: float = 0.5, nms_thr: float = 0.4,
                   is_eval: bool = False, use_nms: bool = True, return_index: bool = False):
    """utils.postprocessing (utils.py:226-258) with the stable tie-break.  Returns a list (len B) of
    [K,7] tensors [x1,y1,x2,y2,obj,score,cls] (shape-[0] tensor for an image with no candidates) or
    [] when nothing in the batch passes.  With return_index also returns per-image int64 arrays of
    the flat box index of every output row."""
    det, index = candidates(det, num_classes, obj_conf_thr, is_eval)
    if index is None:
        return ([], []) if return_index else []
    nB = det.shape[0]
    results, src = [], []
    for b in range(nB):
        sel = index[index[:, 0] == b]
        if len(sel) == 0:
            results.append(torch.Tensor())
            src.append(np.zeros(0, np.int64))
            continue
        rows, idxs = [], []
        if not use_nms:                                             # get_raw_detections (utils.py:204-224)
            box = det[b, sel[:, 1], :5]
            prob = det[b, sel[:, 1], sel[:, 2] + 5]
            results.append(torch.cat((box, prob.unsqueeze(-1), sel[:, 2].float().unsqueeze(-1)), -1))
            src.append(sel[:, 1].numpy().astype(np.int64))
            continue
        for c in sel[:, 2].unique():                                # utils.py:161 (ascending)
            ci = sel[sel[:, 2] == c]
            d = det[b, ci[:, 1]]
            _, order = d[:, 5 + c].sort(descending=True, stable=True)   # utils.py:171 + tie-break
            d = d[order]
            keep = torch.from_numpy(greedy_keep(d[:, :4], nms_thr))
            d = d[keep]
            rows.append(torch.cat((d[:, :5], d[:, 5 + c].view(-1, 1),
                                   torch.full((len(d), 1), float(c))), -1))
            idxs.append(ci[:, 1][order][keep].numpy().astype(np.int64))
        results.append(torch.cat(rows, 0))
        src.append(np.concatenate(idxs))
    return (results, src) if return_index else results


def _reduce_row_by_column(pairs: np.ndarray) -> np.ndarray:
    """reduce_row_by_column (yolo_detect.ipynb cell 33), literally: walk the (row, column) pair list of `iou > thr`;
    a pair (i, j) with j != i deletes every pair whose row is j; the cursor advances by one either way."""
    i = 0
    while i < pairs.shape[0]:
        j = pairs[i, 1]
        if pairs[i, 0] != j:
            pairs = pairs[pairs[:, 0] != j]
        i += 1
    return pairs


def notebook_postprocessing(det: torch.Tensor, num_classes: int, obj_conf_thr: float = 0.5, nms_thr: float = 0.4,
                            return_index: bool = False):
    """The notebook's own inline post-process (yolo_detect.ipynb cell 35 + helpers in cells 30, 33), which is NOT
    utils.postprocessing: objectness threshold (rows failing it are zeroed and dropped by nonzero()), x2 = x1 + w,
    class = arg-max of the raw class probabilities, per class (ascending) sort by objectness (descending; stable =
    the fixed tie-break) and the pair-list NMS of cell 33.  Returns a list (len B) of [K,7] tensors
    [x1,y1,x2,y2,obj,class_prob,cls] (an empty tensor for an image without detections)."""
    det = det.detach().cpu().float()
    results, src = [], []
    for b in range(det.shape[0]):
        d = det[b]
        ci = ((d[:, 4] > obj_conf_thr) & (d[:, 4] != 0)).nonzero().squeeze(1)
        if len(ci) == 0:
            results.append(torch.Tensor())
            src.append(np.zeros(0, np.int64))
            continue
        x1 = d[ci, 0] - d[ci, 2] / 2
        y1 = d[ci, 1] - d[ci, 3] / 2
        x2 = x1 + d[ci, 2]
        y2 = y1 + d[ci, 3]
        prob, cls = torch.max(d[ci, 5:5 + num_classes], 1)
        rows = torch.stack((x1, y1, x2, y2, d[ci, 4], prob, cls.float()), 1)
        out, idxs = [], []
        for c in torch.unique(rows[:, 6]):
            sel = (rows[:, 6] == c).nonzero().squeeze(1)
            r = rows[sel]
            _, order = r[:, 4].sort(descending=True, stable=True)
            r = r[order]
            pairs = (iou_matrix(r[:, :4]) > nms_thr).nonzero().numpy()
            pairs = _reduce_row_by_column(pairs)
            keep = np.unique(pairs[:, 0])
            out.append(r[keep])
            idxs.append(ci[sel][order][keep].numpy().astype(np.int64))
        results.append(torch.cat(out, 0))
        src.append(np.concatenate(idxs))
    return (results, src) if return_index else results


def correct_yolo_boxes(bboxes: torch.Tensor, org_w, org_h, img_w, img_h, is_letterbox=False) -> torch.Tensor:
    """boundingbox.correct_yolo_boxes (boundingbox.py:139-149): letterbox_reverse (:95-116) or rescale_bbox
    (:119-137) with clipping to the original image, then x1y1x2y2 -> xywh (:10-15).  fp32 tensor arithmetic
    with python-float ratios, as the reference."""
    if len(bboxes) == 0:
        return bboxes
    b = bboxes.clone().float()
    if is_letterbox:
        ratio = min(img_w / org_w, img_h / org_h)
        rw, rh = int(org_w * ratio), int(org_h * ratio)
        xp, yp = (img_w - rw) // 2, (img_h - rh) // 2
        rx = ry = ratio
    else:
        rx, ry, xp, yp = img_w / org_w, img_h / org_h, 0, 0
    m = b.sum(-1) != 0
    b[m, 0] = torch.clamp((b[m, 0] - xp) / rx, 0, org_w)
    b[m, 2] = torch.clamp((b[m, 2] - xp) / rx, 0, org_w)
    b[m, 1] = torch.clamp((b[m, 1] - yp) / ry, 0, org_h)
    b[m, 3] = torch.clamp((b[m, 3] - yp) / ry, 0, org_h)
    out = b.clone()
    out[:, 2] = b[:, 2] - b[:, 0]
    out[:, 3] = b[:, 3] - b[:, 1]
    return out


# --------------------------------------------------------------------------------------
# Pre-process: letterbox (utils.letterbox_transforms :34-42, letterbox_image :44-57, load_image :60-72)
# --------------------------------------------------------------------------------------
# The arithmetic of the resize lives in a third-party dependency that is not vendored in the reference:
# OpenCV (cv2.resize(..., interpolation=cv2.INTER_CUBIC), utils.py:50; this image has opencv-python 4.13.0).
# Restated here is OpenCV's own portable implementation for 8-bit images (modules/imgproc/src/resize.cpp:
# resizeGeneric_ with HResizeCubic<uchar,int,short> and VResizeCubic<..., FixedPtCast<int,uchar,22>,
# VResizeCubicVec_32s8u>), pinned bit-for-bit against cv2.resize with cv2.ipp.setUseIPP(False) through
# tests/golden/letterbox_golden.npz.  OpenCV builds that carry Intel IPP (the pip wheel does) route this call to
# ippiResizeCubic instead, whose results differ from OpenCV's own code by at most one grey level on ~4 % of the
# pixels; the golden file holds both variants and the tests check "bit-exact" against the former and
# "<= 1 level" against the latter.
_CUBIC_A = np.float32(-0.75)
_COEF_SCALE = 2048                                   # INTER_RESIZE_COEF_SCALE = 1 << 11
_VEC = 8                                             # elements per iteration of VResizeCubicVec_32s8u (2 x 128-bit)


def letterbox_transforms(inner_dim, outer_dim):
    """utils.letterbox_transforms (utils.py:34-42), python-float arithmetic."""
    outer_w, outer_h = outer_dim
    inner_w, inner_h = inner_dim
    ratio = min(outer_w / inner_w, outer_h / inner_h)
    box_w = int(inner_w * ratio)
    box_h = int(inner_h * ratio)
    return box_w, box_h, (outer_w // 2) - (box_w // 2), (outer_h // 2) - (box_h // 2), ratio


def _cubic_coeffs(x: np.ndarray) -> np.ndarray:
    """interpolateCubic(float x, float* coeffs), A = -0.75, every operation rounded to fp32."""
    A, one = _CUBIC_A, np.float32(1)
    x = x.astype(np.float32)
    c0 = ((A * (x + one) - np.float32(5) * A) * (x + one) + np.float32(8) * A) * (x + one) - np.float32(4) * A
    c1 = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
    c2 = ((A + np.float32(2)) * (one - x) - (A + np.float32(3))) * (one - x) * (one - x) + one
    c3 = one - c0 - c1 - c2
    return np.stack([c0, c1, c2, c3], -1).astype(np.float32)


def _cubic_axis(ssize: int, dsize: int):
    """Source offsets and 11-bit fixed-point taps of one axis (resizeGeneric set-up loop):
    f = (float)((d + 0.5) * scale - 0.5) in double then fp32, s = floor(f), taps = cvRound(coeff * 2048)."""
    scale = 1.0 / (dsize / ssize)                    # resize(): inv_scale = dsize/ssize; scale = 1./inv_scale
    f = ((np.arange(dsize, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f)
    frac = (f - s.astype(np.float32)).astype(np.float32)
    taps = np.clip(np.rint(_cubic_coeffs(frac) * np.float32(_COEF_SCALE)), -32768, 32767).astype(np.int32)
    return s.astype(np.int64), taps


def resize_cubic_u8(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(img, (dw, dh), interpolation=cv2.INTER_CUBIC) for uint8 HxWxC, OpenCV's portable code path.
    Horizontal pass: exact int32 sums of 4 border-replicated taps.  Vertical pass: fp32
    S0*b0 + (S1*b1 + (S2*b2 + S3*b3)) with b = tap/2^22, separate roundings, round-half-even, saturate -- except the
    last (dw*C) % 8 elements of every row, which the scalar tail computes as (sum + 2^21) >> 22."""
    sh, sw, cn = img.shape
    xofs, xa = _cubic_axis(sw, dw)
    yofs, yb = _cubic_axis(sh, dh)
    S = img.astype(np.int32)
    xi = np.clip(xofs[:, None] + np.arange(-1, 3)[None, :], 0, sw - 1)
    H = (S[:, xi, :] * xa[None, :, :, None]).sum(2)                  # [sh, dw, cn] int32, exact
    yi = np.clip(yofs[:, None] + np.arange(-1, 3)[None, :], 0, sh - 1)
    R = H[yi]                                                        # [dh, 4, dw, cn]
    b = yb.astype(np.float32) * (np.float32(1.0) / np.float32(_COEF_SCALE * _COEF_SCALE))
    Rf = R.astype(np.float32)
    t = (Rf[:, 3] * b[:, 3, None, None]).astype(np.float32)
    for k in (2, 1, 0):
        t = ((Rf[:, k] * b[:, k, None, None]).astype(np.float32) + t).astype(np.float32)
    out = np.clip(np.rint(t), 0, 255).astype(np.uint8).reshape(dh, dw * cn)
    tail = (dw * cn) // _VEC * _VEC
    if tail < dw * cn:
        v = (R.astype(np.int64) * yb[:, :, None, None]).sum(1).reshape(dh, dw * cn)
        out[:, tail:] = np.clip((v[:, tail:] + (1 << 21)) >> 22, 0, 255).astype(np.uint8)
    return out.reshape(dh, dw, cn)


def _linear_axis(ssize: int, dsize: int, clamp: bool):
    """Offsets and 11-bit taps of one axis for INTER_LINEAR.  The x axis clamps (s < 0 -> s = 0, f = 0;
    s >= ssize-1 -> s = ssize-1, f = 0); the y axis keeps its fraction and only clips the two row indices."""
    scale = 1.0 / (dsize / ssize)
    f = ((np.arange(dsize, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    frac = (f - s.astype(np.float32)).astype(np.float32)
    if clamp:
        lo, hi = s < 0, s >= ssize - 1
        frac[lo] = 0
        s[lo] = 0
        frac[hi] = 0
        s[hi] = ssize - 1
    taps = np.stack([(np.float32(1) - frac).astype(np.float32), frac], -1)
    return s, np.clip(np.rint(taps * np.float32(_COEF_SCALE)), -32768, 32767).astype(np.int32)


def resize_linear_u8(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(img, (dw, dh)) (default INTER_LINEAR) for uint8 HxWxC, OpenCV's own 8-bit path (resize.cpp:
    HResizeLinear<uchar,int,short> + VResizeLinear<uchar,int,short,...,VResizeLinearVec_32s8u>): exact int32
    horizontal pass, then ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2, vector body and scalar tail alike.
    IPP builds do not divert this call (8-bit linear is excluded there), so one golden covers both."""
    sh, sw, cn = img.shape
    xofs, xa = _linear_axis(sw, dw, True)
    yofs, yb = _linear_axis(sh, dh, False)
    S = img.astype(np.int32)
    H = S[:, xofs, :] * xa[None, :, 0, None] + S[:, np.clip(xofs + 1, 0, sw - 1), :] * xa[None, :, 1, None]
    S0, S1 = H[np.clip(yofs, 0, sh - 1)], H[np.clip(yofs + 1, 0, sh - 1)]
    b0, b1 = yb[:, 0][:, None, None], yb[:, 1][:, None, None]
    t = ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16)
    return np.clip((t + 2) >> 2, 0, 255).astype(np.uint8)


def load_image_resize(img_rgb_u8: np.ndarray, dim) -> torch.Tensor:
    """utils.load_image(path, 'resize', dim) after the file decode (utils.py:68-71): cv2.resize(img, dim), /255, CHW."""
    out = resize_linear_u8(img_rgb_u8, int(dim[0]), int(dim[1]))
    return torch.from_numpy(out).float().permute(2, 0, 1) / 255


def letterbox_image(img: np.ndarray, dim) -> Tuple[np.ndarray, torch.Tensor]:
    """utils.letterbox_image (utils.py:44-57): grey (128) canvas np.full(dim + (3,)), bicubic-resized image pasted
    at the centred offset; returns the canvas (int64, as np.full gives) and Tensor([box_w, box_h, box_x, box_y, ratio])."""
    image = np.full(tuple(dim) + (3,), 128)
    img_dim = (img.shape[1], img.shape[0])
    box_w, box_h, box_x, box_y, ratio = letterbox_transforms(img_dim, dim)
    box_image = resize_cubic_u8(img, box_w, box_h)
    image[box_y:box_y + box_h, box_x:box_x + box_w] = box_image
    return image, torch.Tensor([box_w, box_h, box_x, box_y, ratio])


def iaa_letterbox(img: np.ndarray, dim) -> np.ndarray:
    """IaaLetterbox._augment_images (transforms.py:153-176) for one uint8 image, dim = (width, height): bicubic
    resize (imgaug.imresize_single_image(..., 'cubic') is cv2.resize(..., INTER_CUBIC); imgaug itself is not in
    this image, so this leg is pinned through the cv2-pinned resize above), then np.pad with 128 using
    _compute_height_width_pad's offsets (:203-210)."""
    width, height = int(dim[0]), int(dim[1])
    img_h, img_w = img.shape[:2]
    ratio = min(width / img_w, height / img_h)
    rw, rh = int(img_w * ratio), int(img_h * ratio)
    x_pad, y_pad = (width - rw) // 2, (height - rh) // 2
    rs = resize_cubic_u8(img, rw, rh)
    return np.pad(rs, ((y_pad, height - rh - y_pad), (x_pad, width - rw - x_pad), (0, 0)), mode="constant", constant_values=128)


def load_image_letterbox(img_rgb_u8: np.ndarray, dim) -> Tuple[torch.Tensor, torch.Tensor]:
    """utils.load_image(path, 'letterbox', dim) after the file decode (utils.py:60-72): letterbox, then
    torch.from_numpy(img).float().permute(2,0,1) / 255."""
    image, trans = letterbox_image(img_rgb_u8, dim)
    return torch.from_numpy(image).float().permute(2, 0, 1) / 255, trans


# --------------------------------------------------------------------------------------
# Darknet binary weight stream (WeightManager, darknet.py:249-303)
# --------------------------------------------------------------------------------------
def darknet_blob_from_state_dict(sd, num_classes: int = 80, backbone_only: bool = False) -> np.ndarray:
    """Serialise a state_dict to the float stream WeightManager.loadWeight consumes
    (per BN conv: bn.bias, bn.weight, running_mean, running_var, conv.weight -- darknet.py:279-285;
    per plain conv: bias, weight -- :287-290), without the 5-int header."""
    parts = []
    for e in conv_table(num_classes):
        if backbone_only and not e["key"].startswith("feature."):
            break
        k = e["key"]
        if e["bn"]:
            for s in (".bn.bias", ".bn.weight", ".bn.running_mean", ".bn.running_var", ".conv.weight"):
                parts.append(sd[k + s].detach().cpu().numpy().astype(np.float32).ravel())
        else:
            parts.append(sd[k + ".bias"].detach().cpu().numpy().astype(np.float32).ravel())
            parts.append(sd[k + ".weight"].detach().cpu().numpy().astype(np.float32).ravel())
    return np.concatenate(parts)


def state_dict_from_darknet_blob(blob: np.ndarray, num_classes: int = 80, backbone_only: bool = False):
    """WeightManager.loadWeight (darknet.py:254-263) as a pure function: returns (state_dict, consumed)."""
    sd, ptr = {}, 0

    def take(shape):
        nonlocal ptr
        n = int(np.prod(shape))
        v = torch.from_numpy(np.asarray(blob[ptr:ptr + n], dtype=np.float32).copy()).view(*shape)
        ptr += n
        return v

    for e in conv_table(num_classes):
        if backbone_only and not e["key"].startswith("feature."):
            break
        k, co, ci, ks = e["key"], e["cout"], e["cin"], e["ks"]
        if e["bn"]:
            sd[k + ".bn.bias"] = take((co,))
            sd[k + ".bn.weight"] = take((co,))
            sd[k + ".bn.running_mean"] = take((co,))
            sd[k + ".bn.running_var"] = take((co,))
            sd[k + ".conv.weight"] = take((co, ci, ks, ks))
        else:
            sd[k + ".bias"] = take((co,))
            sd[k + ".weight"] = take((co, ci, ks, ks))
    return sd, ptr


# --------------------------------------------------------------------------------------
# Fast C restatement of the post-process (oracle/nms_oracle.c), for sizes where the
# torch/python loop above would take minutes.  Cross-checked against postprocessing()
# and the goldens in tests/test_oracle_golden.py.
# --------------------------------------------------------------------------------------
_C_LIB = None


def c_lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libyolo_oracle.so")


def _c_lib():
    global _C_LIB
    if _C_LIB is None:
        lib = ctypes.CDLL(c_lib_path())
        lib.oracle_postprocess.restype = ctypes.c_long
        lib.oracle_postprocess.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_long, ctypes.c_long,
                                           ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
        _C_LIB = lib
    return _C_LIB


def postprocessing_c(det: torch.Tensor, num_classes: int, obj_conf_thr=0.5, nms_thr=0.4,
                     is_eval=False, use_nms=True):
    """Same contract as postprocessing(..., return_index=True), computed by oracle/nms_oracle.c."""
    d = np.ascontiguousarray(det.detach().cpu().numpy(), dtype=np.float32)
    nB, nN, nA = d.shape
    assert nA == 5 + num_classes
    cap = nN * (num_classes if is_eval else 1)
    rows = np.empty((nB, cap, 7), np.float32)
    src = np.empty((nB, cap), np.int64)
    counts = np.zeros(nB, np.int64)
    total = _c_lib().oracle_postprocess(d.ctypes.data, nB, nN, num_classes, obj_conf_thr, nms_thr,
                                        int(is_eval), int(use_nms), rows.ctypes.data, src.ctypes.data,
                                        counts.ctypes.data, cap)
    if total < 0:
        return [], []
    out, idx = [], []
    for b in range(nB):
        k = int(counts[b])
        out.append(torch.from_numpy(rows[b, :k].copy()) if k or total < 0 else torch.Tensor())
        if k == 0:
            out[-1] = torch.Tensor()
        idx.append(src[b, :k].copy())
    return out, idx
