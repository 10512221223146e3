"""Host mirror of the reference's post-process entry point (reference utils.py:226-258).

``postprocessing(detections, num_classes, obj_conf_thr=0.5, nms_thr=0.4, is_eval=False, use_nms=True)``
keeps the reference's signature and return convention -- a list (one entry per image) of CPU fp32
tensors [K,7] = x1,y1,x2,y2,obj,score,cls ordered class-ascending / score-descending, ``torch.Tensor()``
for an image with no candidate, and ``[]`` when nothing in the whole batch passes -- but the work
(box convert, score, threshold, sort, IOU, greedy NMS) runs in libyolo_b200.so's kernels
(yb_postprocess).  Only the surviving rows cross PCIe, instead of the whole [B,N,85] tensor the
reference copies with ``.cpu()`` (utils.py:227).

Differences from the reference: the input is never modified (the reference mutates a CPU input in
place); ties in score are broken deterministically (score descending, then
candidate order ascending) where the reference's unstable CPU sort is arbitrary.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Tuple

import torch

from . import _lib

_POST_CTX: Dict[Tuple[int, int], ctypes.c_void_p] = {}


def _ctx_for(device_index: int, num_classes: int):
    key = (device_index, num_classes)
    if key not in _POST_CTX:
        _POST_CTX[key] = _lib.create_ctx(device_index, num_classes, None)
    return _POST_CTX[key]


def rows_to_list(rows: torch.Tensor, counts_h: torch.Tensor, cand_h: torch.Tensor):
    """Device rows [B,cap,7] + host counts -> the reference's list-of-CPU-tensors convention."""
    if int(cand_h.sum()) == 0:
        return []                                      # utils.py:247-251
    mx = int(counts_h.max())
    host = rows[:, :max(mx, 1)].cpu()
    out = []
    for b in range(rows.shape[0]):
        if int(cand_h[b]) == 0:
            out.append(torch.Tensor())                 # utils.py:153-158
        else:
            out.append(host[b, :int(counts_h[b])].clone())
    return out


def postprocessing_raw(detections: torch.Tensor, num_classes: int, obj_conf_thr=0.5, nms_thr=0.4,
                       is_eval=False, use_nms=True, cap=None):
    """Device-side results, nothing synchronised: rows7 [B,cap,7], counts [B] (int32), src_index
    [B,cap] (flat box index of each row), cand_counts [B]."""
    if not torch.cuda.is_available():
        raise RuntimeError("yolo_v3_b200 runs on CUDA devices only (no CPU fallback)")
    if detections.dim() != 3 or detections.shape[2] != 5 + num_classes:
        raise ValueError("expected detections of shape [B, N, 5+num_classes]")
    det = detections
    if not det.is_cuda:
        det = det.cuda(non_blocking=True)              # still the CUDA path: upload, never compute on the host
    det = det.float().contiguous()
    B, N, _ = det.shape
    lib = _lib.load()
    index = det.device.index if det.device.index is not None else torch.cuda.current_device()
    ctx = _ctx_for(index, num_classes)
    if cap is None:
        full = N * (num_classes if is_eval else 1)
        cap = full if not use_nms else min(full, 4096)
    rows = torch.empty(B, cap, 7, device=det.device, dtype=torch.float32)
    counts = torch.empty(B, device=det.device, dtype=torch.int32)
    src = torch.empty(B, cap, device=det.device, dtype=torch.int32)
    cand = torch.empty(B, device=det.device, dtype=torch.int32)
    with torch.cuda.device(det.device):
        _lib.check(lib.yb_postprocess(ctx, ctypes.c_void_p(det.data_ptr()), B, N, float(obj_conf_thr), float(nms_thr),
                                      int(bool(is_eval)), int(bool(use_nms)), ctypes.c_void_p(rows.data_ptr()),
                                      ctypes.c_void_p(counts.data_ptr()), ctypes.c_void_p(src.data_ptr()),
                                      ctypes.c_void_p(cand.data_ptr()), int(cap),
                                      ctypes.c_void_p(torch.cuda.current_stream(det.device).cuda_stream)), ctx)
    return rows, counts, src, cand


def postprocessing(detections, num_classes, obj_conf_thr=0.5, nms_thr=0.4, is_eval=False, use_nms=True,
                   return_index=False):
    cap = None
    while True:
        rows, counts, src, cand = postprocessing_raw(detections, num_classes, obj_conf_thr, nms_thr, is_eval, use_nms, cap)
        counts_h = counts.cpu()
        mx = int(counts_h.max())
        if mx <= rows.shape[1]:
            break
        cap = mx                                       # capacity was too small: one retry with the true maximum
    cand_h = cand.cpu()
    res = rows_to_list(rows, counts_h, cand_h)
    if not return_index:
        return res
    if not res:
        return res, []
    src_h = src[:, :max(mx, 1)].cpu()
    return res, [src_h[b, :int(counts_h[b])].long().numpy() for b in range(rows.shape[0])]
