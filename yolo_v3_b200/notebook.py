"""Host mirror of the post-process the reference's notebook defines inline (yolo_detect.ipynb cell 35, helpers in
cells 30 and 33), so that those cells can be swapped one for one:

    from yolo_v3_b200.notebook import postprocessing

It is not utils.postprocessing: the threshold is on objectness, the class is the arg-max of the raw class
probabilities, boxes of a class are ordered by objectness and x2 = (cx - w/2) + w.  The work runs in
libyolo_b200.so (yb_postprocess_notebook); the notebook's return convention is kept -- a list with one CPU tensor
[K,7] = x1,y1,x2,y2,obj,class_prob,cls per image, an empty tensor for an image without detections.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .utils import _ctx_for


def postprocessing_raw(detections: torch.Tensor, num_classes: int, obj_conf_thr=0.5, nms_thr=0.4, cap=None):
    """Device-side results (rows7 [B,cap,7], counts [B], src_index [B,cap], cand_counts [B]); nothing synchronised."""
    if not torch.cuda.is_available():
        raise RuntimeError("yolo_v3_b200 runs on CUDA devices only (no CPU fallback)")
    if detections.dim() != 3 or detections.shape[2] != 5 + num_classes:
        raise ValueError("expected detections of shape [B, N, 5+num_classes]")
    if obj_conf_thr < 0:
        raise ValueError("obj_conf_thr must be >= 0 (the notebook drops rows whose objectness is zeroed)")
    det = detections if detections.is_cuda else detections.cuda(non_blocking=True)
    det = det.float().contiguous()
    B, N, _ = det.shape
    if cap is None:
        cap = min(N, 4096)
    lib = _lib.load()
    index = det.device.index if det.device.index is not None else torch.cuda.current_device()
    ctx = _ctx_for(index, num_classes)
    rows = torch.empty(B, cap, 7, device=det.device, dtype=torch.float32)
    counts = torch.empty(B, device=det.device, dtype=torch.int32)
    src = torch.empty(B, cap, device=det.device, dtype=torch.int32)
    cand = torch.empty(B, device=det.device, dtype=torch.int32)
    with torch.cuda.device(det.device):
        _lib.check(lib.yb_postprocess_notebook(ctx, ctypes.c_void_p(det.data_ptr()), B, N, float(obj_conf_thr), float(nms_thr),
                                               ctypes.c_void_p(rows.data_ptr()), ctypes.c_void_p(counts.data_ptr()),
                                               ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(cand.data_ptr()), int(cap),
                                               ctypes.c_void_p(torch.cuda.current_stream(det.device).cuda_stream)), ctx)
    return rows, counts, src, cand


def postprocessing(detections, num_classes, obj_conf_thr=0.5, nms_thr=0.4, return_index=False):
    cap = None
    while True:
        rows, counts, src, _ = postprocessing_raw(detections, num_classes, obj_conf_thr, nms_thr, cap)
        counts_h = counts.cpu()
        mx = int(counts_h.max())
        if mx <= rows.shape[1]:
            break
        cap = mx
    host = rows[:, :max(mx, 1)].cpu()
    res = [host[b, :int(counts_h[b])].clone() if int(counts_h[b]) else torch.Tensor() for b in range(rows.shape[0])]
    if not return_index:
        return res
    src_h = src[:, :max(mx, 1)].cpu()
    return res, [src_h[b, :int(counts_h[b])].long().numpy() for b in range(rows.shape[0])]
