"""Host mirror of the reference's box post-correction (reference boundingbox.py:95-149), the step every
caller applies right after postprocessing (test.py:41, evaluate.py:187): undo the letterbox / resize,
clip to the original image and convert x1y1x2y2 -> xywh.  The arithmetic runs in libyolo_b200.so
(yb_correct_boxes); `correct_yolo_boxes` keeps the reference signature for one image,
`correct_yolo_boxes_batch` corrects the fixed-capacity device rows of a whole batch without leaving the GPU.
"""
from __future__ import annotations

import ctypes
from typing import Sequence, Tuple

import torch

from . import _lib
from .utils import _ctx_for


def correct_yolo_boxes_batch(rows: torch.Tensor, counts, org_sizes: Sequence[Tuple[int, int]], img_w: int, img_h: int,
                             is_letterbox: bool = False) -> torch.Tensor:
    """rows: CUDA [B,cap,S>=4] (x1,y1,x2,y2 first), counts: CUDA int32 [B] or None, org_sizes: B x (w,h).
    Returns CUDA [B,cap,4] xywh in original-image pixels (rows past counts[b] are left untouched)."""
    if not rows.is_cuda:
        raise RuntimeError("yolo_v3_b200 runs on CUDA devices only (no CPU fallback)")
    rows = rows.float().contiguous()
    B, cap, stride = rows.shape
    if len(org_sizes) != B:
        raise ValueError("one (w, h) per image expected")
    lib = _lib.load()
    index = rows.device.index if rows.device.index is not None else torch.cuda.current_device()
    ctx = _ctx_for(index, 80)
    flat = (ctypes.c_int * (2 * B))(*[int(v) for wh in org_sizes for v in wh])
    out = torch.zeros(B, cap, 4, device=rows.device, dtype=torch.float32)
    cptr = None
    if counts is not None:
        counts = counts.to(device=rows.device, dtype=torch.int32).contiguous()
        cptr = ctypes.c_void_p(counts.data_ptr())
    with torch.cuda.device(rows.device):
        _lib.check(lib.yb_correct_boxes(ctx, ctypes.c_void_p(rows.data_ptr()), stride, cptr, B, cap, flat, int(img_w), int(img_h),
                                        int(bool(is_letterbox)), ctypes.c_void_p(out.data_ptr()),
                                        ctypes.c_void_p(torch.cuda.current_stream(rows.device).cuda_stream)), ctx)
    return out


def correct_yolo_boxes(bboxes, org_w, org_h, img_w, img_h, is_letterbox=False):
    """Reference signature (boundingbox.py:139): bboxes [K,4] x1y1x2y2 -> [K,4] xywh, same device as the input."""
    if len(bboxes) == 0:
        return bboxes
    if not torch.cuda.is_available():
        raise RuntimeError("yolo_v3_b200 runs on CUDA devices only (no CPU fallback)")
    was_cpu = not bboxes.is_cuda
    b = bboxes.cuda() if was_cpu else bboxes
    out = correct_yolo_boxes_batch(b.reshape(1, -1, b.shape[-1]), None, [(org_w, org_h)], img_w, img_h, is_letterbox)[0]
    return out.cpu() if was_cpu else out
