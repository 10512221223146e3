"""Host mirror of the reference's utils.py entry points on the inference path: ``postprocessing``
(reference utils.py:226-258) and the letterbox pre-process ``letterbox_transforms`` / ``letterbox_image`` /
``load_image`` (utils.py:34-72).

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
import threading
from typing import Dict, Sequence, Tuple

import numpy as np
import torch

from . import _lib

_POST_CTX: Dict[Tuple[int, int, int], ctypes.c_void_p] = {}
_POST_CTX_LOCK = threading.Lock()


def _ctx_for(device_index: int, num_classes: int):
    """One library context per (device, class count, HOST THREAD): a yb_ctx owns scratch buffers that every call reuses
    (include/yolo_b200.h: one ctx per device and host thread, calls serialised by the caller), so threads -- DataLoader
    workers, serving threads -- must not share one."""
    key = (device_index, num_classes, threading.get_ident())
    ctx = _POST_CTX.get(key)
    if ctx is None:
        with _POST_CTX_LOCK:
            ctx = _POST_CTX.get(key)
            if ctx is None:
                ctx = _POST_CTX[key] = _lib.create_ctx(device_index, num_classes, None)
    return ctx


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


# ---- pre-process: letterbox (reference utils.py:34-72) -------------------------------------------------------
def letterbox_transforms(inner_dim, outer_dim):
    """Reference utils.letterbox_transforms (utils.py:34-42); pure host integer / float arithmetic."""
    outer_w, outer_h = outer_dim
    inner_w, inner_h = inner_dim
    ratio = min(outer_w / inner_w, outer_h / inner_h)
    box_w = int(inner_w * ratio)
    box_h = int(inner_h * ratio)
    box_x_offset = (outer_w // 2) - (box_w // 2)
    box_y_offset = (outer_h // 2) - (box_h // 2)
    return box_w, box_h, box_x_offset, box_y_offset, ratio


def _as_cuda_u8(img, device):
    if isinstance(img, torch.Tensor):
        t = img
    else:
        a = np.ascontiguousarray(img)
        if a.dtype != np.uint8:
            raise TypeError("letterbox expects uint8 images (what cv2.imread returns)")
        t = torch.from_numpy(a)
    if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
        raise ValueError("letterbox expects [H,W,3] uint8 images")
    return t.to(device, non_blocking=True).contiguous()


def letterbox_batch(imgs: Sequence, dim, device=None, want_canvas=False, canvas_hw=None, iaa=False):
    """Letterboxes a list of [H,W,3] uint8 RGB images (numpy arrays or tensors, sizes may differ) on the GPU in one
    launch (yb_letterbox).  Returns (x, trans): x = CUDA fp32 [B,3,dim[0],dim[1]] = what ``load_image(path,
    'letterbox', dim)[0]`` holds for every image, trans = CPU fp32 [B,5] = box_w, box_h, box_x, box_y, ratio.
    With want_canvas=True the uint8 [B,dim[0],dim[1],3] canvases are returned instead of x.  The canvas is
    dim[0] rows by dim[1] columns exactly as the reference's np.full(dim + (3,), 128) (utils.py:46) unless canvas_hw
    overrides it.  iaa=True is the dataset transform IaaLetterbox (transforms.py:144-212): the same bicubic resize,
    offsets (w - box_w)//2 instead of w//2 - box_w//2 and a [dim[1], dim[0]] canvas."""
    if not torch.cuda.is_available():
        raise RuntimeError("yolo_v3_b200 runs on CUDA devices only (no CPU fallback)")
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    index = device.index if device.index is not None else torch.cuda.current_device()
    B = len(imgs)
    if B == 0:
        raise ValueError("empty image list")
    dev_imgs = [_as_cuda_u8(im, device) for im in imgs]
    ptrs = (ctypes.c_void_p * B)(*[t.data_ptr() for t in dev_imgs])
    hw = (ctypes.c_int * (2 * B))(*[int(v) for t in dev_imgs for v in t.shape[:2]])
    trans = (ctypes.c_float * (5 * B))()
    dim_w, dim_h = int(dim[0]), int(dim[1])
    if canvas_hw is not None:
        ch, cw = int(canvas_hw[0]), int(canvas_hw[1])
    else:
        ch, cw = (dim_h, dim_w) if iaa else (dim_w, dim_h)
    out = canvas = None
    if want_canvas:
        canvas = torch.empty(B, ch, cw, 3, device=device, dtype=torch.uint8)
    else:
        out = torch.empty(B, 3, ch, cw, device=device, dtype=torch.float32)
    lib = _lib.load()
    ctx = _ctx_for(index, 80)
    with torch.cuda.device(device):
        _lib.check(lib.yb_letterbox(ctx, ptrs, hw, B, dim_w, dim_h, ch, cw, int(bool(iaa)),
                                    ctypes.c_void_p(out.data_ptr()) if out is not None else None,
                                    ctypes.c_void_p(canvas.data_ptr()) if canvas is not None else None,
                                    trans, ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)), ctx)
    t = torch.tensor(list(trans), dtype=torch.float32).reshape(B, 5)
    return (canvas if want_canvas else out), t


def resize_batch(imgs: Sequence, dim, device=None, want_u8=False):
    """cv2.resize(img, dim) (INTER_LINEAR) + float()/255 + HWC->CHW for a list of [H,W,3] uint8 images in one launch
    (yb_resize).  Returns CUDA fp32 [B,3,dim[1],dim[0]] (or the uint8 [B,dim[1],dim[0],3] images with want_u8)."""
    if not torch.cuda.is_available():
        raise RuntimeError("yolo_v3_b200 runs on CUDA devices only (no CPU fallback)")
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    index = device.index if device.index is not None else torch.cuda.current_device()
    B = len(imgs)
    if B == 0:
        raise ValueError("empty image list")
    dev_imgs = [_as_cuda_u8(im, device) for im in imgs]
    ptrs = (ctypes.c_void_p * B)(*[t.data_ptr() for t in dev_imgs])
    hw = (ctypes.c_int * (2 * B))(*[int(v) for t in dev_imgs for v in t.shape[:2]])
    dim_w, dim_h = int(dim[0]), int(dim[1])
    out = u8 = None
    if want_u8:
        u8 = torch.empty(B, dim_h, dim_w, 3, device=device, dtype=torch.uint8)
    else:
        out = torch.empty(B, 3, dim_h, dim_w, device=device, dtype=torch.float32)
    lib = _lib.load()
    ctx = _ctx_for(index, 80)
    with torch.cuda.device(device):
        _lib.check(lib.yb_resize(ctx, ptrs, hw, B, dim_w, dim_h,
                                 ctypes.c_void_p(out.data_ptr()) if out is not None else None,
                                 ctypes.c_void_p(u8.data_ptr()) if u8 is not None else None,
                                 ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)), ctx)
    return u8 if want_u8 else out


def letterbox_image(img, dim):
    """Reference signature (utils.py:44-57): [H,W,3] uint8 image -> (canvas, transform).  The canvas is the numpy
    integer array the reference builds with np.full(dim + (3,), 128); the resize + paste run on the GPU."""
    try:
        canvas, trans = letterbox_batch([img], dim, want_canvas=True)
    except _lib.YbError as e:
        if e.code == -1:                               # numpy's error for a paste that does not fit the canvas
            raise ValueError(str(e)) from None
        raise
    return canvas[0].cpu().numpy().astype(np.int64), torch.Tensor(trans[0].tolist())


def load_image(img_path, mode=None, dim=None):
    """Reference signature (utils.py:60-72).  The file decode (cv2.imread + BGR->RGB) stays on the host -- it is not
    part of the path -- and the letterbox / scaling / HWC->CHW run on the GPU.  Returns (CUDA fp32 [3,H,W], trans)."""
    import cv2
    img = cv2.imread(img_path)
    if img is None:
        raise FileNotFoundError(img_path)
    img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    if mode is not None and dim is not None:
        if mode == "letterbox":
            x, trans = letterbox_batch([img], dim)
            return x[0], torch.Tensor(trans[0].tolist())
        if mode == "resize":
            return resize_batch([img], dim)[0], None
    # no geometric change: the same kernel with a full-size box (taps 0,1,0,0) is an exact copy, /255, HWC->CHW
    h, w = img.shape[:2]
    x, _ = letterbox_batch([img], (w, h), canvas_hw=(h, w))
    return x[0], None
