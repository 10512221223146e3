"""Host mirror of the reference's YoloLayer (reference yololayer.py), inference branch only.

Inside YoloNet the three decodes are one kernel launched by yb_forward; this class exists so that
code which builds or calls a YoloLayer directly keeps working: ``YoloLayer(anchors, mask, img_dim,
numClass)(x, img_dim, None)`` decodes one raw head map [B, 3*(5+C), h, w] on the GPU (yb_decode)
and returns [B, 3*h*w, 5+C] exactly like yololayer.py:31-59, 97-105.  The loss / target branch
(yololayer.py:64-95, 107-172) is training code and out of scope: ``target is not None`` raises.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _lib


class YoloLayer(nn.Module):
    def __init__(self, anchors_all, anchors_mask, img_dim, numClass):
        super().__init__()
        self.anchors_all = anchors_all
        self.anchors_mask = anchors_mask
        self.img_dim = img_dim
        self.numClass = numClass
        self.bbox_attrib = 5 + numClass
        self._ctx = None
        self._ctx_device = None

    def _scale_index(self):
        masks = ([6, 7, 8], [3, 4, 5], [0, 1, 2])
        m = list(self.anchors_mask)
        if m not in [list(x) for x in masks]:
            raise NotImplementedError("only the reference's three anchor masks are supported")
        return [list(x) for x in masks].index(m)

    def forward(self, x, img_dim, target=None):
        if target is not None:
            raise NotImplementedError("training branch of YoloLayer is out of scope for the B200 inference path")
        if not x.is_cuda:
            raise RuntimeError("yolo_v3_b200 runs on CUDA devices only (no CPU fallback)")
        lib = _lib.load()
        index = x.device.index if x.device.index is not None else torch.cuda.current_device()
        if self._ctx is None or self._ctx_device != index:
            flat = [float(v) for pair in self.anchors_all for v in pair]
            self._ctx = _lib.create_ctx(index, self.numClass, flat)
            self._ctx_device = index
        B, ch, h, w = x.shape
        if ch != 3 * self.bbox_attrib:
            raise ValueError("channel count does not match 3*(5+numClass)")
        si = self._scale_index()
        s = (32, 16, 8)[si]
        W, H = int(img_dim[0]), int(img_dim[1])
        if H != h * s or W != w * s:
            raise ValueError(f"img_dim {img_dim} does not match a stride-{s} head of size {(h, w)}")
        # yb_decode decodes the three scales together into one [B,N,5+C] tensor; feed zeros for the
        # other two scales and return this scale's rows.
        x = x.float().contiguous()
        maps = [torch.zeros(B, ch, H // t, W // t, device=x.device) if t != s else x for t in (32, 16, 8)]
        n = [3 * (H // t) * (W // t) for t in (32, 16, 8)]
        det = torch.empty(B, sum(n), self.bbox_attrib, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.yb_decode(self._ctx, *[ctypes.c_void_p(m.data_ptr()) for m in maps], B, H, W,
                                     ctypes.c_void_p(det.data_ptr()),
                                     ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), self._ctx)
        off = sum(n[:si])
        return det[:, off:off + n[si]]

    def __del__(self):
        try:
            if self._ctx is not None:
                _lib.load().yb_destroy(self._ctx)
        except Exception:
            pass


_DECODE_CTX = {}


def decode_heads(maps, img_dim, numClass=80, anchors=None):
    """The three YoloLayer.forward calls of YoloNet.forward plus the callers' torch.cat((det1,det2,det3),1)
    (yololayer.py:31-59, test.py:36) in ONE launch: `maps` = the raw head maps [B,3*(5+C),H/32,W/32], [.., H/16, W/16],
    [.., H/8, W/8] (CUDA, NCHW fp32), img_dim = (W, H).  Returns the concatenated [B,N,5+C] tensor."""
    from .topology import DEFAULT_ANCHORS
    x0 = maps[0]
    if not x0.is_cuda:
        raise RuntimeError("yolo_v3_b200 runs on CUDA devices only (no CPU fallback)")
    lib = _lib.load()
    index = x0.device.index if x0.device.index is not None else torch.cuda.current_device()
    flat = tuple(float(v) for v in (anchors if anchors is not None else DEFAULT_ANCHORS))
    key = (index, numClass, flat)
    if key not in _DECODE_CTX:
        _DECODE_CTX[key] = _lib.create_ctx(index, numClass, list(flat))
    ctx = _DECODE_CTX[key]
    W, H = int(img_dim[0]), int(img_dim[1])
    B = x0.shape[0]
    ms = [m.float().contiguous() for m in maps]
    for m, t in zip(ms, (32, 16, 8)):
        if tuple(m.shape) != (B, 3 * (5 + numClass), H // t, W // t):
            raise ValueError(f"head map of shape {tuple(m.shape)} does not match img_dim {img_dim} at stride {t}")
    n = sum(3 * (H // t) * (W // t) for t in (32, 16, 8))
    det = torch.empty(B, n, 5 + numClass, device=x0.device)
    with torch.cuda.device(x0.device):
        _lib.check(lib.yb_decode(ctx, *[ctypes.c_void_p(m.data_ptr()) for m in ms], B, H, W, ctypes.c_void_p(det.data_ptr()),
                                 ctypes.c_void_p(torch.cuda.current_stream(x0.device).cuda_stream)), ctx)
    return det
