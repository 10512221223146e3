"""Host mirror of the reference's model API (reference darknet.py) over the CUDA library.

`YoloNet` keeps the reference's constructor, `forward(x, target=None)`, `loadWeight` / `saveWeight`
signatures and registers parameters and buffers under the reference's exact names (438 state_dict
entries, e.g. ``feature.mlist.2.conv1.conv.weight``, ``pre_det3.mlist.6.bias``), so
``load_state_dict``, ``state_dict``, ``torch.save`` checkpoints and darknet ``.weights`` files
interoperate.  The modules below only *hold* tensors: no torch operator computes anything.  The
forward pass is one call into libyolo_b200.so (yb_forward), which runs the hand-written sm_100a
kernels; there is no PyTorch/CPU fallback -- without a CUDA device the call raises.

Differences from the reference, all deliberate:
  * inference only: ``target is not None`` (training, yololayer.py:64-95) raises NotImplementedError,
    and so does a forward in ``.train()`` mode;
  * det1/det2/det3 are row-slices of one [B,N,5+C] CUDA tensor (the callers' torch.cat((det1,det2,det3),1)
    still works and yields the same values);
  * ``saveWeight(format='darknet')`` is implemented (the reference raises NotImplementedError,
    darknet.py:237-238);
  * extra keyword ``precision``: 'fp16' = the tcgen05 tensor-core path the benchmark measures; 'fp32' = fp32-grade
    parity with the reference, also on the tensor cores (YB_MODE_FP32_TC: fp16 hi/lo operand pairs, three partial
    products per k-step, fp32 accumulation in TMEM); 'fp32_simt' = the CUDA-core fp32 debugging path.

Weight changes are picked up automatically when they go through ``load_state_dict``, ``loadWeight``, ``.cuda()`` /
``.to()`` / ``.half()``, an optimizer-style in-place op on the parameter itself, or re-assignment.  An in-place write
through ``param.data`` (``p.data.copy_(...)``, what the reference's WeightManager does internally, darknet.py:275) does
NOT bump the tensor's version counter and cannot be seen without reading the weights back: call
``net.refresh_weights()`` after such a write (or construct with ``check_weights=True`` to compare a device-side
checksum of every tensor on each forward, at the cost of one synchronisation per call).
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .topology import BLOCKS, DEFAULT_ANCHORS, darknet_stream_keys, num_boxes
from .yololayer import YoloLayer


class _ConvParams(nn.Module):
    """Holds ``weight`` (and ``bias``) with nn.Conv2d's names and shapes; never executed."""

    def __init__(self, nin, nout, ks, bias):
        super().__init__()
        w = torch.empty(nout, nin, ks, ks)
        nn.init.kaiming_uniform_(w, a=5 ** 0.5)            # nn.Conv2d's default init
        self.weight = nn.Parameter(w)
        if bias:
            bound = 1.0 / (nin * ks * ks) ** 0.5
            self.bias = nn.Parameter(torch.empty(nout).uniform_(-bound, bound))
        else:
            self.register_parameter("bias", None)


class _BNParams(nn.Module):
    """Holds nn.BatchNorm2d's parameters and running statistics; never executed."""

    def __init__(self, n):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n))
        self.bias = nn.Parameter(torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        self.eps = 1e-5


class conv_bn_relu(nn.Module):
    """Parameter container for the reference block of the same name (darknet.py:27-44):
    ``.conv.weight``, ``.bn.{weight,bias,running_mean,running_var,num_batches_tracked}``."""

    def __init__(self, nin, nout, ks, s=1):
        super().__init__()
        self.nin, self.nout, self.ks, self.s = nin, nout, ks, s
        self.conv = _ConvParams(nin, nout, ks, bias=False)
        self.bn = _BNParams(nout)


class res_layer(nn.Module):
    """Container for darknet.py:46-53: conv1 (1x1, halves channels), conv2 (3x3)."""

    def __init__(self, nin):
        super().__init__()
        self.conv1 = conv_bn_relu(nin, nin // 2, 1)
        self.conv2 = conv_bn_relu(nin // 2, nin, 3)


class Darknet(nn.Module):
    """Container for the Darknet-53 backbone (darknet.py:72-104); 29 entries in ``mlist``."""

    def __init__(self, blkList=BLOCKS, nout=32):
        super().__init__()
        mods: List[nn.Module] = [conv_bn_relu(3, nout, 3)]
        for i, nb in enumerate(blkList):
            ch = nout * (2 ** i)
            mods.append(conv_bn_relu(ch, ch * 2, 3, s=2))
            mods += [res_layer(ch * 2) for _ in range(nb)]
        self.mlist = nn.ModuleList(mods)
        self._owner = None
        self._shape = (list(blkList), nout)

    def __getstate__(self):                     # the back-references are not picklable; rebound by YoloNet.__setstate__ / lazily
        st = self.__dict__.copy()
        st["_owner"] = None
        st.pop("_host", None)
        return st

    def _engine_owner(self):
        """The YoloNet whose engine runs this backbone: the net this module is the `feature` of, or -- for a stand-alone
        Darknet(blkList), as darknet.py:72-88 allows -- a hidden host net created on first use that adopts THIS module as
        its `feature` (so the engine reads this module's own parameters; the host's head weights are never executed)."""
        owner = self._owner() if self._owner is not None else None
        if owner is not None:
            return owner
        if self._shape != (list(BLOCKS), 32):
            raise NotImplementedError("the engine runs the Darknet-53 layout only: blkList=[1,2,8,8,4], nout=32")
        import weakref
        host = YoloNet(None)
        host.feature = self
        object.__setattr__(self, "_host", host)             # strong reference, NOT a registered sub-module (no module cycle)
        self._owner = weakref.ref(host)
        return host

    def loadWeight(self, weights_path):
        """Backbone-only darknet stream, e.g. darknet53.conv.74 (darknet.py:102-104)."""
        self._engine_owner()._load_darknet_file(weights_path, backbone_only=True)

    def forward(self, x):
        """darknet.py:83-88: [B,1024,H/32,W/32] fp32 (the cached route outputs stay inside the engine)."""
        if self.training:
            raise RuntimeError("inference path only: call .eval() first (BatchNorm uses running statistics)")
        return self._engine_owner().backbone(x)


class PreDetectionConvGroup(nn.Module):
    """Container for darknet.py:107-118: 3 x (1x1, 3x3) blocks and the plain 1x1 head conv."""

    def __init__(self, nin, nout, num_conv=3, numClass=80):
        super().__init__()
        mods: List[nn.Module] = []
        for i in range(num_conv):
            mods.append(conv_bn_relu(nin, nout, 1))
            mods.append(conv_bn_relu(nout, nout * 2, 3))
            nin = nout * 2
        mods.append(_ConvParams(nin, (numClass + 5) * 3, 1, bias=True))
        self.mlist = nn.ModuleList(mods)


class UpsampleGroup(nn.Module):
    """Container for darknet.py:153-157."""

    def __init__(self, nin):
        super().__init__()
        self.conv = conv_bn_relu(nin, nin // 2, 1)


class YoloNet(nn.Module):
    """Drop-in for the reference's YoloNet (darknet.py:167-246), inference path only."""

    _MODES = {"fp16": _lib.YB_MODE_FP16, "fp32": _lib.YB_MODE_FP32_TC, "fp32_simt": _lib.YB_MODE_FP32}

    def __init__(self, img_dim=None, anchors: Sequence[float] = tuple(DEFAULT_ANCHORS), numClass=80,
                 precision: Optional[str] = None, check_weights: bool = False):
        super().__init__()
        import weakref
        self.numClass = numClass
        self.img_dim = img_dim
        self.stat_keys = ['loss', 'loss_x', 'loss_y', 'loss_w', 'loss_h', 'loss_conf', 'loss_cls',
                          'nCorrect', 'nGT', 'recall']
        self.anchors = [float(a) for a in anchors]
        if len(self.anchors) != 18:
            raise ValueError("anchors must hold 9 (w,h) pairs")
        pairs = [(self.anchors[i], self.anchors[i + 1]) for i in range(0, 18, 2)]
        self.precision = (precision or os.environ.get("YOLO_B200_PRECISION", "fp16")).lower()
        if self.precision not in self._MODES:
            raise ValueError("precision must be 'fp16', 'fp32' or 'fp32_simt'")
        self.check_weights = bool(check_weights)
        self._frozen = False
        self._graph_mode = 2                                # yb_set_graph_mode: auto

        self.feature = Darknet(BLOCKS)
        self.feature._owner = weakref.ref(self)
        self.pre_det1 = PreDetectionConvGroup(1024, 512, numClass=numClass)
        self.yolo1 = YoloLayer(pairs, [6, 7, 8], img_dim, numClass)
        self.up1 = UpsampleGroup(512)
        self.pre_det2 = PreDetectionConvGroup(768, 256, numClass=numClass)
        self.yolo2 = YoloLayer(pairs, [3, 4, 5], img_dim, numClass)
        self.up2 = UpsampleGroup(256)
        self.pre_det3 = PreDetectionConvGroup(384, 128, numClass=numClass)
        self.yolo3 = YoloLayer(pairs, [0, 1, 2], img_dim, numClass)

        self._ctx = None
        self._ctx_device = None
        self._sig = None
        self._in_dtype = _lib.YB_INPUT_F32
        self.header = torch.tensor([0, 2, 0, 0, 0], dtype=torch.int32)
        self.seen = self.header[3]

    # ---- engine plumbing ----------------------------------------------------------------------
    def _named_tensors(self):
        """(state_dict key, current tensor) in state_dict order.  Runs before every forward (the signature below), so it
        does not build a state_dict each time: the (module, slot) of every key is resolved once -- the module tree of
        this class never changes -- and the tensors are read from the modules' own tables, which is where .cuda(),
        load_state_dict() or an assignment put their current versions."""
        slots = self.__dict__.get("_tensor_slots")
        if slots is None:
            where = {}
            for prefix, mod in self.named_modules():
                dot = prefix + "." if prefix else ""
                for name in mod._parameters:
                    where[dot + name] = (mod._parameters, name)
                for name in mod._buffers:
                    where[dot + name] = (mod._buffers, name)
            slots = [(k, *where[k]) for k in self.state_dict(keep_vars=True)]
            self.__dict__["_tensor_slots"] = slots
        for k, table, name in slots:
            yield k, table[name]

    def _signature(self, device):
        sig = (str(device), self.precision) + tuple((v.data_ptr(), v._version) for _, v in self._named_tensors())
        if self.check_weights:
            # content check for writes that bypass the version counter (param.data.copy_()): one multi-tensor reduction
            # and one device-to-host read per forward
            ts = [v.detach().float() for k, v in self._named_tensors() if not k.endswith("num_batches_tracked")]
            sums = torch.stack(torch._foreach_norm(ts)).double()
            w = torch.arange(1, len(ts) + 1, dtype=torch.float64, device=sums.device)
            sig += (float((sums * w).sum()), float(torch.stack([t.reshape(-1)[0] for t in ts]).double().sum()))
        return sig

    def refresh_weights(self):
        """Forget the uploaded weights: the next forward re-reads every tensor, folds BN and re-packs.  Needed after an
        in-place write through ``param.data`` / ``buffer.data`` (e.g. ``p.data.copy_(w)``), which PyTorch does not
        version -- every other way of changing the weights is detected automatically."""
        self._sig = None
        self._frozen = False
        return self

    def set_graph_mode(self, mode="auto"):
        """CUDA-graph replay of the convolution launches after the stem (yb_set_graph_mode): 'never', 'always' or 'auto'
        (default: only for launch-bound shapes such as a single 416x416 image)."""
        self._graph_mode = {"never": 0, "always": 1, "auto": 2}[mode]
        if self._ctx is not None:
            _lib.check(_lib.load().yb_set_graph_mode(self._ctx, self._graph_mode), self._ctx)
        return self

    def freeze_weights(self, frozen: bool = True):
        """Skip the per-forward change detection (a walk over the 438 tensors, ~0.1 ms of host time -- visible at batch 1)
        until refresh_weights() / load_state_dict() / loadWeight() is called."""
        self._frozen = bool(frozen)
        return self

    def load_state_dict(self, *args, **kwargs):
        self._sig = None
        self._frozen = False
        return super().load_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):            # .cuda() / .to() / .half() / .float()
        self._sig = None
        self._frozen = False
        return super()._apply(fn, *args, **kwargs)

    def _engine(self, device: torch.device):
        """Create the context on `device` if needed and (re)upload + finalize when any tensor changed."""
        if device.type != "cuda":
            raise RuntimeError("yolo_v3_b200 runs on CUDA devices only (no CPU fallback); move the input and "
                               "the module to a CUDA device first")
        lib = _lib.load()
        index = device.index if device.index is not None else torch.cuda.current_device()
        if self._ctx is None or self._ctx_device != index:
            self._release()
            self._ctx = _lib.create_ctx(index, self.numClass, self.anchors)
            self._ctx_device = index
            _lib.check(lib.yb_set_graph_mode(self._ctx, self._graph_mode), self._ctx)
            self._sig = None
            self._in_dtype = _lib.YB_INPUT_F32          # a new context reads fp32 images
        if self._frozen and self._sig is not None:
            return lib, self._ctx
        sig = self._signature(device)
        if sig != self._sig:
            for k, v in self._named_tensors():
                if k.endswith("num_batches_tracked"):
                    continue
                t = v.detach()
                if t.dtype != torch.float32:
                    t = t.float()
                t = t.contiguous()
                on_host = 0 if t.is_cuda else 1
                _lib.check(lib.yb_set_tensor(self._ctx, k.encode(), ctypes.c_void_p(t.data_ptr()), t.numel(), on_host),
                           self._ctx)
            _lib.check(lib.yb_finalize(self._ctx, self._MODES[self.precision]), self._ctx)
            self._sig = sig
        return lib, self._ctx

    # copy.deepcopy / pickle / torch.save(net): the engine context is a raw pointer and belongs to THIS object; a copy
    # starts without one (it is created on its first forward) and its backbone points back at the copy
    def __getstate__(self):
        st = self.__dict__.copy()
        st["_ctx"] = None
        st["_ctx_device"] = None
        st["_sig"] = None
        st["_in_dtype"] = _lib.YB_INPUT_F32
        st.pop("_tensor_slots", None)
        st.pop("_last_det", None)
        return st

    def __setstate__(self, st):
        import weakref
        self.__dict__.update(st)
        self.feature._owner = weakref.ref(self)

    def _release(self):
        if self._ctx is not None:
            _lib.load().yb_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _check_input(self, x):
        if not isinstance(x, torch.Tensor) or x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("expected a [B,3,H,W] tensor")
        if not x.is_cuda:
            raise RuntimeError("yolo_v3_b200 runs on CUDA devices only (no CPU fallback): got a CPU tensor")
        if x.shape[2] % 32 or x.shape[3] % 32:
            raise ValueError("H and W must be multiples of 32")
        # fp16 images are read by the stem directly (yb_set_input_dtype): same bits as the fp32 path, which rounds every
        # pixel to fp16 itself, for half the host-to-device bytes.  Any other dtype is widened to fp32, as the reference's
        # callers pass it.
        direct_f16 = x.dtype == torch.float16 and self.precision == "fp16"
        if not direct_f16 and x.dtype != torch.float32:
            x = x.float()
        return x.contiguous()

    def _prepare(self, x):
        """Validated input + (lib, ctx) with the context's input element type matching x."""
        x = self._check_input(x)
        lib, ctx = self._engine(x.device)
        want = _lib.YB_INPUT_F16 if x.dtype == torch.float16 else _lib.YB_INPUT_F32
        if self._in_dtype != want:
            _lib.check(lib.yb_set_input_dtype(ctx, want), ctx)
            self._in_dtype = want
        return x, lib, ctx

    def _stream(self, device):
        return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)

    # ---- reference API ------------------------------------------------------------------------
    def forward(self, x, target=None):
        """darknet.py:198-231 with target=None: returns (det1, det2, det3), fp32 CUDA tensors of shape
        [B,3*(H/32)*(W/32),5+C], [B,3*(H/16)*(W/16),5+C], [B,3*(H/8)*(W/8),5+C]."""
        if target is not None:
            raise NotImplementedError("training (target is not None) is outside the scope of the B200 inference path")
        if self.training:
            raise RuntimeError("inference path only: call .eval() first (BatchNorm uses running statistics)")
        x, lib, ctx = self._prepare(x)
        B, _, H, W = x.shape
        n = num_boxes(H, W)
        det = torch.empty(B, n, 5 + self.numClass, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(lib.yb_forward(ctx, ctypes.c_void_p(x.data_ptr()), B, H, W, ctypes.c_void_p(det.data_ptr()),
                                      self._stream(x.device)), ctx)
        n1 = 3 * (H // 32) * (W // 32)
        n2 = 3 * (H // 16) * (W // 16)
        self._last_det = det
        return det[:, :n1], det[:, n1:n1 + n2], det[:, n1 + n2:]

    def head_logits(self, x):
        """The three raw head maps (pre_detN.mlist[6] outputs), NCHW fp32 -- parity/debug aid."""
        x, lib, ctx = self._prepare(x)
        B, _, H, W = x.shape
        ch = 3 * (5 + self.numClass)
        outs = [torch.empty(B, ch, H // s, W // s, device=x.device, dtype=torch.float32) for s in (32, 16, 8)]
        with torch.cuda.device(x.device):
            _lib.check(lib.yb_forward_logits(ctx, ctypes.c_void_p(x.data_ptr()), B, H, W,
                                             *[ctypes.c_void_p(o.data_ptr()) for o in outs], self._stream(x.device)), ctx)
        return tuple(outs)

    def backbone(self, x):
        """Darknet.forward (darknet.py:83-88): [B,1024,H/32,W/32] fp32."""
        x, lib, ctx = self._prepare(x)
        B, _, H, W = x.shape
        out = torch.empty(B, 1024, H // 32, W // 32, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(lib.yb_backbone(ctx, ctypes.c_void_p(x.data_ptr()), B, H, W, ctypes.c_void_p(out.data_ptr()),
                                       self._stream(x.device)), ctx)
        return out

    def detect_raw(self, x, obj_conf_thr=0.5, nms_thr=0.4, is_eval=False, use_nms=True, cap=None):
        """forward + postprocessing fused on the device (yb_detect). Returns CUDA tensors
        (rows7 [B,cap,7], counts [B], src_index [B,cap], cand_counts [B]); nothing is synchronised."""
        x, lib, ctx = self._prepare(x)
        B, _, H, W = x.shape
        n = num_boxes(H, W)
        if cap is None:
            cap = n * (self.numClass if is_eval else 1) if not use_nms else min(n, 2048)
        rows = torch.empty(B, cap, 7, device=x.device, dtype=torch.float32)
        counts = torch.empty(B, device=x.device, dtype=torch.int32)
        src = torch.empty(B, cap, device=x.device, dtype=torch.int32)
        cand = torch.empty(B, device=x.device, dtype=torch.int32)
        with torch.cuda.device(x.device):
            _lib.check(lib.yb_detect(ctx, ctypes.c_void_p(x.data_ptr()), B, H, W, float(obj_conf_thr), float(nms_thr),
                                     int(bool(is_eval)), int(bool(use_nms)), ctypes.c_void_p(rows.data_ptr()),
                                     ctypes.c_void_p(counts.data_ptr()), ctypes.c_void_p(src.data_ptr()),
                                     ctypes.c_void_p(cand.data_ptr()), int(cap), self._stream(x.device)), ctx)
        return rows, counts, src, cand

    def detect(self, x, obj_conf_thr=0.5, nms_thr=0.4, is_eval=False, use_nms=True):
        """What test.py:35-36 computes per batch -- postprocessing(torch.cat(net(x)), ...) -- in one
        device-side pass.  Same return convention as utils.postprocessing."""
        from .utils import rows_to_list
        cap = None
        while True:
            rows, counts, _, cand = self.detect_raw(x, obj_conf_thr, nms_thr, is_eval, use_nms, cap)
            counts_h = counts.cpu()
            mx = int(counts_h.max())
            if mx <= rows.shape[1]:
                return rows_to_list(rows, counts_h, cand.cpu())
            cap = mx

    # Format : pytorch / darknet  (darknet.py:234-246)
    def saveWeight(self, weights_path, format='pytorch'):
        if format == 'pytorch':
            torch.save(self.state_dict(), weights_path)
        elif format == 'darknet':
            sd = self.state_dict()
            with open(weights_path, "wb") as fp:
                self.header.numpy().astype(np.int32).tofile(fp)
                for k, _ in darknet_stream_keys(self.numClass):
                    sd[k].detach().cpu().numpy().astype(np.float32).ravel().tofile(fp)
        else:
            raise ValueError(format)

    def loadWeight(self, weights_path, format='pytorch'):
        if format == 'pytorch':
            weights = torch.load(weights_path, map_location=lambda storage, loc: storage)
            self.load_state_dict(weights)
        elif format == 'darknet':
            self._load_darknet_file(weights_path, backbone_only=False)
        else:
            raise ValueError(format)

    def _load_darknet_file(self, path, backbone_only):
        """WeightManager.loadWeight (darknet.py:254-290): 5 x int32 header, then fp32 stream."""
        with open(path, "rb") as fp:
            header = np.fromfile(fp, dtype=np.int32, count=5)
            weights = np.fromfile(fp, dtype=np.float32)
        self.header = torch.from_numpy(header.copy())
        self.seen = self.header[3]
        return self.load_darknet_stream(weights, backbone_only)

    def load_darknet_stream(self, weights: np.ndarray, backbone_only=False) -> int:
        sd = self.state_dict(keep_vars=True)
        ptr = 0
        with torch.no_grad():
            for k, shape in darknet_stream_keys(self.numClass, backbone_only):
                n = int(np.prod(shape))
                if ptr + n > len(weights):
                    raise ValueError(f"darknet weight stream ends inside {k}")
                sd[k].copy_(torch.from_numpy(weights[ptr:ptr + n].copy()).view(*shape))
                ptr += n
        self.refresh_weights()
        return ptr


class WeightManager:
    """API-compatible stand-in for darknet.py:249-303 (``WeightManager(model).loadWeight(path)``)."""

    def __init__(self, model):
        self.model = model

    def loadWeight(self, weight_path):
        m = self.model
        if isinstance(m, Darknet):
            owner = m._owner() if m._owner is not None else None
            if owner is None:
                raise RuntimeError("stand-alone Darknet has no engine")
            ptr = owner._load_darknet_file(weight_path, backbone_only=True)
            self.header, self.seen = owner.header, owner.seen
            return ptr
        ptr = m._load_darknet_file(weight_path, backbone_only=False)
        self.header, self.seen = m.header, m.seen
        return ptr
