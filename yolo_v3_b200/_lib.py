"""ctypes binding of libyolo_b200.so (C ABI: include/yolo_b200.h).

This is the only bridge between the Python host mirror and the CUDA path.  There is no fallback:
if the library has not been built, importing this module raises, and if no CUDA device is present
yb_create fails with a RuntimeError.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_longlong, c_size_t, c_uint8, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libyolo_b200.so")

YB_MODE_FP32 = 0
YB_MODE_FP16 = 1
YB_MODE_FP32_TC = 2
YB_INPUT_F32 = 0
YB_INPUT_F16 = 1
YB_E_CAP = -6

# name -> (restype, argtypes); every entry point declared in include/yolo_b200.h
PROTOTYPES = {
    "yb_create": (c_int, [POINTER(c_void_p), c_int, c_int, POINTER(c_float)]),
    "yb_destroy": (None, [c_void_p]),
    "yb_last_error": (c_char_p, [c_void_p]),
    "yb_num_tensors": (c_int, [c_void_p]),
    "yb_tensor_key": (c_char_p, [c_void_p, c_int, POINTER(c_size_t)]),
    "yb_set_tensor": (c_int, [c_void_p, c_char_p, c_void_p, c_size_t, c_int]),
    "yb_get_tensor": (c_int, [c_void_p, c_char_p, c_void_p, c_size_t]),
    "yb_load_darknet_blob": (c_int, [c_void_p, c_void_p, c_size_t, c_int, POINTER(c_size_t)]),
    "yb_save_darknet_blob": (c_int, [c_void_p, c_void_p, c_size_t, c_int, POINTER(c_size_t)]),
    "yb_finalize": (c_int, [c_void_p, c_int]),
    "yb_set_input_dtype": (c_int, [c_void_p, c_int]),
    "yb_set_graph_mode": (c_int, [c_void_p, c_int]),
    "yb_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "yb_forward_logits": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "yb_backbone": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "yb_decode": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "yb_postprocess": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_int, c_int,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "yb_postprocess_notebook": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_float,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "yb_detect": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_int, c_int,
                          c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "yb_correct_boxes": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, POINTER(c_int), c_int, c_int, c_int,
                                 c_void_p, c_void_p]),
    "yb_letterbox": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_int), c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                             c_void_p, POINTER(c_float), c_void_p]),
    "yb_resize": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_int), c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "yb_comm_unique_id": (c_int, [POINTER(c_uint8)]),
    "yb_comm_init": (c_int, [c_void_p, POINTER(c_uint8), c_int, c_int]),
    "yb_bcast_weights": (c_int, [c_void_p, c_int, c_void_p]),
    "yb_allgather_dets": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "yb_launch_count": (c_longlong, [c_void_p]),
    "yb_graph_replays": (c_longlong, [c_void_p]),
    "yb_debug_words": (c_int, [c_void_p, POINTER(c_int), c_int]),
    "yb_set_profiling": (c_int, [c_void_p, c_int]),
    "yb_get_section_ms": (c_int, [c_void_p, POINTER(c_float), POINTER(c_float), POINTER(c_float)]),
    "yb_get_layer_ms": (c_int, [c_void_p, POINTER(c_float), c_int]),
    "yb_run_layer": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "yb_run_stem_block": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
}

_lib = None


def load():
    """Load (once) and return the ctypes library with prototypes installed."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA library first "
            "(python -c 'import __graft_entry__ as g; g.build()'  or  make -C yolo_v3_b200/csrc). "
            "yolo_v3_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class YbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libyolo_b200 error {code}: {msg}")
        self.code = code


def check(rc: int, ctx=None):
    if rc != 0:
        msg = load().yb_last_error(ctx)
        raise YbError(rc, msg.decode() if msg else "")


def create_ctx(device: int, num_classes: int, anchors=None) -> c_void_p:
    lib = load()
    ctx = c_void_p()
    arr = None
    if anchors is not None:
        flat = [float(v) for v in anchors]
        if len(flat) != 18:
            raise ValueError("anchors must hold 9 (w,h) pairs")
        arr = (c_float * 18)(*flat)
    check(lib.yb_create(ctypes.byref(ctx), int(device), int(num_classes), arr), None)
    return ctx
