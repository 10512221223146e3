"""Loads the UNMODIFIED reference modules from oracle/_ref (copied there by `make -C oracle ref`)  --  TEST / BENCH
INFRASTRUCTURE ONLY, like everything under oracle/.

Used by bench.py's CPU arm (`--impl reference`, `cpu_baseline`), so that the number next to ours is the reference's own
code (darknet.YoloNet.forward + utils.postprocessing, test.py:35-36) on the box's host cores, not a restatement.

One harness shim, the same one tests/golden/make_golden.py uses: YoloLayer.forward hard-codes `.cuda()` on three small
tensors (yololayer.py:98-100); for the CPU path `torch.Tensor.cuda` is made the identity WHILE the reference runs
(cpu_only()), so the whole forward stays on the host cores.  The reference sources are not edited.
"""
import contextlib
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
FILES = ("darknet.py", "yololayer.py", "utils.py", "boundingbox.py")


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in FILES)


def load():
    """Returns (darknet, utils) of the reference.  The reference's modules import each other by bare name (`from utils
    import ...`), so its directory goes to the front of sys.path for the import and is removed again."""
    if not available():
        raise ImportError("oracle/_ref is missing: run `make -C oracle ref` where /root/reference exists")
    saved = {k: sys.modules.pop(k) for k in ("darknet", "yololayer", "utils", "boundingbox") if k in sys.modules}
    sys.path.insert(0, REF_DIR)
    try:
        darknet = importlib.import_module("darknet")
        utils = importlib.import_module("utils")
    finally:
        sys.path.remove(REF_DIR)
        for k in ("darknet", "yololayer", "utils", "boundingbox"):
            sys.modules.pop(k, None)
        sys.modules.update(saved)
    return darknet, utils


@contextlib.contextmanager
def cpu_only():
    """Tensor.cuda -> identity while the reference's CPU path runs (see the module docstring)."""
    import torch
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig
