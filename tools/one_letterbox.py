"""Minimal driver for profilers: letterbox 32 synthetic 480x640 uint8 photos to 608x608 (yb_letterbox), a few times."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from yolo_v3_b200 import synth  # noqa: E402
from yolo_v3_b200.utils import letterbox_batch  # noqa: E402

photos = [torch.from_numpy(synth.make_photo(480, 640, 50 + i)).cuda() for i in range(32)]
for _ in range(4):
    x, t = letterbox_batch(photos, (608, 608))
torch.cuda.synchronize()
print("done", tuple(x.shape))
