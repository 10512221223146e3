"""Minimal driver for profilers: W warm-up + K timed detect steps at 608x608 batch 32 (fp16)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from yolo_v3_b200 import YoloNet, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--size", type=int, default=608)
ap.add_argument("--recipe", default="analytic")
ap.add_argument("--precision", default="fp16")
a = ap.parse_args()
sd = synth.make_state_dict(seed=1234, recipe=a.recipe)
net = YoloNet((a.size, a.size), precision=a.precision)
net.load_state_dict(sd)
net = net.cuda().eval()
x = synth.make_images(a.batch, a.size, a.size, seed=0).cuda()
for _ in range(a.warmup + a.steps):
    net.detect_raw(x, 0.5, 0.4, False, True, 512)
torch.cuda.synchronize()
print("done")
