"""BASELINE config 4 (NMS stress: 608x608, batch 64, conf 0.001, ~10 k candidates per image): postprocessing() on a resident
decoded tensor, for profilers.  python tools/cfg4_step.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_v3_b200 import synth
from yolo_v3_b200.utils import postprocessing_raw
from yolo_v3_b200.yololayer import decode_heads
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
logits = [l.cuda() for l in synth.make_head_logits(64, 608, 608, 80, seed=7)]
det = decode_heads(logits, (608, 608))
for _ in range(3 + reps):
    r = postprocessing_raw(det, 80, 0.001, 0.4, False, True, det.shape[1])
torch.cuda.synchronize()
print("candidates/image", float(r[3].float().mean()), "survivors/image", float(r[1].float().mean()))
