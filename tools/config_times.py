"""Device times of the other BASELINE.json configurations (they are parity-test cases, not bench lines; this prints
their throughput for profiles/): cfg 2 Darknet-53 backbone 256x256 batch 64, cfg 4 post-process stress 608x608
batch 64 conf 0.001 (~10k candidates per image), cfg 5 per-GPU shards of the batch-256 job (256/128/64/32 images)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from yolo_v3_b200 import YoloNet, synth, topology  # noqa: E402
from yolo_v3_b200.utils import postprocessing_raw  # noqa: E402


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


sd = synth.make_state_dict(seed=1234, recipe="calibrated")

# cfg 2: backbone only, 256x256, batch 64
net = YoloNet((256, 256), precision="fp16")
net.load_state_dict(sd)
net = net.cuda().eval()
xs = [synth.make_images(64, 256, 256, seed=i).cuda() for i in range(4)]       # 4 x 50 MB, alternated
k = [0]


def bb():
    k[0] += 1
    return net.backbone(xs[k[0] & 3])


ms = timed(bb)
flops = 18.568e9 * 64
print(f"cfg2 backbone 256x256 b64: {ms:.3f} ms  {64 / ms * 1e3:.0f} img/s  {flops / ms / 1e9:.0f} TFLOP/s "
      f"(includes the NHWC->NCHW fp32 export of the [64,1024,8,8] feature map)")
del net, xs

# cfg 4: post-process stress, 608x608, batch 64, conf 0.001
from oracle import yolo_oracle as O  # noqa: E402  (only to decode the synthetic logits once, outside the timed region)
logits = synth.make_head_logits(8, 608, 608, 80, seed=7, obj_mu=-6.5)
anchors = [(10, 13), (16, 30), (33, 23), (30, 61), (62, 45), (59, 119), (116, 90), (156, 198), (373, 326)]
masks = ([6, 7, 8], [3, 4, 5], [0, 1, 2])
det8 = torch.cat([O.decode(l, anchors, masks[i], (608, 608), 80) for i, l in enumerate(logits)], 1)
det = det8.repeat(8, 1, 1).cuda()                                                # 64 images (8 distinct)
rows, counts, src, cand = postprocessing_raw(det, 80, 0.001, 0.4, False, True, 16384)
torch.cuda.synchronize()
ms = timed(lambda: postprocessing_raw(det, 80, 0.001, 0.4, False, True, 16384), reps=5)
nb = det.numel() * 4
print(f"cfg4 post-process stress 608x608 b64 conf 0.001: {ms:.3f} ms  {64 / ms * 1e3:.0f} img/s  "
      f"candidates/img {float(cand.float().mean()):.0f}  survivors/img {float(counts.float().mean()):.0f}  "
      f"{ms / 64 * 1e3:.1f} us/img  first-pass {nb / ms / 1e6:.0f} GB/s over the det tensor")
del det, det8

# cfg 5: shards of the batch-256 job
net = YoloNet((608, 608), precision="fp16")
net.load_state_dict(sd)
net = net.cuda().eval()
for bsz in (32, 64, 128, 256):
    xs = [synth.make_images(bsz, 608, 608, seed=10 + i).cuda() for i in range(2)]
    k = [0]

    def step():
        k[0] += 1
        return net.detect_raw(xs[k[0] & 1], 0.5, 0.4, False, True, 512)

    ms = timed(step, reps=5, warm=2)
    print(f"cfg5 shard of {bsz} images 608x608 detect: {ms:.3f} ms  {bsz / ms * 1e3:.0f} img/s  "
          f"{topology.conv_flops(608, 608) * bsz / ms / 1e9:.0f} TFLOP/s incl. decode+NMS")
    del xs
