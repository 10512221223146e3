"""CPU emulation of candidate tensor-core "fp32-grade" operand splits for the conv stack (round-2 design study).

Question: which operand split lets tcgen05 kind::f16 MMAs (fp32 accumulate in TMEM) reproduce the reference's fp32
conv stack (darknet.py:27-53, :118) within the north-star tolerance (head logits within 1e-4 * max|logit|)?

Schemes (every product term is evaluated exactly in fp64, summed, and rounded ONCE to fp32 -- i.e. this measures the
operand representation error only; the accumulation error of the real tensor core comes on top and is measured on the GPU):
  f16     : x -> fp16(x), w -> fp16(w)                                (1 MMA per k-step; the benched fast path)
  f16x2   : x = xh + xl, w = wh + wl (both fp16), xh*wh + xh*wl + xl*wh  (3 MMAs)
  f16x2s  : as f16x2, weights pre-scaled per output channel by a power of two so that wl stays a normal fp16
  bf16x3  : three-term bf16 split, 6 MMAs
and the fp32 oracle itself (torch fp32 conv), all against an fp64 evaluation of the same network.

Usage: python tools/split_numerics.py [hw=416] [batch=1]
"""
import sys, os, time
import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_v3_b200.synth import make_state_dict, make_images
from yolo_v3_b200.topology import BLOCKS

BN_EPS = 1e-5


def split_f16(t64, n):
    """t64 (fp64 view of fp32 data) -> n fp16 terms as fp64 tensors."""
    out, r = [], t64.clone()
    for _ in range(n):
        h = r.to(torch.float32).to(torch.float16).to(torch.float64)
        out.append(h)
        r = r - h
    return out


def split_bf16(t64, n):
    out, r = [], t64.clone()
    for _ in range(n):
        h = r.to(torch.float32).to(torch.bfloat16).to(torch.float64)
        out.append(h)
        r = r - h
    return out


class Net:
    def __init__(self, sd, scheme):
        self.sd, self.scheme = sd, scheme

    def conv_raw(self, x, w, ks, stride):
        """x fp32 [B,C,H,W], w fp32 -> raw conv result in fp32 (or fp64 for scheme 'f64')."""
        pad = (ks - 1) // 2
        s = self.scheme
        if s == "f32":
            return F.conv2d(x, w, None, stride, pad)
        x64, w64 = x.double(), w.double()
        if s == "f64":
            return F.conv2d(x64, w64, None, stride, pad)
        if s == "f16":
            (xh,), (wh,) = split_f16(x64, 1), split_f16(w64, 1)
            return F.conv2d(xh, wh, None, stride, pad).float()
        if s in ("f16x2", "f16x2s"):
            if s == "f16x2s":
                m = w64.abs().amax(dim=(1, 2, 3), keepdim=True).clamp_min(1e-30)
                sc = torch.exp2(torch.floor(torch.log2(1024.0 / m)))       # max|w'| in [512, 1024]
            else:
                sc = torch.ones_like(w64[:, :1, :1, :1])
            xh, xl = split_f16(x64, 2)
            wh, wl = split_f16(w64 * sc, 2)
            y = F.conv2d(xh, wh, None, stride, pad) + F.conv2d(xh, wl, None, stride, pad) + F.conv2d(xl, wh, None, stride, pad)
            return (y / sc.view(1, -1, 1, 1)).float()
        if s == "bf16x3":
            xs, ws = split_bf16(x64, 3), split_bf16(w64, 3)
            y = 0
            for i in range(3):
                for j in range(3):
                    if i + j <= 2:
                        y = y + F.conv2d(xs[i], ws[j], None, stride, pad)
            return y.float()
        raise ValueError(s)

    def cbr(self, key, x, ks, stride=1):
        sd = self.sd
        y = self.conv_raw(x, sd[key + ".conv.weight"], ks, stride)
        dt = y.dtype
        # eval BN folded as the engine folds it (fp32: alpha = g * invstd, beta = b - mean*alpha)
        var, g, b, mean = (sd[key + ".bn.running_var"].to(dt), sd[key + ".bn.weight"].to(dt), sd[key + ".bn.bias"].to(dt),
                           sd[key + ".bn.running_mean"].to(dt))
        alpha = g / torch.sqrt(var + BN_EPS)
        beta = b - mean * alpha
        y = y * alpha.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)
        y = F.leaky_relu(y, 0.1)
        return y if self.scheme == "f64" else y.float()

    def head(self, key, x):
        y = self.conv_raw(x, self.sd[key + ".weight"], 1, 1)
        return y + self.sd[key + ".bias"].to(y.dtype).view(1, -1, 1, 1)

    def forward(self, x):
        if self.scheme == "f64":
            x = x.double()
        x = self.cbr("feature.mlist.0", x, 3)
        idx, routes = 1, {}
        for nb in BLOCKS:
            x = self.cbr(f"feature.mlist.{idx}", x, 3, 2)
            idx += 1
            for _ in range(nb):
                x = x + self.cbr(f"feature.mlist.{idx}.conv2", self.cbr(f"feature.mlist.{idx}.conv1", x, 1), 3)
                routes[idx] = x
                idx += 1

        def predet(name, x):
            r = None
            for i in range(6):
                x = self.cbr(f"{name}.mlist.{i}", x, 1 if i % 2 == 0 else 3)
                if i == 4:
                    r = x
            return self.head(f"{name}.mlist.6", x), r

        l1, h1 = predet("pre_det1", x)
        x = torch.cat((F.interpolate(self.cbr("up1.conv", h1, 1), scale_factor=2, mode="nearest"), routes[23]), 1)
        l2, h2 = predet("pre_det2", x)
        x = torch.cat((F.interpolate(self.cbr("up2.conv", h2, 1), scale_factor=2, mode="nearest"), routes[14]), 1)
        l3, _ = predet("pre_det3", x)
        return [l1, l2, l3]


def main():
    hw = int(sys.argv[1]) if len(sys.argv) > 1 else 416
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    schemes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["f64", "f32", "f16x2", "f16x2s", "bf16x3", "f16"]
    sd = make_state_dict(seed=1234)
    x = make_images(B, hw, hw, seed=3)
    res = {}
    with torch.no_grad():
        for s in schemes:
            t0 = time.time()
            res[s] = [l.double() for l in Net(sd, s).forward(x)]
            print(f"{s:8s} done in {time.time() - t0:6.1f} s", flush=True)
    truth = res["f64"]
    mx = max(float(l.abs().max()) for l in truth)
    print(f"hw={hw} B={B} max|logit|={mx:.3f}")
    for s in schemes:
        if s == "f64":
            continue
        d64 = max(float((a - b).abs().max()) for a, b in zip(res[s], truth))
        rms = float(torch.sqrt(sum(((a - b) ** 2).sum() for a, b in zip(res[s], truth)) / sum(a.numel() for a in truth)))
        line = f"{s:8s} vs f64: max|d|={d64:.3e} ({d64 / mx:.2e} of max|logit|)  rms={rms:.3e}"
        if "f32" in res and s != "f32":
            d32 = max(float((a - b).abs().max()) for a, b in zip(res[s], res["f32"]))
            line += f" | vs f32 oracle: max|d|={d32:.3e} ({d32 / mx:.2e})"
        print(line)


if __name__ == "__main__":
    main()
