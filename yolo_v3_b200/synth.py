"""Seeded synthetic weights and inputs for tests, smoke() and bench.py.

No weight files ship with the reference (weights/ is git-ignored) and there is no network, so
every measurement uses random-init weights of the real architecture.  Plain He init does not
work for Darknet-53: the 23 residual adds compound the variance until fp16 overflows
(SURVEY.md 8d).  Two recipes are provided, both driven by numpy's RandomState so they do not
depend on the torch version:

* ``analytic``  -- closed-form: He-scaled conv weights, identity-like BN statistics and a damped
  gamma on the second conv of every residual block.  Bit-reproducible on any host; used for the
  committed golden fixtures.
* ``calibrated`` -- the analytic recipe followed by one data-dependent pass that sets every BN's
  running statistics to the batch statistics of a seeded random image (what one training step
  with momentum=1 would record).  Activations stay at std 0.6-1.9 through all 75 layers and the
  head logits exercise the default thresholds.  Used by bench.py and the large GPU tests, where
  the oracle is evaluated live on the same state_dict.

The state_dict keys are the reference's (darknet.py:76-79,112-118,156): 438 entries.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

from .topology import BLOCKS, layer_specs


def make_state_dict(seed: int = 1234, num_classes: int = 80, recipe: str = "calibrated",
                    calib_hw: int = 128) -> Dict[str, torch.Tensor]:
    rs = np.random.RandomState(seed)
    sd: Dict[str, torch.Tensor] = {}
    attrs = num_classes + 5

    def t(a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))

    for e in layer_specs(num_classes):
        k, ci, co, ks = e["key"], e["cin"], e["cout"], e["ks"]
        fan_in = ci * ks * ks
        if e["bn"]:
            sd[k + ".conv.weight"] = t(rs.standard_normal((co, ci, ks, ks)) * np.sqrt(2.0 / (1.01 * fan_in)))
            lo, hi = (0.2, 0.3) if e["res2"] else (0.8, 1.2)
            sd[k + ".bn.weight"] = t(rs.uniform(lo, hi, co))
            sd[k + ".bn.bias"] = t(rs.standard_normal(co) * 0.1)
            sd[k + ".bn.running_mean"] = t(rs.standard_normal(co) * 0.05)
            sd[k + ".bn.running_var"] = t(rs.uniform(0.9, 1.1, co))
            sd[k + ".bn.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
        else:
            sd[k + ".weight"] = t(rs.standard_normal((co, ci, 1, 1)) * (1.5 / np.sqrt(fan_in)))
            b = np.zeros((3, attrs), np.float32)
            b[:, 2:4] = -0.5
            b[:, 4] = -2.0
            b[:, 5:] = -3.0
            sd[k + ".bias"] = t(b.reshape(-1))
    if recipe == "calibrated":
        _calibrate(sd, num_classes, calib_hw, seed + 1)
    elif recipe != "analytic":
        raise ValueError(recipe)
    return sd


def _calibrate(sd, num_classes, hw, seed):
    """One train-mode-BN pass (momentum 1): running_mean/var <- batch statistics.  Host-side weight
    synthesis only; never part of an inference path."""
    import torch.nn.functional as F
    x = torch.from_numpy(np.random.RandomState(seed).rand(2, 3, hw, hw).astype(np.float32))

    def cbr(key, x, ks, s=1):
        y = F.conv2d(x, sd[key + ".conv.weight"], None, s, (ks - 1) // 2)
        mean = y.mean((0, 2, 3))
        sd[key + ".bn.running_mean"] = mean.clone()
        sd[key + ".bn.running_var"] = y.var((0, 2, 3), unbiased=True).clone()
        var_b = y.var((0, 2, 3), unbiased=False)
        y = (y - mean.view(1, -1, 1, 1)) / torch.sqrt(var_b.view(1, -1, 1, 1) + 1e-5)
        y = y * sd[key + ".bn.weight"].view(1, -1, 1, 1) + sd[key + ".bn.bias"].view(1, -1, 1, 1)
        return F.leaky_relu(y, 0.1)

    with torch.no_grad():
        x = cbr("feature.mlist.0", x, 3)
        idx, routes = 1, {}
        for nb in BLOCKS:
            x = cbr(f"feature.mlist.{idx}", x, 3, 2)
            idx += 1
            for _ in range(nb):
                x = x + cbr(f"feature.mlist.{idx}.conv2", cbr(f"feature.mlist.{idx}.conv1", x, 1), 3)
                routes[idx] = x
                idx += 1

        def predet(name, x):
            r = None
            for i in range(6):
                x = cbr(f"{name}.mlist.{i}", x, 1 if i % 2 == 0 else 3)
                if i == 4:
                    r = x
            return r

        h1 = predet("pre_det1", x)
        x = torch.cat((F.interpolate(cbr("up1.conv", h1, 1), scale_factor=2, mode="nearest"), routes[23]), 1)
        h2 = predet("pre_det2", x)
        x = torch.cat((F.interpolate(cbr("up2.conv", h2, 1), scale_factor=2, mode="nearest"), routes[14]), 1)
        predet("pre_det3", x)


def make_images(batch: int, h: int, w: int, seed: int = 0) -> torch.Tensor:
    """[B,3,H,W] fp32 in [0,1) (what test.py feeds the net after ToTensor, transforms.py:25)."""
    return torch.from_numpy(np.random.RandomState(seed).rand(batch, 3, h, w).astype(np.float32))


def make_photo(h: int, w: int, seed: int = 0) -> np.ndarray:
    """Seeded uint8 RGB image [h,w,3] with photo-like content (smooth structure, edges, mild noise): the
    input of the letterbox pre-process (what cv2.imread + cvtColor hand to utils.letterbox_image, utils.py:61-67)."""
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    ph = rs.uniform(0, 6.28, (3, 3))
    img = np.empty((h, w, 3), np.float64)
    for c in range(3):
        img[..., c] = (128 + 55 * np.sin(xx / 17.0 + ph[c, 0]) * np.cos(yy / 23.0 + ph[c, 1])
                       + 35 * np.sin((xx + 2 * yy) / 7.0 + ph[c, 2]) + 40 * (((xx // 37) + (yy // 29)) % 2))
    img += rs.randint(-3, 4, (h, w, 3))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def make_head_logits(batch: int, h: int, w: int, num_classes: int = 80, seed: int = 7,
                     obj_mu: float = -6.5, dense: bool = False):
    """Seeded raw head maps [B,255,h/32..h/8,...] for the post-process stress configs
    (cfg 4): t_xy~N(0,1), t_wh~N(0,0.5), t_obj~N(mu,3), t_cls~N(-4,1.5);
    `dense` piles the class mass on 4 classes so suppression is heavy."""
    rs = np.random.RandomState(seed)
    attrs = num_classes + 5
    out = []
    for s in (32, 16, 8):
        gh, gw = h // s, w // s
        a = np.empty((batch, 3, attrs, gh, gw), np.float32)
        a[:, :, 0:2] = rs.standard_normal((batch, 3, 2, gh, gw))
        a[:, :, 2:4] = rs.standard_normal((batch, 3, 2, gh, gw)) * 0.5
        a[:, :, 4] = rs.standard_normal((batch, 3, gh, gw)) * 3.0 + obj_mu
        a[:, :, 5:] = rs.standard_normal((batch, 3, num_classes, gh, gw)) * 1.5 - 4.0
        if dense:
            a[:, :, 9:] -= 6.0
        out.append(torch.from_numpy(a.reshape(batch, 3 * attrs, gh, gw)))
    return out
