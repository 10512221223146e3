"""The 75 convolutions of YOLOv3 in darknet-cfg order, with the reference's state_dict prefixes.

Restates the module structure built by YoloNet.__init__ (reference darknet.py:167-195): Darknet
([1,2,8,8,4] residual blocks, darknet.py:72-79), three PreDetectionConvGroups (darknet.py:107-118)
and two UpsampleGroups (darknet.py:153-157).  The order is the order WeightManager walks
(darknet.py:292-303) and therefore the order of a darknet .weights stream.
"""
from __future__ import annotations

from typing import List

BLOCKS = [1, 2, 8, 8, 4]
DEFAULT_ANCHORS = [10, 13, 16, 30, 33, 23, 30, 61, 62, 45, 59, 119, 116, 90, 156, 198, 373, 326]


def layer_specs(num_classes: int = 80) -> List[dict]:
    """key, cin, cout, ks, stride, bn, res2 (second conv of a residual block)."""
    t: List[dict] = []

    def add(key, cin, cout, ks, s=1, bn=True, res2=False):
        t.append(dict(key=key, cin=cin, cout=cout, ks=ks, stride=s, bn=bn, res2=res2))

    add("feature.mlist.0", 3, 32, 3)
    idx, ch = 1, 32
    for nb in BLOCKS:
        add(f"feature.mlist.{idx}", ch, 2 * ch, 3, 2)
        idx, ch = idx + 1, 2 * ch
        for _ in range(nb):
            add(f"feature.mlist.{idx}.conv1", ch, ch // 2, 1)
            add(f"feature.mlist.{idx}.conv2", ch // 2, ch, 3, res2=True)
            idx += 1
    for name, nin, nout in (("pre_det1", 1024, 512), ("up1", 512, 256), ("pre_det2", 768, 256),
                            ("up2", 256, 128), ("pre_det3", 384, 128)):
        if name.startswith("up"):
            add(f"{name}.conv", nin, nout, 1)
            continue
        for i in range(3):
            add(f"{name}.mlist.{2 * i}", nin, nout, 1)
            add(f"{name}.mlist.{2 * i + 1}", nout, 2 * nout, 3)
            nin = 2 * nout
        add(f"{name}.mlist.6", nin, (num_classes + 5) * 3, 1, bn=False)
    return t


def darknet_stream_keys(num_classes: int = 80, backbone_only: bool = False):
    """(state_dict key, shape) in the order a darknet weight stream stores them
    (darknet.py:279-290): BN conv -> bn.bias, bn.weight, running_mean, running_var, conv.weight;
    plain conv -> bias, weight."""
    out = []
    for e in layer_specs(num_classes):
        if backbone_only and not e["key"].startswith("feature."):
            break
        k, co, ci, ks = e["key"], e["cout"], e["cin"], e["ks"]
        if e["bn"]:
            out += [(k + ".bn.bias", (co,)), (k + ".bn.weight", (co,)), (k + ".bn.running_mean", (co,)),
                    (k + ".bn.running_var", (co,)), (k + ".conv.weight", (co, ci, ks, ks))]
        else:
            out += [(k + ".bias", (co,)), (k + ".weight", (co, ci, ks, ks))]
    return out


def num_boxes(h: int, w: int) -> int:
    return 3 * ((h // 32) * (w // 32) + (h // 16) * (w // 16) + (h // 8) * (w // 8))


def conv_flops(h: int, w: int, num_classes: int = 80, backbone_only: bool = False) -> float:
    """2*MAC of the convolutions for one image (140.692 GFLOP at 608x608)."""
    total = 0.0
    size = {}
    cur = (h, w)
    for e in layer_specs(num_classes):
        k = e["key"]
        if k.startswith("feature."):
            if e["stride"] == 2:
                cur = (cur[0] // 2, cur[1] // 2)
            size[k] = cur
        elif k.startswith("pre_det1") or k.startswith("up1"):
            size[k] = (h // 32, w // 32)
        elif k.startswith("pre_det2") or k.startswith("up2"):
            size[k] = (h // 16, w // 16)
        else:
            size[k] = (h // 8, w // 8)
    for e in layer_specs(num_classes):
        if backbone_only and not e["key"].startswith("feature."):
            break
        oh, ow = size[e["key"]]
        total += 2.0 * oh * ow * e["cout"] * e["cin"] * e["ks"] * e["ks"]
    return total
