"""yolo_v3_b200 -- B200-native (sm_100a) YOLOv3 inference path behind the reference's Python API.

    from yolo_v3_b200 import YoloNet, postprocessing      # instead of darknet.YoloNet / utils.postprocessing

The compute lives in libyolo_b200.so (C ABI: include/yolo_b200.h), built by __graft_entry__.build().
Importing the package does not need a GPU; running the path does, and fails loudly without one.
"""
from .topology import DEFAULT_ANCHORS, conv_flops, layer_specs, num_boxes  # noqa: F401


def __getattr__(name):
    # torch-dependent members are imported lazily so that `import yolo_v3_b200` stays cheap
    if name in ("YoloNet", "Darknet", "WeightManager", "conv_bn_relu", "res_layer",
                "PreDetectionConvGroup", "UpsampleGroup"):
        from . import darknet
        return getattr(darknet, name)
    if name in ("postprocessing", "postprocessing_raw", "letterbox_transforms", "letterbox_image", "letterbox_batch", "resize_batch", "load_image"):
        from . import utils
        return getattr(utils, name)
    if name == "YoloLayer":
        from .yololayer import YoloLayer
        return YoloLayer
    raise AttributeError(name)
