"""Host mirror of the detection-results half of the reference's evaluate.py (evaluate.py:117-121, 150-206): the
COCO results writer and the predict loop, with the compute on the device -- `net.detect(..., is_eval=True)` (forward +
eval-mode post-process in one call, yb_detect) and `correct_yolo_boxes` (yb_correct_boxes).  The JSON text is what the
reference writes: one `json.dump(entry, indent=4, separators=(',', ':'))` per detection, comma separated, in `[...]`.
Dataset / annotation tooling (COCOEvalDataset, create_annotations*, pycocotools) is outside the path and not mirrored.
"""
from __future__ import annotations

import json
import os
import os.path as osp
import re
from collections import OrderedDict
from contextlib import contextmanager

import torch

from .boundingbox import correct_yolo_boxes


def get_image_id_from_path(image_path):
    """Reference utils.get_image_id_from_path (utils.py:294-297)."""
    image_path = osp.splitext(image_path)[0]
    m = re.search(r"\d+$", image_path)
    return int(m.group())


def create_results_entry(image_id, category_id, bbox, score):
    """Reference evaluate.create_results_entry (evaluate.py:117-121)."""
    return OrderedDict({"image_id": image_id, "category_id": category_id, "bbox": bbox, "score": score})


class BatchHandler:
    def process_batch(self, sample, predictions):
        raise NotImplementedError


class JsonPredictionWriter(BatchHandler):
    """Same interface and output text as the reference's writer (evaluate.py:164-195): `[`, one
    `json.dump(entry, indent=4, separators=(',', ':'))` per detection separated by commas, `]`.  The whole batch is
    box-corrected on the device in one call and formatted from host lists; an empty result set gives `[]` (the
    reference's seek-back-and-truncate leaves a lone `]` there)."""

    def __init__(self, out_path, classes_names, is_letterbox=False):
        self.out_path = out_path
        self.classes_names = classes_names
        self.is_letterbox = is_letterbox
        self.file = open(out_path, "w")
        self._entries = 0

    def write_start(self):
        self.file.write("[")

    def write_end(self):
        self.file.write("]")
        self.file.close()

    def _emit(self, entry):
        if self._entries:
            self.file.write(",")
        json.dump(entry, self.file, indent=4, separators=(",", ":"))
        self._entries += 1

    def process_batch(self, sample, predictions):
        net_hw = [tuple(t.shape[1:3]) for t in sample["img"]]
        org_hw = [tuple(t.shape[1:3]) for t in sample["org_img"]]
        for (ih, iw), (oh, ow), path, pred in zip(net_hw, org_hw, sample["img_path"], predictions):
            if pred is None or len(pred) == 0:
                continue
            xywh = correct_yolo_boxes(pred[..., 0:4], ow, oh, iw, ih, self.is_letterbox).tolist()
            image_id = get_image_id_from_path(path)
            for box, score, cls in zip(xywh, pred[..., 5].tolist(), pred[..., 6].tolist()):
                self._emit(create_results_entry(image_id, int(cls), box, score))


@contextmanager
def open_json_pred_writer(out_path, classes_names, is_letterbox=False):
    pred_writer = JsonPredictionWriter(out_path, classes_names, is_letterbox)
    try:
        pred_writer.write_start()
        yield pred_writer
    finally:
        pred_writer.write_end()


def predict_and_process(data, net, num_classes, batch_handler=None):
    """Reference evaluate.predict_and_process (evaluate.py:197-206): conf 0.005, nms 0.45, eval mode, one fused device
    call per batch instead of net(...) + torch.cat + postprocessing."""
    with torch.no_grad():
        for sample in data:
            predictions = net.detect(sample["img"].cuda(), 0.005, 0.45, is_eval=True, use_nms=True)
            if not predictions:                                   # [] when nothing passes in the whole batch
                predictions = [torch.Tensor()] * len(sample["img"])
            batch_handler.process_batch(sample, predictions)
