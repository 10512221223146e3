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
    """Reference evaluate.JsonPredictionWriter (evaluate.py:164-195)."""

    def __init__(self, out_path, classes_names, is_letterbox=False):
        self.out_path = out_path
        self.file = open(out_path, "w")
        self.classes_names = classes_names
        self.is_letterbox = is_letterbox

    def write_start(self):
        self.file.write("[")

    def write_end(self):
        self.file.seek(self.file.tell() - 1, os.SEEK_SET)
        self.file.truncate()
        self.file.write("]")
        self.file.close()

    def process_batch(self, sample, predictions):
        imgs, org_imgs, img_paths = sample["img"], sample["org_img"], sample["img_path"]
        for img, org_img, img_path, prediction in zip(imgs, org_imgs, img_paths, predictions):
            img_w, img_h, org_w, org_h = img.shape[2], img.shape[1], org_img.shape[2], org_img.shape[1]
            image_id = get_image_id_from_path(img_path)
            if prediction is not None and len(prediction) != 0:
                bboxes = correct_yolo_boxes(prediction[..., 0:4], org_w, org_h, img_w, img_h, self.is_letterbox)
                category_ids = prediction[..., 6]
                scores = prediction[..., 5]
                for category_id, bbox, score in zip(category_ids.tolist(), bboxes.tolist(), scores.tolist()):
                    res = create_results_entry(image_id, int(category_id), bbox, score)
                    json.dump(res, self.file, indent=4, separators=(",", ":"))
                    self.file.write(",")


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
