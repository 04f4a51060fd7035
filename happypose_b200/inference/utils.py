"""Detection helpers (happypose/toolbox/inference/utils.py:164-206)."""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import pandas as pd
import torch

from ..utils.tensor_collection import PandasTensorCollection


def add_instance_id(inputs: PandasTensorCollection) -> PandasTensorCollection:
    """Adds `instance_id` = running index of the detection inside its (batch_im_id, label) group.
    The reference's groupby(...).apply(...) (utils.py:175-183) drops the grouping columns under pandas >= 2.2/3.0;
    cumcount() gives the pinned pandas-2.2.2 result on every version."""
    if "instance_id" in inputs.infos:
        return inputs
    df = inputs.infos
    # cumcount of (batch_im_id, label) in row order, without a GroupBy object (this runs once per frame on the hot path)
    seen: dict = {}
    inst = np.empty(len(df), np.int64)
    for i, key in enumerate(zip(df["batch_im_id"].tolist(), df["label"].tolist())):
        inst[i] = seen.get(key, 0)
        seen[key] = inst[i] + 1
    df["instance_id"] = inst
    inputs.infos = df
    return inputs


def filter_detections(detections: PandasTensorCollection, labels: Optional[List[str]] = None, one_instance_per_class: bool = False) -> PandasTensorCollection:
    if labels is not None:
        df = detections.infos
        df = df[df.label.isin(labels)]
        detections = detections[df.index.tolist()]
    if one_instance_per_class:
        df = detections.infos
        df = df.sort_values("score", ascending=False, kind="stable").groupby(["batch_im_id", "label"]).head(1)
        detections = detections[df.index.tolist()]
    return detections


def make_detections_from_object_data(object_data: List[dict], device="cuda") -> PandasTensorCollection:
    """[{label, bbox_modal=[x1,y1,x2,y2]}, ...] -> detections of frame 0 (inference/utils.py:230-244 analogue)."""
    infos = pd.DataFrame({"label": [d["label"] for d in object_data], "batch_im_id": 0, "instance_id": np.arange(len(object_data))})
    bboxes = torch.as_tensor(np.stack([np.asarray(d["bbox_modal"], np.float32) for d in object_data])).float()
    return PandasTensorCollection(infos=infos, bboxes=bboxes.to(device))
