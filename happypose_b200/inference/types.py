"""Inference-time types (happypose/toolbox/inference/types.py:88-235)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from ..utils.tensor_collection import PandasTensorCollection

# infos columns: label, batch_im_id, instance_id[, hypothesis_id, pose_score, coarse_logit, ...]; tensors: poses [B,4,4]
PoseEstimatesType = PandasTensorCollection
# infos columns: label, batch_im_id, instance_id[, score]; tensors: bboxes [B,4]
DetectionsType = PandasTensorCollection


def assert_detections_valid(detections: DetectionsType) -> None:
    df = detections.infos
    for f in ["batch_im_id", "label", "instance_id"]:
        assert f in df, f"detections.infos missing column {f}"
    assert "bboxes" in detections.tensors, "detections missing tensor bboxes."


@dataclass
class InferenceConfig:
    detection_type: str = "detector"
    coarse_estimation_type: str = "SO3_grid"
    SO3_grid_size: int = 576
    n_refiner_iterations: int = 5
    n_pose_hypotheses: int = 5
    run_depth_refiner: bool = False
    depth_refiner: Optional[str] = None
    bsz_objects: int = 16
    bsz_images: int = 576


@dataclass
class ObservationTensor:
    """images: [B,C,H,W], C = 3 (rgb in [0,1]) or 4 (rgb + depth in metres); K: [B,3,3]."""

    images: torch.Tensor
    K: Optional[torch.Tensor] = None

    def cuda(self) -> "ObservationTensor":
        return self.to("cuda")

    def cpu(self) -> "ObservationTensor":
        return self.to("cpu")

    def to(self, device):
        self.images = self.images.to(device)
        if self.K is not None:
            self.K = self.K.to(device)
        return self

    @property
    def batch_size(self) -> int:
        return self.images.shape[0]

    @property
    def depth(self) -> torch.Tensor:
        assert self.channel_dim == 4
        return self.images[:, 3]

    @property
    def channel_dim(self) -> int:
        return self.images.shape[1]

    def is_valid(self) -> bool:
        if self.images.ndim != 4:
            return False
        if self.channel_dim not in (3, 4):
            return False
        if self.K is not None and self.K.shape != torch.Size([self.batch_size, 3, 3]):
            return False
        if self.images.dtype != torch.float:
            return False
        return not bool(torch.max(self.images[:, :3]) > 1)

    @staticmethod
    def from_numpy(rgb: np.ndarray, depth: Optional[np.ndarray] = None, K: Optional[np.ndarray] = None) -> "ObservationTensor":
        assert rgb.dtype == np.uint8
        img = torch.as_tensor(rgb).float() / 255
        if img.shape[-1] == 3:
            img = img.permute(2, 0, 1)
        if depth is not None:
            img = torch.cat((img, torch.as_tensor(depth).unsqueeze(0)), dim=0)
        return ObservationTensor(img.unsqueeze(0), torch.as_tensor(K).float().unsqueeze(0))

    @staticmethod
    def from_torch_batched(rgb: torch.Tensor, depth: torch.Tensor, K: torch.Tensor) -> "ObservationTensor":
        assert rgb.dtype == torch.uint8
        img = torch.as_tensor(rgb).float() / 255
        if depth is not None:
            if depth.ndim == 3:
                depth = depth.unsqueeze(1)
            img = torch.cat((img, depth), dim=1)
        return ObservationTensor(img, torch.as_tensor(K).float())
