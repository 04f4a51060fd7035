"""Renderer data types with the reference's names and fields (happypose/toolbox/renderer/types.py:37-150),
minus everything that needs Panda3D objects."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import torch

RgbaColor = Tuple[float, float, float, float]
Resolution = Tuple[int, int]  # all callers pass (height, width) (types.py:118 unpacks h, w)
NodeFunction = Callable[..., None]


@dataclass
class BatchRenderOutput:
    """types.py:44-56.
    rgbs: (bsz, 3, h, w) float32 in [0, 1]; normals: (bsz, 3, h, w) float32 in [0, 1] or None;
    depths: (bsz, 1, h, w) float32 metres or None; binary_masks: (bsz, 1, h, w) bool or None.
    """

    rgbs: torch.Tensor
    normals: Optional[torch.Tensor]
    depths: Optional[torch.Tensor]
    binary_masks: Optional[torch.Tensor]


@dataclass
class Panda3dLightData:
    """types.py:140-150.  light_type: ambient, point or directional (alpha is irrelevant)."""

    light_type: str
    color: RgbaColor = (1.0, 1.0, 1.0, 1.0)
    positioning_function: Optional[NodeFunction] = None


@dataclass
class Panda3dCameraData:
    """types.py:91-102 (fields only; the lens maths lives in the rasteriser kernel)."""

    K: object
    resolution: Resolution
    TWC: object = None
    z_near: float = 0.1
    z_far: float = 10
    node_name: str = "camera"
    positioning_function: Optional[NodeFunction] = None
