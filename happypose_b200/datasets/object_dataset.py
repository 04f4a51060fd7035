"""RigidObject / RigidObjectDataset with the reference's constructor and attributes
(happypose/toolbox/datasets/object_dataset.py:32-173); symmetry pose generation is out of scope
(training / evaluation only) and returns the identity."""
from __future__ import annotations

from pathlib import Path
from typing import List, Optional, Set, Tuple

import numpy as np


class RigidObject:
    def __init__(
        self,
        label: str,
        mesh_path: Path,
        category: Optional[str] = None,
        mesh_diameter: Optional[float] = None,
        mesh_units: str = "m",
        symmetries_discrete: Optional[list] = None,
        symmetries_continuous: Optional[list] = None,
        ypr_offset_deg: Tuple[float, float, float] = (0.0, 0.0, 0.0),
        scaling_factor: float = 1.0,
        scaling_factor_mesh_units_to_meters: Optional[float] = None,
    ):
        self.label = label
        self.category = category
        self.mesh_path = mesh_path
        self.mesh_units = mesh_units
        if scaling_factor_mesh_units_to_meters is not None:
            self.scaling_factor_mesh_units_to_meters = scaling_factor_mesh_units_to_meters
        else:
            self.scaling_factor_mesh_units_to_meters = {"m": 1.0, "mm": 0.001}[self.mesh_units]
        self.scaling_factor = scaling_factor
        self._mesh_diameter = None  # the reference never stores mesh_diameter (object_dataset.py:113-120)
        self.diameter_meters = None
        self.symmetries_discrete = symmetries_discrete or []
        self.symmetries_continuous = symmetries_continuous or []
        self.ypr_offset_deg = ypr_offset_deg

    @property
    def is_symmetric(self) -> bool:
        return len(self.symmetries_discrete) > 0 or len(self.symmetries_continuous) > 0

    @property
    def scale(self) -> float:
        """Scale factor that converts the mesh to metres."""
        return self.scaling_factor_mesh_units_to_meters * self.scaling_factor

    def make_symmetry_poses(self, n_symmetries_continuous: int = 64) -> np.ndarray:
        return np.eye(4, dtype=np.float32)[None]


class RigidObjectDataset:
    def __init__(self, objects: List[RigidObject]):
        self.list_objects = objects
        self.label_to_objects = {obj.label: obj for obj in objects}
        if len(self.list_objects) != len(self.label_to_objects):
            raise RuntimeError("There are objects with duplicate labels")

    def __getitem__(self, idx: int) -> RigidObject:
        return self.list_objects[idx]

    def get_object_by_label(self, label: str) -> RigidObject:
        return self.label_to_objects[label]

    def __len__(self) -> int:
        return len(self.list_objects)

    @property
    def objects(self) -> List[RigidObject]:
        return self.list_objects

    def filter_objects(self, keep_labels: Set[str]) -> "RigidObjectDataset":
        return RigidObjectDataset([obj for obj in self.list_objects if obj.label in keep_labels])
