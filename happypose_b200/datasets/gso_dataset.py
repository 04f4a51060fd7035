"""Google-Scanned-Objects folders as a RigidObjectDataset (mirror of happypose/toolbox/datasets/gso_dataset.py:31-83):
<root>/models_<split>/<object id>/meshes/model.obj (+ .mtl + texture), ids listed in <root>/invalid_meshes.json skipped,
labels `gso_<id>`; the "normalized" / "pointcloud" splits are stored 10x larger than metres (scaling_factor 0.1)."""
from __future__ import annotations

import json
from pathlib import Path
from typing import List

from .object_dataset import RigidObject, RigidObjectDataset


def make_gso_infos(gso_dir: Path, model_name: str = "model.obj") -> List[str]:
    gso_dir = Path(gso_dir)
    invalid_file = gso_dir.parent / "invalid_meshes.json"
    invalid = set(json.loads(invalid_file.read_text())) if invalid_file.exists() else set()
    return sorted(d.name for d in gso_dir.iterdir() if (d / "meshes" / model_name).exists() and d.name not in invalid)


class GoogleScannedObjectDataset(RigidObjectDataset):
    def __init__(self, gso_root: Path, split: str = "orig"):
        gso_root = Path(gso_root)
        self.gso_dir = gso_root / f"models_{split}"
        if split == "orig":
            scaling_factor = 1.0
        elif split in {"normalized", "pointcloud"}:
            scaling_factor = 0.1
        else:
            raise ValueError(split)
        objects = [RigidObject(label=f"gso_{oid}", mesh_path=self.gso_dir / oid / "meshes" / "model.obj", scaling_factor=scaling_factor)
                   for oid in make_gso_infos(self.gso_dir)]
        super().__init__(objects)
