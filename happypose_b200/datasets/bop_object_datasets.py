"""BOP-format object folders as a RigidObjectDataset (mirror of happypose/toolbox/datasets/bop_object_datasets.py:31-62):
`models_info.json` keyed by integer object id, meshes `obj_%06d.ply` in millimetres, optional symmetry lists.  Symmetries
are carried as plain arrays / dicts (the symmetry classes belong to the training / evaluation code, out of scope)."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from .object_dataset import RigidObject, RigidObjectDataset


class BOPObjectDataset(RigidObjectDataset):
    def __init__(self, ds_dir: Path, label_format: str = "{label}"):
        ds_dir = Path(ds_dir)
        infos = json.loads((ds_dir / "models_info.json").read_text())
        objects = []
        for obj_id, bop_info in infos.items():
            obj_label = f"obj_{int(obj_id):06d}"
            objects.append(RigidObject(
                label=label_format.format(label=obj_label),
                mesh_path=(ds_dir / obj_label).with_suffix(".ply"),
                mesh_units="mm",
                symmetries_discrete=[np.asarray(x, np.float64).reshape(4, 4) for x in bop_info.get("symmetries_discrete", [])],
                symmetries_continuous=[dict(offset=d["offset"], axis=d["axis"]) for d in bop_info.get("symmetries_continuous", [])],
                mesh_diameter=bop_info["diameter"],
            ))
        self.ds_dir = ds_dir
        super().__init__(objects)
