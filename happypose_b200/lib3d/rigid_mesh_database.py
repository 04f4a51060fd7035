"""MeshDataBase / BatchedMeshes / Meshes (happypose/toolbox/lib3d/rigid_mesh_database.py:52-200,
lib3d/mesh_ops.py:74-84) without trimesh: vertices come from happypose_b200.io.mesh_io.

The point sets feed the crop / TCO-init kernels.  Because sample_points(n, deterministic=True) draws the SAME
RandomState(0) index set on every call, the 2000- and 200-point subsets are gathered once per mesh database and
cached on the device (the reference re-gathers b x N_max x 3 floats on every iteration).
"""
from __future__ import annotations

from copy import deepcopy
from typing import Dict, List

import numpy as np
import torch

from ..datasets.object_dataset import RigidObject
from ..io import mesh_io
from ..utils.tensor_collection import TensorCollection


def sample_points(points: torch.Tensor, n_points: int, deterministic: bool = False) -> torch.Tensor:
    """mesh_ops.py:74-84."""
    assert points.dim() == 3
    assert n_points <= points.shape[1]
    rs = np.random.RandomState(0) if deterministic else np.random
    ids = torch.as_tensor(rs.choice(points.shape[1], size=n_points, replace=False)).to(points.device)
    return torch.index_select(points, 1, ids)


def pad_stack_tensors(tensor_list, fill="select_random", deterministic=True):
    """rigid_mesh_database.py:172-200: pad every tensor to the longest one, re-drawing own rows (one shared
    RandomState(0)) or repeating a fill tensor."""
    n_max = max(t.shape[0] for t in tensor_list)
    rs = np.random.RandomState(0) if deterministic else np.random
    out = []
    for t in tensor_list:
        n_pad = n_max - len(t)
        if n_pad > 0:
            if isinstance(fill, torch.Tensor):
                assert fill.shape == t.shape[1:]
                pad = fill.unsqueeze(0).repeat(n_pad, *[1 for _ in fill.shape]).to(t.device).to(t.dtype)
            else:
                assert fill == "select_random"
                pad = t[rs.choice(np.arange(len(t)), size=n_pad)]
            t = torch.cat((t, pad), dim=0)
        out.append(t)
    return torch.stack(out)


class MeshDataBase:
    def __init__(self, obj_list: List[RigidObject]):
        self.obj_dict = {obj.label: obj for obj in obj_list}
        self.obj_list = obj_list
        self.infos: Dict[str, dict] = {obj.label: {} for obj in obj_list}
        self.meshes = {label: mesh_io.load_mesh(obj.mesh_path) for label, obj in self.obj_dict.items()}
        for label, obj in self.obj_dict.items():
            if obj.diameter_meters is None:
                pts = np.asarray(self.meshes[label].verts) * obj.scale
                obj.diameter_meters = float(np.linalg.norm(pts.max(0) - pts.min(0)))

    @staticmethod
    def from_object_ds(object_ds) -> "MeshDataBase":
        return MeshDataBase([object_ds[n] for n in range(len(object_ds))])

    def batched(self, aabb=False, resample_n_points=None, n_sym=64) -> "BatchedMeshes":
        if aabb or resample_n_points:
            raise NotImplementedError("aabb / surface resampling are training-time options (out of scope)")
        labels, points, symmetries = [], [], []
        infos = deepcopy(self.infos)
        for label, mesh in self.meshes.items():
            obj = self.obj_dict[label]
            pts = torch.tensor(np.asarray(mesh.verts, np.float64)) * obj.scale  # float64, like trimesh vertices
            sym = torch.as_tensor(obj.make_symmetry_poses(n_symmetries_continuous=n_sym))
            infos[label]["n_points"] = pts.shape[0]
            infos[label]["n_sym"] = sym.shape[0]
            labels.append(label)
            points.append(pts)
            symmetries.append(sym)
        points = pad_stack_tensors(points, fill="select_random", deterministic=True)
        symmetries = pad_stack_tensors(symmetries, fill=torch.eye(4), deterministic=True)
        return BatchedMeshes(infos, np.array(labels), points, symmetries).float()


class BatchedMeshes(TensorCollection):
    def __init__(self, infos, labels, points, symmetries):
        super().__init__()
        self.infos = infos
        self.label_to_id = {label: n for n, label in enumerate(labels)}
        self.labels = np.asarray(labels)
        self.register_tensor("points", points)
        self.register_tensor("symmetries", symmetries)
        self.__dict__["_subset_cache"] = {}

    def __deepcopy__(self, memo):  # read-only database: copies of a model keep sharing it (and its device buffers)
        return self

    @property
    def n_sym_mapping(self):
        return {label: obj["n_sym"] for label, obj in self.infos.items()}

    def select(self, labels) -> "Meshes":
        ids = [self.label_to_id[label] for label in labels]  # KeyError on unknown labels, like the reference
        return Meshes(
            infos=[self.infos[label] for label in labels],
            labels=self.labels[ids],
            points=self.points[ids],
            symmetries=self.symmetries[ids],
        )

    # ---- device-side helpers used by the kernels (no per-hypothesis gather) ----
    def label_ids(self, labels, device=None) -> torch.Tensor:
        ids = torch.tensor([self.label_to_id[label] for label in labels], dtype=torch.int32)
        return ids.to(device if device is not None else self.points.device)

    def points_subset(self, n_points: int) -> torch.Tensor:
        """[n_obj, n_points, 3]: the deterministic RandomState(0) subset, gathered once per device/dtype."""
        key = (n_points, self.points.device, self.points.data_ptr())
        cache = self.__dict__["_subset_cache"]
        if key not in cache:
            n_max = self.points.shape[1]
            if n_points >= n_max:
                assert n_points == n_max, "sample_points asks for more points than the padded meshes have"
                cache[key] = self.points.contiguous()
            else:
                cache[key] = sample_points(self.points, n_points, deterministic=True).contiguous()
        return cache[key]


class Meshes(TensorCollection):
    def __init__(self, infos, labels, points, symmetries):
        super().__init__()
        self.infos = infos
        self.labels = np.asarray(labels)
        self.register_tensor("points", points)
        self.register_tensor("symmetries", symmetries)

    def select_labels(self, labels):
        raise NotImplementedError

    def sample_points(self, n_points, deterministic=False):
        return sample_points(self.points, n_points, deterministic=deterministic)
