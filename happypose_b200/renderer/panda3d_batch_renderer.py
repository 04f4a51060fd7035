"""Drop-in replacement for happypose.toolbox.renderer.panda3d_batch_renderer.Panda3dBatchRenderer
(panda3d_batch_renderer.py:128-348): same constructor, same render() signature, same BatchRenderOutput layout --
but the b scenes are rasterised by ONE CUDA launch (hpb_render) instead of b Panda3D/OpenGL frames in CPU worker
processes, and nothing crosses the host: poses stay on the device, images are produced on the device.

Differences a caller can observe (all documented in DESIGN.md):
  * `n_workers`, `preload_cache`, `split_objects` are accepted and ignored (there are no worker processes;
    meshes are always uploaded at construction).
  * outputs are contiguous NCHW tensors (the reference returns permuted views of NHWC memory, :249-279).
  * only ambient lights are evaluated (the hot path uses a single ambient light, pose_rigid.py:415-420);
    point / directional lights raise NotImplementedError instead of being shaded.
  * render_binary_mask without render_depth raises AssertionError eagerly (the reference loses the worker's
    assertion and hangs, test_batch_renderer_panda3d.py:244-256; Panda3dSceneRenderer asserts, :331-332).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .. import ops
from .._capi import Context
from ..datasets.object_dataset import RigidObjectDataset
from ..io import mesh_io
from .types import BatchRenderOutput, Panda3dLightData, Resolution


def ambient_from_light_datas(light_datas: Sequence[Sequence[Panda3dLightData]]) -> Optional[np.ndarray]:
    """[b,3] summed ambient colour per scene; None when every scene is lit by exactly ambient (1,1,1)."""
    amb = np.zeros((len(light_datas), 3), np.float32)
    trivial = True
    for n, lights in enumerate(light_datas):
        for light in lights:
            if light.light_type != "ambient":
                raise NotImplementedError(
                    f"light_type={light.light_type!r}: only ambient lights are evaluated by the CUDA rasteriser")
            amb[n] += np.asarray(light.color[:3], np.float32)
        if not np.array_equal(amb[n], np.ones(3, np.float32)):
            trivial = False
    return None if trivial else amb


class Panda3dBatchRenderer:
    def __init__(
        self,
        asset_dataset: RigidObjectDataset,
        n_workers: int = 8,
        preload_cache: bool = True,
        split_objects: bool = False,
        device=None,
    ):
        assert n_workers >= 1
        self._object_dataset = asset_dataset
        self._n_workers = n_workers
        self._split_objects = split_objects
        self._is_closed = False
        self._ctx = Context.get(device)
        self._label_to_mesh_id: Dict[str, int] = {}
        for obj in asset_dataset.list_objects:
            self._upload(obj)

    # ------------------------------------------------------------------------------------------
    def _upload(self, obj) -> None:
        mesh = mesh_io.load_mesh(obj.mesh_path)
        scale = obj.scaling_factor_mesh_units_to_meters * obj.scaling_factor  # panda3d_scene_renderer.py:211
        verts = np.asarray(mesh.verts, np.float64)
        normals = mesh.normals
        if tuple(float(a) for a in obj.ypr_offset_deg) != (0.0, 0.0, 0.0):  # node.setHpr (:216), renderer only
            R = mesh_io.hpr_matrix(obj.ypr_offset_deg)
            verts = verts @ R.T
            if normals is not None:
                normals = (np.asarray(normals, np.float64) @ R.T).astype(np.float32)
        verts_m = (verts * float(scale)).astype(np.float32)
        self._label_to_mesh_id[obj.label] = ops.mesh_upload(
            self._ctx, verts_m, mesh.faces, normals, mesh.uv, mesh.vcolor, mesh.texture)

    def __deepcopy__(self, memo):  # meshes live in the per-device context: model copies share the renderer
        return self

    @property
    def device(self) -> torch.device:
        return self._ctx.device

    def mesh_ids(self, labels: Sequence[str]) -> torch.Tensor:
        """int32 device tensor of mesh ids; unknown labels raise KeyError like the reference (:220)."""
        ids = [self._label_to_mesh_id[label] for label in labels]
        return torch.tensor(ids, dtype=torch.int32).to(self._ctx.device, non_blocking=True)

    # ------------------------------------------------------------------------------------------
    def render(
        self,
        labels: List[str],
        TCO: torch.Tensor,
        K: torch.Tensor,
        light_datas: List[List[Panda3dLightData]],
        resolution: Resolution,
        render_normals: bool = False,
        render_depth: bool = False,
        render_binary_mask: bool = False,
    ) -> BatchRenderOutput:
        bsz = TCO.shape[0]
        assert TCO.shape == (bsz, 4, 4)
        assert K.shape == (bsz, 3, 3)
        assert bsz == len(labels), "Need same number of labels as TCO/K batch size"
        if render_binary_mask:
            assert render_depth, "Binary mask can only be rendered if depth is rendered"
        assert not self._is_closed, "renderer was stopped"
        ambient = ambient_from_light_datas(light_datas) if light_datas is not None else None
        rgbs, normals, depths, masks = ops.render(
            self._ctx,
            self.mesh_ids(labels),
            TCO.detach(),
            K,
            resolution,
            ambient=None if ambient is None else torch.as_tensor(ambient),
            render_normals=render_normals,
            render_depth=render_depth,
            render_binary_mask=render_binary_mask,
        )
        return BatchRenderOutput(rgbs=rgbs, normals=normals, depths=depths, binary_masks=masks)

    def render_into(self, mesh_ids, TCO, K, resolution, out, out_channel_offset, render_normals, render_depth, views=1, ambient=None):
        """Fast path used by PosePredictor: device mesh ids in, network-input slice out (no cat, no label lookup)."""
        return ops.render(
            self._ctx, mesh_ids, TCO, K, resolution, ambient=ambient, render_normals=render_normals,
            render_depth=render_depth, out=out, out_channel_offset=out_channel_offset, views=views)

    # ------------------------------------------------------------------------------------------
    def stop(self) -> None:
        """Idempotent (:332-345).  Mesh buffers live as long as the per-device context."""
        self._is_closed = True

    def __del__(self) -> None:
        self._is_closed = True
