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

import hashlib

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .. import ops
from .._capi import Context
from ..datasets.object_dataset import RigidObjectDataset
from ..io import mesh_io
from .types import BatchRenderOutput, Panda3dLightData, Resolution


MAX_LIGHTS = 8  # point / directional lights per scene evaluated by hpb_render (HPB_MAX_LIGHTS)


class _Bounds:
    def __init__(self, radius: float):
        self.radius = radius

    def get_radius(self) -> float:
        return self.radius

    getRadius = get_radius


class _RootNode:
    """What a Panda3dLightData.positioning_function may ask the scene root for (panda3d_scene_renderer.py:121-129:
    `root_node.getBounds().radius`): the bounding sphere of the scene = of the single object, which sits at the identity."""

    def __init__(self, radius: float):
        self._bounds = _Bounds(radius)

    def getBounds(self):
        return self._bounds

    get_bounds = getBounds


class _LightNode:
    """Records where a positioning_function puts a light: setPos / set_pos (point lights), lookAt / look_at from the
    current position or setHpr-free direction via set_direction (directional lights)."""

    def __init__(self):
        self.pos = np.zeros(3, np.float64)
        self.direction = np.array([0.0, 1.0, 0.0])  # Panda3D lights shine along +Y of their node by default

    def setPos(self, *xyz):
        self.pos = np.asarray(xyz[0] if len(xyz) == 1 else xyz, np.float64).reshape(3)

    set_pos = setPos

    def lookAt(self, *xyz):
        tgt = np.asarray(xyz[0] if len(xyz) == 1 else xyz, np.float64).reshape(3)
        d = tgt - self.pos
        n = np.linalg.norm(d)
        if n > 0:
            self.direction = d / n

    look_at = lookAt

    def set_direction(self, xyz):
        d = np.asarray(xyz, np.float64).reshape(3)
        self.direction = d / max(np.linalg.norm(d), 1e-30)


def lights_from_light_datas(light_datas: Sequence[Sequence[Panda3dLightData]], radii: Sequence[float]):
    """-> (ambient [b,3] or None, lights [b, n, 8] or None).
    ambient: summed ambient colour per scene; None when every scene is lit by exactly ambient (1,1,1) and nothing else.
    lights: the scenes' point / directional lights as hpb_render wants them (type, xyz in the world = object frame, rgb, -),
    padded with black lights to the largest count.  Positions come from each light's positioning_function, called with
    stand-ins for Panda3D's root and light NodePaths (panda3d_scene_renderer.py:294-318); radii[n] is the bounding-sphere
    radius of scene n's object (what root_node.getBounds().radius returns there)."""
    b = len(light_datas)
    amb = np.zeros((b, 3), np.float32)
    per_scene = []
    trivial = True
    for n, lights in enumerate(light_datas):
        rows = []
        for light in lights:
            if light.light_type == "ambient":
                amb[n] += np.asarray(light.color[:3], np.float32)
                continue
            if light.light_type not in ("point", "directional"):
                raise NotImplementedError(light.light_type)  # panda3d_scene_renderer.py:309-310
            assert light.positioning_function is not None  # :303, :308
            node = _LightNode()
            light.positioning_function(_RootNode(float(radii[n])), node)
            xyz = node.pos if light.light_type == "point" else node.direction
            rows.append([0.0 if light.light_type == "point" else 1.0, *xyz, *light.color[:3], 0.0])
        if len(rows) > MAX_LIGHTS:
            raise NotImplementedError(f"{len(rows)} point / directional lights in one scene (at most {MAX_LIGHTS} are evaluated)")
        per_scene.append(rows)
        if rows or not np.array_equal(amb[n], np.ones(3, np.float32)):
            trivial = False
    n_max = max((len(r) for r in per_scene), default=0)
    lights_arr = None
    if n_max > 0:
        lights_arr = np.zeros((b, n_max, 8), np.float32)
        for n, rows in enumerate(per_scene):
            if rows:
                lights_arr[n, :len(rows)] = np.asarray(rows, np.float32)
    return (None if trivial else amb), lights_arr


def ambient_from_light_datas(light_datas: Sequence[Sequence[Panda3dLightData]]) -> Optional[np.ndarray]:
    """[b,3] summed ambient colour per scene for ambient-only scenes; None when every scene is lit by exactly (1,1,1)."""
    amb, lights = lights_from_light_datas(light_datas, [1.0] * len(light_datas))
    assert lights is None, "scene has point / directional lights: use lights_from_light_datas"
    return amb


def make_scene_lights(ambient_light_color=(0.1, 0.1, 0.1, 1.0), point_lights_color=(0.4, 0.4, 0.4, 1.0)) -> List[Panda3dLightData]:
    """1 ambient light + 6 point lights at +-10 bounding radii along the world axes (panda3d_scene_renderer.py:105-141):
    the light rig of models that render without normals (pose_rigid.py:421-422)."""
    from functools import partial

    def pos_fn(root_node, light_node, pos):
        radius = root_node.getBounds().radius
        light_node.setPos(tuple((np.asarray(pos, np.float64) * radius * 10).tolist()))

    light_datas = [Panda3dLightData(light_type="ambient", color=ambient_light_color)]
    for pos_n in ([1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]):
        light_datas.append(Panda3dLightData(light_type="point", color=point_lights_color, positioning_function=partial(pos_fn, pos=pos_n)))
    return light_datas


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
        self._label_to_radius: Dict[str, float] = {}  # bounding-sphere radius (metres): scale of the light rig
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
        # Panda3D's bounding sphere of a GeomNode: centred on the bounding box, reaching the farthest vertex
        centre = 0.5 * (verts_m.min(0).astype(np.float64) + verts_m.max(0).astype(np.float64))
        self._label_to_radius[obj.label] = float(np.linalg.norm(verts_m.astype(np.float64) - centre, axis=1).max())
        # Meshes live in the per-device context for its lifetime: a second renderer over the same assets (the coarse and the
        # refiner model each build one in the reference's set-up code) reuses the uploaded copy instead of adding another.
        h = hashlib.blake2b(digest_size=16)
        for a in (verts_m, mesh.faces, normals, mesh.uv, mesh.vcolor, mesh.texture):
            h.update(b"-" if a is None else np.ascontiguousarray(a).tobytes() + str(np.asarray(a).shape).encode())
        cache = self._ctx.__dict__.setdefault("_uploaded_meshes", {})
        key = h.digest()
        if key not in cache:
            cache[key] = ops.mesh_upload(self._ctx, verts_m, mesh.faces, normals, mesh.uv, mesh.vcolor, mesh.texture)
        self._label_to_mesh_id[obj.label] = cache[key]

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
        ambient, lights = (lights_from_light_datas(light_datas, [self._label_to_radius[lb] for lb in labels])
                           if light_datas is not None else (None, None))
        rgbs, normals, depths, masks = ops.render(
            self._ctx,
            self.mesh_ids(labels),
            TCO.detach(),
            K,
            resolution,
            ambient=None if ambient is None else torch.as_tensor(ambient),
            lights=None if lights is None else torch.as_tensor(lights),
            render_normals=render_normals,
            render_depth=render_depth,
            render_binary_mask=render_binary_mask,
        )
        return BatchRenderOutput(rgbs=rgbs, normals=normals, depths=depths, binary_masks=masks)

    def render_into(self, mesh_ids, TCO, K, resolution, out, out_channel_offset, render_normals, render_depth, views=1, ambient=None,
                    lights=None):
        """Fast path used by PosePredictor: device mesh ids in, network-input slice out (no cat, no label lookup)."""
        return ops.render(
            self._ctx, mesh_ids, TCO, K, resolution, ambient=ambient, render_normals=render_normals,
            render_depth=render_depth, out=out, out_channel_offset=out_channel_offset, views=views, lights=lights)

    def scene_light_rig(self, mesh_ids: torch.Tensor) -> torch.Tensor:
        """[b,6,8] device tensor: make_scene_lights()'s six point lights for every scene, scaled by each mesh's bounding
        radius (device-side table lookup: no label round trip).  The rig's ambient term is (0.1, 0.1, 0.1)."""
        if getattr(self, "_rig_table", None) is None or self._rig_table.shape[0] != len(self._label_to_mesh_id):
            by_id = sorted((mid, self._label_to_radius[lb]) for lb, mid in self._label_to_mesh_id.items())
            n_ids = max(m for m, _ in by_id) + 1
            table = np.zeros((n_ids, 6, 8), np.float32)
            axes = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], np.float64)
            for mid, radius in by_id:
                table[mid, :, 1:4] = (axes * radius * 10).astype(np.float32)
                table[mid, :, 4:7] = 0.4
            self._rig_table = torch.as_tensor(table).to(self._ctx.device)
        return self._rig_table[mesh_ids.long()]

    # ------------------------------------------------------------------------------------------
    def stop(self) -> None:
        """Idempotent (:332-345).  Mesh buffers live as long as the per-device context."""
        self._is_closed = True

    def __del__(self) -> None:
        self._is_closed = True
