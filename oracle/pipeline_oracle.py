"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the MegaPose / CosyPose render-and-compare pipeline, assembled from
the pinned numpy oracle (np_oracle.py), the rasteriser oracle (raster.py) and the torch networks run on the CPU in
float32.  Never imported by happypose_b200/.  Used by tests/ (end-to-end parity: final poses within 1 mm ADD),
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.

Follows:
  PosePredictor.forward_coarse    megapose/models/pose_rigid.py:708-788
  PosePredictor.forward           megapose/models/pose_rigid.py:546-674
  PoseEstimator.forward_coarse_model / forward_refiner / forward_scoring_model / run_inference_pipeline
                                  megapose/inference/pose_estimator.py:328-485,105-220,223-325,516-668
  CosyPose PosePredictor.forward  cosypose/models/pose.py:116-199
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import np_oracle as O
from . import raster


@dataclass
class OracleScene:
    """Everything the pipeline needs about the objects: oracle meshes for rendering, the padded mesh-db points."""

    meshes: List[raster.OracleMesh]          # indexed by object id
    points: np.ndarray                       # [n_obj, N_max, 3] float32 (rigid_mesh_database.batched().points)

    def subset(self, n: int) -> np.ndarray:
        n_max = self.points.shape[1]
        if n >= n_max:
            return self.points
        return self.points[:, O.sample_point_ids(n_max, n)]


def make_scene(mesh_arrays: Sequence[dict], scales: Sequence[float]) -> OracleScene:
    meshes = [raster.OracleMesh(d["verts"], d["faces"], d.get("normals"), d.get("uv"), d.get("vcolor"), d.get("texture"), scale=s)
              for d, s in zip(mesh_arrays, scales)]
    points = O.mesh_db_points([d["verts"] for d in mesh_arrays], scales)
    return OracleScene(meshes, points)


def cpu_model(model, net_device="cpu"):
    """Float32 copy of a PosePredictor's torch network + the configuration flags the pipeline reads.

    net_device: where the (reference, unchanged) torch network is evaluated.  "cpu" = everything on the host (bench
    baseline, smoke).  The parity tests pass the GPU here: the ResNet is not part of the ported path, and evaluating
    it with the same cuDNN kernels on both sides makes the comparison measure the kernels under test (raster, crop,
    pose update) instead of cuDNN-vs-oneDNN summation order, which a random-init ResNet amplifies across iterations.
    Every other operation of the oracle pipeline stays on the CPU."""
    import copy
    from types import SimpleNamespace

    backbone = copy.deepcopy(model.backbone).to(net_device).float().eval()
    heads = {k: copy.deepcopy(h).to(net_device).float().eval() for k, h in model.heads.items()}
    ns = SimpleNamespace(backbone=backbone, heads=heads, render_size=tuple(model.render_size), net_device=net_device)
    for k, default in (("input_depth", False), ("render_normals", False), ("render_depth", False), ("n_rendered_views", 1),
                       ("multiview_type", "TCO+front_3views"), ("remove_TCO_rendering", False),
                       ("depth_normalization_type", "none"), ("pose_dim", 9)):
        setattr(ns, k, getattr(model, k, default))
    return ns


def _net(model, x: np.ndarray) -> Dict[str, np.ndarray]:
    """PosePredictor.net_forward on the CPU in float32 (pose_rigid.py:352-374)."""
    with torch.no_grad():
        f = model.backbone(torch.as_tensor(x).to(getattr(model, "net_device", "cpu")))
        if f.dim() == 4:
            f = f.flatten(2).mean(dim=-1)
        return {k: head(f).cpu().numpy() for k, head in model.heads.items()}


def _render(scene: OracleScene, obj_ids, TCO, K, size, normals, depth, n_threads):
    out = raster.render(scene.meshes, obj_ids, TCO, K, size, render_normals=normals, render_depth=depth, n_threads=n_threads)
    planes = [out["rgb"]]
    if normals:
        planes.append(out["normals"])
    if depth:
        planes.append(out["depth"])
    return np.concatenate(planes, 1)


def _depth_channels(model) -> List[int]:
    ch = [3] if model.input_depth else []
    c_in = 3 + len(ch)
    if model.render_depth:
        c_r = 3 + (3 if model.render_normals else 0) + 1
        ch += [c_in + (c_r - 1) + c_r * v for v in range(model.n_rendered_views)]
    return ch


def _normalize(model, x: np.ndarray, tCR: np.ndarray) -> np.ndarray:
    ch = _depth_channels(model)
    if ch:
        x[:, ch] = O.normalize_depth(x[:, ch], tCR[:, 2], model.depth_normalization_type)
    return x


def forward_coarse(model, scene: OracleScene, images, K_rows, im_ids, obj_ids, TCO, n_threads=1):
    """-> dict(logits [b,1], scores, x [b,C,h,w], K_crop, boxes_crop)."""
    if not model.input_depth:
        images = images[:, :3]
    TCO = O.normalize_T(np.asarray(TCO, np.float32))
    tCR = TCO[:, :3, 3].copy()
    pts = scene.subset(2000)[np.asarray(obj_ids)]
    crops, K_crop, boxes_rend, boxes_crop = O.crop_inputs(images, K_rows, TCO, tCR, pts, model.render_size, im_ids=im_ids)
    renders = _render(scene, obj_ids, TCO, K_crop, model.render_size, model.render_normals, model.render_depth, n_threads)
    x = _normalize(model, np.concatenate([crops, renders], 1).astype(np.float32), tCR)
    logits = _net(model, x)["renderings_logits"]
    return {"logits": logits, "scores": 1.0 / (1.0 + np.exp(-logits)), "x": x, "K_crop": K_crop, "boxes_crop": boxes_crop, "boxes_rend": boxes_rend}


def forward_refiner(model, scene: OracleScene, images, K_rows, im_ids, obj_ids, TCO, n_iterations, n_threads=1):
    """-> list of per-iteration dicts (TCO_input, TCO_output, K_crop, KV_crop, TCV_O, x, pose)."""
    if not model.input_depth:
        images = images[:, :3]
    obj_ids = np.asarray(obj_ids)
    b = len(TCO)
    V = model.n_rendered_views
    TCO_input = np.asarray(TCO, np.float32)
    out = []
    for _ in range(n_iterations):
        TCO_input = O.normalize_T(TCO_input)
        tCR = TCO_input[:, :3, 3].copy()
        TCV_O = O.make_TCO_multiview(TCO_input, tCR, model.multiview_type, V, model.remove_TCO_rendering)
        tCV_R = TCV_O[:, :, :3, 3].copy()
        pts = scene.subset(2000)[obj_ids]
        crops, K_crop, boxes_rend, boxes_crop = O.crop_inputs(images, K_rows, TCO_input, tCR, pts, model.render_size, im_ids=im_ids)
        if V > 1 or model.remove_TCO_rendering:
            pts200 = np.repeat(scene.subset(200)[obj_ids], V, 0)
            Kmv = np.repeat(np.asarray(K_rows, np.float32), V, 0)
            Tmv = TCV_O.reshape(b * V, 4, 4)
            uv = O.project_points_robust(pts200, Kmv, Tmv)
            br = O.boxes_from_uv(uv)
            bc, _ = O.deepim_crops_robust(np.zeros((1, 3) + tuple(images.shape[-2:]), np.float32), br, Kmv, Tmv,
                                          tCV_R.reshape(b * V, 3), pts200, model.render_size, return_crops=False)
            KV_crop = O.get_K_crop_resize(Kmv, bc, images.shape[-2:], model.render_size).reshape(b, V, 3, 3)
            if not model.remove_TCO_rendering:
                KV_crop[:, 0] = K_crop
        else:
            KV_crop = K_crop[:, None]
        r = _render(scene, np.repeat(obj_ids, V), TCV_O.reshape(b * V, 4, 4), KV_crop.reshape(b * V, 3, 3), model.render_size,
                    model.render_normals, model.render_depth, n_threads)
        renders = r.reshape((b, V * r.shape[1]) + r.shape[2:])
        x = _normalize(model, np.concatenate([crops, renders], 1).astype(np.float32), tCR)
        pose = _net(model, x)["pose"]
        TCO_output = O.update_pose(TCO_input, K_crop, pose, tCR)
        out.append({"TCO_input": TCO_input, "TCO_output": TCO_output, "K_crop": K_crop, "KV_crop": KV_crop, "TCV_O": TCV_O, "x": x, "pose": pose})
        TCO_input = TCO_output
    return out


def run_inference_pipeline(coarse, refiner, scene: OracleScene, images, K, det_obj_ids, det_im_ids, det_boxes, so3_grid,
                           n_refiner_iterations=5, n_pose_hypotheses=1, n_threads=1):
    """One group per detection (instances are distinct rows).  Returns dict with coarse logits [B,M], the kept rows,
    refined poses, pose logits and the final pose per detection (row order = detections)."""
    det_obj_ids = np.asarray(det_obj_ids)
    det_im_ids = np.asarray(det_im_ids)
    B, M = len(det_obj_ids), len(so3_grid)
    obj_ids = np.repeat(det_obj_ids, M)
    im_ids = np.repeat(det_im_ids, M)
    K_rows = np.asarray(K, np.float32)[im_ids]
    boxes = np.repeat(np.asarray(det_boxes, np.float32), M, 0)
    R = np.tile(np.asarray(so3_grid, np.float32), (B, 1, 1))
    TCO0 = O.TCO_init_from_boxes_autodepth_with_R(boxes, scene.points[obj_ids], K_rows, R)
    c = forward_coarse(coarse, scene, images, K_rows, im_ids, obj_ids, TCO0, n_threads)
    logits = c["logits"].reshape(B, M)
    groups = np.repeat(np.arange(B), M)
    keep = O.filter_top_k(logits.reshape(-1), groups, n_pose_hypotheses)
    it = forward_refiner(refiner, scene, images, K_rows[keep], im_ids[keep], obj_ids[keep], TCO0[keep], n_refiner_iterations, n_threads)
    refined = it[-1]["TCO_output"]
    s = forward_coarse(coarse, scene, images, K_rows[keep], im_ids[keep], obj_ids[keep], refined, n_threads)
    pose_logits = s["logits"].reshape(-1)
    best = O.filter_top_k(pose_logits, groups[keep], 1)
    return {"coarse_logits": logits, "TCO_init": TCO0, "keep": keep, "iterations": it, "refined": refined,
            "pose_logits": pose_logits, "final_rows": best, "final_poses": refined[best], "final_groups": groups[keep][best]}


def add_error(points: np.ndarray, T_a: np.ndarray, T_b: np.ndarray) -> float:
    """ADD: mean distance between the model points under the two poses (metres)."""
    pa = points @ T_a[:3, :3].T + T_a[:3, 3]
    pb = points @ T_b[:3, :3].T + T_b[:3, 3]
    return float(np.linalg.norm(pa - pb, axis=1).mean())


# --------------------------------------------------------------------------------------------------------------
# CosyPose (cosypose/models/pose.py:58-199): single view, RGB only, 6-channel network input, no normalize_T
# --------------------------------------------------------------------------------------------------------------
def cosypose_forward(model, scene: OracleScene, images, K_rows, im_ids, obj_ids, TCO, n_iterations, n_threads=1):
    obj_ids = np.asarray(obj_ids)
    TCO_input = np.asarray(TCO, np.float32)
    out = []
    for _ in range(n_iterations):
        pts = scene.subset(2000)[obj_ids]
        tCR = TCO_input[:, :3, 3].copy()  # deepim_crops_robust (cosypose/lib3d/cropping.py:98-135) centres on the object origin
        crops, K_crop, boxes_rend, boxes_crop = O.crop_inputs(images[:, :3], K_rows, TCO_input, tCR, pts, model.render_size, im_ids=im_ids)
        renders = _render(scene, obj_ids, TCO_input, K_crop, model.render_size, False, False, n_threads)
        x = np.concatenate([crops, renders], 1).astype(np.float32)
        pose = _net(model, x)["pose"]
        if model.pose_dim == 9:
            dR = O.compute_rotation_matrix_from_ortho6d(pose[:, 0:6])
            v = pose[:, 6:9]
        else:
            dR = O.compute_rotation_matrix_from_quaternions(pose[:, 0:4])
            v = pose[:, 4:7]
        TCO_output = O.apply_imagespace_predictions(TCO_input, K_crop, v, dR)
        out.append({"TCO_input": TCO_input, "TCO_output": TCO_output, "K_crop": K_crop, "x": x, "pose": pose})
        TCO_input = TCO_output
    return out
