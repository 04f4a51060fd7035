"""ICP depth refiner re-hosted on the B200 kernels (mirror of
happypose/pose_estimators/megapose/inference/icp_refiner.py:220-303 and depth_refiner.py).

    reference                                                         here
    ----------------------------------------------------------------  ----------------------------------------------------
    renderer.render(render_depth=True) at 480x640, N Panda3D frames   hpb_render, one launch, depth stays on the device
    per object: depth maps .cpu().numpy(), compute_masks, getXYZ of   hpb_icp_points: masks + both point clouds of all N
    both maps, boolean-mask selections (host, one object at a time)   objects in one launch, ONE device->host copy of
                                                                      the (counts, points) the registration needs
    get_normal (cv2.inpaint + Gaussian filter + gradients), host      unchanged, host (OpenCV / SciPy)
    cv2.ppf_match_3d_ICP(100, tolerence=0.05, numLevels=4), host      unchanged, host: OpenCV-contrib's registration is a
                                                                      third-party CPU library, like the reference uses it

`registration` lets a caller plug another point-to-plane ICP with the same contract
(points_src [n,6], points_tgt [m,6]) -> (retval, residual, pose 4x4); by default cv2.ppf_match_3d_ICP is used and its
absence is an ImportError at construction (no silent replacement).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch

from .. import ops
from ..renderer.types import Panda3dLightData


class DepthRefiner(ABC):
    """depth_refiner.py: refine_poses(predictions, masks [B,H,W], depth [B,H,W], K [B,3,3]) -> (refined, extra_data)."""

    @abstractmethod
    def refine_poses(self, predictions, masks=None, depth=None, K=None):
        ...


def get_normal(depth_refine: np.ndarray, fx: float, fy: float, cx: float, cy: float, refine: bool = True) -> np.ndarray:
    """Per-pixel surface normals of a depth map, icp_refiner.py:33-103 (whole-image branch): holes in-painted
    (Navier-Stokes, radius 2) and smoothed (Gaussian, sigma 2), central differences with spacing 2, cross product of the
    two tangent vectors built with the int16-truncated pixel offsets."""
    import cv2
    from scipy import ndimage

    res_y, res_x = depth_refine.shape
    constant_x, constant_y = 1 / fx, 1 / fy
    if refine:
        depth_refine = np.nan_to_num(depth_refine)
        mask = np.zeros_like(depth_refine).astype(np.uint8)
        mask[depth_refine == 0] = 1
        depth_refine = cv2.inpaint(depth_refine.astype(np.float32), mask, 2, cv2.INPAINT_NS).astype(np.float32)
        depth_refine = ndimage.gaussian_filter(depth_refine, 2)
    uv = np.zeros((res_y, res_x, 2), dtype=np.int16)
    uv[:, :, 1] = np.arange(0, res_x) - cx
    uv[:, :, 0] = np.arange(0, res_y)[:, np.newaxis] - cy
    dig = np.gradient(depth_refine, 2, edge_order=2)
    v_x = np.zeros((res_y, res_x, 3))
    v_y = np.zeros((res_y, res_x, 3))
    v_y[:, :, 0] = uv[:, :, 1] * constant_x * dig[0]
    v_y[:, :, 1] = depth_refine * constant_y + (uv[:, :, 0] * constant_y) * dig[0]
    v_y[:, :, 2] = dig[0]
    v_x[:, :, 0] = depth_refine * constant_x + uv[:, :, 1] * constant_x * dig[1]
    v_x[:, :, 1] = uv[:, :, 0] * constant_y * dig[1]
    v_x[:, :, 2] = dig[1]
    cross = np.cross(v_x.reshape(-1, 3), v_y.reshape(-1, 3))
    norm = np.expand_dims(np.linalg.norm(cross, axis=1), axis=1)
    norm[norm == 0] = 1
    return np.nan_to_num((cross / norm).reshape(res_y, res_x, 3))


def _opencv_registration() -> Callable:
    import cv2

    if not hasattr(cv2, "ppf_match_3d_ICP"):
        raise ImportError("ICPRefiner needs cv2.ppf_match_3d_ICP (opencv-contrib-python), or pass registration=<callable>")

    def register(points_src: np.ndarray, points_tgt: np.ndarray):
        icp = cv2.ppf_match_3d_ICP(100, tolerence=0.05, numLevels=4)  # icp_refiner.py:195-196
        return icp.registerModelToScene(points_src.reshape(-1, 6), points_tgt.reshape(-1, 6))

    return register


class ICPRefiner(DepthRefiner):
    def __init__(self, mesh_db, renderer, registration: Optional[Callable] = None, n_min_points: int = 1000,
                 tolerance: float = 0.05) -> None:
        self.mesh_db = mesh_db
        self.renderer = renderer
        self.light_datas = [Panda3dLightData("ambient")]
        self.registration = registration if registration is not None else _opencv_registration()
        self.n_min_points = n_min_points
        self.tolerance = tolerance

    # -- device stage ---------------------------------------------------------------------------------------------
    def icp_inputs(self, predictions, depth: torch.Tensor, K: torch.Tensor, masks: Optional[torch.Tensor] = None):
        """Everything icp_refinement needs in front of the registration, for all N estimates at once, on the device:
        (depth_rendered [N,H,W], points_tgt [N,cap,3], points_src [N,cap,3], counts [N,2], index_tgt, index_src [N,cap])."""
        df = predictions.infos
        labels = df["label"].tolist()
        dev = predictions.poses.device
        im_ids = torch.as_tensor(df["batch_im_id"].to_numpy().astype(np.int64)).to(dev)
        depth = depth.to(dev)
        if depth.dim() == 4:
            depth = depth[:, 0]
        K_rows = K.to(dev)[im_ids]
        out = self.renderer.render(labels, TCO=predictions.poses, K=K_rows, light_datas=[self.light_datas] * len(labels),
                                   resolution=tuple(depth.shape[-2:]), render_depth=True)
        depth_rendered = out.depths[:, 0]
        pt, ps, counts, it, isrc = ops.icp_points(self.renderer._ctx, depth, depth_rendered, im_ids.int(), K_rows, masks=masks,
                                                  depth_delta_thresh=0.1, return_index=True)
        return depth_rendered, pt, ps, counts, it, isrc

    # -- reference API ---------------------------------------------------------------------------------------------
    def refine_poses(self, predictions, masks: Optional[torch.Tensor] = None, depth: Optional[torch.Tensor] = None,
                     K: Optional[torch.Tensor] = None) -> Tuple[object, Dict]:
        """icp_refiner.py:232-303."""
        assert depth is not None
        assert K is not None
        predictions_refined = predictions.clone()
        if "poses_input" not in predictions_refined.tensors:
            predictions_refined.register_tensor("poses_input", predictions.poses.clone())
        N = len(predictions)
        depth_rendered, pt, ps, counts, it, isrc = self.icp_inputs(predictions, depth, K, masks)
        counts_h = counts.cpu().numpy()  # the stage's one synchronising copy; point clouds follow, trimmed to their counts
        n_max = int(counts_h.max()) if N else 0
        pt_h, ps_h = pt[:, :n_max].cpu().numpy(), ps[:, :n_max].cpu().numpy()
        it_h, is_h = it[:, :n_max].cpu().numpy(), isrc[:, :n_max].cpu().numpy()
        depth_h = depth[:, 0].cpu().numpy() if depth.dim() == 4 else depth.cpu().numpy()
        dr_h = depth_rendered.cpu().numpy()
        K_h = K.cpu().numpy()
        im_ids = predictions.infos["batch_im_id"].to_numpy()
        retvals = np.full(N, -1, np.int64)
        normals_cache: Dict[int, np.ndarray] = {}
        for n in range(N):
            predictions_refined.poses_input[n] = predictions.poses[n].clone()
            nt, ns = int(counts_h[n, 0]), int(counts_h[n, 1])
            if nt < self.n_min_points or ns < self.n_min_points:
                continue
            view, cam_K = int(im_ids[n]), K_h[int(im_ids[n])]
            kw = dict(fx=cam_K[0, 0], fy=cam_K[1, 1], cx=cam_K[0, 2], cy=cam_K[1, 2])
            if view not in normals_cache:  # the measured map's normals depend on the frame only
                normals_cache[view] = get_normal(depth_h[view], refine=True, **kw).astype(np.float32)
            n_src_map = get_normal(dr_h[n], refine=True, **kw).astype(np.float32)
            # the kernel's pixel indices pick the normals of exactly the selected points (the reference indexes with the
            # same boolean masks, :150-176)
            points_tgt = np.concatenate([pt_h[n, :nt], normals_cache[view].reshape(-1, 3)[it_h[n, :nt]]], 1).astype(np.float32)
            points_src = np.concatenate([ps_h[n, :ns], n_src_map.reshape(-1, 3)[is_h[n, :ns]]], 1).astype(np.float32)
            TCO = predictions.poses[n].cpu().numpy().copy()
            shift = np.mean(points_tgt[:, :3], axis=0) - np.mean(points_src[:, :3], axis=0)  # :188-192
            TCO[:3, -1] += shift.reshape(-1)
            points_src[:, :3] += shift[None]
            retval, residual, pose = self.registration(points_src, points_tgt)
            if residual > self.tolerance or residual < 0:
                retval = -1
            retvals[n] = retval
            if retval != -1:
                predictions_refined.poses[n] = torch.as_tensor(np.asarray(pose) @ TCO, dtype=torch.float32).to(predictions.poses.device)
        return predictions_refined, {"retval": retvals, "n_points": counts_h}
