"""TEST INFRASTRUCTURE ONLY -- numpy restatement ("oracle") of the reference's torch math on the
MegaPose / CosyPose render-and-compare path.  Never imported by happypose_b200/.

Each function cites the reference file:line it follows (paths relative to /root/reference/happypose).
The restatement is pinned against the *real* reference functions: tests/golden/generate_golden.py
imports them unchanged through oracle/ref_shim.py, runs them on seeded inputs and commits the
outputs as tests/golden/ref_*.npz; tests/test_oracle_np.py checks this file against those vectors.
(The rasteriser oracle is separate: raster_oracle.c, parity unpinned.)
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------------------------
# toolbox/lib3d/camera_geometry.py
# ----------------------------------------------------------------------------------------------
def project_points_robust(points_3d, K, TCO, z_min=0.1):
    """camera_geometry.py:40-56: P = K @ TCO[:3]; z clamped to >= z_min before the divide."""
    points_3d = np.asarray(points_3d, F32)
    K = np.asarray(K, F32)
    TCO = np.asarray(TCO, F32)
    b, n = points_3d.shape[:2]
    ph = np.concatenate([points_3d, np.ones((b, n, 1), F32)], -1)
    P = (K @ TCO[:, :3]).astype(F32)
    suv = np.einsum("bij,bnj->bni", P, ph).astype(F32)
    z = np.maximum(F32(z_min), suv[..., 2])
    return (suv[..., :2] / z[..., None]).astype(F32)


def project_points(points_3d, K, TCO):
    """camera_geometry.py:24-37 (no clamp)."""
    points_3d = np.asarray(points_3d, F32)
    b, n = points_3d.shape[:2]
    ph = np.concatenate([points_3d, np.ones((b, n, 1), F32)], -1)
    P = (np.asarray(K, F32) @ np.asarray(TCO, F32)[:, :3]).astype(F32)
    suv = np.einsum("bij,bnj->bni", P, ph).astype(F32)
    return (suv[..., :2] / suv[..., 2:3]).astype(F32)


def boxes_from_uv(uv):
    """camera_geometry.py:59-67: (x1,y1,x2,y2) = min/max over the point axis."""
    return np.concatenate([uv.min(1), uv.max(1)], 1).astype(F32)


def get_K_crop_resize(K, boxes, orig_size, crop_resize):
    """camera_geometry.py:70-122 (skew not handled; orig_size unused by the maths)."""
    K = np.asarray(K, F32)
    boxes = np.asarray(boxes, F32)
    new_K = K.copy()
    final_width, final_height = F32(max(crop_resize)), F32(min(crop_resize))
    crop_width = boxes[:, 2] - boxes[:, 0]
    crop_height = boxes[:, 3] - boxes[:, 1]
    crop_cj = (boxes[:, 0] + boxes[:, 2]) / F32(2)
    crop_ci = (boxes[:, 1] + boxes[:, 3]) / F32(2)
    cx = K[:, 0, 2] + (crop_width - F32(1)) / F32(2) - crop_cj
    cy = K[:, 1, 2] + (crop_height - F32(1)) / F32(2) - crop_ci
    center_x = (crop_width - F32(1)) / F32(2)
    center_y = (crop_height - F32(1)) / F32(2)
    orig_cx_diff = cx - center_x
    orig_cy_diff = cy - center_y
    scale_x = final_width / crop_width
    scale_y = final_height / crop_height
    scaled_center_x = (final_width - F32(1)) / F32(2)
    scaled_center_y = (final_height - F32(1)) / F32(2)
    new_K[:, 0, 0] = scale_x * K[:, 0, 0]
    new_K[:, 1, 1] = scale_y * K[:, 1, 1]
    new_K[:, 0, 2] = scaled_center_x + scale_x * orig_cx_diff
    new_K[:, 1, 2] = scaled_center_y + scale_y * orig_cy_diff
    return new_K.astype(F32)


# ----------------------------------------------------------------------------------------------
# toolbox/lib3d/cropping.py  (+ torchvision.ops.roi_align, aligned=False, sampling_ratio=4)
# ----------------------------------------------------------------------------------------------
def deepim_boxes(rend_center_uv, obs_boxes, rend_boxes, lamb=1.4, im_size=(240, 320)):
    """cropping.py:27-75.  rend_center_uv [b,1,2]; boxes (x1,y1,x2,y2); never clamped (:71-74)."""
    obs_boxes = np.asarray(obs_boxes, F32)
    rend_boxes = np.asarray(rend_boxes, F32)
    xc = np.asarray(rend_center_uv, F32)[:, 0, 0]
    yc = np.asarray(rend_center_uv, F32)[:, 0, 1]
    w, h = max(im_size), min(im_size)
    r = F32(w / h)
    xdist = np.max(np.abs(np.stack([obs_boxes[:, 0] - xc, rend_boxes[:, 0] - xc, obs_boxes[:, 2] - xc, rend_boxes[:, 2] - xc], 1)), 1)
    ydist = np.max(np.abs(np.stack([obs_boxes[:, 1] - yc, rend_boxes[:, 1] - yc, obs_boxes[:, 3] - yc, rend_boxes[:, 3] - yc], 1)), 1)
    width = (np.maximum(xdist, ydist * r) * F32(2) * F32(lamb)).astype(F32)
    height = (np.maximum(xdist / r, ydist) * F32(2) * F32(lamb)).astype(F32)
    return np.stack([xc - width / F32(2), yc - height / F32(2), xc + width / F32(2), yc + height / F32(2)], 1).astype(F32)


def roi_align(images, rois, output_size, sampling_ratio=4):
    """torchvision.ops.roi_align(images, rois[K,5], output_size, spatial_scale=1, sampling_ratio, aligned=False).

    torchvision 0.14.1 (pyproject.toml:59-66) csrc/ops/cpu/roi_align_kernel.cpp, restated: roi_w/h floored at 1,
    bin = roi/out, sample (y,x) = start + p*bin + (i+.5)*bin/S; a sample with y<-1 | y>H | x<-1 | x>W adds 0,
    otherwise clamp to [0,H-1]x[0,W-1] and bilinear; mean over S*S samples.  Called from cropping.py:167,174,187.
    """
    images = np.asarray(images, F32)
    rois = np.asarray(rois, F32)
    _, C, H, W = images.shape
    oh, ow = output_size
    S = int(sampling_ratio)
    out = np.zeros((len(rois), C, oh, ow), F32)
    for k, roi in enumerate(rois):
        img = images[int(roi[0])]
        x1, y1, x2, y2 = roi[1:5]
        roi_w = max(F32(x2 - x1), F32(1.0))
        roi_h = max(F32(y2 - y1), F32(1.0))
        bin_h, bin_w = F32(roi_h / F32(oh)), F32(roi_w / F32(ow))
        ys = (y1 + np.arange(oh, dtype=F32)[:, None] * bin_h + (np.arange(S, dtype=F32)[None, :] + F32(0.5)) * bin_h / F32(S)).reshape(-1).astype(F32)
        xs = (x1 + np.arange(ow, dtype=F32)[:, None] * bin_w + (np.arange(S, dtype=F32)[None, :] + F32(0.5)) * bin_w / F32(S)).reshape(-1).astype(F32)

        def prep(c, n):
            invalid = (c < -1.0) | (c > n)
            c = np.maximum(c, F32(0))
            lo = np.floor(c).astype(np.int64)
            hi = lo + 1
            top = lo >= n - 1
            lo = np.where(top, n - 1, lo)
            hi = np.where(top, n - 1, hi)
            c = np.where(top, lo.astype(F32), c)
            l = (c - lo.astype(F32)).astype(F32)
            return invalid, lo, hi, l, (F32(1) - l).astype(F32)

        iy, ylo, yhi, ly, hy = prep(ys, H)
        ix, xlo, xhi, lx, hx = prep(xs, W)
        v = (
            img[:, ylo[:, None], xlo[None, :]] * (hy[:, None] * hx[None, :])
            + img[:, ylo[:, None], xhi[None, :]] * (hy[:, None] * lx[None, :])
            + img[:, yhi[:, None], xlo[None, :]] * (ly[:, None] * hx[None, :])
            + img[:, yhi[:, None], xhi[None, :]] * (ly[:, None] * lx[None, :])
        ).astype(F32)
        v = np.where((iy[:, None] | ix[None, :])[None], F32(0), v)
        out[k] = v.reshape(C, oh, S, ow, S).sum((2, 4), dtype=F32) / F32(S * S)
    return out


def crop_images(images, boxes5, output_size, sampling_ratio=4):
    """cropping.py:155-197: RGB -> roi_align; RGB-D -> also roi_align a validity map (depth>0) and zero the
    cropped depth where validity < 0.99 (DEPTH_DIMS = [3])."""
    images = np.asarray(images, F32)
    crops = roi_align(images, boxes5, output_size, sampling_ratio)
    if images.shape[1] == 4:
        valid = (images[:, 3:4] > 0).astype(F32)
        vc = roi_align(valid, boxes5, output_size, 4)
        crops[:, 3:4] *= (vc >= F32(0.99)).astype(F32)
    return crops


def deepim_crops_robust(images, obs_boxes, K, TCO_pred, tCR_in, O_vertices, output_size, lamb=1.4, return_crops=True, im_ids=None):
    """cropping.py:113-152.  `im_ids` (extension) lets a caller index one shared frame instead of the
    reference's pre-expanded images[batch_im_ids] (pose_estimator.py:390); default = arange(b)."""
    images = np.asarray(images, F32)
    h, w = images.shape[-2:]
    b = len(TCO_pred)
    uv = project_points_robust(O_vertices, K, TCO_pred)
    rend_boxes = boxes_from_uv(uv)
    TCR = np.asarray(TCO_pred, F32).copy()
    TCR[:, :3, 3] = tCR_in
    center = project_points_robust(np.zeros((b, 1, 3), F32), K, TCR)
    boxes = deepim_boxes(center, obs_boxes, rend_boxes, lamb=lamb, im_size=(h, w))
    crops = None
    if return_crops:
        ids = np.arange(b, dtype=F32) if im_ids is None else np.asarray(im_ids, F32)
        crops = crop_images(images, np.concatenate([ids[:, None], boxes], 1), output_size, 4)
    return boxes, crops


def crop_inputs(images, K, TCO, tCR, points, render_size=(240, 320), im_ids=None):
    """megapose/models/pose_rigid.py:199-277 -> (images_cropped, K_crop, boxes_rend, boxes_crop).
    `points` = mesh_db.select(labels).sample_points(2000, deterministic=True)."""
    uv = project_points_robust(points, K, TCO)
    boxes_rend = boxes_from_uv(uv)
    boxes_crop, crops = deepim_crops_robust(images, boxes_rend, K, TCO, tCR, points, render_size, 1.4, True, im_ids)
    K_crop = get_K_crop_resize(K, boxes_crop, np.asarray(images).shape[-2:], render_size)
    return crops, K_crop, boxes_rend, boxes_crop


# ----------------------------------------------------------------------------------------------
# toolbox/lib3d/rotations.py, transform_ops.py, cosypose_ops.py
# ----------------------------------------------------------------------------------------------
def compute_rotation_matrix_from_ortho6d(poses):
    """rotations.py:22-36: x = a/|a|; z = (x X b)/|x X b|; y = z X x; columns (x,y,z)."""
    poses = np.asarray(poses)
    x_raw, y_raw = poses[..., 0:3], poses[..., 3:6]
    x = x_raw / np.linalg.norm(x_raw, axis=-1, keepdims=True)
    z = np.cross(x, y_raw)
    z = z / np.linalg.norm(z, axis=-1, keepdims=True)
    y = np.cross(z, x)
    return np.stack([x, y, z], -1).astype(poses.dtype)


def normalize_T(T):
    """transform_ops.py:107-120: rebuild T from columns 0,1 of R (ortho6d) and t."""
    T = np.asarray(T)
    pose9 = np.concatenate([T[..., :3, 0], T[..., :3, 1], T[..., :3, 3]], -1)
    out = np.zeros(T.shape, T.dtype)
    out[..., :3, :3] = compute_rotation_matrix_from_ortho6d(pose9[..., :6])
    out[..., :3, 3] = pose9[..., 6:]
    out[..., 3, 3] = 1
    return out


def compute_rotation_matrix_from_quaternions(quats):
    """rotations.py:186-229 (CosyPose pose_dim=7): quaternion (w,x,y,z)-normalised -> R."""
    q = np.asarray(quats)
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    # reference layout: quats[..., 0:4] = (x, y, z, w) after normalisation (see rotations.py)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.stack(
        [
            1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
            2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
            2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y),
        ],
        -1,
    ).reshape(q.shape[:-1] + (3, 3))
    return R.astype(q.dtype)


def pose_update_with_reference_point(TCO, K, vxvyvz, dRCO, tCR):
    """cosypose_ops.py:34-62 (MegaPose)."""
    TCO, K, vxvyvz, dRCO, tCR = (np.asarray(a) for a in (TCO, K, vxvyvz, dRCO, tCR))
    zsrc = tCR[:, 2:3]
    ztgt = vxvyvz[:, 2:3] * zsrc
    fxfy = np.stack([K[:, 0, 0], K[:, 1, 1]], 1)
    tCR_out = tCR.copy()
    tCR_out[:, 2] = ztgt[:, 0]
    tCR_out[:, :2] = ((vxvyvz[:, :2] / fxfy) + (tCR[:, :2] / zsrc)) * ztgt
    tCO_out = (dRCO @ (TCO[:, :3, 3] - tCR)[..., None])[..., 0] + tCR_out
    out = TCO.copy()
    out[:, :3, 3] = tCO_out
    out[:, :3, :3] = dRCO @ TCO[:, :3, :3]
    return out


def apply_imagespace_predictions(TCO, K, vxvyvz, dRCO):
    """cosypose/lib3d/cosypose_ops.py:18-42 (CosyPose; reference point = object origin)."""
    TCO, K, vxvyvz, dRCO = (np.asarray(a) for a in (TCO, K, vxvyvz, dRCO))
    out = TCO.copy()
    zsrc = TCO[:, 2, 3:4]
    ztgt = vxvyvz[:, 2:3] * zsrc
    fxfy = np.stack([K[:, 0, 0], K[:, 1, 1]], 1)
    out[:, 2, 3] = ztgt[:, 0]
    out[:, :2, 3] = ((vxvyvz[:, :2] / fxfy) + (TCO[:, :2, 3] / zsrc)) * ztgt
    out[:, :3, :3] = dRCO @ TCO[:, :3, :3]
    return out


def update_pose(TCO, K_crop, pose_outputs, tCR):
    """megapose/models/pose_rigid.py:339-350."""
    dR = compute_rotation_matrix_from_ortho6d(np.asarray(pose_outputs)[:, 0:6])
    return pose_update_with_reference_point(TCO, K_crop, np.asarray(pose_outputs)[:, 6:9], dR, tCR)


def transform_pts(T, pts):
    """transform_ops.py:28-55 (3-D T only)."""
    T, pts = np.asarray(T), np.asarray(pts)
    return (np.einsum("bij,bnj->bni", T[:, :3, :3], pts) + T[:, None, :3, 3]).astype(pts.dtype)


def TCO_init_from_boxes_autodepth_with_R(boxes_2d, model_points_3d, K, R):
    """cosypose_ops.py:184-238."""
    boxes_2d, K, R = np.asarray(boxes_2d, F32), np.asarray(K, F32), np.asarray(R, F32)
    pts = np.asarray(model_points_3d, F32)
    b = len(boxes_2d)
    fxfy = np.stack([K[:, 0, 0], K[:, 1, 1]], 1)
    cxcy = np.stack([K[:, 0, 2], K[:, 1, 2]], 1)
    TCO = np.tile(np.array([[0, 1, 0, 0], [0, 0, -1, 0], [-1, 0, 0, 1], [0, 0, 0, 1]], F32), (b, 1, 1))
    TCO[:, :3, :3] = R
    centers = (boxes_2d[:, [0, 1]] + boxes_2d[:, [2, 3]]) / F32(2)
    TCO[:, :2, 3] = ((centers - cxcy) * F32(1.0)) / fxfy
    C = transform_pts(TCO, pts)
    dx = C[:, :, 0].max(1) - C[:, :, 0].min(1)
    dy = C[:, :, 1].max(1) - C[:, :, 1].min(1)
    bb_dx = (boxes_2d[:, 2] - boxes_2d[:, 0]) + F32(1)
    bb_dy = (boxes_2d[:, 3] - boxes_2d[:, 1]) + F32(1)
    z = ((fxfy[:, 1] * dy / bb_dy) + (fxfy[:, 0] * dx / bb_dx)) / F32(2)
    TCO[:, :2, 3] = ((centers - cxcy) * z[:, None]) / fxfy
    TCO[:, 2, 3] = z
    return TCO


def TCO_init_from_boxes_zup_autodepth(boxes_2d, model_points_3d, K):
    """cosypose_ops.py:241-283: same with the fixed z-up rotation [[0,1,0],[0,0,-1],[-1,0,0]]."""
    R = np.tile(np.array([[0, 1, 0], [0, 0, -1], [-1, 0, 0]], F32), (len(boxes_2d), 1, 1))
    return TCO_init_from_boxes_autodepth_with_R(boxes_2d, model_points_3d, K, R)


def TCO_init_from_boxes(z_range, boxes, K):
    """cosypose_ops.py:159-181: identity rotation, z = mean(z_range), xy from the bbox centre."""
    boxes, K = np.asarray(boxes, F32), np.asarray(K, F32)
    b = len(boxes)
    z = F32(np.asarray(z_range, F32).mean())
    centers = (boxes[:, [0, 1]] + boxes[:, [2, 3]]) / F32(2)
    fxfy = np.stack([K[:, 0, 0], K[:, 1, 1]], 1)
    cxcy = np.stack([K[:, 0, 2], K[:, 1, 2]], 1)
    TCO = np.tile(np.eye(4, dtype=F32), (b, 1, 1))
    TCO[:, :2, 3] = ((centers - cxcy) * z) / fxfy
    TCO[:, 2, 3] = z
    return TCO


# ----------------------------------------------------------------------------------------------
# toolbox/utils/transform_utils.py (roma.unitquat_to_rotmat, roma 1.5.0)
# ----------------------------------------------------------------------------------------------
def unitquat_to_rotmat(quat_xyzw):
    """roma.unitquat_to_rotmat: (x,y,z,w) unit quaternion -> R; transform_utils.py:46-47 feeds it a
    float32 tensor built from the text file."""
    q = np.asarray(quat_xyzw, F32)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    tx, ty, tz = F32(2) * x, F32(2) * y, F32(2) * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    R = np.stack(
        [
            F32(1) - (tyy + tzz), txy - twz, txz + twy,
            txy + twz, F32(1) - (txx + tzz), tyz - twx,
            txz - twy, tyz + twx, F32(1) - (txx + tyy),
        ],
        -1,
    ).reshape(q.shape[:-1] + (3, 3))
    return R.astype(F32)


# ----------------------------------------------------------------------------------------------
# toolbox/lib3d/multiview.py  (Panda3D NodePath.lookAt restated -- parity unpinned, Panda3D absent)
# ----------------------------------------------------------------------------------------------
_TCCGL = np.array([[1, 0, 0, 0], [0, 0, -1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], np.float64)


def _look_at(pos, target, up):
    """Panda3D look_at(): +Y forward exactly at the target, +Z as close to `up` as possible, +X = Y x Z.
    Returns the node's rotation with the node axes as COLUMNS (node -> parent)."""
    fwd = target - pos
    fwd = fwd / np.linalg.norm(fwd)
    right = np.cross(fwd, up)
    right = right / np.linalg.norm(right)
    upv = np.cross(right, fwd)
    return np.stack([right, fwd, upv], 1)


def _views_TC0_CV(TCO, tCR, cam_positions_wrt_cam0):
    """multiview.py:28-92 in closed form, float64 like the reference's numpy path."""
    TCO = np.asarray(TCO, np.float64)
    tCR = np.asarray(tCR, np.float64)
    TOC = np.linalg.inv(TCO)
    if not np.isfinite(TOC).all():
        TOC = np.eye(4)
        tCR = np.zeros(3)
    T_W_C0 = TOC @ _TCCGL  # camera node (Panda axes) in the world = object frame; compute_view_mat
    ref = TOC[:3, :3] @ tCR + TOC[:3, 3]
    radius = np.linalg.norm(tCR)
    up = T_W_C0[:3, 2]  # camera's +Z (Panda up) in world (:65-66)
    c0 = T_W_C0[:3, 3]
    R_p = _look_at(c0, ref, up)  # "camera-pointing-to-ref" (:69-72)
    out = []
    for cam_pos in np.asarray(cam_positions_wrt_cam0, np.float64) * radius:
        pos = c0 + R_p @ cam_pos  # setPos(cam_pointing_to_ref, *cam_pos) (:77)
        R_n = _look_at(pos, ref, up)
        T_W_N = np.eye(4)
        T_W_N[:3, :3] = R_n
        T_W_N[:3, 3] = pos
        T_C0_N = np.linalg.inv(T_W_C0) @ T_W_N  # node.getMat(cam), transposed to column convention (:79)
        out.append(_TCCGL @ T_C0_N @ np.linalg.inv(_TCCGL))  # (:80)
    return out


_MV_POSITIONS = {
    "TCO+front_1view": [[0, 0, 0]],
    "TCO+front_3views": [[0, 0, 0], [1, 0, 0], [-1, 0, 0]],
    "sphere_26views": [
        [x, y, z] for y in [0, 1, 2] for x in [0, -1, 1] for z in [0, 1, -1] if not (x == 0 and y == 1 and z == 0)
    ],
}


def make_TCO_multiview(TCO, tCR, multiview_type="TCO+front_3views", n_views=4, remove_TCO_rendering=False,
                       views_inplane_rotations=False):
    """multiview.py:166-251: TCV_O [b,V,4,4] ([b,4V,4,4] with views_inplane_rotations, :239-250) in TCO's dtype."""
    TCO = np.asarray(TCO)
    tCR = np.asarray(tCR)
    b = len(TCO)
    if n_views == 1:
        TC0_CV = np.tile(np.eye(4), (b, 1, 1, 1))
    else:
        if multiview_type not in _MV_POSITIONS:
            raise ValueError(multiview_type)
        rows = []
        for n in range(b):
            views = [] if remove_TCO_rendering else [np.eye(4)]
            views += _views_TC0_CV(TCO[n], tCR[n], _MV_POSITIONS[multiview_type])
            rows.append(np.stack(views))
        TC0_CV = np.stack(rows)
    TC0_CV = TC0_CV.astype(TCO.dtype)
    # invert_transform_matrices (transform_ops.py:58-67): R^T, -R^T t
    inv = TC0_CV.copy()
    Rt = np.swapaxes(TC0_CV[..., :3, :3], -1, -2)
    inv[..., :3, :3] = Rt
    inv[..., :3, 3] = -(Rt @ TC0_CV[..., :3, 3:4])[..., 0]
    TCV_O = (inv @ TCO[:, None]).astype(TCO.dtype)
    if views_inplane_rotations:
        # multiview.py:239-250: every view 4x, copies 1..3 with the ROTATION part pre-multiplied by a rotation about the
        # camera z axis by 90 / 180 / 270 degrees (transforms3d.euler.euler2mat(0, 0, angle), cast to TCO's dtype); the
        # translation is left as it is
        assert remove_TCO_rendering
        TCV_O = np.repeat(TCV_O[:, :, None], 4, axis=2)
        for idx, angle in enumerate([np.pi / 2, np.pi, 3 * np.pi / 2]):
            c, s = np.cos(angle), np.sin(angle)
            dR = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]]).astype(TCO.dtype)
            TCV_O[:, :, idx + 1, :3, :3] = dR @ TCV_O[:, :, idx + 1, :3, :3]
        TCV_O = TCV_O.reshape(b, -1, 4, 4)
    return TCV_O


# ----------------------------------------------------------------------------------------------
# toolbox/utils/tensor_collection.py : filter_top_pose_estimates
# ----------------------------------------------------------------------------------------------
def filter_top_k(scores, group_ids, top_k):
    """tensor_collection.py:201-230: df.sort_values(field, ascending=False).groupby(cols).head(K).index.

    Rows come out in GLOBAL descending score order, at most K per group.  pandas' default quicksort
    leaves ties unspecified; this build defines ties as lowest original index first.
    NaN scores sort last, as pandas does (na_position='last').
    """
    scores = np.asarray(scores, np.float64)
    group_ids = np.asarray(group_ids)
    key = np.where(np.isnan(scores), -np.inf, scores)
    nan_last = np.isnan(scores)
    order = np.lexsort((np.arange(len(scores)), -key, nan_last))
    taken = {}
    out = []
    for i in order:
        g = int(group_ids[i])
        c = taken.get(g, 0)
        if c < top_k:
            taken[g] = c + 1
            out.append(int(i))
    return np.asarray(out, np.int64)


# ----------------------------------------------------------------------------------------------
# toolbox/lib3d/rigid_mesh_database.py, mesh_ops.py
# ----------------------------------------------------------------------------------------------
def pad_stack_points(points_list):
    """rigid_mesh_database.py:172-200 (fill='select_random', deterministic): one RandomState(0) shared by
    the whole list; objects shorter than N_max are padded with re-drawn own vertices."""
    n_max = max(len(p) for p in points_list)
    rs = np.random.RandomState(0)
    out = []
    for p in points_list:
        n_pad = n_max - len(p)
        if n_pad > 0:
            ids = rs.choice(np.arange(len(p)), size=n_pad)
            p = np.concatenate([p, p[ids]], 0)
        out.append(p)
    return np.stack(out)


def mesh_db_points(verts_list, scales):
    """rigid_mesh_database.py:80-127: float64 vertices * scale, padded, cast to float32."""
    pts = [np.asarray(v, np.float64) * float(s) for v, s in zip(verts_list, scales)]
    return pad_stack_points(pts).astype(F32)


def sample_point_ids(n_total, n_points):
    """mesh_ops.py:74-84 (deterministic=True): RandomState(0).choice(N, n, replace=False) -- the same
    index set on every call."""
    assert n_points <= n_total
    return np.random.RandomState(0).choice(n_total, size=n_points, replace=False)


# ----------------------------------------------------------------------------------------------
# megapose/models/pose_rigid.py : normalize_depth
# ----------------------------------------------------------------------------------------------
def normalize_depth(depth, tCR_z, kind):
    """pose_rigid.py:510-544."""
    depth = np.asarray(depth, F32)
    z = np.asarray(tCR_z, F32).reshape((-1,) + (1,) * (depth.ndim - 1))
    if kind == "tCR_scale":
        return depth / z
    if kind == "tCR_scale_clamp_center":
        return np.clip(depth / z, 0, 2) - F32(1)
    if kind == "tCR_center_clamp":
        return np.clip(depth - z, -2, 2)
    if kind == "none":
        return depth
    raise ValueError(kind)


# ----------------------------------------------------------------------------------------------
# megapose/inference/icp_refiner.py (input stage of the ICP depth refiner), refiner_utils.py
# ----------------------------------------------------------------------------------------------
def icp_get_xyz(depth, fx, fy, cx, cy):
    """icp_refiner.py:106-135 (getXYZ, whole image): the pixel offsets live in an int16 table, so (x - cx) and (y - cy)
    are TRUNCATED towards zero before use; x = u * depth * 1 / fx evaluated left to right in float32."""
    depth = np.asarray(depth, F32)
    H, W = depth.shape
    u = (np.arange(W) - np.float64(cx)).astype(np.int16)
    v = (np.arange(H) - np.float64(cy)).astype(np.int16)
    xyz = np.zeros((H, W, 3), np.float64)
    xyz[:, :, 0] = u[None, :] * depth * 1 / F32(fx)
    xyz[:, :, 1] = v[:, None] * depth * 1 / F32(fy)
    xyz[:, :, 2] = depth
    return xyz.astype(F32)


def icp_compute_masks(depth_rendered, depth_measured, depth_delta_thresh=0.1):
    """refiner_utils.py compute_masks(mask_type="threshold"): measured > 0, rendered > 0, |measured - rendered| <= thresh."""
    m = np.logical_and(depth_measured > 0, depth_rendered > 0)
    m[np.abs(depth_measured - depth_rendered) > depth_delta_thresh] = False
    return m


def icp_input_points(depth_measured, depth_rendered, K, mask=None, depth_delta_thresh=0.1):
    """icp_refinement (icp_refiner.py:138-176) up to the ICP call: (points_tgt [n_t,3], points_src [n_s,3]) in row-major
    pixel order.  Target = measured points with 0.2 < d < 5 inside the mask; source = rendered points at the same pixels
    where something was rendered."""
    dm, dr = np.asarray(depth_measured, F32), np.asarray(depth_rendered, F32)
    if mask is None:
        mask = icp_compute_masks(dr, dm, depth_delta_thresh)
    K = np.asarray(K, F32)
    valid = np.logical_and(np.logical_and(dm > 0.2, dm < 5), mask)
    pt = icp_get_xyz(dm, K[0, 0], K[1, 1], K[0, 2], K[1, 2])[valid]
    ps = icp_get_xyz(dr, K[0, 0], K[1, 1], K[0, 2], K[1, 2])[np.logical_and(valid, dr > 0)]
    return pt, ps
