"""Generates tests/golden/ref_*.npz by running the REAL reference functions (imported unchanged
from /root/reference through oracle/ref_shim.py) on seeded inputs.

Build-container only (the GPU box has no /root/reference):

    python tests/golden/generate_golden.py

The committed .npz files pin oracle/np_oracle.py (tests/test_oracle_np.py, CPU) and the CUDA path
(tests/test_gpu_*.py).  Reference versions used here: torch 2.11, torchvision 0.26 (roi_align
aligned=False semantics unchanged since the reference's pinned 0.14.1), pandas 3.0.2 (pinned 2.2.2).
"""
import os
import sys

import numpy as np
import pandas as pd
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

cam = ref_shim.ref("toolbox.lib3d.camera_geometry")
crop = ref_shim.ref("toolbox.lib3d.cropping")
cops = ref_shim.ref("toolbox.lib3d.cosypose_ops")
rot = ref_shim.ref("toolbox.lib3d.rotations")
tops = ref_shim.ref("toolbox.lib3d.transform_ops")
mops = ref_shim.ref("toolbox.lib3d.mesh_ops")
mdb = ref_shim.ref("toolbox.lib3d.rigid_mesh_database")
tcoll = ref_shim.ref("toolbox.utils.tensor_collection")
ccops = ref_shim.ref("pose_estimators.cosypose.cosypose.lib3d.cosypose_ops")

torch.set_num_threads(1)
T = torch.as_tensor


def random_rotations(rs, n):
    q = rs.randn(n, 4)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    x, y, z, w = q.T
    return np.stack(
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
         2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
         2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1).reshape(n, 3, 3)


def mesh_points():
    d = np.load(f"{HERE}/obj_000001.npz")
    pts64 = torch.tensor(d["verts"].astype(np.float64)) * 0.001  # rigid_mesh_database.py:104-106
    return pts64.float()


def gen_geometry():
    """R7: crop_inputs path (pose_rigid.py:199-277) on small frames + one full-size sub-sampled case."""
    rs = np.random.RandomState(1)
    pts_all = mesh_points()  # [9951,3]
    b = 6
    ids = mops.sample_points(pts_all.unsqueeze(0), 2000, deterministic=True)[0]
    points = ids.unsqueeze(0).repeat(b, 1, 1)
    # small frames: 3 images 4 channels (RGB-D), 120x160 -> crops 60x80
    H, W, oh, ow = 120, 160, 60, 80
    images = rs.rand(3, 4, H, W).astype(np.float32)
    images[:, 3] = np.where(rs.rand(3, H, W) < 0.15, 0.0, 0.3 + images[:, 3])  # depth with holes
    im_ids = np.array([0, 1, 2, 2, 1, 0])
    K = np.tile(np.array([[150.0, 0, 80.3], [0, 151.0, 60.7], [0, 0, 1]], np.float32), (b, 1, 1))
    K[:, 0, 0] += rs.rand(b).astype(np.float32) * 5
    TCO = np.tile(np.eye(4, dtype=np.float32), (b, 1, 1))
    TCO[:, :3, :3] = random_rotations(rs, b)
    TCO[:, :3, 3] = np.array([[0.02, -0.01, 0.45], [-0.15, 0.1, 0.6], [0.3, 0.2, 0.5], [0.0, 0.0, 0.25], [-0.4, -0.3, 0.9], [0.05, 0.02, 0.12]], np.float32)
    tCR = TCO[:, :3, 3].copy()
    tCR[2] += np.array([0.01, -0.01, 0.02], np.float32)  # reference point != object origin for one row
    imgs_exp = T(images)[im_ids]  # the reference expands the frames (pose_estimator.py:390)
    out = {}
    for C, tag in ((4, "rgbd"), (3, "rgb")):
        uv = cam.project_points_robust(points, T(K), T(TCO))
        boxes_rend = cam.boxes_from_uv(uv)
        boxes_crop, crops = crop.deepim_crops_robust(
            images=imgs_exp[:, :C].contiguous(), obs_boxes=boxes_rend, K=T(K), TCO_pred=T(TCO), tCR_in=T(tCR),
            O_vertices=points, output_size=(oh, ow), lamb=1.4)
        K_crop = cam.get_K_crop_resize(K=T(K).clone(), boxes=boxes_crop, orig_size=(H, W), crop_resize=(oh, ow))
        out[f"crops_{tag}"] = crops.numpy()
        out["uv"] = uv.numpy()
        out["boxes_rend"] = boxes_rend.numpy()
        out["boxes_crop"] = boxes_crop.numpy()
        out["K_crop"] = K_crop.numpy()
    np.savez_compressed(f"{HERE}/ref_crop_small.npz", images=images, im_ids=im_ids, K=K, TCO=TCO, tCR=tCR,
                        point_ids=np.random.RandomState(0).choice(len(pts_all), 2000, replace=False), **out)

    # full-size case (BASELINE config #1 geometry): 480x640 -> 240x320, stored sub-sampled [::5, ::5]
    rs = np.random.RandomState(2)
    b = 4
    image = rs.rand(1, 3, 480, 640).astype(np.float32)
    K = np.tile(np.array([[605.95, 0, 319.03], [0, 605.01, 249.68], [0, 0, 1]], np.float32), (b, 1, 1))
    TCO = np.tile(np.eye(4, dtype=np.float32), (b, 1, 1))
    TCO[:, :3, :3] = random_rotations(rs, b)
    TCO[:, :3, 3] = np.array([[0.1, 0.07, 0.45], [-0.2, 0.12, 0.7], [0.25, -0.18, 0.55], [0.0, 0.0, 0.3]], np.float32)
    tCR = TCO[:, :3, 3].copy()
    points = ids.unsqueeze(0).repeat(b, 1, 1)
    uv = cam.project_points_robust(points, T(K), T(TCO))
    boxes_rend = cam.boxes_from_uv(uv)
    boxes_crop, crops = crop.deepim_crops_robust(
        images=T(image).repeat(b, 1, 1, 1), obs_boxes=boxes_rend, K=T(K), TCO_pred=T(TCO), tCR_in=T(tCR),
        O_vertices=points, output_size=(240, 320), lamb=1.4)
    K_crop = cam.get_K_crop_resize(K=T(K).clone(), boxes=boxes_crop, orig_size=(480, 640), crop_resize=(240, 320))
    np.savez_compressed(f"{HERE}/ref_crop_full.npz", image_seed=2, K=K, TCO=TCO, tCR=tCR,
                        boxes_rend=boxes_rend.numpy(), boxes_crop=boxes_crop.numpy(), K_crop=K_crop.numpy(),
                        crops_sub=crops.numpy()[:, :, ::5, ::5], crops_sum=crops.double().sum((1, 2, 3)).numpy())


def gen_pose():
    """R4, R6, R13."""
    rs = np.random.RandomState(3)
    b = 64
    TCO = np.tile(np.eye(4, dtype=np.float32), (b, 1, 1))
    TCO[:, :3, :3] = random_rotations(rs, b) + rs.randn(b, 3, 3) * 1e-3  # slightly non-orthonormal
    TCO[:, :3, 3] = rs.uniform([-0.3, -0.3, 0.2], [0.3, 0.3, 1.5], (b, 3))
    TCO = TCO.astype(np.float32)
    K_crop = np.tile(np.eye(3, dtype=np.float32), (b, 1, 1))
    K_crop[:, 0, 0] = rs.uniform(400, 1500, b)
    K_crop[:, 1, 1] = rs.uniform(400, 1500, b)
    K_crop[:, 0, 2] = rs.uniform(100, 220, b)
    K_crop[:, 1, 2] = rs.uniform(80, 160, b)
    out9 = rs.randn(b, 9).astype(np.float32)
    out9[:, 6:8] *= 20
    out9[:, 8] = 1 + 0.2 * out9[:, 8]
    out7 = rs.randn(b, 7).astype(np.float32)
    out7[:, 6] = 1 + 0.2 * out7[:, 6]
    tCR = (TCO[:, :3, 3] + rs.randn(b, 3) * 0.01).astype(np.float32)
    res = {}
    res["ortho6d"] = rot.compute_rotation_matrix_from_ortho6d(T(out9[:, :6])).numpy()
    res["normalize_T"] = tops.normalize_T(T(TCO)).numpy()
    TCOn = tops.normalize_T(T(TCO))
    dR = rot.compute_rotation_matrix_from_ortho6d(T(out9[:, :6]))
    res["pose_update_megapose"] = cops.pose_update_with_reference_point(TCOn, T(K_crop), T(out9[:, 6:9]), dR, T(tCR)).numpy()
    res["pose_update_cosypose6d"] = ccops.apply_imagespace_predictions(TCOn, T(K_crop), T(out9[:, 6:9]), dR).numpy()
    dRq = rot.compute_rotation_matrix_from_quaternions(T(out7[:, :4]))
    res["quat_R"] = dRq.numpy()
    res["pose_update_cosyposequat"] = ccops.apply_imagespace_predictions(TCOn, T(K_crop), T(out7[:, 4:7]), dRq).numpy()
    # float64 inputs are legal too (tests pass f64 TCO, test_batch_renderer_panda3d.py:86-91)
    res["normalize_T_f64"] = tops.normalize_T(T(TCO.astype(np.float64))).numpy()

    # TCO init (R4)
    pts_all = mesh_points()
    bi = 24
    boxes = np.stack([rs.uniform(50, 300, bi), rs.uniform(40, 200, bi)], 1)
    boxes = np.concatenate([boxes, boxes + rs.uniform(30, 250, (bi, 2))], 1).astype(np.float32)
    boxes[0] = [384, 234, 522, 455]  # barbecue-sauce bbox of the reference example
    Ki = np.tile(np.array([[605.95, 0, 319.03], [0, 605.01, 249.68], [0, 0, 1]], np.float32), (bi, 1, 1))
    Ri = random_rotations(rs, bi).astype(np.float32)
    points = pts_all.unsqueeze(0).repeat(bi, 1, 1)
    res["tco_init_autodepth_with_R"] = cops.TCO_init_from_boxes_autodepth_with_R(T(boxes), points, T(Ki), T(Ri)).numpy()
    res["tco_init_zup_autodepth"] = cops.TCO_init_from_boxes_zup_autodepth(T(boxes), points, T(Ki)).numpy()
    res["tco_init_from_boxes"] = cops.TCO_init_from_boxes((1.0, 1.0), T(boxes), T(Ki)).numpy()
    np.savez_compressed(f"{HERE}/ref_pose.npz", TCO=TCO, K_crop=K_crop, out9=out9, out7=out7, tCR=tCR,
                        boxes=boxes, K_init=Ki, R_init=Ri, **res)


def gen_topk():
    """R16: filter_top_pose_estimates (tensor_collection.py:201-230)."""
    rs = np.random.RandomState(4)
    cases = {}
    for name, (n_det, M, K) in {"c1": (1, 576, 1), "k5": (3, 72, 5), "multi": (7, 40, 2), "kbig": (2, 4, 9)}.items():
        rows = []
        for d in range(n_det):
            for m in range(M):
                rows.append({"batch_im_id": d % 2, "label": f"obj_{d % 3}", "instance_id": d // 3, "hypothesis_id": m})
        df = pd.DataFrame(rows)
        scores = rs.randn(len(df)).astype(np.float32)  # float32 logits written into a float column
        df["coarse_logit"] = scores
        coll = tcoll.PandasTensorCollection(df, poses=torch.arange(len(df)).float())
        filt = tcoll.filter_top_pose_estimates(coll, top_K=K, group_cols=["batch_im_id", "label", "instance_id"], filter_field="coarse_logit")
        cases[f"{name}_scores"] = scores
        gid = df.groupby(["batch_im_id", "label", "instance_id"], sort=False).ngroup().values
        cases[f"{name}_groups"] = gid.astype(np.int32)
        cases[f"{name}_K"] = K
        cases[f"{name}_idx"] = filt.poses.numpy().astype(np.int64)
    np.savez_compressed(f"{HERE}/ref_topk.npz", **cases)


def gen_meshdb():
    """R3: pad_stack_tensors + sample_points (rigid_mesh_database.py:172-200, mesh_ops.py:74-84)."""
    rs = np.random.RandomState(5)
    lens = [2500, 3100, 2800]
    tensors = [torch.tensor(rs.rand(n, 3)) for n in lens]
    padded = mdb.pad_stack_tensors(tensors, fill="select_random", deterministic=True).float()
    s2000 = mops.sample_points(padded, 2000, deterministic=True)
    s200 = mops.sample_points(padded, 200, deterministic=True)
    np.savez_compressed(f"{HERE}/ref_meshdb.npz", lens=np.asarray(lens), seed=5, padded_tail=padded.numpy()[:, 2400:],
                        s2000_head=s2000.numpy()[:, :64], s200=s200.numpy())


if __name__ == "__main__":
    gen_geometry()
    gen_pose()
    gen_topk()
    gen_meshdb()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
