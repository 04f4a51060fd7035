"""CPU: oracle/np_oracle.py against the golden vectors produced by the real reference functions
(tests/golden/generate_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from oracle import np_oracle as O


def mesh_points(can_mesh_arrays):
    return (can_mesh_arrays["verts"].astype(np.float64) * 0.001).astype(np.float32)


def test_crop_small_matches_reference(golden, can_mesh_arrays):
    g = golden("ref_crop_small.npz")
    pts = mesh_points(can_mesh_arrays)
    ids = O.sample_point_ids(len(pts), 2000)
    assert (ids == g["point_ids"]).all()
    b = len(g["TCO"])
    points = np.tile(pts[ids][None], (b, 1, 1))
    uv = O.project_points_robust(points, g["K"], g["TCO"])
    np.testing.assert_allclose(uv, g["uv"], rtol=0, atol=2e-3)
    for C, tag in ((4, "rgbd"), (3, "rgb")):
        crops, K_crop, boxes_rend, boxes_crop = O.crop_inputs(
            g["images"][:, :C], g["K"], g["TCO"], g["tCR"], points, render_size=(60, 80), im_ids=g["im_ids"])
        np.testing.assert_allclose(boxes_rend, g["boxes_rend"], atol=2e-3)
        np.testing.assert_allclose(boxes_crop, g["boxes_crop"], atol=5e-3)
        np.testing.assert_allclose(K_crop, g["K_crop"], rtol=2e-5, atol=2e-3)
        # crops within 1e-3 absolute of the real roi_align
        ref = g[f"crops_{tag}"]
        if C == 4:
            # depth validity threshold (cropping.py:191-193) can flip on a handful of pixels when boxes move by 1e-3 px
            d = np.abs(crops[:, 3] - ref[:, 3])
            assert (d > 1e-3).mean() < 2e-3
            np.testing.assert_allclose(crops[:, :3], ref[:, :3], atol=1e-3)
        else:
            np.testing.assert_allclose(crops, ref, atol=1e-3)


def test_roi_align_exact_on_reference_boxes(golden):
    """Same boxes in -> crops must agree to float rounding (SURVEY probe 2: 1.5e-6)."""
    g = golden("ref_crop_small.npz")
    rois = np.concatenate([g["im_ids"][:, None].astype(np.float32), g["boxes_crop"]], 1)
    crops = O.roi_align(g["images"][:, :3], rois, (60, 80), 4)
    np.testing.assert_allclose(crops, g["crops_rgb"], atol=5e-6)
    crops4 = O.crop_images(g["images"], rois, (60, 80), 4)
    np.testing.assert_allclose(crops4, g["crops_rgbd"], atol=5e-6)


def test_crop_full_size_matches_reference(golden, can_mesh_arrays):
    g = golden("ref_crop_full.npz")
    pts = mesh_points(can_mesh_arrays)
    ids = O.sample_point_ids(len(pts), 2000)
    b = len(g["TCO"])
    points = np.tile(pts[ids][None], (b, 1, 1))
    image = np.random.RandomState(int(g["image_seed"])).rand(1, 3, 480, 640).astype(np.float32)
    crops, K_crop, boxes_rend, boxes_crop = O.crop_inputs(image, g["K"], g["TCO"], g["tCR"], points, (240, 320), im_ids=np.zeros(b))
    np.testing.assert_allclose(boxes_rend, g["boxes_rend"], atol=5e-3)
    np.testing.assert_allclose(boxes_crop, g["boxes_crop"], atol=1e-2)
    np.testing.assert_allclose(K_crop, g["K_crop"], rtol=3e-5, atol=5e-3)
    np.testing.assert_allclose(crops[:, :, ::5, ::5], g["crops_sub"], atol=1e-3)
    np.testing.assert_allclose(crops.astype(np.float64).sum((1, 2, 3)), g["crops_sum"], rtol=1e-5)


def test_pose_math_matches_reference(golden):
    g = golden("ref_pose.npz")
    np.testing.assert_allclose(O.compute_rotation_matrix_from_ortho6d(g["out9"][:, :6]), g["ortho6d"], atol=1e-6)
    Tn = O.normalize_T(g["TCO"])
    np.testing.assert_allclose(Tn, g["normalize_T"], atol=1e-6)
    np.testing.assert_allclose(O.normalize_T(g["TCO"].astype(np.float64)), g["normalize_T_f64"], atol=1e-12)
    dR = O.compute_rotation_matrix_from_ortho6d(g["out9"][:, :6])
    up = O.pose_update_with_reference_point(Tn, g["K_crop"], g["out9"][:, 6:9], dR, g["tCR"])
    np.testing.assert_allclose(up, g["pose_update_megapose"], atol=1e-5)  # BASELINE: pose updates within 1e-5
    np.testing.assert_allclose(O.update_pose(Tn, g["K_crop"], g["out9"], g["tCR"]), g["pose_update_megapose"], atol=1e-5)
    np.testing.assert_allclose(O.apply_imagespace_predictions(Tn, g["K_crop"], g["out9"][:, 6:9], dR), g["pose_update_cosypose6d"], atol=1e-5)
    dRq = O.compute_rotation_matrix_from_quaternions(g["out7"][:, :4])
    np.testing.assert_allclose(dRq, g["quat_R"], atol=2e-6)
    np.testing.assert_allclose(O.apply_imagespace_predictions(Tn, g["K_crop"], g["out7"][:, 4:7], dRq), g["pose_update_cosyposequat"], atol=1e-5)


def test_tco_init_matches_reference(golden, can_mesh_arrays):
    g = golden("ref_pose.npz")
    pts = mesh_points(can_mesh_arrays)
    b = len(g["boxes"])
    points = np.tile(pts[None], (b, 1, 1))
    np.testing.assert_allclose(O.TCO_init_from_boxes_autodepth_with_R(g["boxes"], points, g["K_init"], g["R_init"]), g["tco_init_autodepth_with_R"], atol=1e-5)
    np.testing.assert_allclose(O.TCO_init_from_boxes_zup_autodepth(g["boxes"], points, g["K_init"]), g["tco_init_zup_autodepth"], atol=1e-5)
    np.testing.assert_allclose(O.TCO_init_from_boxes((1.0, 1.0), g["boxes"], g["K_init"]), g["tco_init_from_boxes"], atol=1e-6)


@pytest.mark.parametrize("name", ["c1", "k5", "multi", "kbig"])
def test_topk_matches_pandas(golden, name):
    g = golden("ref_topk.npz")
    idx = O.filter_top_k(g[f"{name}_scores"], g[f"{name}_groups"], int(g[f"{name}_K"]))
    assert idx.dtype == np.int64
    assert (idx == g[f"{name}_idx"]).all()  # bit-exact indices, global descending order


def test_topk_ties_lowest_index_first():
    s = np.array([1.0, 3.0, 3.0, 2.0, 3.0, np.nan], np.float32)
    g = np.array([0, 0, 1, 1, 0, 0])
    assert O.filter_top_k(s, g, 2).tolist() == [1, 2, 4, 3]
    assert O.filter_top_k(s, g, 10).tolist() == [1, 2, 4, 3, 0, 5]
    assert O.filter_top_k(s[:0], g[:0], 3).tolist() == []


def test_meshdb_padding_and_sampling_match_reference(golden):
    g = golden("ref_meshdb.npz")
    rs = np.random.RandomState(int(g["seed"]))
    pts = [rs.rand(int(n), 3) for n in g["lens"]]
    padded = O.pad_stack_points(pts).astype(np.float32)
    np.testing.assert_array_equal(padded[:, 2400:], g["padded_tail"])
    np.testing.assert_array_equal(padded[:, O.sample_point_ids(padded.shape[1], 2000)][:, :64], g["s2000_head"])
    np.testing.assert_array_equal(padded[:, O.sample_point_ids(padded.shape[1], 200)], g["s200"])


def test_so3_grid_is_orthonormal():
    import os

    q = np.load(os.path.join(os.path.dirname(__file__), "..", "happypose_b200", "data", "so3_grid_576.npy"))
    R = O.unitquat_to_rotmat(q)
    assert R.shape == (576, 3, 3) and R.dtype == np.float32
    np.testing.assert_allclose(R @ R.transpose(0, 2, 1), np.tile(np.eye(3), (576, 1, 1)), atol=1e-5)  # the .qua rows carry 6 digits
    np.testing.assert_allclose(np.linalg.det(R), 1.0, atol=1e-5)
    # first row of data_576.qua: (0.809511, 0.106574, 0.351469, 0.458043) -> R[0,0] = 1-2(y^2+z^2)
    assert abs(R[0, 0, 0] - (1 - 2 * (0.106574**2 + 0.351469**2))) < 1e-6


def test_multiview_invariants():
    """multiview.py:28-92 restated in closed form (Panda3D absent -> parity unpinned); check what the
    construction guarantees: view 0 = TCO, every extra camera looks straight at the reference point,
    view 1 keeps the distance, views 2/3 sit sqrt(2) further on opposite sides."""
    rs = np.random.RandomState(0)
    b = 5
    TCO = np.tile(np.eye(4, dtype=np.float32), (b, 1, 1))
    q = rs.randn(b, 4)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    TCO[:, :3, :3] = O.unitquat_to_rotmat(q)
    TCO[:, :3, 3] = rs.uniform([-0.2, -0.2, 0.3], [0.2, 0.2, 1.0], (b, 3))
    tCR = TCO[:, :3, 3].copy()
    TCV_O = O.make_TCO_multiview(TCO, tCR, "TCO+front_3views", n_views=4)
    assert TCV_O.shape == (b, 4, 4, 4) and TCV_O.dtype == np.float32
    np.testing.assert_allclose(TCV_O[:, 0], TCO, atol=1e-6)
    r = np.linalg.norm(tCR, axis=1)
    for v, dist in ((1, r), (2, np.sqrt(2) * r), (3, np.sqrt(2) * r)):
        t = TCV_O[:, v, :3, 3]  # object origin (= reference point) in view v
        np.testing.assert_allclose(t[:, :2], 0, atol=1e-5)
        np.testing.assert_allclose(t[:, 2], dist, rtol=1e-5)
        R = TCV_O[:, v, :3, :3]
        np.testing.assert_allclose(R @ R.transpose(0, 2, 1), np.tile(np.eye(3), (b, 1, 1)), atol=1e-5)
    # views 2 and 3 are mirror images about view 1's optical axis: their camera centres differ
    c2 = -np.einsum("bji,bj->bi", TCV_O[:, 2, :3, :3], TCV_O[:, 2, :3, 3])
    c3 = -np.einsum("bji,bj->bi", TCV_O[:, 3, :3, :3], TCV_O[:, 3, :3, 3])
    assert (np.linalg.norm(c2 - c3, axis=1) > 1.9 * r).all()
    one = O.make_TCO_multiview(TCO, tCR, n_views=1)
    np.testing.assert_allclose(one[:, 0], TCO, atol=0)
    bad = TCO.copy()
    bad[0, 0, 0] = np.nan
    out = O.make_TCO_multiview(bad, tCR, "TCO+front_3views", n_views=4)
    assert np.isfinite(out[1:]).all()


def test_icp_input_points_match_reference(golden):
    """ref_icp.npz: getXYZ / compute_masks / the validity rules of icp_refinement, run from the reference's own source
    (tests/golden/generate_golden_icp.py)."""
    g = golden("ref_icp.npz")
    for case in (0, 1):
        dm, dr, K = g[f"depth_measured{case}"], g[f"depth_rendered{case}"], g[f"K{case}"]
        np.testing.assert_array_equal(O.icp_compute_masks(dr, dm, 0.1), g[f"mask{case}"])
        pt, ps = O.icp_input_points(dm, dr, K)
        np.testing.assert_array_equal(pt, g[f"points_tgt{case}"])
        np.testing.assert_array_equal(ps, g[f"points_src{case}"])
