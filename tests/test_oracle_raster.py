"""CPU: structural assertions of the reference's renderer tests, re-asserted on the rasteriser oracle
(tests/test_batch_renderer_panda3d.py:71-242, tests/test_scene_renderer_panda3d.py:206-214).
Panda3D cannot run here, so these -- not pixel goldens -- are what the reference itself pins."""
import numpy as np
import pytest

from oracle import raster
from tests.scenes import icosphere, reference_test_scene


@pytest.fixture(scope="module")
def can(can_mesh_arrays):
    d = can_mesh_arrays
    return raster.OracleMesh(d["verts"], d["faces"], d["normals"], d["uv"], texture=d["texture"], scale=0.001)


def test_reference_scene_structure(can):
    T, K, res = reference_test_scene()
    Nc = 4
    out = raster.render([can], [0] * Nc, np.stack([T] * Nc), np.stack([K] * Nc), res,
                        render_normals=True, render_depth=True, render_binary_mask=True)
    rgb, nrm, dep, msk = out["rgb"], out["normals"], out["depth"], out["mask"]
    assert rgb.shape == (Nc, 3, 480, 640) and nrm.shape == (Nc, 3, 480, 640)
    assert dep.shape == (Nc, 1, 480, 640) and msk.shape == (Nc, 1, 480, 640)
    assert rgb.dtype == np.float32 and nrm.dtype == np.float32 and dep.dtype == np.float32 and msk.dtype == bool
    for a in (rgb, nrm, dep, msk):
        assert (a[0] == a[1]).all()  # identical cameras -> identical renders (:116-122)
    assert (rgb[0, :, 0, 0] == 0).all() and (rgb[0, :, 240, 320] > 0).any()
    assert dep[0, 0, 0, 0] == 0 and 0 < dep[0, 0, 240, 320] < 0.3
    assert (nrm[0, :, 0, 0] == 0).all() and (nrm[0, :, 240, 320] > 0).any()
    assert not msk[0, 0, 0, 0] and msk[0, 0, 240, 320]
    # values are 8-bit quantised (uint8 framebuffer / 255, panda3d_batch_renderer.py:249)
    assert np.abs(rgb * 255 - np.round(rgb * 255)).max() < 1e-4
    assert nrm.max() <= 247 / 255 + 1e-6  # texel values stop at floor(31*255/32)
    # can: radius ~51 mm at z=0.3 -> depth of the front surface ~0.249
    assert abs(dep[0, 0, 240, 320] - (0.3 - 0.0512)) < 2e-3
    assert (msk == (dep > 0)).all()


def test_flag_combinations_return_none(can):
    T, K, res = reference_test_scene()
    for n, d in ((False, False), (True, False), (False, True), (True, True)):
        out = raster.render([can], [0], T[None], K[None], (120, 160), render_normals=n, render_depth=d)
        assert out["rgb"] is not None
        assert (out["normals"] is not None) == n and (out["depth"] is not None) == d and out["mask"] is None


def test_mask_requires_depth(can):
    T, K, res = reference_test_scene()
    with pytest.raises(AssertionError):
        raster.render([can], [0], T[None], K[None], res, render_binary_mask=True)


def test_non_finite_pose_gives_zero_images(can):
    T, K, res = reference_test_scene()
    Tb = np.stack([T, T])
    Tb[1, 0, 3] = np.nan
    K = K.copy()
    K[:2] /= 4
    out = raster.render([can], [0, 0], Tb, np.stack([K, K]), (120, 160), render_normals=True, render_depth=True)
    assert out["rgb"][0].max() > 0
    assert out["rgb"][1].max() == 0 and out["normals"][1].max() == 0 and out["depth"][1].max() == 0


def test_sphere_depth_is_analytic():
    """Untextured icosphere: depth at the centre = z - r (vertex on the axis), silhouette radius ~ f*r/sqrt(z^2-r^2)."""
    v, f, n = icosphere(3, 0.05)
    m = raster.OracleMesh(v, f, n)
    T = np.eye(4)
    T[2, 3] = 0.5
    K = np.array([[400.0, 0, 80.5], [0, 400, 60.5], [0, 0, 1]])  # pixel (60,80) samples the optical axis
    out = raster.render([m], [0], T[None], K[None], (120, 160), render_depth=True, render_binary_mask=True, render_normals=True)
    dep, msk = out["depth"][0, 0], out["mask"][0, 0]
    assert abs(dep[60, 80] - 0.45) < 5e-4
    area = msk.sum()
    r_pix = 400 * 0.05 / np.sqrt(0.5**2 - 0.05**2)
    assert abs(area - np.pi * r_pix**2) / (np.pi * r_pix**2) < 0.03
    assert (out["rgb"][0][:, msk] == 1.0).all()  # no texture, no vertex colour -> white * ambient 1
    # normal facing the camera: n_cv = (0,0,-1) -> Panda axes (0,-1,0) -> frac = (0,0,0) -> every channel is the
    # half-way blend of texel 31 (247) and texel 0 (0) across the repeat seam = 124
    c = out["normals"][0][:, 60, 80] * 255
    assert (np.abs(c - 124) <= 1).all()


def test_far_and_near_rules():
    v, f, n = icosphere(2, 0.05)
    m = raster.OracleMesh(v, f, n)
    K = np.array([[400.0, 0, 80], [0, 400, 60], [0, 0, 1]])
    T = np.tile(np.eye(4), (3, 1, 1))
    T[0, 2, 3] = 9.5   # beyond d > 0.999 (z > ~9.17 m): drawn in rgb, depth/mask 0 (renderer/utils.py:56-59)
    T[1, 2, 3] = 10.5  # beyond the far plane: clipped everywhere
    T[2, 2, 3] = 0.12  # straddles the near plane: triangles with a vertex closer than 0.1 m are dropped
    K3 = np.stack([K] * 3)
    K3[0, 0, 0] = K3[0, 1, 1] = K3[1, 0, 0] = K3[1, 1, 1] = 8000
    out = raster.render([m], [0, 0, 0], T, K3, (120, 160), render_depth=True, render_binary_mask=True)
    assert out["rgb"][0].max() == 1.0 and out["depth"][0].max() == 0 and not out["mask"][0].any()
    assert out["rgb"][1].max() == 0
    assert out["depth"][2].max() > 0 and out["depth"][2][out["depth"][2] > 0].min() >= 0.1


def test_closed_surface_analysis():
    """Back faces are skipped only on meshes proven closed (welded by position) and consistently oriented."""
    v, f, _ = icosphere(2, 0.05)
    assert raster.closed_surface_sign(v, f) == -1          # outward winding
    assert raster.closed_surface_sign(v, f[:, ::-1]) == 1  # inside-out
    assert raster.closed_surface_sign(v, f[:-1]) == 0      # one face missing: open
    g = f.copy()
    g[0] = g[0, ::-1]
    assert raster.closed_surface_sign(v, g) == 0           # one face flipped: inconsistent
    quad_v = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], np.float32)
    quad_f = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    assert raster.closed_surface_sign(quad_v, quad_f) == 0
    # a two-sided sheet is a closed chain of zero volume: no orientation, no culling
    assert raster.closed_surface_sign(quad_v, np.concatenate([quad_f, quad_f[:, ::-1]])) == 0
    # texture seams duplicate vertices: welding by position keeps the surface closed
    v2 = np.concatenate([v, v[:5]])
    f2 = f.copy()
    f2[f2 == 3] = len(v) + 3
    assert raster.closed_surface_sign(v2, f2) == -1
    # two spheres, one inside-out: mixed orientation -> no culling
    f3 = np.concatenate([f, f[:, ::-1] + len(v)])
    assert raster.closed_surface_sign(np.concatenate([v, v + 0.2]), f3) == 0


def test_backface_skipping_is_invisible(can, can_mesh_arrays):
    """The can is closed (seams welded): rendering with back faces skipped == rendering every triangle two-sided."""
    d = can_mesh_arrays
    assert can.closed_sign == -1 and can.cull_sign == -1
    two_sided = raster.OracleMesh(d["verts"], d["faces"], d["normals"], d["uv"], texture=d["texture"], scale=0.001, cull=False)
    assert two_sided.cull_sign == 0
    rs = np.random.RandomState(11)
    from tests.scenes import random_crop_scene

    T, K = random_crop_scene(rs, 16)
    T[0, 2, 3] = 0.12  # the near plane (0.1 m) cuts the can open: culling must switch itself off for this scene
    a = raster.render([can], [0] * 16, T, K, (240, 320), render_normals=True, render_depth=True, n_threads=8)
    b = raster.render([two_sided], [0] * 16, T, K, (240, 320), render_normals=True, render_depth=True, n_threads=8)
    assert (a["depth"][0] > 0).any()
    for k in ("rgb", "normals", "depth"):
        assert (a[k] != b[k]).mean() <= 1e-4, k  # silhouette depth ties may pick the other face of a shared edge
    assert (a["depth"][0] == b["depth"][0]).all() and (a["rgb"][0] == b["rgb"][0]).all()


def test_point_light_rig_lambert_shading(can_mesh_arrays):
    """The render_normals=False light rig (1 ambient 0.1 + 6 point lights 0.4 at +-10 bounding radii,
    panda3d_scene_renderer.py:105-141) as per-pixel Lambert shading: same geometry as the ambient render, colour =
    albedo * min(1, 0.1 + 0.4 * sum max(0, n.l)); black lights change nothing; a single light from the camera side lights
    the centre of the can more than its limbs."""
    d = can_mesh_arrays
    mesh = raster.OracleMesh(d["verts"], d["faces"], d["normals"], d["uv"], texture=d["texture"], scale=0.001)
    T, K, (H, W) = reference_test_scene()
    T32, K32 = T[None].astype(np.float32), K[None].astype(np.float32)
    amb1 = raster.render([mesh], np.zeros(1, int), T32, K32, (H, W), render_depth=True)
    radius = float(np.linalg.norm(mesh.pos - 0.5 * (mesh.pos.min(0) + mesh.pos.max(0)), axis=1).max())
    axes = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], np.float32)
    rig = np.zeros((1, 6, 8), np.float32)
    rig[0, :, 1:4] = axes * radius * 10
    rig[0, :, 4:7] = 0.4
    lit = raster.render([mesh], np.zeros(1, int), T32, K32, (H, W), ambient=np.full((1, 3), 0.1, np.float32), lights=rig, render_depth=True)
    assert np.array_equal(lit["depth"], amb1["depth"])
    cov = amb1["depth"][0, 0] > 0
    ratio = lit["rgb"][0][:, cov].sum(0) / np.maximum(amb1["rgb"][0][:, cov].sum(0), 1e-6)
    bright = amb1["rgb"][0][:, cov].sum(0) > 0.5
    assert 0.1 <= ratio[bright].min() and ratio[bright].max() <= 1.0 + 1e-6 and 0.3 < ratio[bright].mean() < 0.9
    dark = raster.render([mesh], np.zeros(1, int), T32, K32, (H, W), ambient=np.full((1, 3), 0.1, np.float32), lights=np.zeros((1, 2, 8), np.float32))
    only_amb = raster.render([mesh], np.zeros(1, int), T32, K32, (H, W), ambient=np.full((1, 3), 0.1, np.float32))
    assert np.array_equal(dark["rgb"], only_amb["rgb"])
    # one point light at the camera (object frame position of the camera centre = -R^T t): n.l ~ 1 at the centre of the can
    cam_in_obj = -(T[:3, :3].T @ T[:3, 3])
    one = np.zeros((1, 1, 8), np.float32)
    one[0, 0, 1:4] = cam_in_obj
    one[0, 0, 4:7] = 1.0
    head = raster.render([mesh], np.zeros(1, int), T32, K32, (H, W), ambient=np.zeros((1, 3), np.float32), lights=one)["rgb"][0].sum(0)
    full = amb1["rgb"][0].sum(0)
    ys, xs = np.nonzero(cov)
    cy, cx = int(ys.mean()), int(xs.mean())
    centre = head[cy - 5:cy + 5, cx - 5:cx + 5].sum() / full[cy - 5:cy + 5, cx - 5:cx + 5].sum()
    limb = head[cy - 5:cy + 5, xs.min() + 1:xs.min() + 6].sum() / max(full[cy - 5:cy + 5, xs.min() + 1:xs.min() + 6].sum(), 1e-6)
    assert centre > 0.9 and limb < 0.6
