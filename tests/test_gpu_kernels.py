"""GPU parity tests proper: every kernel called through the C ABI (happypose_b200.ops -> libhpb200.so) and compared
with the CPU oracle on the same seeded inputs, plus the golden vectors produced by the real reference functions."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from oracle import raster as oraster
from tests.scenes import icosphere, random_crop_scene, random_rotations, reference_test_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from happypose_b200._capi import Context

    return Context.get("cuda:0")


@pytest.fixture(scope="module")
def can(ctx, can_mesh_arrays):
    from happypose_b200 import ops

    d = can_mesh_arrays
    om = oraster.OracleMesh(d["verts"], d["faces"], d["normals"], d["uv"], texture=d["texture"], scale=0.001)
    mid = ops.mesh_upload(ctx, om.pos, d["faces"], d["normals"], d["uv"], texture=d["texture"])
    return om, mid


def _render_both(ctx, om, mid, T, K, res, ambient=None, **kw):
    from happypose_b200 import ops

    b = len(T)
    ref = oraster.render([om], [0] * b, T, K, res, ambient=ambient, render_normals=True, render_depth=True, render_binary_mask=True, n_threads=8, **kw)
    rgb, nrm, dep, msk = ops.render(ctx, torch.full((b,), mid), torch.as_tensor(T), torch.as_tensor(K), res,
                                    ambient=None if ambient is None else torch.as_tensor(ambient),
                                    render_normals=True, render_depth=True, render_binary_mask=True, **kw)
    torch.cuda.synchronize()
    return ref, (rgb.cpu().numpy(), nrm.cpu().numpy(), dep.cpu().numpy(), msk.cpu().numpy())


def _check_render_parity(ref, got, name=""):
    rgb, nrm, dep, msk = got
    # mask / coverage: bit-exact is expected (same integer edge functions); BASELINE bar: <= 0.5 % silhouette mismatches
    mism = (msk != ref["mask"]).mean()
    assert mism <= 0.005, f"{name}: mask mismatch {mism}"
    both = msk & ref["mask"]
    rel = np.abs(dep - ref["depth"])[both] / ref["depth"][both]
    assert rel.size > 0
    # BASELINE bar: depth within 1e-4 relative on interior pixels
    assert (rel > 1e-4).mean() <= 0.005, f"{name}: depth rel err max {rel.max()}"
    # colour / normals: 8-bit values; allow one quantisation step on <= 0.5 % of the pixels
    for a, r, tag in ((rgb, ref["rgb"], "rgb"), (nrm, ref["normals"], "normals")):
        d = np.abs(a - r) * 255
        assert (d > 0.5).mean() <= 0.005, f"{name}: {tag} differs on {(d > 0.5).mean():.4%} of values (max {d.max():.1f}/255)"
    return mism, rel.max()


def test_mip_chain_matches_oracle(ctx, can):
    from happypose_b200 import ops

    om, mid = can
    for lvl in range(len(om.tex_w)):
        got = ops.mesh_get_mip(ctx, mid, lvl)
        w, h, off = int(om.tex_w[lvl]), int(om.tex_h[lvl]), int(om.tex_off[lvl])
        ref = om.tex[off:off + w * h].reshape(h, w, 4)
        assert got.shape == ref.shape
        assert (got == ref).all(), f"mip level {lvl} differs"


def test_render_reference_scene_bit_exact(ctx, can):
    """The reference's own renderer test scene (tests/test_batch_renderer_panda3d.py:43-91), f64 TCO / K inputs."""
    om, mid = can
    T, K, res = reference_test_scene()
    ref, got = _render_both(ctx, om, mid, np.stack([T] * 4), np.stack([K] * 4), res)
    rgb, nrm, dep, msk = got
    assert rgb.shape == (4, 3, 480, 640) and dep.shape == (4, 1, 480, 640) and msk.dtype == bool
    assert (msk == ref["mask"]).all()
    assert (dep == ref["depth"]).all(), "depth must be bit-identical to the oracle"
    assert (nrm == ref["normals"]).all()
    assert (rgb == ref["rgb"]).all()
    for a in got:
        assert (a[0] == a[1]).all() and (a[0] == a[3]).all()
    assert rgb[0, :, 0, 0].max() == 0 and rgb[0, :, 240, 320].max() > 0
    assert dep[0, 0, 0, 0] == 0 and 0 < dep[0, 0, 240, 320] < 0.3
    assert not msk[0, 0, 0, 0] and msk[0, 0, 240, 320]


def test_render_random_poses_parity(ctx, can):
    om, mid = can
    rs = np.random.RandomState(7)
    T, K = random_crop_scene(rs, 24)
    ref, got = _render_both(ctx, om, mid, T, K, (240, 320), ambient=rs.uniform(0.7, 1.0, (24, 1)).repeat(3, 1))
    mism, relmax = _check_render_parity(ref, got, "random")
    assert (got[3] == ref["mask"]).all() and (got[2] == ref["depth"]).all()  # stronger than the bar: bit-exact
    assert got[3].mean() > 0.1


def test_render_more_scenes_than_sms(ctx, can):
    """Persistent CTAs loop over scenes: 320 scenes > 148 SMs; the visibility buffer must be re-armed correctly."""
    om, mid = can
    rs = np.random.RandomState(8)
    T, K = random_crop_scene(rs, 320, res=(60, 80))
    ref, got = _render_both(ctx, om, mid, T, K, (60, 80))
    assert (got[3] == ref["mask"]).all() and (got[2] == ref["depth"]).all()
    _check_render_parity(ref, got, "many")


def test_render_edge_cases(ctx, can):
    """non-finite pose -> zero images; object behind / across the near plane; far rule; off-screen; empty batch."""
    from happypose_b200 import ops

    om, mid = can
    T0, K0, _ = reference_test_scene()
    K0 = K0.copy()
    K0[:2] /= 4
    T = np.stack([T0] * 7)
    T[1, 0, 3] = np.nan
    T[2, 2, 3] = -0.5    # behind the camera
    T[3, 2, 3] = 0.12    # straddles the near plane (triangles with a vertex closer than z_near are dropped)
    T[4, 2, 3] = 9.6     # d > 0.999: colour drawn, depth/mask 0
    T[5, 2, 3] = 12.0    # beyond far: clipped
    T[6, 0, 3] = 5.0     # off screen
    K = np.stack([K0] * 7)
    K[4, 0, 0] = K[4, 1, 1] = K[5, 0, 0] = K[5, 1, 1] = 3000
    ref, got = _render_both(ctx, om, mid, T, K, (120, 160))
    rgb, nrm, dep, msk = got
    assert (msk == ref["mask"]).all() and (dep == ref["depth"]).all()
    assert (rgb == ref["rgb"]).all() and (nrm == ref["normals"]).all()
    assert rgb[1].max() == 0 and dep[1].max() == 0 and nrm[1].max() == 0
    assert rgb[2].max() == 0 and rgb[5].max() == 0 and rgb[6].max() == 0
    assert rgb[4].max() > 0 and dep[4].max() == 0 and not msk[4].any()
    Kn = K.copy()
    Kn[0, 1, 1] = np.inf
    _, got2 = _render_both(ctx, om, mid, T[:1], Kn[:1], (120, 160))
    assert got2[0].max() == 0
    out = ops.render(ctx, torch.zeros(0, dtype=torch.int32), torch.zeros(0, 4, 4), torch.zeros(0, 3, 3), (120, 160), render_depth=True)
    assert out[0].shape == (0, 3, 120, 160) and out[2].shape == (0, 1, 120, 160)


def test_render_untextured_and_two_meshes(ctx, can):
    from happypose_b200 import ops

    om, mid = can
    v, f, n = icosphere(3, 0.05)
    sph = oraster.OracleMesh(v, f, None)  # normals generated (area-weighted) on both sides
    sid = ops.mesh_upload(ctx, sph.pos, f)
    colors = (np.random.RandomState(0).rand(len(v), 3) * 255).astype(np.uint8)
    sphc = oraster.OracleMesh(v, f, n, vcolor=colors)
    cid = ops.mesh_upload(ctx, sphc.pos, f, n, vcolor=colors)
    rs = np.random.RandomState(9)
    T, K = random_crop_scene(rs, 6, res=(120, 160))
    ids_o = [0, 1, 2, 1, 0, 2]
    ids_g = torch.tensor([mid, sid, cid, sid, mid, cid])
    ref = oraster.render([om, sph, sphc], ids_o, T, K, (120, 160), render_normals=True, render_depth=True, render_binary_mask=True)
    rgb, nrm, dep, msk = ops.render(ctx, ids_g, torch.as_tensor(T), torch.as_tensor(K), (120, 160), render_normals=True, render_depth=True, render_binary_mask=True)
    got = (rgb.cpu().numpy(), nrm.cpu().numpy(), dep.cpu().numpy(), msk.cpu().numpy())
    assert (got[3] == ref["mask"]).all() and (got[2] == ref["depth"]).all()
    _check_render_parity(ref, got, "mixed meshes")
    assert (got[0][1][:, got[3][1, 0]] == 1.0).all()  # untextured, uncoloured mesh is white under ambient 1


def test_render_big_triangles_int64_path(ctx):
    """A 2-triangle quad filling the screen exercises the 64-bit edge-function path and the guard band."""
    from happypose_b200 import ops

    v = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], np.float32)
    f = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    n = np.tile(np.array([[0, 0, -1]], np.float32), (4, 1))
    om = oraster.OracleMesh(v, f, n)
    mid = ops.mesh_upload(ctx, om.pos, f, n)
    T = np.tile(np.eye(4, dtype=np.float32), (3, 1, 1))
    T[:, 2, 3] = [0.5, 1.0, 0.2]
    T[1, :3, :3] = random_rotations(np.random.RandomState(3), 1)[0] @ np.diag([1, 1, 1])
    T[1, 2, 3] = 3.0
    K = np.tile(np.array([[300, 0, 160.3], [0, 300, 120.1], [0, 0, 1]], np.float32), (3, 1, 1))
    ref = oraster.render([om], [0, 0, 0], T, K, (240, 320), render_normals=True, render_depth=True, render_binary_mask=True)
    rgb, nrm, dep, msk = ops.render(ctx, torch.full((3,), mid), torch.as_tensor(T), torch.as_tensor(K), (240, 320), render_normals=True, render_depth=True, render_binary_mask=True)
    assert (msk.cpu().numpy() == ref["mask"]).all() and (dep.cpu().numpy() == ref["depth"]).all()
    assert msk[0].all()  # the quad covers the whole frame at z = 0.5
    assert (nrm.cpu().numpy() == ref["normals"]).all() and (rgb.cpu().numpy() == ref["rgb"]).all()


def test_render_gso_scale_mesh_uses_global_vertex_scratch(can_mesh_arrays):
    """BASELINE config #5 class: a mesh whose 12 B / vertex screen-space arrays exceed shared memory (40 962 vertices,
    81 920 triangles) is staged in the CTA's global scratch slice; a small mesh in the same launch keeps using shared
    memory (the choice is per scene).  Own context: the scratch buffer is sized by the largest mesh of a context."""
    from happypose_b200 import ops
    from happypose_b200._capi import Context

    ctx = Context(torch.device("cuda:0"))
    v, f, n = icosphere(6, 0.06)
    rs = np.random.RandomState(21)
    bump = 1.0 + 0.04 * np.sin(9 * n[:, :1]) * np.cos(7 * n[:, 1:2]) + 0.01 * rs.randn(len(v), 1).astype(np.float32)
    v = (v * bump).astype(np.float32)
    colors = (rs.rand(len(v), 3) * 255).astype(np.uint8)
    assert len(v) * 12 > 227 * 1024
    big = oraster.OracleMesh(v, f, None, vcolor=colors)  # normals generated from the bumped surface
    bid = ops.mesh_upload(ctx, big.pos, f, big.nrm, vcolor=colors)
    d = can_mesh_arrays
    om = oraster.OracleMesh(d["verts"], d["faces"], d["normals"], d["uv"], texture=d["texture"], scale=0.001)
    mid = ops.mesh_upload(ctx, om.pos, d["faces"], d["normals"], d["uv"], texture=d["texture"])
    T, K = random_crop_scene(rs, 5, res=(240, 320))
    ids_o = [0, 1, 0, 0, 1]
    ids_g = torch.tensor([bid, mid, bid, bid, mid])
    ref = oraster.render([big, om], ids_o, T, K, (240, 320), render_normals=True, render_depth=True, render_binary_mask=True, n_threads=8)
    rgb, nrm, dep, msk = ops.render(ctx, ids_g, torch.as_tensor(T), torch.as_tensor(K), (240, 320), render_normals=True, render_depth=True, render_binary_mask=True)
    got = (rgb.cpu().numpy(), nrm.cpu().numpy(), dep.cpu().numpy(), msk.cpu().numpy())
    assert (got[3] == ref["mask"]).all() and (got[2] == ref["depth"]).all()
    assert got[3][0].mean() > 0.05
    _check_render_parity(ref, got, "gso-scale mesh")


def test_backface_skipping_matches_two_sided(ctx, can):
    """Closed-surface analysis agrees with the oracle's; skipping back faces of the (closed) can changes nothing."""
    from happypose_b200 import ops

    om, mid = can
    assert ops.mesh_closed_sign(ctx, mid) == om.closed_sign == -1
    v, f, _ = icosphere(2, 0.05)
    for faces in (f, f[:, ::-1].copy(), f[:-1].copy()):
        assert ops.mesh_closed_sign(ctx, ops.mesh_upload(ctx, v, faces)) == oraster.closed_surface_sign(v, faces)
    rs = np.random.RandomState(12)
    T, K = random_crop_scene(rs, 40)
    T[0, 2, 3] = 0.12  # cut by the near plane: two-sided automatically
    ids = torch.full((40,), mid)
    a = ops.render(ctx, ids, torch.as_tensor(T), torch.as_tensor(K), (240, 320), render_normals=True, render_depth=True)
    ops.mesh_set_cull(ctx, mid, False)
    try:
        b = ops.render(ctx, ids, torch.as_tensor(T), torch.as_tensor(K), (240, 320), render_normals=True, render_depth=True)
    finally:
        ops.mesh_set_cull(ctx, mid, True)
    for x, y in zip(a[:3], b[:3]):
        assert (x != y).float().mean().item() <= 1e-4
    assert torch.equal(a[2][0], b[2][0]) and (a[2][0] > 0).any()
    # the two-sided path itself is still oracle-exact
    om2 = oraster.OracleMesh(om.pos, om.faces, om.nrm, om.uv, scale=1.0, cull=False)
    om2.tex, om2.tex_w, om2.tex_h, om2.tex_off = om.tex, om.tex_w, om.tex_h, om.tex_off
    ref = oraster.render([om2], [0] * 8, T[:8], K[:8], (240, 320), render_normals=True, render_depth=True, n_threads=8)
    assert (b[2][:8].cpu().numpy() == ref["depth"]).all() and (b[0][:8].cpu().numpy() == ref["rgb"]).all()


def test_render_into_network_input_slice(ctx, can):
    from happypose_b200 import ops

    om, mid = can
    rs = np.random.RandomState(10)
    T, K = random_crop_scene(rs, 5, res=(120, 160))
    x = torch.full((5, 9, 120, 160), -7.0, device="cuda")
    rgb, nrm, _, _ = ops.render(ctx, torch.full((5,), mid), torch.as_tensor(T), torch.as_tensor(K), (120, 160), render_normals=True, out=x, out_channel_offset=3)
    rgb2, nrm2, _, _ = ops.render(ctx, torch.full((5,), mid), torch.as_tensor(T), torch.as_tensor(K), (120, 160), render_normals=True)
    assert (x[:, :3] == -7.0).all()
    assert torch.equal(x[:, 3:6], rgb2) and torch.equal(x[:, 6:9], nrm2)
    assert rgb.data_ptr() == x[:, 3:6].data_ptr()


# ------------------------------------------------------------------------------------------------------------
def _mesh_points(can_mesh_arrays):
    return (can_mesh_arrays["verts"].astype(np.float64) * 0.001).astype(np.float32)


def test_crop_matches_reference_golden_small(ctx, golden, can_mesh_arrays):
    from happypose_b200 import ops

    g = golden("ref_crop_small.npz")
    pts = _mesh_points(can_mesh_arrays)[g["point_ids"]]
    b = len(g["TCO"])
    for C, tag in ((4, "rgbd"), (3, "rgb")):
        crops, K_crop, boxes_rend, boxes_crop = ops.crop(
            ctx, torch.as_tensor(g["images"][:, :C].copy()), torch.as_tensor(g["im_ids"]), torch.as_tensor(pts[None]),
            torch.zeros(b, dtype=torch.int32), g["K"], g["TCO"], g["tCR"], (60, 80))
        np.testing.assert_allclose(boxes_rend.cpu().numpy(), g["boxes_rend"], atol=2e-3)
        np.testing.assert_allclose(boxes_crop.cpu().numpy(), g["boxes_crop"], atol=5e-3)
        np.testing.assert_allclose(K_crop.cpu().numpy(), g["K_crop"], rtol=2e-5, atol=2e-3)
        ref = g[f"crops_{tag}"]
        got = crops.cpu().numpy()
        np.testing.assert_allclose(got[:, :3], ref[:, :3], atol=1e-3)  # BASELINE bar: crops within 1e-3 absolute
        if C == 4:
            assert (np.abs(got[:, 3] - ref[:, 3]) > 1e-3).mean() < 2e-3  # validity threshold flips (cropping.py:191-193)


def test_crop_full_size_matches_reference_golden(ctx, golden, can_mesh_arrays):
    from happypose_b200 import ops

    g = golden("ref_crop_full.npz")
    pts = _mesh_points(can_mesh_arrays)[O.sample_point_ids(9951, 2000)]
    b = len(g["TCO"])
    image = np.random.RandomState(int(g["image_seed"])).rand(1, 3, 480, 640).astype(np.float32)
    crops, K_crop, boxes_rend, boxes_crop = ops.crop(
        ctx, torch.as_tensor(image), torch.zeros(b, dtype=torch.int32), torch.as_tensor(pts[None]), torch.zeros(b, dtype=torch.int32),
        g["K"], g["TCO"], g["tCR"], (240, 320))
    np.testing.assert_allclose(boxes_crop.cpu().numpy(), g["boxes_crop"], atol=1e-2)
    np.testing.assert_allclose(K_crop.cpu().numpy(), g["K_crop"], rtol=3e-5, atol=5e-3)
    got = crops.cpu().numpy()
    np.testing.assert_allclose(got[:, :, ::5, ::5], g["crops_sub"], atol=1e-3)
    np.testing.assert_allclose(got.astype(np.float64).sum((1, 2, 3)), g["crops_sum"], rtol=1e-5)


def test_crop_vs_oracle_downsampling_and_out_of_frame(ctx, can_mesh_arrays):
    """Boxes much larger than the frame (generic roi_align path) and boxes hanging off the frame (zero padding rule)."""
    from happypose_b200 import ops

    rs = np.random.RandomState(11)
    pts = _mesh_points(can_mesh_arrays)[O.sample_point_ids(9951, 2000)]
    b = 6
    images = rs.rand(2, 4, 96, 128).astype(np.float32)
    images[:, 3] *= rs.rand(2, 96, 128) > 0.2
    K = np.tile(np.array([[120.0, 0, 64], [0, 120, 48], [0, 0, 1]], np.float32), (b, 1, 1))
    TCO = np.tile(np.eye(4, dtype=np.float32), (b, 1, 1))
    TCO[:, :3, :3] = random_rotations(rs, b)
    TCO[:, :3, 3] = [[0, 0, 0.15], [0.3, 0.2, 0.5], [-0.4, 0.0, 0.6], [0, 0, 0.11], [0.0, -0.35, 0.7], [0.01, 0.01, 2.5]]
    tCR = TCO[:, :3, 3].copy()
    im_ids = np.array([0, 1, 0, 1, 1, 0])
    points = np.tile(pts[None], (b, 1, 1))
    ref_crops, ref_K, ref_br, ref_bc = O.crop_inputs(images, K, TCO, tCR, points, (48, 64), im_ids=im_ids)
    crops, K_crop, boxes_rend, boxes_crop = ops.crop(ctx, torch.as_tensor(images), torch.as_tensor(im_ids), torch.as_tensor(pts[None]),
                                                     torch.zeros(b, dtype=torch.int32), K, TCO, tCR, (48, 64))
    np.testing.assert_allclose(boxes_crop.cpu().numpy(), ref_bc, rtol=1e-5, atol=5e-3)
    np.testing.assert_allclose(K_crop.cpu().numpy(), ref_K, rtol=3e-5, atol=5e-3)
    # resample with the oracle on the GPU's own boxes so that only the roi_align arithmetic is compared
    rois = np.concatenate([im_ids[:, None].astype(np.float32), boxes_crop.cpu().numpy()], 1)
    ref2 = O.crop_images(images, rois, (48, 64), 4)
    got = crops.cpu().numpy()
    np.testing.assert_allclose(got[:, :3], ref2[:, :3], atol=2e-5)
    assert (np.abs(got[:, 3] - ref2[:, 3]) > 1e-3).mean() < 2e-3


def test_crop_wide_output_uses_column_tiles(ctx, can_mesh_arrays):
    """Outputs wider than one CTA's 320 columns (e.g. the 480x640 crops of a depth refiner) are split into column tiles
    (blockIdx.z); odd widths leave a ragged last tile.  Up-sampling and mild down-sampling, few rows and many rows, packed
    (many rows per frame) and planar (few rows per frame) source layouts."""
    from happypose_b200 import ops

    rs = np.random.RandomState(12)
    pts = _mesh_points(can_mesh_arrays)[O.sample_point_ids(9951, 2000)]
    for b, size in ((2, (96, 648)), (9, (50, 333)), (3, (480, 640))):
        images = rs.rand(1, 3, 240, 320).astype(np.float32)
        K = np.tile(np.array([[300.0, 0, 160], [0, 300, 120], [0, 0, 1]], np.float32), (b, 1, 1))
        TCO = np.tile(np.eye(4, dtype=np.float32), (b, 1, 1))
        TCO[:, :3, :3] = random_rotations(rs, b)
        TCO[:, :3, 3] = np.stack([rs.uniform(-0.05, 0.05, b), rs.uniform(-0.05, 0.05, b), rs.uniform(0.25, 0.9, b)], 1)
        tCR = TCO[:, :3, 3].copy()
        im_ids = np.zeros(b, np.int64)
        crops, K_crop, boxes_rend, boxes_crop = ops.crop(ctx, torch.as_tensor(images), torch.as_tensor(im_ids), torch.as_tensor(pts[None]),
                                                         torch.zeros(b, dtype=torch.int32), K, TCO, tCR, size)
        rois = np.concatenate([im_ids[:, None].astype(np.float32), boxes_crop.cpu().numpy()], 1)
        ref = O.crop_images(images, rois, size, 4)
        np.testing.assert_allclose(crops.cpu().numpy(), ref, atol=2e-5)


def test_crop_fp16_taps_equal_float_crop_of_the_fp16_frame(ctx, golden, can_mesh_arrays):
    """hpb_set_crop_tap_precision(16): the kernel samples an fp16 copy of the RGB frame.  Exactness statement: the result
    is the 32-bit-tap crop of fp16(frame), bit for bit (same boxes, same weights, same summation order); against the
    reference's float32 golden crops (real torchvision roi_align) it stays inside the BASELINE bar of 1e-3 absolute."""
    from happypose_b200 import ops

    pts = _mesh_points(can_mesh_arrays)[O.sample_point_ids(9951, 2000)]
    g = golden("ref_crop_full.npz")
    rep = 3  # 12 rows of one frame: enough rows per frame for the interleaved-copy path the option applies to
    K, TCO, tCR = (np.tile(g[k], (rep,) + (1,) * (g[k].ndim - 1)) for k in ("K", "TCO", "tCR"))
    b = len(TCO)
    image = torch.as_tensor(np.random.RandomState(int(g["image_seed"])).rand(1, 3, 480, 640).astype(np.float32)).cuda()
    zero = torch.zeros(b, dtype=torch.int32)
    args = (zero, torch.as_tensor(pts[None]), zero, K, TCO, tCR, (240, 320))
    c16, K16, _, bc16 = ops.crop(ctx, image, *args, tap_bits=16)
    c32_of_half, K32, _, bc32 = ops.crop(ctx, image.half().float(), *args, tap_bits=32)
    assert torch.equal(c16, c32_of_half) and torch.equal(K16, K32) and torch.equal(bc16, bc32)
    c32, _, _, _ = ops.crop(ctx, image, *args, tap_bits=32)
    assert 0 < float((c16 - c32).abs().max()) <= 2.5e-4  # fp16 spacing on [0.5, 1) is 4.9e-4; convex combinations keep the bound
    np.testing.assert_allclose(c16.cpu().numpy()[:4, :, ::5, ::5], g["crops_sub"], atol=1e-3)
    np.testing.assert_allclose(c32.cpu().numpy()[:4, :, ::5, ::5], g["crops_sub"], atol=1e-3)
    # RGB-D frames keep float32 taps whatever the option says
    imaged = torch.cat([image, torch.rand(1, 1, 480, 640, device="cuda") + 0.2], 1)
    d16, _, _, _ = ops.crop(ctx, imaged, *args, tap_bits=16)
    d32, _, _, _ = ops.crop(ctx, imaged, *args, tap_bits=32)
    assert torch.equal(d16, d32)


def test_crop_boxes_multiview_200_points(ctx, can_mesh_arrays):
    from happypose_b200 import ops

    rs = np.random.RandomState(12)
    pts = _mesh_points(can_mesh_arrays)[O.sample_point_ids(9951, 200)]
    T, K = random_crop_scene(rs, 16, res=(480, 640))
    tCR = T[:, :3, 3].copy()
    points = np.tile(pts[None], (16, 1, 1))
    uv = O.project_points_robust(points, K, T)
    br = O.boxes_from_uv(uv)
    bc, _ = O.deepim_crops_robust(np.zeros((1, 3, 480, 640), np.float32), br, K, T, tCR, points, (240, 320), return_crops=False)
    Kc = O.get_K_crop_resize(K, bc, (480, 640), (240, 320))
    K_crop, boxes_rend, boxes_crop = ops.crop_boxes(ctx, (480, 640), torch.as_tensor(pts[None]), torch.zeros(16, dtype=torch.int32), K, T, tCR, (240, 320))
    np.testing.assert_allclose(boxes_rend.cpu().numpy(), br, atol=5e-3)
    np.testing.assert_allclose(boxes_crop.cpu().numpy(), bc, atol=1e-2)
    np.testing.assert_allclose(K_crop.cpu().numpy(), Kc, rtol=3e-5, atol=5e-3)


# ------------------------------------------------------------------------------------------------------------
def test_pose_kernels_match_reference_golden(ctx, golden):
    from happypose_b200 import _capi, ops

    g = golden("ref_pose.npz")
    Tn = ops.normalize_T(ctx, torch.as_tensor(g["TCO"]))
    np.testing.assert_allclose(Tn.cpu().numpy(), g["normalize_T"], atol=1e-6)
    # BASELINE bar: pose updates within 1e-5
    up = ops.pose_update(ctx, Tn, g["K_crop"], g["out9"], g["tCR"], _capi.POSE_MEGAPOSE)
    np.testing.assert_allclose(up.cpu().numpy(), g["pose_update_megapose"], atol=1e-5)
    up = ops.pose_update(ctx, Tn, g["K_crop"], g["out9"], None, _capi.POSE_COSYPOSE_6D)
    np.testing.assert_allclose(up.cpu().numpy(), g["pose_update_cosypose6d"], atol=1e-5)
    up = ops.pose_update(ctx, Tn, g["K_crop"], g["out7"], None, _capi.POSE_COSYPOSE_QUAT)
    np.testing.assert_allclose(up.cpu().numpy(), g["pose_update_cosyposequat"], atol=1e-5)
    with pytest.raises(ValueError):
        ops.pose_update(ctx, Tn, g["K_crop"], g["out9"], None, _capi.POSE_MEGAPOSE)  # tCR missing


def test_tco_init_matches_reference_golden(ctx, golden, can_mesh_arrays):
    from happypose_b200 import _capi, ops

    g = golden("ref_pose.npz")
    pts = torch.as_tensor(_mesh_points(can_mesh_arrays)[None])
    b = len(g["boxes"])
    ids = torch.zeros(b, dtype=torch.int32)
    out = ops.tco_init(ctx, _capi.TCO_INIT_AUTODEPTH_WITH_R, g["boxes"], g["K_init"], pts, ids, g["R_init"])
    np.testing.assert_allclose(out.cpu().numpy(), g["tco_init_autodepth_with_R"], atol=1e-5)
    out = ops.tco_init(ctx, _capi.TCO_INIT_ZUP_AUTODEPTH, g["boxes"], g["K_init"], pts, ids)
    np.testing.assert_allclose(out.cpu().numpy(), g["tco_init_zup_autodepth"], atol=1e-5)
    out = ops.tco_init(ctx, _capi.TCO_INIT_FROM_BOXES, g["boxes"], g["K_init"], z_mean=1.0)
    np.testing.assert_allclose(out.cpu().numpy(), g["tco_init_from_boxes"], atol=1e-6)


def test_multiview_matches_oracle(ctx):
    from happypose_b200 import ops

    rs = np.random.RandomState(13)
    T, _ = random_crop_scene(rs, 32)
    T[:, :2, 3] += rs.uniform(-0.2, 0.2, (32, 2)).astype(np.float32)
    T = O.normalize_T(T)
    tCR = T[:, :3, 3].copy()
    for mv, nv in (("TCO+front_3views", 4), ("TCO+front_1view", 2), ("sphere_26views", 27)):
        ref = O.make_TCO_multiview(T, tCR, mv, nv)
        got = ops.multiview(ctx, torch.as_tensor(T), torch.as_tensor(tCR), mv, nv).cpu().numpy()
        np.testing.assert_allclose(got, ref, atol=1e-5)
    got = ops.multiview(ctx, torch.as_tensor(T), torch.as_tensor(tCR), "TCO+front_3views", 3, remove_TCO_rendering=True).cpu().numpy()
    np.testing.assert_allclose(got, O.make_TCO_multiview(T, tCR, "TCO+front_3views", 3, remove_TCO_rendering=True), atol=1e-5)
    # the training-time option of lib3d.multiview.make_TCO_multiview (multiview.py:239-250; megapose_forward_loss.py:116-123):
    # 26 sphere views x 4 in-plane rotations
    from happypose_b200.lib3d.multiview import make_TCO_multiview

    got = make_TCO_multiview(torch.as_tensor(T).cuda(), torch.as_tensor(tCR).cuda(), "sphere_26views", 26, remove_TCO_rendering=True,
                             views_inplane_rotations=True).cpu().numpy()
    ref = O.make_TCO_multiview(T, tCR, "sphere_26views", 26, remove_TCO_rendering=True, views_inplane_rotations=True)
    assert got.shape == ref.shape == (32, 104, 4, 4)
    np.testing.assert_allclose(got, ref, atol=1e-5)
    one = ops.multiview(ctx, torch.as_tensor(T), torch.as_tensor(tCR), "TCO+front_3views", 1).cpu().numpy()
    assert (one[:, 0] == T).all()
    bad = T.copy()
    bad[0, 1, 1] = np.nan
    got = ops.multiview(ctx, torch.as_tensor(bad), torch.as_tensor(tCR), "TCO+front_3views", 4).cpu().numpy()
    assert np.isfinite(got[1:]).all()


def test_refiner_prologue_equals_the_separate_kernels(ctx, can):
    """hpb_refiner_prologue (one launch per refiner iteration) is bit-identical to hpb_normalize_T -> hpb_multiview ->
    hpb_crop_boxes (2000 points) -> hpb_crop_boxes per view (200 points) -> KV_crop[:, 0] = K_crop, for every multiview
    type, with and without the TCO view, and hpb_crop_pixels at its boxes equals hpb_crop's pixels."""
    from happypose_b200 import ops

    om, _ = can
    rs = np.random.RandomState(17)
    dev = torch.device("cuda")
    b = 7
    T, _ = random_crop_scene(rs, b, z_range=(0.35, 1.0))
    T[:, :2, 3] += rs.uniform(-0.1, 0.1, (b, 2)).astype(np.float32)
    T[:, :3, :3] += rs.uniform(-0.01, 0.01, (b, 3, 3)).astype(np.float32)  # not orthonormal: normalize_T has work to do
    K = np.tile(np.array([[605.95, 0, 319.03], [0, 605.01, 249.68], [0, 0, 1]], np.float32), (b, 1, 1))
    pts2000 = torch.as_tensor(np.stack([om.pos[rs.choice(len(om.pos), 2000, replace=False)] for _ in range(2)])).to(dev)
    pts200 = torch.as_tensor(np.stack([om.pos[rs.choice(len(om.pos), 200, replace=False)] for _ in range(2)])).to(dev)
    obj_ids = torch.as_tensor(rs.randint(0, 2, b).astype(np.int32)).to(dev)
    Tt, Kt = torch.as_tensor(T).to(dev), torch.as_tensor(K).to(dev)
    for mv, nv, remove in (("TCO+front_3views", 4, False), ("TCO+front_3views", 3, True), ("TCO+front_1view", 2, False),
                           ("sphere_26views", 27, False), ("TCO+front_3views", 1, False)):
        got = ops.refiner_prologue(ctx, Tt, Kt, obj_ids, pts2000, pts200, (480, 640), (240, 320), mv, nv, remove)
        Tn = ops.normalize_T(ctx, Tt)
        tCR = Tn[:, :3, 3].contiguous()
        TV = ops.multiview(ctx, Tn, tCR, mv, nv, remove)
        Kc, br, bc = ops.crop_boxes(ctx, (480, 640), pts2000, obj_ids, Kt, Tn, tCR, (240, 320))
        KV, _, _ = ops.crop_boxes(ctx, (480, 640), pts200, obj_ids.repeat_interleave(nv), Kt.repeat_interleave(nv, 0), TV.flatten(0, 1),
                                  TV[:, :, :3, 3].reshape(-1, 3).contiguous(), (240, 320))
        KV = KV.view(b, nv, 3, 3).clone()
        if not remove:
            KV[:, 0] = Kc
        for name, a, w in (("T_norm", got["T_norm"], Tn), ("tCR", got["tCR"], tCR), ("TCV_O", got["TCV_O"], TV), ("K_crop", got["K_crop"], Kc),
                           ("boxes_rend", got["boxes_rend"], br), ("boxes_crop", got["boxes_crop"], bc), ("KV_crop", got["KV_crop"], KV)):
            assert torch.equal(a, w), f"{name} differs for {mv} / {nv} views"
    img = torch.as_tensor(rs.rand(2, 3, 480, 640).astype(np.float32)).to(dev)
    im_ids = torch.as_tensor(rs.randint(0, 2, b).astype(np.int32)).to(dev)
    for tap in (32, 16):
        want, _, _, bc2 = ops.crop(ctx, img, im_ids, pts2000, obj_ids, Kt, Tn, tCR, (240, 320), tap_bits=tap)
        x = torch.zeros((b, 9, 240, 320), device=dev)
        crops = ops.crop_pixels(ctx, img, im_ids, bc2, (240, 320), out=x, tap_bits=tap)
        assert torch.equal(crops, want) and crops.data_ptr() == x.data_ptr() and not x[:, 3:].any()


def test_normalize_depth(ctx):
    from happypose_b200 import ops

    rs = np.random.RandomState(14)
    x = rs.rand(3, 9, 20, 30).astype(np.float32) * 2
    tCR = rs.uniform(0.3, 1.0, (3, 3)).astype(np.float32)
    for kind in ("tCR_scale", "tCR_scale_clamp_center", "tCR_center_clamp", "none"):
        t = torch.as_tensor(x.copy()).cuda()
        ops.normalize_depth_(ctx, t, [3, 8], tCR, kind)
        ref = x.copy()
        ref[:, [3, 8]] = O.normalize_depth(x[:, [3, 8]], tCR[:, 2], kind)
        np.testing.assert_allclose(t.cpu().numpy(), ref, atol=1e-6)
    with pytest.raises(ValueError):
        ops.normalize_depth_(ctx, torch.zeros(1, 4, 2, 2).cuda(), [3], tCR[:1], "bogus")


# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c1", "k5", "multi", "kbig"])
def test_topk_matches_pandas_golden(ctx, golden, name):
    from happypose_b200 import ops

    g = golden("ref_topk.npz")
    groups = g[f"{name}_groups"]
    idx = ops.topk_segmented(ctx, g[f"{name}_scores"], groups, int(groups.max()) + 1, int(g[f"{name}_K"]))
    assert idx.dtype == torch.int64
    assert (idx.cpu().numpy() == g[f"{name}_idx"]).all()  # bit-exact indices


def test_topk_large_and_ties(ctx):
    from happypose_b200 import ops

    rs = np.random.RandomState(15)
    n_groups, M = 240, 576  # BASELINE config #4: 240 detections x 576 hypotheses
    scores = rs.randn(n_groups * M).astype(np.float32)
    scores[rs.randint(0, len(scores), 5000)] = 0.25   # many exact ties
    scores[rs.randint(0, len(scores), 50)] = np.nan
    scores[:7] = [0.0, -0.0, np.inf, -np.inf, 0.0, -0.0, np.inf]
    groups = rs.permutation(np.repeat(np.arange(n_groups), M)).astype(np.int32)  # interleaved, not contiguous
    for K in (1, 5):
        idx = ops.topk_segmented(ctx, scores, groups, n_groups, K).cpu().numpy()
        assert (idx == O.filter_top_k(scores, groups, K)).all()
    s = np.array([1.0, 3.0, 3.0, 2.0, 3.0, np.nan], np.float32)
    g = np.array([0, 0, 1, 1, 0, 0], np.int32)
    assert ops.topk_segmented(ctx, s, g, 2, 2).tolist() == [1, 2, 4, 3]
    assert ops.topk_segmented(ctx, s, g, 2, 10).tolist() == [1, 2, 4, 3, 0, 5]
    assert ops.topk_segmented(ctx, s[:0], g[:0], 0, 3).tolist() == []
    # a group bigger than the rank kernel's shared-memory tile (SO(3) grid 4608)
    big = rs.randn(4608 * 2).astype(np.float32)
    gb = np.repeat(np.arange(2), 4608).astype(np.int32)
    assert (ops.topk_segmented(ctx, big, gb, 2, 7).cpu().numpy() == O.filter_top_k(big, gb, 7)).all()


# ----------------------------------------------------------------------------------------------------------------
# network-input hand-off kernels (bit-exact: bf16 rounding and max are exact operations)
# ----------------------------------------------------------------------------------------------------------------
def test_pack_input_bf16_is_bit_exact(ctx):
    from happypose_b200 import ops

    torch.manual_seed(0)
    for C, cp in ((9, 16), (27, 32), (3, 8)):
        x = torch.randn(3, C, 24, 40, device="cuda")
        got = ops.pack_input_bf16(ctx, x, cp)
        assert got.shape == (3, cp, 24, 40) and got.is_contiguous(memory_format=torch.channels_last)
        assert torch.equal(got[:, :C], x.to(torch.bfloat16)) and (got[:, C:] == 0).all()


def test_pack_input_s2d_matches_torch_statement(ctx):
    from happypose_b200 import ops
    from happypose_b200.megapose.fast_resnet import s2d_reference

    torch.manual_seed(1)
    for C, H, W, cz in ((9, 240, 320, 64), (27, 16, 24, 128), (6, 10, 14, 32), (9, 12, 16, 96)):
        x = torch.randn(2, C, H, W, device="cuda")
        got = ops.pack_input_s2d_bf16(ctx, x, cz)
        ref = s2d_reference(x, cz).to(torch.bfloat16)
        assert got.shape == ref.shape == (2, cz, H // 2 + 3, W // 2 + 3)
        assert torch.equal(got, ref)


def test_point_lights_match_oracle(ctx, can):
    """hpb_render with point / directional lights (the render_normals=False light rig of pose_rigid.py:105-141,421-422) is
    bit-identical to the oracle's per-pixel Lambert shading: rig of 6 point lights, a directional light, mixed batch with
    an unlit scene padded with black lights, written into a network-input slice with 4 views."""
    from happypose_b200 import ops

    om, mid = can
    rs = np.random.RandomState(41)
    b = 8
    H, W = 240, 320
    T, K = random_crop_scene(rs, b)
    radius = float(np.linalg.norm(om.pos - 0.5 * (om.pos.min(0) + om.pos.max(0)), axis=1).max())
    axes = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], np.float32)
    lights = np.zeros((b, 7, 8), np.float32)
    lights[:, :6, 1:4] = axes * radius * 10
    lights[:, :6, 4:7] = 0.4
    lights[:, 6, 0] = 1.0                      # a directional light
    lights[:, 6, 1:4] = rs.randn(b, 3)
    lights[:, 6, 1:4] /= np.linalg.norm(lights[:, 6, 1:4], axis=1, keepdims=True)
    lights[:, 6, 4:7] = rs.uniform(0, 0.5, (b, 3))
    lights[3] = 0.0                            # scene 3: black lights only
    amb = np.full((b, 3), 0.1, np.float32)
    ref = oraster.render([om], np.zeros(b, int), T, K, (H, W), ambient=amb, lights=lights, render_depth=True, n_threads=4)
    ids = torch.full((b,), mid, dtype=torch.int32)
    rgb, _, dep, _ = ops.render(ctx, ids, torch.as_tensor(T), torch.as_tensor(K), (H, W), ambient=torch.as_tensor(amb),
                                lights=torch.as_tensor(lights), render_depth=True)
    assert np.array_equal(dep.cpu().numpy(), ref["depth"])
    d = np.abs(rgb.cpu().numpy() - ref["rgb"]) * 255
    assert d.max() <= 1.0 + 1e-3 and (d > 0.5).mean() < 1e-3  # 8-bit levels: identical up to rare rounding flips
    # 4 views interleaved into a [b/4, 3 + 4*3, h, w] input (render_normals=False refiner layout)
    x = torch.zeros((2, 15, H, W), device="cuda")
    ops.render(ctx, ids, torch.as_tensor(T), torch.as_tensor(K), (H, W), ambient=torch.as_tensor(amb), lights=torch.as_tensor(lights),
               out=x, out_channel_offset=3, views=4)
    assert torch.equal(x[:, 3:].reshape(8, 3, H, W), rgb)


def test_render_s2d_bf16_equals_render_then_pack(ctx, can):
    """hpb_render_s2d_bf16 (the rasteriser writing the stem's bf16 space-to-depth input itself) is bit-identical to
    hpb_crop -> hpb_render into the float32 network input -> hpb_pack_input_s2d_bf16, including the zero border, the
    zero pad channels, a scene with a non-finite pose (zero render, crop still present) and > #SM scenes."""
    from happypose_b200 import ops

    om, mid = can
    rs = np.random.RandomState(31)
    for b, res in ((5, (240, 320)), (160, (60, 80)), (3, (118, 162))):
        T, K = random_crop_scene(rs, b, res=res)
        if b == 5:
            T[3, 0, 3] = np.nan
        crops = torch.as_tensor(rs.rand(b, 3, *res).astype(np.float32)).cuda()
        ids = torch.full((b,), mid, dtype=torch.int32)
        x = torch.empty((b, 9, *res), device="cuda")
        x[:, :3] = crops
        ops.render(ctx, ids, torch.as_tensor(T), torch.as_tensor(K), res, render_normals=True, out=x, out_channel_offset=3)
        want = ops.pack_input_s2d_bf16(ctx, x, 64)
        got = ops.render_s2d_bf16(ctx, ids, torch.as_tensor(T), torch.as_tensor(K), crops, 64)
        assert got.shape == want.shape and got.is_contiguous(memory_format=torch.channels_last)
        assert torch.equal(got.view(torch.int16), want.view(torch.int16)), f"b={b} res={res}"
        # crops given as the first 3 channels of a wider tensor (row stride 9 planes), 96 padded channels (24 per sub-pixel)
        want96 = ops.pack_input_s2d_bf16(ctx, x, 96)
        got96 = ops.render_s2d_bf16(ctx, ids, torch.as_tensor(T), torch.as_tensor(K), x[:, :3], 96)
        assert torch.equal(got96.view(torch.int16), want96.view(torch.int16))
        # a second call: the visibility buffer must have been re-armed by the first
        again = ops.render_s2d_bf16(ctx, ids, torch.as_tensor(T), torch.as_tensor(K), crops, 64)
        assert torch.equal(again.view(torch.int16), want.view(torch.int16))
        # the shipped form: crop handed over as bf16 pixels (r,g,b,0), persistent pre-zeroed output; written twice into the
        # same buffer with different crops to show that nothing stale survives (also with 128 padded channels, where the
        # kernel really skips the padding groups)
        crops_h = torch.zeros((b, *res, 4), dtype=torch.bfloat16, device="cuda")
        crops_h[..., :3] = crops.permute(0, 2, 3, 1).to(torch.bfloat16)
        buf = torch.empty((b + 2, 64, res[0] // 2 + 3, res[1] // 2 + 3), dtype=torch.bfloat16, device="cuda", memory_format=torch.channels_last).zero_()
        ops.render_s2d_bf16(ctx, ids, torch.as_tensor(T), torch.as_tensor(K), torch.flip(crops_h, (1,)).contiguous(), 64, out=buf[:b], pad_prezeroed=True)
        got_h = ops.render_s2d_bf16(ctx, ids, torch.as_tensor(T), torch.as_tensor(K), crops_h, 64, out=buf[:b], pad_prezeroed=True)
        assert got_h.data_ptr() == buf.data_ptr()
        assert torch.equal(got_h.view(torch.int16), want.view(torch.int16)), f"bf16x4 / pre-zeroed, b={b} res={res}"
        assert not buf[b:].any()
        buf128 = torch.empty((b, 128, res[0] // 2 + 3, res[1] // 2 + 3), dtype=torch.bfloat16, device="cuda", memory_format=torch.channels_last).zero_()
        ops.render_s2d_bf16(ctx, ids, torch.as_tensor(T), torch.as_tensor(K), torch.flip(crops_h, (1,)).contiguous(), 128, out=buf128, pad_prezeroed=True)
        got128 = ops.render_s2d_bf16(ctx, ids, torch.as_tensor(T), torch.as_tensor(K), crops_h, 128, out=buf128, pad_prezeroed=True)
        assert torch.equal(got128.view(torch.int16), ops.pack_input_s2d_bf16(ctx, x, 128).view(torch.int16))


def test_crop_tma_ring_equals_per_lane_kernel(ctx, can):
    """hpb_crop_tma.cu (source rows streamed by TMA into a shared-memory ring) returns bit for bit what the per-lane gather
    kernel returns, for fp16 taps of RGB frames: 576-row coarse batch, boxes hanging over every frame edge, narrow outputs,
    several frames, tiny batches (short bands), both output formats; wide boxes take the generic path in both."""
    from happypose_b200 import ops

    om, _ = can
    rs = np.random.RandomState(52)
    dev = torch.device("cuda")
    pts = torch.as_tensor(om.pos[rs.choice(len(om.pos), 2000, replace=False)][None]).to(dev)

    def both(fn):
        out = []
        for tma in (1, 0):
            ctx.check(ctx.lib.hpb_set_crop_tma(ctx.handle, tma), "hpb_set_crop_tma")
            out.append(fn())
        ctx.check(ctx.lib.hpb_set_crop_tma(ctx.handle, 0), "hpb_set_crop_tma")  # the default
        return out

    for b, n_im, res, spread in ((576, 1, (240, 320), 0.02), (64, 2, (240, 320), 0.35), (9, 1, (60, 80), 0.2), (40, 1, (118, 162), 0.1), (1, 1, (240, 320), 0.0)):
        T, _ = random_crop_scene(rs, b, z_range=(0.3, 1.1))
        T[:, :2, 3] += rs.uniform(-spread, spread, (b, 2)).astype(np.float32)  # large spread: boxes leave the frame on all sides
        if b == 64:
            T[:4, 2, 3] = 0.13  # a few very wide boxes: generic path
        K = np.tile(np.array([[605.95, 0, 319.03], [0, 605.01, 249.68], [0, 0, 1]], np.float32), (b, 1, 1))
        img = torch.as_tensor(rs.rand(n_im, 3, 480, 640).astype(np.float32)).to(dev)
        im_ids = torch.as_tensor(rs.randint(0, n_im, b).astype(np.int32)).to(dev)
        zero = torch.zeros(b, dtype=torch.int32, device=dev)
        Tt, Kt = torch.as_tensor(T).to(dev), torch.as_tensor(K).to(dev)
        tCR = Tt[:, :3, 3].contiguous()
        a, c = both(lambda: ops.crop(ctx, img, im_ids, pts, zero, Kt, Tt, tCR, res, tap_bits=16)[0])
        assert torch.equal(a, c), f"planar b={b} res={res}"
        assert a.abs().sum() > 0
        a, c = both(lambda: ops.crop_bf16x4(ctx, img, im_ids, pts, zero, Kt, Tt, tCR, res, tap_bits=16)[0])
        assert torch.equal(a.view(torch.int16), c.view(torch.int16)), f"bf16x4 b={b} res={res}"


def test_crop_bf16x4_is_the_rounded_float32_crop(ctx, can):
    """hpb_crop_bf16x4 = hpb_crop rounded to bfloat16 (nearest even), pixel-interleaved, for float32 and fp16 taps, the
    fast (few frames, many rows) and the generic (wide box) paths; K_crop / boxes identical."""
    from happypose_b200 import ops

    om, _ = can
    rs = np.random.RandomState(32)
    dev = torch.device("cuda")
    for b, n_im, tap in ((40, 1, 32), (40, 1, 16), (3, 3, 32)):
        T, _ = random_crop_scene(rs, b, z_range=(0.35, 0.9))
        T[:, :2, 3] += rs.uniform(-0.05, 0.05, (b, 2)).astype(np.float32)
        if b == 3:
            T[0, 2, 3] = 0.12  # very close: a crop box several times wider than the output -> generic path
        K = np.tile(np.array([[605.95, 0, 319.03], [0, 605.01, 249.68], [0, 0, 1]], np.float32), (b, 1, 1))
        img = torch.as_tensor(rs.rand(n_im, 3, 480, 640).astype(np.float32)).to(dev)
        im_ids = torch.as_tensor(rs.randint(0, n_im, b).astype(np.int32)).to(dev)
        pts = torch.as_tensor(om.pos[rs.choice(len(om.pos), 2000, replace=False)][None]).to(dev)
        zero = torch.zeros(b, dtype=torch.int32, device=dev)
        Tt, Kt = torch.as_tensor(T).to(dev), torch.as_tensor(K).to(dev)
        tCR = Tt[:, :3, 3].contiguous()
        c32, K32, br32, bc32 = ops.crop(ctx, img, im_ids, pts, zero, Kt, Tt, tCR, (240, 320), tap_bits=tap)
        ch, Kh, brh, bch = ops.crop_bf16x4(ctx, img, im_ids, pts, zero, Kt, Tt, tCR, (240, 320), tap_bits=tap)
        assert ch.shape == (b, 240, 320, 4) and ch.dtype == torch.bfloat16
        assert torch.equal(K32, Kh) and torch.equal(br32, brh) and torch.equal(bc32, bch)
        want = c32.permute(0, 2, 3, 1).to(torch.bfloat16)
        assert torch.equal(ch[..., :3].contiguous().view(torch.int16), want.contiguous().view(torch.int16)), f"b={b} tap={tap}"
        assert not ch[..., 3].any()


def test_maxpool_bf16_nhwc_is_bit_exact(ctx):
    from happypose_b200 import ops

    torch.manual_seed(2)
    # C = 64 goes through the TMA-staged tile kernel (hpb_maxpool_tma.cu: tiles of 16 x 8 outputs, so odd sizes exercise
    # partial tiles and the masked padding taps; negative inputs show that the TMA's zero fill never leaks into a result),
    # other channel counts through the plain kernel; both against torch
    for shape in ((2, 64, 120, 160), (3, 64, 37, 53), (1, 64, 5, 3), (150, 64, 16, 24), (1, 8, 7, 9), (3, 16, 2, 2)):
        x = (torch.randn(shape, device="cuda") - 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        ref = torch.nn.functional.max_pool2d(x, 3, 2, 1)
        for tma in (True, False):
            ctx.check(ctx.lib.hpb_set_maxpool_tma(ctx.handle, 1 if tma else 0), "hpb_set_maxpool_tma")
            got = ops.maxpool3x3s2_bf16(ctx, x)
            assert got.shape == ref.shape and torch.equal(got, ref), f"{shape} tma={tma}"
    ctx.check(ctx.lib.hpb_set_maxpool_tma(ctx.handle, 1), "hpb_set_maxpool_tma")
    x = torch.full((1, 64, 6, 6), float("nan"), device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    assert torch.isnan(ops.maxpool3x3s2_bf16(ctx, x)).all()  # NaN propagates like torch


def test_folded_resnet_bf16_close_to_module(ctx):
    """The shipped executor (BN folded, space-to-depth stem, fused epilogues, libhpb200 max-pool) against the plain module."""
    from happypose_b200.megapose.backbones import make_backbone
    from happypose_b200.megapose.fast_resnet import try_fold

    torch.manual_seed(3)
    net = make_backbone("vanilla_resnet34", 9).cuda().eval()
    folded = try_fold(net, torch.bfloat16, ctx)
    assert folded is not None and folded.stem_s2d is not None and folded.fast_pool
    x = torch.rand(4, 9, 240, 320, device="cuda")
    with torch.no_grad():
        ref = net(x)
        got = folded(x).float()
    assert got.shape == ref.shape
    assert (got - ref).abs().max() <= 0.03 * ref.abs().max()  # bf16 end to end


@pytest.mark.parametrize("halo,shape", [(1, (2, 19, 35)), (1, (2, 30, 41)), (1, (5, 123, 163)), (2, (2, 30, 41)), (2, (5, 123, 163)),
                                        (0, (2, 19, 35)), (0, (5, 123, 163))])
def test_stem_tensor_core_conv_matches_float_conv(ctx, halo, shape):
    """hpb_stem_conv4x4_relu_bf16_nhwc (tcgen05 implicit GEMM) vs relu(conv2d + bias) in float32 on the same bf16 operands:
    fp32 accumulation on both sides, so the results differ by summation order + the final bf16 rounding only.  The big shape
    is the real stem (123 x 163 cells): several tiles per CTA, every ring stage and both accumulators reused.  halo = 1 is the
    shipped variant (one box per tile, taps by descriptor offset, staged TMA-store epilogue; 27 x 38 outputs exercise the clipped
    edge tiles), 2 the same with the register-store epilogue, 0 the box-per-tap cross-check."""
    from happypose_b200 import ops

    b, Hz, Wz = shape
    g = torch.Generator(device="cpu").manual_seed(21)
    z = torch.randn(b, 64, Hz, Wz, generator=g).to(torch.bfloat16).cuda().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(64, 64, 4, 4, generator=g) * 0.05).to(torch.bfloat16).cuda().contiguous(memory_format=torch.channels_last)
    bias = torch.randn(64, generator=g).cuda()
    ctx.check(ctx.lib.hpb_set_stem_tc_halo(ctx.handle, halo), "hpb_set_stem_tc_halo")
    try:
        out = ops.stem_conv4x4_relu_bf16(ctx, z, w, bias)
        declined = ops.stem_conv4x4_relu_bf16(ctx, z[:, :, :Hz - 1].contiguous(memory_format=torch.channels_last), w, bias) if not halo else None
    finally:
        ctx.check(ctx.lib.hpb_set_stem_tc_halo(ctx.handle, 1), "hpb_set_stem_tc_halo")
    assert out is not None and out.shape == (b, 64, Hz - 3, Wz - 3) and out.is_contiguous(memory_format=torch.channels_last)
    assert declined is None  # the box-per-tap scheme declines shapes it cannot tile instead of mangling them
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = torch.relu(torch.nn.functional.conv2d(z.float(), w.float(), bias))
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    got = out.float()
    assert (ref > 0).float().mean() > 0.3
    # one bf16 ulp (2^-8 relative) of the result + the fp32 summation noise of 1024 products
    err = (got - ref).abs()
    assert bool((err <= ref.abs() * 2.0 ** -8 + 1e-3).all()), float((err - ref.abs() * 2.0 ** -8).max())
    # and it is the correctly rounded value almost everywhere
    assert float((got == ref.to(torch.bfloat16).float()).float().mean()) > 0.99


@pytest.mark.parametrize("residual", [False, True])
@pytest.mark.parametrize("shape", [(2, 18, 16), (3, 37, 45), (4, 60, 80)])
def test_conv3x3_tensor_core_matches_float_conv(ctx, shape, residual):
    """hpb_conv3x3_bias_relu_bf16_nhwc (tcgen05 implicit GEMM, zero padding = TMA out-of-bounds fill) vs
    relu(conv2d(pad 1) + bias [+ residual]) in float32 on the same bf16 operands.  37 x 45 exercises clipped edge tiles in both
    directions; 60 x 80 is ResNet-34 layer1 at the 240 x 320 render size (several tiles per CTA)."""
    from happypose_b200 import ops

    b, H, W = shape
    g = torch.Generator(device="cpu").manual_seed(31)
    x = torch.randn(b, 64, H, W, generator=g).to(torch.bfloat16).cuda().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(64, 64, 3, 3, generator=g) * 0.05).to(torch.bfloat16).cuda().contiguous(memory_format=torch.channels_last)
    bias = torch.randn(64, generator=g).cuda()
    res = torch.randn(b, 64, H, W, generator=g).to(torch.bfloat16).cuda().contiguous(memory_format=torch.channels_last) if residual else None
    out = ops.conv3x3_bias_relu_bf16(ctx, x, w, bias, res)
    assert out is not None and out.shape == (b, 64, H, W) and out.is_contiguous(memory_format=torch.channels_last)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = torch.nn.functional.conv2d(x.float(), w.float(), bias, padding=1)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    if residual:
        ref = ref + res.float()
    ref = torch.relu(ref)
    got = out.float()
    assert (ref > 0).float().mean() > 0.3
    err = (got - ref).abs()
    assert bool((err <= ref.abs() * 2.0 ** -8 + 1e-3).all()), float((err - ref.abs() * 2.0 ** -8).max())
    assert float((got == ref.to(torch.bfloat16).float()).float().mean()) > 0.99
    assert ops.conv3x3_bias_relu_bf16(ctx, x[:, :, :17].contiguous(memory_format=torch.channels_last), w, bias) is None  # H < 18: declined
