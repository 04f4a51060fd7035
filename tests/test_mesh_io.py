"""CPU: on-disk mesh formats -> MeshData (happypose_b200/io/mesh_io.py; replaces assimp inside Panda3D,
toolbox/renderer/panda3d_scene_renderer.py:206-219, and trimesh, toolbox/lib3d/rigid_mesh_database.py:52-78), and the
BOP / GSO object-folder readers (toolbox/datasets/bop_object_datasets.py:31-62, gso_dataset.py:31-83).
Fixtures: tests/golden/meshes/ (tests/golden/make_mesh_fixtures.py)."""
import json
import os

import numpy as np
import pytest

from happypose_b200.io import mesh_io

HERE = os.path.dirname(os.path.abspath(__file__))
MESHES = os.path.join(HERE, "golden", "meshes")


@pytest.fixture(scope="module")
def expected():
    d = np.load(os.path.join(MESHES, "ico_expected.npz"))
    return {k: d[k] for k in d.files}


def test_ascii_ply_in_the_reference_layout(expected):
    """Same header layout as the reference's tests/data/obj_000001.ply: normals, texture_u/v, `comment TextureFile`."""
    m = mesh_io.load_mesh(os.path.join(MESHES, "ico_ascii.ply"))
    np.testing.assert_array_equal(m.verts, expected["verts"])
    np.testing.assert_array_equal(m.faces, expected["faces"])
    np.testing.assert_array_equal(m.normals, expected["normals"])
    np.testing.assert_array_equal(m.uv, expected["uv"])
    assert m.vcolor is None
    np.testing.assert_array_equal(m.texture, expected["texture"])  # the file named by the TextureFile comment, as RGB


def test_binary_ply_with_colours_doubles_and_a_quad(expected):
    m = mesh_io.load_mesh(os.path.join(MESHES, "ico_binary.ply"))
    np.testing.assert_array_equal(m.verts, expected["verts"])
    np.testing.assert_array_equal(m.faces, expected["binary_faces"])  # the quad is fan-triangulated
    np.testing.assert_array_equal(m.vcolor, expected["vcolor"])
    assert m.normals is None and m.uv is None and m.texture is None


def test_obj_with_mtl_texture_separate_indices_and_negative_indices(expected):
    m = mesh_io.load_mesh(os.path.join(MESHES, "ico.obj"))
    # every (v, vt, vn) corner triple here maps vertex i to itself, so the welded output is a permutation of the source
    assert len(m.verts) == len(expected["verts"]) and len(m.faces) == len(expected["binary_faces"])
    order = {tuple(np.round(v, 4)): i for i, v in enumerate(expected["verts"])}
    perm = np.array([order[tuple(np.round(v, 4))] for v in m.verts])
    np.testing.assert_array_equal(m.verts, expected["verts"][perm])
    np.testing.assert_array_equal(m.normals, expected["normals"][perm])
    np.testing.assert_array_equal(m.uv, expected["uv"][perm])
    np.testing.assert_array_equal(perm[m.faces], expected["binary_faces"])
    np.testing.assert_array_equal(m.texture, expected["texture"])  # map_Kd of the .mtl


def test_reference_test_mesh_round_trips_through_ascii_ply(tmp_path, can_mesh_arrays):
    """The reference's obj_000001 (9 951 vertices, 15 728 triangles) written in its own PLY layout and read back."""
    d = can_mesh_arrays
    path = tmp_path / "obj_000001.ply"
    with open(path, "w") as fh:
        fh.write("ply\nformat ascii 1.0\ncomment TextureFile obj_000001.png\n")
        fh.write(f"element vertex {len(d['verts'])}\nproperty float x\nproperty float y\nproperty float z\nproperty float nx\n"
                 "property float ny\nproperty float nz\nproperty float texture_u\nproperty float texture_v\n")
        fh.write(f"element face {len(d['faces'])}\nproperty list uchar int vertex_indices\nend_header\n")
        rows = np.concatenate([d["verts"], d["normals"], d["uv"]], 1).astype(np.float32)
        for r in rows:
            fh.write(" ".join(repr(float(x)) for x in r) + "\n")
        for t in d["faces"]:
            fh.write(f"3 {t[0]} {t[1]} {t[2]}\n")
    import cv2

    cv2.imwrite(str(tmp_path / "obj_000001.png"), np.ascontiguousarray(d["texture"][:, :, ::-1]))
    m = mesh_io.load_mesh(str(path))
    np.testing.assert_array_equal(m.verts.astype(np.float32), d["verts"].astype(np.float32))
    np.testing.assert_array_equal(m.faces, d["faces"])
    np.testing.assert_array_equal(m.normals, d["normals"].astype(np.float32))
    np.testing.assert_array_equal(m.uv, d["uv"].astype(np.float32))
    np.testing.assert_array_equal(m.texture, d["texture"])


def test_bop_and_gso_object_folders(tmp_path):
    import shutil

    from happypose_b200.datasets.bop_object_datasets import BOPObjectDataset
    from happypose_b200.datasets.gso_dataset import GoogleScannedObjectDataset
    from happypose_b200.lib3d.rigid_mesh_database import MeshDataBase

    bop = tmp_path / "bop_models"
    bop.mkdir()
    for k in (1, 5):
        shutil.copy(os.path.join(MESHES, "ico_ascii.ply"), bop / f"obj_{k:06d}.ply")
    shutil.copy(os.path.join(MESHES, "ico_tex.png"), bop / "ico_tex.png")
    (bop / "models_info.json").write_text(json.dumps({
        "1": {"diameter": 100.0, "symmetries_continuous": [{"axis": [0, 0, 1], "offset": [0, 0, 0]}]},
        "5": {"diameter": 100.0, "symmetries_discrete": [list(np.eye(4).reshape(-1))]}}))
    ds = BOPObjectDataset(bop, label_format="ycbv-{label}")
    assert [o.label for o in ds.list_objects] == ["ycbv-obj_000001", "ycbv-obj_000005"]
    assert ds[0].mesh_units == "mm" and ds[0].scale == 0.001 and ds[0].is_symmetric and ds[1].is_symmetric
    db = MeshDataBase.from_object_ds(ds)
    assert db.infos["ycbv-obj_000005"]["n_points"] == 42 if hasattr(db, "infos") and "n_points" in db.infos["ycbv-obj_000005"] else True
    root = tmp_path / "gso"
    for oid in ("Mug_A", "Shoe_B", "Broken_C"):
        (root / "models_normalized" / oid / "meshes").mkdir(parents=True)
        for f in ("ico.obj", "ico.mtl", "ico_tex.png"):
            shutil.copy(os.path.join(MESHES, f), root / "models_normalized" / oid / "meshes" / ("model.obj" if f == "ico.obj" else f))
    (root / "invalid_meshes.json").write_text(json.dumps(["Broken_C"]))
    gso = GoogleScannedObjectDataset(root, split="normalized")
    assert [o.label for o in gso.list_objects] == ["gso_Mug_A", "gso_Shoe_B"] and gso[0].scaling_factor == 0.1 and gso[0].scale == 0.1
    m = mesh_io.load_mesh(gso[0].mesh_path)
    assert m.texture is not None and len(m.verts) == 42
