"""Regenerates the DATA fixtures that come from the reference's own data files.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_fixtures_from_reference_data.py

Outputs
  tests/golden/obj_000001.npz          mesh of /root/reference/tests/data/obj_000001.ply (ASCII PLY:
                                       x y z nx ny nz texture_u texture_v; 9951 verts, 15728 triangles,
                                       mesh units = mm) + its texture obj_000001.png area-resampled from
                                       4096^2 to 512^2 RGB so the fixture stays small.
  happypose_b200/data/so3_grid_{72,512,576,4608}.npy
                                       the SO(3) grids happypose/pose_estimators/megapose/data/data_*.qua
                                       (x y z w text rows, toolbox/utils/transform_utils.py:24-48) as float64.
"""
import os

import cv2
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def read_ascii_ply(path):
    with open(path) as f:
        assert f.readline().strip() == "ply"
        nv = nf = 0
        while True:
            line = f.readline().strip()
            if line.startswith("element vertex"):
                nv = int(line.split()[-1])
            elif line.startswith("element face"):
                nf = int(line.split()[-1])
            elif line == "end_header":
                break
        v = np.loadtxt(f, max_rows=nv, dtype=np.float64)
        fc = np.loadtxt(f, max_rows=nf, dtype=np.int64)
    assert (fc[:, 0] == 3).all()
    return v, fc[:, 1:4]


def main():
    v, faces = read_ascii_ply(f"{REF}/tests/data/obj_000001.ply")
    tex = cv2.imread(f"{REF}/tests/data/obj_000001.png", cv2.IMREAD_COLOR)[:, :, ::-1]
    tex = cv2.resize(tex, (512, 512), interpolation=cv2.INTER_AREA)
    np.savez_compressed(
        f"{HERE}/obj_000001.npz",
        verts=v[:, 0:3].astype(np.float32),
        normals=v[:, 3:6].astype(np.float32),
        uv=v[:, 6:8].astype(np.float32),
        faces=faces.astype(np.int32),
        texture=np.ascontiguousarray(tex),
    )
    for n in (72, 512, 576, 4608):
        q = np.loadtxt(f"{REF}/happypose/pose_estimators/megapose/data/data_{n}.qua", dtype=np.float64)
        assert q.ndim == 2 and q.shape[1] == 4  # note: the reference's data_512.qua holds 576 rows
        np.save(f"{ROOT}/happypose_b200/data/so3_grid_{n}.npy", q)


if __name__ == "__main__":
    main()
