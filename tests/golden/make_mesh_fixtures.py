"""Writes the small on-disk mesh fixtures of tests/test_mesh_io.py into tests/golden/meshes/ (seeded; no reference data):
  ico_ascii.ply        ASCII PLY in the layout of the reference's tests/data/obj_000001.ply (x y z nx ny nz texture_u
                       texture_v, `comment TextureFile ico_tex.png`, `property list uchar int vertex_indices`)
  ico_binary.ply       binary_little_endian PLY with vertex colours, double positions and one quad face
  ico.obj / ico.mtl    Wavefront OBJ with v / vt / vn corners (separate indices, one negative), a quad, map_Kd ico_tex.png
  ico_tex.png          16 x 8 RGB texture
  ico_expected.npz     the arrays a correct reader must return
"""
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.scenes import icosphere  # noqa: E402

OUT = os.path.join(HERE, "meshes")


def main():
    import cv2

    os.makedirs(OUT, exist_ok=True)
    v, f, n = icosphere(1, 50.0)  # 42 vertices, 80 faces, millimetres
    v = np.round(v.astype(np.float64), 4)
    n = np.round(n.astype(np.float64), 5)
    uv = np.round(np.stack([np.arctan2(n[:, 1], n[:, 0]) / (2 * np.pi) + 0.5, np.arccos(np.clip(n[:, 2], -1, 1)) / np.pi], 1), 5)
    rs = np.random.RandomState(0)
    tex = rs.randint(0, 256, (8, 16, 3)).astype(np.uint8)
    cv2.imwrite(os.path.join(OUT, "ico_tex.png"), tex[:, :, ::-1])
    col = rs.randint(0, 256, (len(v), 4)).astype(np.uint8)
    # ---- ASCII PLY (reference layout) ----
    with open(os.path.join(OUT, "ico_ascii.ply"), "w") as fh:
        fh.write("ply\nformat ascii 1.0\ncomment VCGLIB generated\ncomment TextureFile ico_tex.png\n")
        fh.write(f"element vertex {len(v)}\nproperty float x\nproperty float y\nproperty float z\nproperty float nx\nproperty float ny\n"
                 "property float nz\nproperty float texture_u\nproperty float texture_v\n")
        fh.write(f"element face {len(f)}\nproperty list uchar int vertex_indices\nend_header\n")
        for i in range(len(v)):
            fh.write(" ".join(repr(float(x)) for x in (*v[i], *n[i], *uv[i])) + "\n")
        for t in f:
            fh.write(f"3 {t[0]} {t[1]} {t[2]}\n")
    # ---- binary PLY: double positions, uchar colours, first two triangles merged into one quad ----
    quad = [int(f[0][0]), int(f[0][1]), int(f[0][2]), int(f[1][2])]
    with open(os.path.join(OUT, "ico_binary.ply"), "wb") as fh:
        hdr = ("ply\nformat binary_little_endian 1.0\n"
               f"element vertex {len(v)}\nproperty double x\nproperty double y\nproperty double z\n"
               "property uchar red\nproperty uchar green\nproperty uchar blue\nproperty uchar alpha\n"
               f"element face {len(f) - 1}\nproperty list uchar uint vertex_index\nend_header\n")
        fh.write(hdr.encode())
        for i in range(len(v)):
            fh.write(struct.pack("<3d4B", *v[i], *col[i]))
        fh.write(struct.pack("<B4I", 4, *quad))
        for t in f[2:]:
            fh.write(struct.pack("<B3I", 3, *[int(x) for x in t]))
    binary_faces = np.asarray([[quad[0], quad[1], quad[2]], [quad[0], quad[2], quad[3]]] + [list(t) for t in f[2:]], np.int32)
    # ---- OBJ: positions, uvs and normals indexed separately; uv shared by index, normals reversed order ----
    nn = len(v)
    with open(os.path.join(OUT, "ico.mtl"), "w") as fh:
        fh.write("newmtl m0\nKd 1 1 1\nmap_Kd ico_tex.png\n")
    with open(os.path.join(OUT, "ico.obj"), "w") as fh:
        fh.write("mtllib ico.mtl\nusemtl m0\n")
        for p in v:
            fh.write("v " + " ".join(repr(float(x)) for x in p) + "\n")
        for t in uv:
            fh.write("vt " + " ".join(repr(float(x)) for x in t) + "\n")
        for q in n[::-1]:
            fh.write("vn " + " ".join(repr(float(x)) for x in q) + "\n")
        ref = lambda i: f"{i + 1}/{i + 1}/{nn - i}"  # noqa: E731
        fh.write("f " + " ".join(ref(i) for i in quad) + "\n")
        for k, t in enumerate(f[2:]):
            if k == 0:  # negative (relative) indices
                fh.write("f " + " ".join(f"{int(i) - nn}/{int(i) - nn}/{-(int(i) + 1)}" for i in t) + "\n")
            else:
                fh.write("f " + " ".join(ref(int(i)) for i in t) + "\n")
    np.savez(os.path.join(OUT, "ico_expected.npz"), verts=v, faces=f.astype(np.int32), normals=n.astype(np.float32), uv=uv.astype(np.float32),
             texture=tex, vcolor=col, binary_faces=binary_faces)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
