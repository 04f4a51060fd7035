"""Minimal mesh ingestion (replaces assimp inside Panda3D and trimesh for this path).

Formats: PLY (ascii / binary_little_endian; x y z [nx ny nz] [texture_u texture_v | s t] [red green blue [alpha]],
triangles or polygons, `comment TextureFile <name>`), Wavefront OBJ (v / vn / vt / f, polygons fan-triangulated,
map_Kd from the .mtl), and .npz bundles with arrays verts, faces[, normals, uv, vcolor, texture] (the test fixture
format).  Returns a MeshData in MESH UNITS; scaling to metres happens at upload, in float64, like
rigid_mesh_database.py:104-106.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import numpy as np


@dataclass
class MeshData:
    verts: np.ndarray  # [nv,3] float64 (mesh units)
    faces: np.ndarray  # [nf,3] int32
    normals: Optional[np.ndarray] = None  # [nv,3] float32
    uv: Optional[np.ndarray] = None  # [nv,2] float32
    vcolor: Optional[np.ndarray] = None  # [nv,4] uint8
    texture: Optional[np.ndarray] = None  # [H,W,3|4] uint8


def _load_image(path: str) -> Optional[np.ndarray]:
    if not os.path.exists(path):
        return None
    try:
        import cv2

        im = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if im is None:
            return None
        if im.ndim == 2:
            im = np.stack([im] * 3, -1)
        if im.dtype != np.uint8:
            im = (im.astype(np.float32) / np.iinfo(im.dtype).max * 255).astype(np.uint8)
        im = im[:, :, [2, 1, 0] + ([3] if im.shape[2] == 4 else [])]
        return np.ascontiguousarray(im)
    except ImportError:
        from PIL import Image

        return np.ascontiguousarray(np.asarray(Image.open(path).convert("RGB")))


_PLY_TYPES = {
    "char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4", "float": "f4", "double": "f8",
    "int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2", "int32": "i4", "uint32": "u4", "float32": "f4", "float64": "f8",
}


def _triangulate(polys) -> np.ndarray:
    tris = []
    for p in polys:
        for k in range(1, len(p) - 1):
            tris.append((p[0], p[k], p[k + 1]))
    return np.asarray(tris, np.int32).reshape(-1, 3)


def load_ply(path: str) -> MeshData:
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements, texfile = None, [], None
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii", "replace").strip().split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "comment" and len(tok) >= 3 and tok[1] == "TextureFile":
                texfile = tok[2]
            elif tok[0] == "element":
                elements.append({"name": tok[1], "count": int(tok[2]), "props": []})
            elif tok[0] == "property":
                if tok[1] == "list":
                    elements[-1]["props"].append(("list", tok[2], tok[3], tok[4]))
                else:
                    elements[-1]["props"].append((tok[1], tok[2]))
            elif tok[0] == "end_header":
                break
        if fmt not in ("ascii", "binary_little_endian"):
            raise ValueError(f"{path}: unsupported PLY format {fmt}")
        vdata, polys = None, None
        for el in elements:
            names = [p[-1] for p in el["props"]]
            if el["name"] == "vertex":
                if fmt == "ascii":
                    arr = np.loadtxt(f, max_rows=el["count"], dtype=np.float64, ndmin=2)
                else:
                    dt = np.dtype([(p[1], "<" + _PLY_TYPES[p[0]]) for p in el["props"]])
                    raw = np.frombuffer(f.read(dt.itemsize * el["count"]), dt)
                    arr = np.stack([raw[n].astype(np.float64) for n in names], 1)
                vdata = {n: arr[:, i] for i, n in enumerate(names)}
            elif el["name"] == "face":
                polys = []
                if fmt == "ascii":
                    for _ in range(el["count"]):
                        tok = f.readline().split()
                        k = int(tok[0])
                        polys.append([int(t) for t in tok[1:1 + k]])
                else:
                    for _ in range(el["count"]):
                        rec = None
                        for p in el["props"]:
                            if p[0] == "list":
                                k = int(np.frombuffer(f.read(np.dtype(_PLY_TYPES[p[1]]).itemsize), "<" + _PLY_TYPES[p[1]])[0])
                                it = np.dtype(_PLY_TYPES[p[2]]).itemsize
                                vals = np.frombuffer(f.read(it * k), "<" + _PLY_TYPES[p[2]])
                                if p[3] in ("vertex_indices", "vertex_index"):
                                    rec = vals.astype(np.int64).tolist()
                            else:
                                f.read(np.dtype(_PLY_TYPES[p[0]]).itemsize)
                        polys.append(rec)
            else:  # skip unknown elements (ascii only needs line skipping)
                if fmt == "ascii":
                    for _ in range(el["count"]):
                        f.readline()
                else:
                    raise ValueError(f"{path}: unsupported binary element {el['name']}")
    if vdata is None or polys is None:
        raise ValueError(f"{path}: PLY needs vertex and face elements")
    verts = np.stack([vdata["x"], vdata["y"], vdata["z"]], 1)
    normals = np.stack([vdata["nx"], vdata["ny"], vdata["nz"]], 1).astype(np.float32) if "nx" in vdata else None
    uv = None
    for a, b in (("texture_u", "texture_v"), ("s", "t"), ("u", "v")):
        if a in vdata:
            uv = np.stack([vdata[a], vdata[b]], 1).astype(np.float32)
            break
    vcolor = None
    if "red" in vdata:
        alpha = vdata["alpha"] if "alpha" in vdata else np.full(len(verts), 255.0)
        vcolor = np.stack([vdata["red"], vdata["green"], vdata["blue"], alpha], 1).astype(np.uint8)
    texture = _load_image(os.path.join(os.path.dirname(path), texfile)) if texfile else None
    if texture is None and uv is not None:
        texture = _load_image(os.path.splitext(path)[0] + ".png")
    return MeshData(verts, _triangulate(polys), normals, uv, vcolor, texture)


def load_obj(path: str) -> MeshData:
    v, vn, vt, corners, polys, mtllib = [], [], [], {}, [], None
    out_v, out_n, out_t = [], [], []
    with open(path) as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                v.append([float(x) for x in tok[1:4]])
            elif tok[0] == "vn":
                vn.append([float(x) for x in tok[1:4]])
            elif tok[0] == "vt":
                vt.append([float(x) for x in tok[1:3]])
            elif tok[0] == "mtllib":
                mtllib = tok[1]
            elif tok[0] == "f":
                poly = []
                for c in tok[1:]:
                    parts = (c.split("/") + ["", ""])[:3]
                    key = tuple(int(p) if p else 0 for p in parts)
                    key = tuple(k + (len(src) + 1 if k < 0 else 0) for k, src in zip(key, (v, vt, vn)))
                    if key not in corners:  # one output vertex per distinct (v, vt, vn) corner
                        corners[key] = len(out_v)
                        out_v.append(v[key[0] - 1])
                        out_t.append(vt[key[1] - 1] if key[1] else None)
                        out_n.append(vn[key[2] - 1] if key[2] else None)
                    poly.append(corners[key])
                polys.append(poly)
    if not out_v:
        raise ValueError(f"{path}: no faces")
    normals = np.asarray(out_n, np.float32) if all(n is not None for n in out_n) else None
    uv = np.asarray(out_t, np.float32) if all(t is not None for t in out_t) else None
    texture = None
    if mtllib and uv is not None:
        mpath = os.path.join(os.path.dirname(path), mtllib)
        if os.path.exists(mpath):
            with open(mpath) as f:
                for line in f:
                    tok = line.split()
                    if tok and tok[0] == "map_Kd":
                        texture = _load_image(os.path.join(os.path.dirname(path), tok[-1]))
                        break
    return MeshData(np.asarray(out_v, np.float64), _triangulate(polys), normals, uv, None, texture)


def load_npz(path: str) -> MeshData:
    d = np.load(path)
    g = lambda k: d[k] if k in d.files else None  # noqa: E731
    return MeshData(d["verts"].astype(np.float64), d["faces"].astype(np.int32), g("normals"), g("uv"), g("vcolor"), g("texture"))


def load_mesh(path) -> MeshData:
    path = str(path)
    ext = os.path.splitext(path)[1].lower()
    if ext == ".ply":
        return load_ply(path)
    if ext == ".obj":
        return load_obj(path)
    if ext == ".npz":
        return load_npz(path)
    raise ValueError(f"unsupported mesh format: {path}")


def hpr_matrix(ypr_offset_deg) -> np.ndarray:
    """Panda3D NodePath.setHpr(h, p, r) as a rotation of column vectors in the mesh frame (z up):
    heading about +Z, then pitch about +X, then roll about +Y (panda3d_scene_renderer.py:211-217)."""
    h, p, r = (np.deg2rad(float(a)) for a in ypr_offset_deg)
    Rz = np.array([[np.cos(h), -np.sin(h), 0], [np.sin(h), np.cos(h), 0], [0, 0, 1]])
    Rx = np.array([[1, 0, 0], [0, np.cos(p), -np.sin(p)], [0, np.sin(p), np.cos(p)]])
    Ry = np.array([[np.cos(r), 0, np.sin(r)], [0, 1, 0], [-np.sin(r), 0, np.cos(r)]])
    return Rz @ Rx @ Ry
