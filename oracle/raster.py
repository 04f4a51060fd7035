"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the CPU rasteriser oracle (raster_oracle.c).

Never imported by happypose_b200/.  Callers: tests/, __graft_entry__.smoke(), bench.py
(cpu_baseline / --impl reference).  See raster_oracle.c for the restated reference semantics
(toolbox/renderer/panda3d_scene_renderer.py:55-102,320-390, renderer/types.py:91-137,254-299,
renderer/utils.py:46-79) and for what is / is not pinned against real Panda3D pixels (oracle/figure_pin.py,
tests/test_oracle_figure_pin.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libhpb_oracle.so")

FLAG_RGB, FLAG_NORMALS, FLAG_DEPTH, FLAG_MASK = 1, 2, 4, 8


class _HpoMesh(ctypes.Structure):
    _fields_ = [
        ("n_verts", ctypes.c_int64),
        ("n_faces", ctypes.c_int64),
        ("pos", ctypes.c_void_p),
        ("nrm", ctypes.c_void_p),
        ("uv", ctypes.c_void_p),
        ("vcol", ctypes.c_void_p),
        ("faces", ctypes.c_void_p),
        ("tex", ctypes.c_void_p),
        ("tex_levels", ctypes.c_int32),
        ("tex_w", ctypes.c_void_p),
        ("tex_h", ctypes.c_void_p),
        ("tex_off", ctypes.c_void_p),
        ("cull_sign", ctypes.c_int32),
    ]


def build_oracle(force: bool = False) -> str:
    """Compile raster_oracle.c with the committed Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "raster_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "_build/libhpb_oracle.so"])
    return _LIB_PATH


_lib = None


def _load():
    global _lib
    if _lib is None:
        build_oracle()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.hpo_render_batch.restype = ctypes.c_int
        _lib.hpo_render_batch.argtypes = (
            [ctypes.c_void_p] * 6 + [ctypes.c_int] * 5 + [ctypes.c_float] * 2 + [ctypes.c_uint32] + [ctypes.c_void_p] * 4
        )
        _lib.hpo_mip_downsample.restype = None
        _lib.hpo_mip_downsample.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    return _lib


def vertex_normals(pos: np.ndarray, faces: np.ndarray) -> np.ndarray:
    """Area-weighted smooth normals, used when a mesh file carries none."""
    p = pos.astype(np.float64)
    fn = np.cross(p[faces[:, 1]] - p[faces[:, 0]], p[faces[:, 2]] - p[faces[:, 0]])
    n = np.zeros_like(p)
    for k in range(3):
        np.add.at(n, faces[:, k], fn)
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    n = np.where(ln > 0, n / np.maximum(ln, 1e-300), 0.0)
    return n.astype(np.float32)


def closed_surface_sign(pos: np.ndarray, faces: np.ndarray) -> int:
    """Sign of the screen-space area2 of FRONT faces when the mesh is a closed, consistently oriented surface, else 0.

    Vertices are welded by exact position; every undirected edge of the welded mesh must be used as often forwards
    as backwards (a closed 2-chain) and every connected component must enclose a volume of the same sign.  Then along
    any viewing ray #front hits == #back hits, so a covered pixel is always covered by a front face and skipping back
    faces cannot change the two-sided image (panda3d_scene_renderer.py:102) except where a back face would have won a
    depth tie against the front face it shares a silhouette edge with.  Restates hpb_closed_surface_sign (product).
    """
    p = np.ascontiguousarray(np.asarray(pos, np.float32)) + np.float32(0.0)  # -0.0 -> +0.0
    f = np.asarray(faces, np.int64)
    if not np.isfinite(p).all() or len(f) == 0:
        return 0
    _, wid = np.unique(p, axis=0, return_inverse=True)
    wid = wid.reshape(-1)
    w = wid[f]
    ok = (w[:, 0] != w[:, 1]) & (w[:, 1] != w[:, 2]) & (w[:, 0] != w[:, 2])
    w, fk = w[ok], f[ok]
    if len(w) == 0:
        return 0
    u = np.concatenate([w[:, 0], w[:, 1], w[:, 2]])
    v = np.concatenate([w[:, 1], w[:, 2], w[:, 0]])
    lo, hi = np.minimum(u, v), np.maximum(u, v)
    sgn = np.where(u < v, 1, -1)
    n_w = int(wid.max()) + 1
    bal = np.zeros(0)
    keys, inv = np.unique(lo * n_w + hi, return_inverse=True)
    bal = np.bincount(inv.reshape(-1), weights=sgn, minlength=len(keys))
    if np.any(bal != 0):
        return 0
    # connected components of the welded mesh
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components

    g = coo_matrix((np.ones(len(u)), (u, v)), shape=(n_w, n_w))
    _, comp = connected_components(g, directed=False)
    pd_ = p.astype(np.float64)
    a, b, c = pd_[fk[:, 0]], pd_[fk[:, 1]], pd_[fk[:, 2]]
    det = np.einsum("ij,ij->i", a, np.cross(b, c)) / 6.0
    vol = np.bincount(comp[w[:, 0]], weights=det, minlength=int(comp.max()) + 1)
    used = np.bincount(comp[w[:, 0]], minlength=int(comp.max()) + 1) > 0
    diag = float(np.linalg.norm(pd_.max(0) - pd_.min(0)))
    tol = 1e-9 * diag ** 3
    vol = vol[used]
    if not np.all(np.abs(vol) > tol):
        return 0
    if np.all(vol > 0):
        return -1  # outward winding: front faces project with area2 < 0 in (x right, y down) pixel coordinates
    if np.all(vol < 0):
        return 1
    return 0


def mip_chain(tex_rgb: np.ndarray):
    """RGBA8 box-filter mip chain (level l+1 = max(1, size//2)), flattened into one buffer."""
    lib = _load()
    assert tex_rgb.dtype == np.uint8 and tex_rgb.ndim == 3 and tex_rgb.shape[2] in (3, 4)
    h, w = tex_rgb.shape[:2]
    lvl = np.full((h, w, 4), 255, np.uint8)
    lvl[..., : tex_rgb.shape[2]] = tex_rgb
    lvl = np.ascontiguousarray(lvl)
    levels, ws, hs = [lvl], [w], [h]
    while w > 1 or h > 1:
        nw, nh = max(1, w // 2), max(1, h // 2)
        dst = np.empty((nh, nw, 4), np.uint8)
        lib.hpo_mip_downsample(levels[-1].ctypes.data, w, h, dst.ctypes.data, nw, nh)
        levels.append(dst)
        w, h = nw, nh
        ws.append(w)
        hs.append(h)
    offs = np.cumsum([0] + [l.shape[0] * l.shape[1] for l in levels[:-1]]).astype(np.int64)
    buf = np.concatenate([l.reshape(-1, 4) for l in levels], 0)
    return np.ascontiguousarray(buf), np.asarray(ws, np.int32), np.asarray(hs, np.int32), offs


class OracleMesh:
    """Mesh in metres as the oracle consumes it.

    pos is pre-scaled on the host exactly like the product does at upload:
    float32(float64(v) * float64(scale)) (rigid_mesh_database.py:104-106 scales in float64).
    """

    def __init__(self, verts, faces, normals=None, uv=None, vcolor=None, texture=None, scale: float = 1.0, cull: bool = True):
        self.pos = np.ascontiguousarray((np.asarray(verts, np.float64) * float(scale)).astype(np.float32))
        self.faces = np.ascontiguousarray(np.asarray(faces, np.int32))
        self.closed_sign = closed_surface_sign(self.pos, self.faces)
        self.cull_sign = self.closed_sign if cull else 0
        assert self.pos.ndim == 2 and self.pos.shape[1] == 3 and self.faces.ndim == 2 and self.faces.shape[1] == 3
        if normals is None:
            normals = vertex_normals(self.pos, self.faces)
        self.nrm = np.ascontiguousarray(np.asarray(normals, np.float32))
        self.uv = None if uv is None else np.ascontiguousarray(np.asarray(uv, np.float32))
        self.vcol = None
        if vcolor is not None:
            vc = np.asarray(vcolor, np.uint8)
            if vc.shape[1] == 3:
                vc = np.concatenate([vc, np.full((len(vc), 1), 255, np.uint8)], 1)
            self.vcol = np.ascontiguousarray(vc)
        self.tex = None
        if texture is not None and self.uv is not None:
            self.tex, self.tex_w, self.tex_h, self.tex_off = mip_chain(np.asarray(texture))

    def c_struct(self) -> _HpoMesh:
        s = _HpoMesh()
        s.n_verts, s.n_faces = len(self.pos), len(self.faces)
        s.pos, s.nrm, s.faces = self.pos.ctypes.data, self.nrm.ctypes.data, self.faces.ctypes.data
        s.uv = self.uv.ctypes.data if self.uv is not None else None
        s.vcol = self.vcol.ctypes.data if self.vcol is not None else None
        s.cull_sign = int(self.cull_sign)
        if self.tex is not None:
            s.tex, s.tex_levels = self.tex.ctypes.data, len(self.tex_w)
            s.tex_w, s.tex_h, s.tex_off = self.tex_w.ctypes.data, self.tex_h.ctypes.data, self.tex_off.ctypes.data
        else:
            s.tex, s.tex_levels = None, 0
        return s


def render(
    meshes: Sequence[OracleMesh],
    mesh_ids,
    TCO,
    K,
    resolution,
    ambient=None,
    render_rgb: bool = True,
    render_normals: bool = False,
    render_depth: bool = False,
    render_binary_mask: bool = False,
    z_near: float = 0.1,
    z_far: float = 10.0,
    n_threads: int = 1,
    lights=None,
):
    """Render b hypotheses.  Returns dict(rgb [b,3,h,w] f32, normals, depth [b,1,h,w] f32, mask bool).
    lights: [b, n_lights, 8] (type 0 point / 1 directional, xyz in the object = world frame, rgb, -) or None."""
    lib = _load()
    if render_binary_mask:
        assert render_depth, "Binary mask can only be rendered if depth is rendered"  # panda3d_scene_renderer.py:331-332
    h, w = int(resolution[0]), int(resolution[1])
    TCO = np.ascontiguousarray(np.asarray(TCO, np.float32).reshape(-1, 16))
    K = np.ascontiguousarray(np.asarray(K, np.float32).reshape(-1, 9))
    b = len(TCO)
    assert len(K) == b
    mesh_ids = np.ascontiguousarray(np.asarray(mesh_ids, np.int32))
    assert len(mesh_ids) == b
    arr = (_HpoMesh * len(meshes))(*[m.c_struct() for m in meshes])
    amb = None if ambient is None else np.ascontiguousarray(np.asarray(ambient, np.float32).reshape(b, 3))
    lts = None if lights is None else np.ascontiguousarray(np.asarray(lights, np.float32).reshape(b, -1, 8))
    n_lights = 0 if lts is None else lts.shape[1]
    flags = (FLAG_RGB if render_rgb else 0) | (FLAG_NORMALS if render_normals else 0) | (FLAG_DEPTH if render_depth else 0) | (FLAG_MASK if render_binary_mask else 0)
    rgb = np.empty((b, 3, h, w), np.float32) if render_rgb else None
    nrm = np.empty((b, 3, h, w), np.float32) if render_normals else None
    dep = np.empty((b, 1, h, w), np.float32) if render_depth else None
    msk = np.empty((b, 1, h, w), np.uint8) if render_binary_mask else None

    def ptr(a):
        return None if a is None else a.ctypes.data

    def run(n0, n1):
        return lib.hpo_render_batch(
            ctypes.addressof(arr), mesh_ids.ctypes.data, TCO.ctypes.data, K.ctypes.data, ptr(amb), ptr(lts), n_lights,
            n0, n1, h, w, z_near, z_far, flags, ptr(rgb), ptr(nrm), ptr(dep), ptr(msk),
        )

    n_threads = max(1, min(int(n_threads), b))
    if n_threads == 1:
        rc = run(0, b)
    else:
        # interleaved single-hypothesis jobs keep the threads balanced
        with ThreadPoolExecutor(n_threads) as ex:
            rcs = list(ex.map(lambda n: run(n, n + 1), range(b)))
        rc = min(rcs) if rcs else 0
    if rc != 0:
        raise MemoryError("oracle rasteriser allocation failed")
    return {
        "rgb": rgb,
        "normals": nrm,
        "depth": dep,
        "mask": None if msk is None else msk.astype(bool),
    }
