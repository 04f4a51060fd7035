"""Torch-tensor front end of the C ABI (include/hpb200.h): one Python function per entry point.

All functions enqueue on torch's current CUDA stream and never synchronise the host.  Inputs are cast to
contiguous float32 / int32 on the context's device; outputs are freshly allocated torch tensors owned by the
caller (the reference returns fresh tensors too, panda3d_batch_renderer.py:245-286).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _capi
from ._capi import Context, ptr, stream_ptr


class KernelTimer:
    """Optional CUDA-event bracket around individual C-ABI launches (bench.py's live roofline measurement).
    Events are recorded on torch's current stream -- the stream the kernels are launched on."""

    def __init__(self):
        self.records = {}
        self.flop_records = {}

    def bracket(self, name: str, algorithmic_bytes: int, fp32_equivalent_bytes: Optional[int] = None):
        """algorithmic_bytes: bytes the launch has to move; fp32_equivalent_bytes: what the reference's float32 layout of
        the same result would be (differs only for the fused bf16 hand-off)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eq = int(algorithmic_bytes if fp32_equivalent_bytes is None else fp32_equivalent_bytes)
        self.records.setdefault(name, []).append((e0, e1, int(algorithmic_bytes), eq))
        return e0, e1

    def bracket_flops(self, name: str, issued_flops: float, algorithmic_flops: float):
        """For a tensor-core launch: issued_flops = what the MMAs it issues compute, algorithmic_flops = the reference op's."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.flop_records.setdefault(name, []).append((e0, e1, float(issued_flops), float(algorithmic_flops)))
        return e0, e1

    def flop_summary(self):
        """name -> dict(launches, ms_total, ms_avg, tflops, algorithmic_tflops, largest{...}); call after torch.cuda.synchronize()."""
        out = {}
        for name, recs in self.flop_records.items():
            ms = [r[0].elapsed_time(r[1]) for r in recs]
            tot = sum(ms)
            big = max(r[2] for r in recs)
            sel = [(m, r) for m, r in zip(ms, recs) if r[2] == big]
            big_ms = sum(m for m, _ in sel) / len(sel)
            out[name] = {"launches": len(recs), "ms_total": tot, "ms_avg": tot / len(recs),
                         "tflops": sum(r[2] for r in recs) / 1e12 / (tot / 1e3) if tot > 0 else 0.0,
                         "algorithmic_tflops": sum(r[3] for r in recs) / 1e12 / (tot / 1e3) if tot > 0 else 0.0,
                         "largest": {"launches": len(sel), "ms_avg": big_ms, "tflops": big / 1e12 / (big_ms / 1e3) if big_ms > 0 else 0.0,
                                     "algorithmic_tflops": sel[0][1][3] / 1e12 / (big_ms / 1e3) if big_ms > 0 else 0.0}}
        return out

    def summary(self):
        """name -> dict(launches, ms_total, ms_avg, bytes_avg, gbps); call after torch.cuda.synchronize()."""
        out = {}
        for name, recs in self.records.items():
            ms = [r[0].elapsed_time(r[1]) for r in recs]
            byts = [r[2] for r in recs]
            eqs = [r[3] for r in recs]
            tot = sum(ms)
            out[name] = {"launches": len(recs), "ms_total": tot, "ms_avg": tot / len(recs), "bytes_avg": sum(byts) / len(recs),
                         "gbps": (sum(byts) / 1e9) / (tot / 1e3) if tot > 0 else 0.0,
                         "fp32_equivalent_gbps": (sum(eqs) / 1e9) / (tot / 1e3) if tot > 0 else 0.0}
            # the launches of the largest size on their own (the step mixes one big coarse launch with many tiny
            # refiner launches that cannot fill the machine)
            big = max(byts)
            sel = [m for m, c in zip(ms, byts) if c == big]
            big_eq = max(q for q, c in zip(eqs, byts) if c == big)
            out[name]["largest"] = {"launches": len(sel), "bytes": big, "ms_avg": sum(sel) / len(sel),
                                    "gbps": big / 1e9 / (sum(sel) / len(sel) / 1e3) if sum(sel) > 0 else 0.0,
                                    "fp32_equivalent_gbps": big_eq / 1e9 / (sum(sel) / len(sel) / 1e3) if sum(sel) > 0 else 0.0}
        return out


_kernel_timer: Optional[KernelTimer] = None


def set_kernel_timer(timer: Optional[KernelTimer]) -> None:
    global _kernel_timer
    _kernel_timer = timer


def kernel_timer_active() -> bool:
    return _kernel_timer is not None


def _f32(t, device) -> torch.Tensor:
    return torch.as_tensor(t).to(device=device, dtype=torch.float32).contiguous()


def _i32(t, device) -> torch.Tensor:
    return torch.as_tensor(t).to(device=device, dtype=torch.int32).contiguous()


# ------------------------------------------------------------------------------------------------
# meshes
# ------------------------------------------------------------------------------------------------
def mesh_upload(ctx: Context, verts_m, faces, normals=None, uv=None, vcolor=None, texture=None) -> int:
    """verts_m: [nv,3] float32 METRES.  Returns the dense mesh id."""
    v = np.ascontiguousarray(np.asarray(verts_m, np.float32))
    f = np.ascontiguousarray(np.asarray(faces, np.int32))
    if v.ndim != 2 or v.shape[1] != 3 or f.ndim != 2 or f.shape[1] != 3:
        raise ValueError("verts must be [nv,3] and faces [nf,3] (triangles only)")
    n = None if normals is None else np.ascontiguousarray(np.asarray(normals, np.float32))
    t = None if uv is None else np.ascontiguousarray(np.asarray(uv, np.float32))
    c = None
    if vcolor is not None:
        c = np.asarray(vcolor, np.uint8)
        if c.shape[1] == 3:
            c = np.concatenate([c, np.full((len(c), 1), 255, np.uint8)], 1)
        c = np.ascontiguousarray(c)
    tex = None if texture is None else np.ascontiguousarray(np.asarray(texture, np.uint8))
    for a, name, width in ((n, "normals", 3), (t, "uv", 2), (c, "vcolor", 4)):
        if a is not None and a.shape != (len(v), width):
            raise ValueError(f"{name} must be [nv,{width}]")
    th, tw, tc = (tex.shape[0], tex.shape[1], tex.shape[2]) if tex is not None else (0, 0, 0)
    mid = ctypes.c_int32(-1)
    rc = ctx.lib.hpb_mesh_upload(
        ctx.handle, v.ctypes.data, None if n is None else n.ctypes.data, None if t is None else t.ctypes.data,
        None if c is None else c.ctypes.data, len(v), f.ctypes.data, len(f),
        None if tex is None else tex.ctypes.data, th, tw, tc, ctypes.byref(mid))
    ctx.check(rc, "hpb_mesh_upload")
    return int(mid.value)


def mesh_get_mip(ctx: Context, mesh_id: int, level: int) -> Optional[np.ndarray]:
    w, h, levels = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    ctx.check(ctx.lib.hpb_mesh_get_mip(ctx.handle, mesh_id, level, None, ctypes.byref(w), ctypes.byref(h), ctypes.byref(levels)), "hpb_mesh_get_mip")
    if levels.value == 0:
        return None
    out = np.empty((h.value, w.value, 4), np.uint8)
    ctx.check(ctx.lib.hpb_mesh_get_mip(ctx.handle, mesh_id, level, out.ctypes.data, None, None, None), "hpb_mesh_get_mip")
    return out


def mesh_closed_sign(ctx: Context, mesh_id: int) -> int:
    """-1 / +1 when the uploaded mesh is a closed, consistently oriented surface (its back faces are skipped), else 0."""
    sign = ctypes.c_int(0)
    ctx.check(ctx.lib.hpb_mesh_closed_sign(ctx.handle, mesh_id, ctypes.byref(sign)), "hpb_mesh_closed_sign")
    return int(sign.value)


def mesh_set_cull(ctx: Context, mesh_id: int, enable: bool) -> None:
    """enable=False forces two-sided rendering of every triangle of the mesh (tests: culling must be invisible)."""
    ctx.check(ctx.lib.hpb_mesh_set_cull(ctx.handle, mesh_id, 1 if enable else 0), "hpb_mesh_set_cull")


# ------------------------------------------------------------------------------------------------
# rasteriser
# ------------------------------------------------------------------------------------------------
def render(
    ctx: Context,
    mesh_ids: torch.Tensor,
    TCO: torch.Tensor,
    K: torch.Tensor,
    resolution: Tuple[int, int],
    ambient: Optional[torch.Tensor] = None,
    render_rgb: bool = True,
    render_normals: bool = False,
    render_depth: bool = False,
    render_binary_mask: bool = False,
    z_near: float = 0.1,
    z_far: float = 10.0,
    out: Optional[torch.Tensor] = None,
    out_channel_offset: int = 0,
    views: int = 1,
    lights: Optional[torch.Tensor] = None,
):
    """Renders b scenes.  Returns (rgb, normals, depth, mask) tensors (None when not requested).

    lights: [b, n_lights, 8] float32 (type 0 point / 1 directional, xyz in the object = world frame, rgb, -): per-pixel Lambert
    shading on top of the ambient term (hpb_render's lights_dev); None = ambient only (the hot path).

    If `out` ([b, C_total, h, w] float32, contiguous) is given, rgb / normals / depth are written into consecutive
    channels of `out` starting at `out_channel_offset` (rgb 3, then normals 3, then depth 1) and the returned
    tensors are views of it -- the rasteriser then writes the network input in place (no torch.cat).
    With views = V > 1 (multi-view refiner), the b = n*V scenes are hypothesis-major and `out` is [n, C_total, h, w]:
    view v of hypothesis i lands in channels offset + v*C_r ... of out[i] (C_r = channels of one render), i.e. the
    layout of render_images_multiview (pose_rigid.py:447-452).  The returned tensors are then None.
    """
    dev = ctx.device
    h, w = int(resolution[0]), int(resolution[1])
    TCO = _f32(TCO, dev).reshape(-1, 16)
    K = _f32(K, dev).reshape(-1, 9)
    b = TCO.shape[0]
    assert K.shape[0] == b, "K and TCO batch sizes differ"
    mesh_ids = _i32(mesh_ids, dev)
    assert mesh_ids.numel() == b, "Need same number of labels as TCO/K batch size"
    amb = None if ambient is None else _f32(ambient, dev).reshape(b, 3)
    lts, n_lights = None, 0
    if lights is not None and lights.numel() > 0:
        lts = _f32(lights, dev)
        assert lts.dim() == 3 and lts.shape[0] == b and lts.shape[2] == 8 and lts.shape[1] <= 8, "lights must be [b, n<=8, 8]"
        n_lights = lts.shape[1]
    flags = (1 if render_rgb else 0) | (2 if render_normals else 0) | (4 if render_depth else 0) | (8 if render_binary_mask else 0)
    rgb = nrm = dep = msk = None
    view_stride = 0
    if out is not None:
        assert b % views == 0
        assert out.is_contiguous() and out.dtype == torch.float32 and out.shape[0] == b // views and tuple(out.shape[2:]) == (h, w)
        c, bs = out_channel_offset, out.stride(0)
        if render_rgb:
            rgb = out[:, c:c + 3]; c += 3
        if render_normals:
            nrm = out[:, c:c + 3]; c += 3
        if render_depth:
            dep = out[:, c:c + 1]; c += 1
        view_stride = (c - out_channel_offset) * h * w
        assert out_channel_offset + views * (c - out_channel_offset) <= out.shape[1]
        strides = (bs, bs, bs)
    else:
        assert views == 1, "views > 1 needs an `out` tensor"

        if render_rgb:
            rgb = torch.empty((b, 3, h, w), dtype=torch.float32, device=dev)
        if render_normals:
            nrm = torch.empty((b, 3, h, w), dtype=torch.float32, device=dev)
        if render_depth:
            dep = torch.empty((b, 1, h, w), dtype=torch.float32, device=dev)
        strides = (3 * h * w, 3 * h * w, h * w)
    if render_binary_mask:
        msk = torch.empty((b, 1, h, w), dtype=torch.bool, device=dev)
    if b > 0:
        ev = None
        if _kernel_timer is not None:
            n_planes = (3 if render_rgb else 0) + (3 if render_normals else 0) + (1 if render_depth else 0)
            ev = _kernel_timer.bracket("hpb_raster_kernel", b * h * w * (4 * n_planes + (1 if render_binary_mask else 0)))
            ev[0].record()
        rc = ctx.lib.hpb_render(
            ctx.handle, ptr(mesh_ids), ptr(TCO), ptr(K), ptr(amb), b, h, w, z_near, z_far, flags,
            ptr(rgb), strides[0], ptr(nrm), strides[1], ptr(dep), strides[2], ptr(msk), h * w,
            views, view_stride if views > 1 else 0, ptr(lts), n_lights, stream_ptr(dev))
        if ev is not None:
            ev[1].record()
        ctx.check(rc, "hpb_render")
    if views > 1:
        return None, None, None, msk
    return rgb, nrm, dep, msk


# ------------------------------------------------------------------------------------------------
# crop
# ------------------------------------------------------------------------------------------------
def crop(
    ctx: Context,
    images: torch.Tensor,
    im_ids: torch.Tensor,
    points: torch.Tensor,
    obj_ids: torch.Tensor,
    K: torch.Tensor,
    TCO: torch.Tensor,
    tCR: torch.Tensor,
    render_size: Tuple[int, int],
    lamb: float = 1.4,
    out: Optional[torch.Tensor] = None,
    tap_bits: int = 32,
):
    """crop_inputs: returns (images_cropped [b,C,h,w], K_crop [b,3,3], boxes_rend [b,4], boxes_crop [b,4]).

    tap_bits = 16 lets the kernel sample an fp16 copy of an RGB frame (hpb_set_crop_tap_precision: the result is the
    float32 crop of fp16(frame), |error| <= 2.5e-4 on [0,1] pixel values); used for the bf16 network hand-off.

    images [n_im,C,H,W] float32 (not expanded per hypothesis), im_ids [b]; points [n_obj,n_pts,3], obj_ids [b].
    With `out` ([b,C_total,h,w]) the crop is written into channels [0,C) of it.
    """
    dev = ctx.device
    images = _f32(images, dev)
    n_im, C, H, W = images.shape
    h, w = int(render_size[0]), int(render_size[1])
    K = _f32(K, dev).reshape(-1, 9)
    TCO = _f32(TCO, dev).reshape(-1, 16)
    tCR = _f32(tCR, dev).reshape(-1, 3)
    b = TCO.shape[0]
    assert K.shape[0] == b and tCR.shape[0] == b
    im_ids = _i32(im_ids, dev)
    obj_ids = _i32(obj_ids, dev)
    points = _f32(points, dev)
    assert points.dim() == 3 and points.shape[2] == 3
    if out is not None:
        assert out.is_contiguous() and out.dtype == torch.float32 and out.shape[0] == b and tuple(out.shape[2:]) == (h, w)
        crops, bs = out[:, :C], out.stride(0)
    else:
        crops = torch.empty((b, C, h, w), dtype=torch.float32, device=dev)
        bs = C * h * w
    K_crop = torch.empty((b, 3, 3), dtype=torch.float32, device=dev)
    boxes_rend = torch.empty((b, 4), dtype=torch.float32, device=dev)
    boxes_crop = torch.empty((b, 4), dtype=torch.float32, device=dev)
    ev = None
    if _kernel_timer is not None:
        ev = _kernel_timer.bracket("hpb_crop", b * C * h * w * 4)
        ev[0].record()
    rc = ctx.lib.hpb_crop(
        ctx.handle, ptr(images), n_im, C, H, W, ptr(im_ids), ptr(points), points.shape[0], points.shape[1], ptr(obj_ids),
        ptr(K), ptr(TCO), ptr(tCR), b, h, w, lamb, ptr(crops), bs, ptr(K_crop), ptr(boxes_rend), ptr(boxes_crop),
        16 if (tap_bits == 16 and C == 3) else 32, stream_ptr(dev))  # tap precision travels with the launch
    if ev is not None:
        ev[1].record()
    ctx.check(rc, "hpb_crop")
    return crops, K_crop, boxes_rend, boxes_crop


def crop_bf16x4(ctx: Context, images, im_ids, points, obj_ids, K, TCO, tCR, render_size, lamb: float = 1.4, tap_bits: int = 32):
    """crop_inputs for RGB frames with the crop delivered as [b,h,w,4] bfloat16 pixels (r,g,b,0) -- the float32 crop of
    `crop` rounded to nearest even, 8 bytes per pixel (hpb_crop_bf16x4).  Feeds render_s2d_bf16.
    Returns (crops [b,h,w,4] bf16, K_crop, boxes_rend, boxes_crop)."""
    dev = ctx.device
    images = _f32(images, dev)
    n_im, C, H, W = images.shape
    assert C == 3, "the bf16x4 crop format is for RGB frames"
    h, w = int(render_size[0]), int(render_size[1])
    K = _f32(K, dev).reshape(-1, 9)
    TCO = _f32(TCO, dev).reshape(-1, 16)
    tCR = _f32(tCR, dev).reshape(-1, 3)
    b = TCO.shape[0]
    assert K.shape[0] == b and tCR.shape[0] == b
    im_ids = _i32(im_ids, dev)
    obj_ids = _i32(obj_ids, dev)
    points = _f32(points, dev)
    crops = torch.empty((b, h, w, 4), dtype=torch.bfloat16, device=dev)
    K_crop = torch.empty((b, 3, 3), dtype=torch.float32, device=dev)
    boxes_rend = torch.empty((b, 4), dtype=torch.float32, device=dev)
    boxes_crop = torch.empty((b, 4), dtype=torch.float32, device=dev)
    ev = None
    if _kernel_timer is not None:
        ev = _kernel_timer.bracket("hpb_crop", b * h * w * 8, fp32_equivalent_bytes=b * 3 * h * w * 4)
        ev[0].record()
    rc = ctx.lib.hpb_crop_bf16x4(
        ctx.handle, ptr(images), n_im, H, W, ptr(im_ids), ptr(points), points.shape[0], points.shape[1], ptr(obj_ids),
        ptr(K), ptr(TCO), ptr(tCR), b, h, w, lamb, ptr(crops), h * w, ptr(K_crop), ptr(boxes_rend), ptr(boxes_crop),
        16 if tap_bits == 16 else 32, stream_ptr(dev))
    if ev is not None:
        ev[1].record()
    ctx.check(rc, "hpb_crop_bf16x4")
    return crops, K_crop, boxes_rend, boxes_crop


def crop_pixels(ctx: Context, images, im_ids, boxes_crop, render_size, out: Optional[torch.Tensor] = None, tap_bits: int = 32) -> torch.Tensor:
    """The resampling half of `crop`: roi_align of the frames at boxes_crop [b,4] (hpb_crop_pixels).  With `out` ([b,C_total,h,w])
    the crop is written into its channels [0,C)."""
    dev = ctx.device
    images = _f32(images, dev)
    n_im, C, H, W = images.shape
    h, w = int(render_size[0]), int(render_size[1])
    boxes = _f32(boxes_crop, dev).reshape(-1, 4)
    b = boxes.shape[0]
    im_ids = _i32(im_ids, dev)
    if out is not None:
        assert out.is_contiguous() and out.dtype == torch.float32 and out.shape[0] == b and tuple(out.shape[2:]) == (h, w)
        crops, bs = out[:, :C], out.stride(0)
    else:
        crops = torch.empty((b, C, h, w), dtype=torch.float32, device=dev)
        bs = C * h * w
    ev = None
    if _kernel_timer is not None:
        ev = _kernel_timer.bracket("hpb_crop", b * C * h * w * 4)
        ev[0].record()
    rc = ctx.lib.hpb_crop_pixels(ctx.handle, ptr(images), n_im, C, H, W, ptr(im_ids), ptr(boxes), b, h, w, ptr(crops), bs,
                                 16 if (tap_bits == 16 and C == 3) else 32, stream_ptr(dev))
    if ev is not None:
        ev[1].record()
    ctx.check(rc, "hpb_crop_pixels")
    return crops


def refiner_prologue(ctx: Context, TCO, K, obj_ids, points_crop, points_mv, image_size, render_size, multiview_type: str, n_views: int,
                     remove_TCO_rendering: bool = False, lamb: float = 1.4):
    """Everything of one refiner iteration in front of the image kernels, in one launch (hpb_refiner_prologue):
    dict(T_norm [b,4,4], tCR [b,3], TCV_O [b,V,4,4], K_crop [b,3,3], boxes_rend [b,4], boxes_crop [b,4], KV_crop [b,V,3,3])."""
    dev = ctx.device
    TCO = _f32(TCO, dev).reshape(-1, 16)
    b = TCO.shape[0]
    K = _f32(K, dev).reshape(-1, 9)
    assert K.shape[0] == b
    obj_ids = _i32(obj_ids, dev)
    pc, pm = _f32(points_crop, dev), _f32(points_mv, dev)
    assert pc.dim() == 3 and pm.dim() == 3 and pc.shape[0] == pm.shape[0]
    if n_views > 1 and multiview_type not in _capi.MV_TYPES:
        raise ValueError(multiview_type)
    mv = _capi.MV_TYPES.get(multiview_type, 0)
    H, W = int(image_size[0]), int(image_size[1])
    h, w = int(render_size[0]), int(render_size[1])
    f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)  # noqa: E731
    out = {"T_norm": f(b, 4, 4), "tCR": f(b, 3), "TCV_O": f(b, n_views, 4, 4), "K_crop": f(b, 3, 3), "boxes_rend": f(b, 4),
           "boxes_crop": f(b, 4), "KV_crop": f(b, n_views, 3, 3)}
    rc = ctx.lib.hpb_refiner_prologue(
        ctx.handle, ptr(TCO), ptr(K), ptr(obj_ids), ptr(pc), pc.shape[0], pc.shape[1], ptr(pm), pm.shape[1], b, H, W, h, w, lamb, mv,
        n_views, int(remove_TCO_rendering), ptr(out["T_norm"]), ptr(out["tCR"]), ptr(out["TCV_O"]), ptr(out["K_crop"]),
        ptr(out["boxes_rend"]), ptr(out["boxes_crop"]), ptr(out["KV_crop"]), stream_ptr(dev))
    ctx.check(rc, "hpb_refiner_prologue")
    return out


def crop_boxes(ctx: Context, image_size, points, obj_ids, K, TCO, tCR, render_size, lamb: float = 1.4):
    """compute_crops_multiview maths: (K_crop, boxes_rend, boxes_crop) without resampling any pixels."""
    dev = ctx.device
    H, W = int(image_size[0]), int(image_size[1])
    h, w = int(render_size[0]), int(render_size[1])
    K = _f32(K, dev).reshape(-1, 9)
    TCO = _f32(TCO, dev).reshape(-1, 16)
    tCR = _f32(tCR, dev).reshape(-1, 3)
    b = TCO.shape[0]
    obj_ids = _i32(obj_ids, dev)
    points = _f32(points, dev)
    K_crop = torch.empty((b, 3, 3), dtype=torch.float32, device=dev)
    boxes_rend = torch.empty((b, 4), dtype=torch.float32, device=dev)
    boxes_crop = torch.empty((b, 4), dtype=torch.float32, device=dev)
    rc = ctx.lib.hpb_crop_boxes(
        ctx.handle, H, W, ptr(points), points.shape[0], points.shape[1], ptr(obj_ids), ptr(K), ptr(TCO), ptr(tCR),
        b, h, w, lamb, ptr(K_crop), ptr(boxes_rend), ptr(boxes_crop), stream_ptr(dev))
    ctx.check(rc, "hpb_crop_boxes")
    return K_crop, boxes_rend, boxes_crop


# ------------------------------------------------------------------------------------------------
# pose maths
# ------------------------------------------------------------------------------------------------
def normalize_T(ctx: Context, T: torch.Tensor) -> torch.Tensor:
    dev = ctx.device
    shape = T.shape
    Tf = _f32(T, dev).reshape(-1, 16)
    out = torch.empty_like(Tf)
    ctx.check(ctx.lib.hpb_normalize_T(ctx.handle, ptr(Tf), Tf.shape[0], ptr(out), stream_ptr(dev)), "hpb_normalize_T")
    return out.reshape(shape)


def pose_update(ctx: Context, TCO, K_crop, pose_outputs, tCR=None, variant: int = _capi.POSE_MEGAPOSE) -> torch.Tensor:
    dev = ctx.device
    TCO = _f32(TCO, dev).reshape(-1, 16)
    b = TCO.shape[0]
    K_crop = _f32(K_crop, dev).reshape(-1, 9)
    width = 7 if variant == _capi.POSE_COSYPOSE_QUAT else 9
    o = _f32(pose_outputs, dev)
    assert o.shape == (b, width), f"pose outputs must be [b,{width}]"
    t = None if tCR is None else _f32(tCR, dev).reshape(b, 3)
    out = torch.empty_like(TCO)
    ctx.check(ctx.lib.hpb_pose_update(ctx.handle, ptr(TCO), ptr(K_crop), ptr(o), ptr(t), b, variant, ptr(out), stream_ptr(dev)), "hpb_pose_update")
    return out.reshape(b, 4, 4)


def tco_init(ctx: Context, variant: int, boxes, K, points=None, obj_ids=None, R=None, z_mean: float = 1.0) -> torch.Tensor:
    dev = ctx.device
    boxes = _f32(boxes, dev)
    assert boxes.dim() == 2 and boxes.shape[-1] == 4
    b = boxes.shape[0]
    K = _f32(K, dev).reshape(-1, 9)
    pts = None if points is None else _f32(points, dev)
    ids = None if obj_ids is None else _i32(obj_ids, dev)
    Rm = None if R is None else _f32(R, dev).reshape(-1, 9)
    out = torch.empty((b, 16), dtype=torch.float32, device=dev)
    n_obj, n_pts = (pts.shape[0], pts.shape[1]) if pts is not None else (0, 0)
    rc = ctx.lib.hpb_tco_init(ctx.handle, variant, ptr(boxes), ptr(pts), n_obj, n_pts, ptr(ids), ptr(K), ptr(Rm), float(z_mean), b, ptr(out), stream_ptr(dev))
    ctx.check(rc, "hpb_tco_init")
    return out.reshape(b, 4, 4)


def multiview(ctx: Context, TCO, tCR, multiview_type: str, n_views: int, remove_TCO_rendering: bool = False) -> torch.Tensor:
    dev = ctx.device
    TCO = _f32(TCO, dev).reshape(-1, 16)
    b = TCO.shape[0]
    tCR = _f32(tCR, dev).reshape(b, 3)
    if n_views > 1 and multiview_type not in _capi.MV_TYPES:
        raise ValueError(multiview_type)
    mv = _capi.MV_TYPES.get(multiview_type, 0)
    out = torch.empty((b, n_views, 4, 4), dtype=torch.float32, device=dev)
    rc = ctx.lib.hpb_multiview(ctx.handle, ptr(TCO), ptr(tCR), b, mv, n_views, int(remove_TCO_rendering), ptr(out), stream_ptr(dev))
    ctx.check(rc, "hpb_multiview")
    return out


def normalize_depth_(ctx: Context, x: torch.Tensor, channels: Sequence[int], tCR, kind: str) -> torch.Tensor:
    """In place on x [b,C,h,w] float32 contiguous: normalises the listed depth channels by tCR_z."""
    if kind not in _capi.DEPTH_NORM:
        raise ValueError(f"Unknown depth_normalization_type = {kind}")
    if kind == "none" or len(channels) == 0:
        return x
    dev = ctx.device
    assert x.is_contiguous() and x.dtype == torch.float32
    b, C, h, w = x.shape
    tCR = _f32(tCR, dev).reshape(b, 3)
    ch = (ctypes.c_int32 * len(channels))(*[int(c) for c in channels])
    rc = ctx.lib.hpb_normalize_depth(ctx.handle, ptr(x), x.stride(0), ctypes.addressof(ch), len(channels), ptr(tCR), b, h, w, _capi.DEPTH_NORM[kind], stream_ptr(dev))
    ctx.check(rc, "hpb_normalize_depth")
    return x


def pack_input_bf16(ctx: Context, x: torch.Tensor, c_padded: int) -> torch.Tensor:
    """x [b,C,h,w] float32 (contiguous) -> [b,c_padded,h,w] bfloat16 in channels_last memory, extra channels zero."""
    dev = ctx.device
    assert x.is_contiguous() and x.dtype == torch.float32 and x.dim() == 4
    b, C, h, w = x.shape
    out = torch.empty((b, c_padded, h, w), dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last)
    rc = ctx.lib.hpb_pack_input_bf16(ctx.handle, ptr(x), x.stride(0), b, C, h, w, ptr(out), c_padded, stream_ptr(dev))
    ctx.check(rc, "hpb_pack_input_bf16")
    return out


def pack_input_s2d_bf16(ctx: Context, x: torch.Tensor, c_padded: int) -> torch.Tensor:
    """x [b,C,H,W] float32 -> z [b,c_padded,H/2+3,W/2+3] bfloat16 channels_last: 2x2 space-to-depth of x zero-padded by 3
    (see hpb_pack_input_s2d_bf16); a 7x7/s2/p3 convolution of x is a 4x4/s1/p0 convolution of z."""
    dev = ctx.device
    assert x.is_contiguous() and x.dtype == torch.float32 and x.dim() == 4
    b, C, H, W = x.shape
    out = torch.empty((b, c_padded, H // 2 + 3, W // 2 + 3), dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last)
    rc = ctx.lib.hpb_pack_input_s2d_bf16(ctx.handle, ptr(x), x.stride(0), b, C, H, W, ptr(out), c_padded, stream_ptr(dev))
    ctx.check(rc, "hpb_pack_input_s2d_bf16")
    return out


def render_s2d_bf16(ctx: Context, mesh_ids: torch.Tensor, TCO: torch.Tensor, K: torch.Tensor, crops: torch.Tensor, c_padded: int,
                    ambient: Optional[torch.Tensor] = None, z_near: float = 0.1, z_far: float = 10.0,
                    out: Optional[torch.Tensor] = None, pad_prezeroed: bool = False) -> torch.Tensor:
    """Renders rgb + normals of b scenes and writes the stem's input directly: z [b,c_padded,h/2+3,w/2+3] bfloat16
    channels_last = pack_input_s2d_bf16(cat(crops, rgb, normals)) without the float32 network input or the packing pass
    (hpb_render_s2d_bf16).  crops: [b,3,h,w] float32 (a contiguous tensor or the first 3 channels of a wider one), or
    [b,h,w,4] bfloat16 (crop_bf16x4).  `out`: a caller-owned result buffer; with pad_prezeroed its channels >= 48 must
    already be zero and stay untouched by anyone else (the kernel then writes only the 32 data bytes per sub-pixel)."""
    dev = ctx.device
    TCO = _f32(TCO, dev).reshape(-1, 16)
    K = _f32(K, dev).reshape(-1, 9)
    b = TCO.shape[0]
    assert K.shape[0] == b, "K and TCO batch sizes differ"
    mesh_ids = _i32(mesh_ids, dev)
    assert mesh_ids.numel() == b
    if crops.dtype == torch.bfloat16:
        assert crops.dim() == 4 and crops.shape[0] == b and crops.shape[3] == 4 and crops.is_contiguous()
        h, w = int(crops.shape[1]), int(crops.shape[2])
        fmt, crops_bs, crop_bytes = _capi.CROPS_BF16X4, h * w, h * w * 8
    else:
        assert crops.dtype == torch.float32 and crops.dim() == 4 and crops.shape[0] == b and crops.shape[1] == 3
        h, w = int(crops.shape[2]), int(crops.shape[3])
        assert crops.stride(3) == 1 and crops.stride(2) == w and crops.stride(1) == h * w, "crop planes must be dense"
        fmt, crops_bs, crop_bytes = _capi.CROPS_F32_PLANAR, crops.stride(0), 3 * h * w * 4
    amb = None if ambient is None else _f32(ambient, dev).reshape(b, 3)
    shape = (b, c_padded, h // 2 + 3, w // 2 + 3)
    if out is None:
        assert not pad_prezeroed, "pad_prezeroed needs a caller-owned, pre-zeroed `out`"
        out = torch.empty(shape, dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last)
    else:
        assert tuple(out.shape) == shape and out.dtype == torch.bfloat16 and out.is_contiguous(memory_format=torch.channels_last)
    ev = None
    if _kernel_timer is not None:
        # bytes this launch moves: the bf16 cells of the stem input it writes (4 x 32 B per cell into a pre-zeroed buffer, else
        # the whole padded cell) + the crop it reads; fp32-equivalent = the 6 float32 planes per view of hpb_render
        cell_bytes = 128 if pad_prezeroed else c_padded * 2
        ev = _kernel_timer.bracket("hpb_raster_kernel", b * ((h // 2 + 3) * (w // 2 + 3) * cell_bytes + crop_bytes),
                                   fp32_equivalent_bytes=b * 6 * h * w * 4)
        ev[0].record()
    rc = ctx.lib.hpb_render_s2d_bf16(ctx.handle, ptr(mesh_ids), ptr(TCO), ptr(K), ptr(amb), b, h, w, z_near, z_far,
                                     ptr(crops), crops_bs, fmt, ptr(out), c_padded, 1 if pad_prezeroed else 0, stream_ptr(dev))
    if ev is not None:
        ev[1].record()
    ctx.check(rc, "hpb_render_s2d_bf16")
    return out


def maxpool3x3s2_bf16(ctx: Context, x: torch.Tensor) -> torch.Tensor:
    """F.max_pool2d(x, 3, 2, 1) for a bfloat16 channels_last [b,C,H,W] tensor (C % 8 == 0); returns channels_last."""
    assert x.dtype == torch.bfloat16 and x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)
    b, C, H, W = x.shape
    out = torch.empty((b, C, (H - 1) // 2 + 1, (W - 1) // 2 + 1), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    rc = ctx.lib.hpb_maxpool3x3s2_bf16_nhwc(ctx.handle, ptr(x), b, H, W, C, ptr(out), stream_ptr(ctx.device))
    ctx.check(rc, "hpb_maxpool3x3s2_bf16_nhwc")
    return out


def stem_k_slice_mask(weight: torch.Tensor) -> int:
    """64-bit mask of the non-zero 16-channel weight slices of a [O,64,4,4] stem weight: bit 4 * (4 * kh + kw) + k."""
    O, C, kh, kw = weight.shape
    assert (C, kh, kw) == (64, 4, 4)
    nz = (weight.detach().float().permute(2, 3, 1, 0).reshape(16, 4, 16 * O) != 0).any(dim=2).cpu().numpy().reshape(-1)
    return int(sum(1 << i for i, v in enumerate(nz) if v))


def stem_conv4x4_relu_bf16(ctx: Context, z: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor,
                           k_slice_mask: int = (1 << 64) - 1) -> Optional[torch.Tensor]:
    """relu(conv2d(z, weight) + bias) for the space-to-depth stem on the tensor cores (tcgen05 implicit GEMM, hpb_stem_tc.cu):
    z [b,64,Hz,Wz] bf16 channels_last, weight [64,64,4,4] bf16 channels_last, bias [64] float32 -> [b,64,Hz-3,Wz-3] bf16
    channels_last.  Returns None when the library does not serve the shape (the caller keeps its cuDNN convolution)."""
    assert z.dtype == torch.bfloat16 and z.dim() == 4 and z.is_contiguous(memory_format=torch.channels_last)
    assert weight.dtype == torch.bfloat16 and weight.is_contiguous(memory_format=torch.channels_last) and tuple(weight.shape[2:]) == (4, 4)
    assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == weight.shape[0]
    b, C, Hz, Wz = z.shape
    O = weight.shape[0]
    assert weight.shape[1] == C
    if C != 64 or O != 64 or Hz < 4 or Wz < 4:
        return None
    out = torch.empty((b, O, Hz - 3, Wz - 3), dtype=torch.bfloat16, device=z.device, memory_format=torch.channels_last)
    ev = None
    if _kernel_timer is not None:
        # issued: tiles of 16 x 8 outputs, one M128 x N64 x K16 MMA per multiplied weight slice; algorithmic: the 7x7 kernel over
        # the (at most 16) real channels per sub-pixel the slices stand for is not known here, so the 4x4x64 form is quoted
        tiles = b * ((Hz - 3 + 15) // 16) * ((Wz - 3 + 7) // 8)
        slices = bin(int(k_slice_mask) & ((1 << 64) - 1)).count("1")
        ev = _kernel_timer.bracket_flops("hpb_stem_tc_kernel", tiles * slices * 2.0 * 128 * 64 * 16, b * (Hz - 3) * (Wz - 3) * 2.0 * O * slices * 16)
        ev[0].record()
    rc = ctx.lib.hpb_stem_conv4x4_relu_bf16_nhwc(ctx.handle, ptr(z), b, Hz, Wz, C, ptr(weight), ptr(bias), O, int(k_slice_mask) & ((1 << 64) - 1), ptr(out), stream_ptr(ctx.device))
    if ev is not None:
        ev[1].record()
    if rc == -4:  # HPB_ENOTFOUND: shape / driver entry point not served
        return None
    ctx.check(rc, "hpb_stem_conv4x4_relu_bf16_nhwc")
    return out


def conv3x3_residual_weight(weight: torch.Tensor) -> torch.Tensor:
    """[64,64,3,3] weight -> the [64 out][10 taps][64 ch] bf16 operand of the residual form of hpb_conv3x3_bias_relu_bf16_nhwc:
    the nine taps followed by a 64 x 64 identity, through which the tensor core adds the residual tile."""
    O, C = weight.shape[:2]
    taps = weight.detach().permute(0, 2, 3, 1).reshape(O, 9, C)
    eye = torch.eye(O, C, dtype=taps.dtype, device=taps.device).reshape(O, 1, C)
    return torch.cat([taps, eye], dim=1).to(torch.bfloat16).contiguous()


def conv3x3_bias_relu_bf16(ctx: Context, x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor,
                           residual: Optional[torch.Tensor] = None, residual_weight: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """relu(conv2d(x, weight, padding=1) + bias [+ residual]) for 64 -> 64 channels on the tensor cores (hpb_conv3x3_tc.cu):
    x, residual [b,64,H,W] bf16 channels_last, weight [64,64,3,3] bf16 channels_last, bias [64] float32 -> [b,64,H,W] bf16
    channels_last.  Returns None when the library does not serve the shape (the caller keeps its cuDNN convolution).
    `residual_weight`: conv3x3_residual_weight(weight), cached by the caller."""
    assert x.dtype == torch.bfloat16 and x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)
    assert weight.dtype == torch.bfloat16 and weight.is_contiguous(memory_format=torch.channels_last) and tuple(weight.shape[2:]) == (3, 3)
    assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == weight.shape[0]
    b, C, H, W = x.shape
    O = weight.shape[0]
    assert weight.shape[1] == C
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and tuple(residual.shape) == (b, O, H, W) and residual.is_contiguous(memory_format=torch.channels_last)
    if C != 64 or O != 64 or H < 18 or W < 16:
        return None
    out = torch.empty((b, O, H, W), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    if residual is not None:  # the residual form multiplies a tenth, identity tap (built here unless the caller caches it)
        weight = residual_weight if residual_weight is not None else conv3x3_residual_weight(weight)
        assert weight.dtype == torch.bfloat16 and tuple(weight.shape) == (O, 10, C) and weight.is_contiguous()
    rc = ctx.lib.hpb_conv3x3_bias_relu_bf16_nhwc(ctx.handle, ptr(x), b, H, W, C, ptr(weight), ptr(bias), O,
                                                 ptr(residual) if residual is not None else None, ptr(out), stream_ptr(ctx.device))
    if rc == -4:
        return None
    ctx.check(rc, "hpb_conv3x3_bias_relu_bf16_nhwc")
    return out


# ------------------------------------------------------------------------------------------------
# ICP depth refiner, input stage
# ------------------------------------------------------------------------------------------------
def icp_points(ctx: Context, depth_measured, depth_rendered, im_ids, K, masks=None, depth_delta_thresh: float = 0.1,
               capacity: Optional[int] = None, return_mask: bool = False, return_index: bool = False):
    """Masks and point clouds of icp_refinement (icp_refiner.py:138-176) for N pose estimates in one launch.
    depth_measured [n_im,H,W], depth_rendered [N,H,W] (or [N,1,H,W]), im_ids [N], K [N,3,3], masks [n_im,H,W] bool/uint8 or
    None (threshold mask).  Returns (points_tgt [N,cap,3], points_src [N,cap,3], counts [N,2] int32[, mask [N,H,W] bool]
    [, index_tgt [N,cap] int32, index_src [N,cap] int32 = linear pixel index of every point]); rows beyond counts are
    undefined.  Nothing synchronises: read `counts` when you need the sizes."""
    dev = ctx.device
    dm = _f32(depth_measured, dev)
    if dm.dim() == 4:
        dm = dm[:, 0]
    dr = _f32(depth_rendered, dev)
    if dr.dim() == 4:
        dr = dr[:, 0]
    dm, dr = dm.contiguous(), dr.contiguous()
    n_im, H, W = dm.shape
    N = dr.shape[0]
    assert tuple(dr.shape[1:]) == (H, W), "rendered and measured depth maps must have the same resolution"
    im_ids = _i32(im_ids, dev)
    K = _f32(K, dev).reshape(-1, 9)
    assert im_ids.numel() == N and K.shape[0] == N
    mk = None
    if masks is not None:
        mk = torch.as_tensor(masks).to(device=dev)
        if mk.dim() == 4:
            mk = mk[:, 0]
        mk = mk.to(torch.uint8).contiguous()
        assert tuple(mk.shape) == (n_im, H, W)
    cap = int(capacity) if capacity is not None else H * W
    pt = torch.empty((N, cap, 3), dtype=torch.float32, device=dev)
    ps = torch.empty((N, cap, 3), dtype=torch.float32, device=dev)
    counts = torch.zeros((N, 2), dtype=torch.int32, device=dev)
    mo = torch.empty((N, H, W), dtype=torch.uint8, device=dev) if return_mask else None
    it = torch.empty((N, cap), dtype=torch.int32, device=dev) if return_index else None
    isrc = torch.empty((N, cap), dtype=torch.int32, device=dev) if return_index else None
    rc = ctx.lib.hpb_icp_points(ctx.handle, ptr(dm), n_im, ptr(dr), ptr(mk), ptr(im_ids), ptr(K), N, H, W, float(depth_delta_thresh), cap,
                                ptr(pt), ptr(ps), ptr(counts), ptr(mo), ptr(it), ptr(isrc), stream_ptr(dev))
    ctx.check(rc, "hpb_icp_points")
    out = (pt, ps, counts)
    if return_mask:
        out += (mo.bool(),)
    if return_index:
        out += (it, isrc)
    return out


# ------------------------------------------------------------------------------------------------
# top-K
# ------------------------------------------------------------------------------------------------
def topk_segmented(ctx: Context, scores, group_ids, n_groups: int, K: int, expected_count: Optional[int] = None) -> torch.Tensor:
    """Row indices (int64, on device) kept by filter_top_pose_estimates, in global descending score order.
    The single host sync is reading the survivor count, which sizes the result; a caller that knows the count (every
    group has at least K rows => n_groups * K) passes `expected_count` and nothing is read back."""
    dev = ctx.device
    s = _f32(scores, dev).reshape(-1)
    g = _i32(group_ids, dev).reshape(-1)
    n = s.numel()
    assert g.numel() == n
    cap = max(1, min(n, max(0, n_groups) * max(0, K)))
    out = torch.empty((cap,), dtype=torch.int64, device=dev)
    cnt = torch.zeros((1,), dtype=torch.int32, device=dev)
    rc = ctx.lib.hpb_topk_segmented(ctx.handle, ptr(s), ptr(g), n, n_groups, K, ptr(out), ptr(cnt), stream_ptr(dev))
    ctx.check(rc, "hpb_topk_segmented")
    if expected_count is not None:
        assert 0 <= expected_count <= cap
        return out[:expected_count]
    return out[: int(cnt.item())]
