"""ctypes binding of libhpb200.so (include/hpb200.h).  torch is used only for device memory and streams.

There is NO CPU fallback: if the library is missing or the device is not a B200-class GPU, calls raise.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Dict, Optional

import torch

from . import _build

c_void_p, c_int, c_int32, c_int64, c_float, c_uint32 = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_uint32,
)

RENDER_RGB, RENDER_NORMALS, RENDER_DEPTH, RENDER_MASK = 1, 2, 4, 8
POSE_MEGAPOSE, POSE_COSYPOSE_6D, POSE_COSYPOSE_QUAT = 0, 1, 2
TCO_INIT_AUTODEPTH_WITH_R, TCO_INIT_ZUP_AUTODEPTH, TCO_INIT_FROM_BOXES = 0, 1, 2
CROPS_F32_PLANAR, CROPS_BF16X4 = 0, 1
DEPTH_NORM = {"none": 0, "tCR_scale": 1, "tCR_scale_clamp_center": 2, "tCR_center_clamp": 3}
MV_TYPES = {"TCO+front_1view": 0, "TCO+front_3views": 1, "sphere_26views": 2}

# name -> (restype, argtypes); must list every symbol include/hpb200.h declares (tests/test_capi_symbols.py)
SIGNATURES = {
    "hpb_version": (c_int, []),
    "hpb_last_error": (ctypes.c_char_p, []),
    "hpb_create": (c_int, [c_int, ctypes.POINTER(c_void_p)]),
    "hpb_destroy": (c_int, [c_void_p]),
    "hpb_launch_count": (c_int64, [c_void_p]),
    "hpb_workspace_epoch": (c_int64, [c_void_p]),
    "hpb_raster_clipped_scenes": (c_int, [c_void_p, ctypes.POINTER(c_int64), c_int]),
    "hpb_reserve": (c_int, [c_void_p, c_int, c_int, c_int64, c_int64, c_int64]),
    "hpb_mesh_upload": (c_int, [c_void_p] * 5 + [c_int64, c_void_p, c_int64, c_void_p, c_int, c_int, c_int, ctypes.POINTER(c_int32)]),
    "hpb_mesh_count": (c_int, [c_void_p]),
    "hpb_mesh_get_mip": (c_int, [c_void_p, c_int32, c_int, c_void_p, ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "hpb_mesh_closed_sign": (c_int, [c_void_p, c_int32, ctypes.POINTER(c_int)]),
    "hpb_mesh_set_cull": (c_int, [c_void_p, c_int32, c_int]),
    "hpb_render": (c_int, [c_void_p] * 5 + [c_int, c_int, c_int, c_float, c_float, c_uint32,
                           c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int64, c_void_p, c_int,
                           c_void_p]),
    "hpb_crop": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p,
                         c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_int64, c_void_p,
                         c_void_p, c_void_p, c_int, c_void_p]),
    "hpb_set_crop_tap_precision": (c_int, [c_void_p, c_int]),
    "hpb_crop_boxes": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hpb_crop_pixels": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int64,
                                c_int, c_void_p]),
    "hpb_refiner_prologue": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                     c_int, c_int, c_float, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p]),
    "hpb_normalize_T": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "hpb_pose_update": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "hpb_tco_init": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float,
                             c_int, c_void_p, c_void_p]),
    "hpb_multiview": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "hpb_normalize_depth": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "hpb_pack_input_bf16": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "hpb_pack_input_s2d_bf16": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "hpb_render_s2d_bf16": (c_int, [c_void_p] * 5 + [c_int, c_int, c_int, c_float, c_float, c_void_p, c_int64, c_int, c_void_p, c_int,
                                    c_int, c_void_p]),
    "hpb_crop_bf16x4": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_int64, c_void_p,
                                c_void_p, c_void_p, c_int, c_void_p]),
    "hpb_maxpool3x3s2_bf16_nhwc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "hpb_icp_points": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int64,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hpb_set_maxpool_tma": (c_int, [c_void_p, c_int]),
    "hpb_set_stem_tc_halo": (c_int, [c_void_p, c_int]),
    "hpb_conv3x3_bias_relu_bf16_nhwc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "hpb_stem_conv4x4_relu_bf16_nhwc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, ctypes.c_uint64, c_void_p, c_void_p]),
    "hpb_set_crop_tma": (c_int, [c_void_p, c_int]),
    "hpb_topk_segmented": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
}

_lib = None
_lib_lock = threading.Lock()


class HpbError(RuntimeError):
    pass


def load_library(build_if_missing: bool = True) -> ctypes.CDLL:
    """Loads libhpb200.so; raises (never falls back) if it cannot be found or built."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        path = _build.LIB_PATH
        if not os.path.exists(path):
            if not build_if_missing:
                raise HpbError(f"{path} is missing; run `python -m happypose_b200._build`")
            _build.build()
        lib = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError = the library does not export what the header declares
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load_library().hpb_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise HpbError(f"{what} failed ({rc}): {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class Context:
    """One hpb_ctx per device (owns mesh buffers and the rasteriser workspace)."""

    _by_device: Dict[int, "Context"] = {}

    def __init__(self, device: torch.device):
        if not torch.cuda.is_available():
            raise HpbError("happypose_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.device = torch.device(device)
        assert self.device.type == "cuda"
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", index)
        self.lib = load_library()
        h = c_void_p()
        _check(self.lib.hpb_create(index, ctypes.byref(h)), "hpb_create")
        self.handle = h

    @classmethod
    def get(cls, device=None) -> "Context":
        device = torch.device("cuda" if device is None else device)
        index = device.index if device.index is not None else torch.cuda.current_device()
        ctx = cls._by_device.get(index)
        if ctx is None:
            ctx = cls(torch.device("cuda", index))
            cls._by_device[index] = ctx
        return ctx

    def __deepcopy__(self, memo):  # one context per device: models that are deep-copied keep sharing it
        return self

    def __copy__(self):
        return self

    def launch_count(self) -> int:
        return int(self.lib.hpb_launch_count(self.handle))

    def workspace_epoch(self) -> int:
        """Bumped whenever a workspace buffer was replaced by a larger one (captured graphs may want re-capturing)."""
        return int(self.lib.hpb_workspace_epoch(self.handle))

    def clipped_scenes(self, reset: bool = False) -> int:
        """Scenes rendered so far in which the near plane cut the mesh (their cut triangles were dropped, not clipped)."""
        n = c_int64(0)
        _check(self.lib.hpb_raster_clipped_scenes(self.handle, ctypes.byref(n), 1 if reset else 0), "hpb_raster_clipped_scenes")
        return int(n.value)

    def reserve(self, render_size=(0, 0), frame_pixels: int = 0, topk_rows: int = 0, topk_groups: int = 0) -> None:
        """Sizes the workspaces up front (hpb_reserve) so that no launch has to allocate, e.g. under graph capture."""
        _check(self.lib.hpb_reserve(self.handle, int(render_size[0]), int(render_size[1]), int(frame_pixels), int(topk_rows),
                                    int(topk_groups)), "hpb_reserve")

    def check(self, rc: int, what: str) -> None:
        _check(rc, what)
