"""happypose/toolbox/lib3d/cosypose_ops.py:34-62,159-283 and cosypose/lib3d/cosypose_ops.py:18-42 on the GPU."""
import torch

from .. import _capi, ops
from .._capi import Context


def _ctx(t: torch.Tensor) -> Context:
    return Context.get(t.device if t.is_cuda else None)


def _six_d(dRCO: torch.Tensor) -> torch.Tensor:
    # the first two columns of a rotation are its own 6-D representation
    return torch.cat([dRCO[:, :3, 0], dRCO[:, :3, 1]], dim=-1)


def pose_update_with_reference_point(TCO, K, vxvyvz, dRCO, tCR):
    bsz = len(TCO)
    assert TCO.shape[-2:] == (4, 4) and K.shape[-2:] == (3, 3) and dRCO.shape[-2:] == (3, 3)
    assert vxvyvz.shape[-1] == 3 and tCR.shape == (bsz, 3)
    out9 = torch.cat([_six_d(dRCO).to(vxvyvz.dtype), vxvyvz], dim=-1)
    return ops.pose_update(_ctx(TCO), TCO, K, out9, tCR, _capi.POSE_MEGAPOSE)


def apply_imagespace_predictions(TCO, K, vxvyvz, dRCO):
    assert TCO.shape[-2:] == (4, 4) and K.shape[-2:] == (3, 3) and dRCO.shape[-2:] == (3, 3)
    assert vxvyvz.shape[-1] == 3
    out9 = torch.cat([_six_d(dRCO).to(vxvyvz.dtype), vxvyvz], dim=-1)
    return ops.pose_update(_ctx(TCO), TCO, K, out9, None, _capi.POSE_COSYPOSE_6D)


def _points_args(model_points_3d, bsz, obj_ids):
    """Accepts the reference layout [B,N,3] (one point set per row) or a shared [n_obj,N,3] table + obj_ids."""
    if obj_ids is None:
        assert model_points_3d.shape[0] == bsz
        obj_ids = torch.arange(bsz, dtype=torch.int32)
    return model_points_3d, obj_ids


def TCO_init_from_boxes_autodepth_with_R(boxes_2d, model_points_3d, K, R, obj_ids=None):
    assert boxes_2d.shape[-1] == 4 and boxes_2d.dim() == 2
    pts, ids = _points_args(model_points_3d, boxes_2d.shape[0], obj_ids)
    return ops.tco_init(_ctx(boxes_2d), _capi.TCO_INIT_AUTODEPTH_WITH_R, boxes_2d, K, pts, ids, R)


def TCO_init_from_boxes_zup_autodepth(boxes_2d, model_points_3d, K, obj_ids=None):
    assert boxes_2d.shape[-1] == 4 and boxes_2d.dim() == 2
    pts, ids = _points_args(model_points_3d, boxes_2d.shape[0], obj_ids)
    return ops.tco_init(_ctx(boxes_2d), _capi.TCO_INIT_ZUP_AUTODEPTH, boxes_2d, K, pts, ids)


def TCO_init_from_boxes(z_range, boxes, K):
    assert len(z_range) == 2 and boxes.shape[-1] == 4 and boxes.dim() == 2
    z = float(torch.as_tensor(z_range, dtype=torch.float32).mean())
    return ops.tco_init(_ctx(boxes), _capi.TCO_INIT_FROM_BOXES, boxes, K, z_mean=z)
