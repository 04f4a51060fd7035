"""happypose/toolbox/lib3d/transform_ops.py (hot-path subset): normalize_T on the GPU; the trivial helpers stay torch."""
import torch

from .. import ops
from .._capi import Context


def normalize_T(T: torch.Tensor) -> torch.Tensor:
    """transform_ops.py:118-120 -> hpb_normalize_T (one launch instead of ~12 eager ops).  float32 on CUDA."""
    return ops.normalize_T(Context.get(T.device if T.is_cuda else None), T)


def transform_pts(T: torch.Tensor, pts: torch.Tensor) -> torch.Tensor:
    """transform_ops.py:28-55."""
    bsz, n_pts = T.shape[0], pts.shape[1]
    assert pts.shape == (bsz, n_pts, 3)
    if T.dim() == 4:
        pts = pts.unsqueeze(1)
        assert T.shape[-2:] == (4, 4)
    elif T.dim() == 3:
        assert T.shape == (bsz, 4, 4)
    else:
        raise ValueError("Unsupported shape for T", T.shape)
    return (T.unsqueeze(-3)[..., :3, :3] @ pts.unsqueeze(-1) + T.unsqueeze(-3)[..., :3, [-1]]).squeeze(-1)


def invert_transform_matrices(T: torch.Tensor) -> torch.Tensor:
    """transform_ops.py:58-67: (R, t) -> (R^T, -R^T t)."""
    R_inv = T[..., :3, :3].transpose(-2, -1)
    T_inv = T.clone()
    T_inv[..., :3, :3] = R_inv
    T_inv[..., :3, [-1]] = -R_inv @ T[..., :3, [-1]]
    return T_inv
