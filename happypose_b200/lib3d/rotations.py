"""happypose/toolbox/lib3d/rotations.py:22-36 (hot-path subset)."""
import torch

from .. import ops
from .._capi import Context


def compute_rotation_matrix_from_ortho6d(poses: torch.Tensor) -> torch.Tensor:
    """6-D -> SO(3) (Zhou et al.): x = a/|a|, z = (x X b)/|x X b|, y = z X x, columns (x, y, z).
    Runs the Gram-Schmidt of hpb_normalize_T on a [b,4,4] buffer whose first two columns are a and b."""
    assert poses.shape[-1] == 6
    flat = poses.reshape(-1, 6)
    T = torch.zeros(flat.shape[0], 4, 4, dtype=torch.float32, device=flat.device)
    T[:, :3, 0] = flat[:, 0:3]
    T[:, :3, 1] = flat[:, 3:6]
    R = ops.normalize_T(Context.get(poses.device if poses.is_cuda else None), T)[:, :3, :3]
    return R.reshape(poses.shape[:-1] + (3, 3))
