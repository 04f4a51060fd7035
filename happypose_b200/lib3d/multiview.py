"""happypose/toolbox/lib3d/multiview.py:166-251 on the GPU: the per-sample CPU loop over Panda3D NodePath.lookAt
(multiview.py:28-92) becomes one kernel launch (hpb_multiview); poses never leave the device."""
import numpy as np
import torch

from .. import ops
from .._capi import Context


def make_TCO_multiview(
    TCO: torch.Tensor,
    tCR: torch.Tensor,
    multiview_type: str = "front_3views",
    n_views: int = 4,
    remove_TCO_rendering: bool = False,
    views_inplane_rotations: bool = False,
) -> torch.Tensor:
    """TCO (bsz,4,4), tCR (bsz,3) -> TCV_O (bsz, n_views[*4], 4, 4)."""
    ctx = Context.get(TCO.device if TCO.is_cuda else None)
    if n_views > 1 and multiview_type not in ("TCO+front_1view", "TCO+front_3views", "sphere_26views"):
        raise ValueError(multiview_type)
    TCV_O = ops.multiview(ctx, TCO, tCR, multiview_type, n_views, remove_TCO_rendering).to(TCO.dtype)
    if views_inplane_rotations:  # multiview.py:239-250 (training-time option)
        assert remove_TCO_rendering
        TCV_O = TCV_O.unsqueeze(2).repeat(1, 1, 4, 1, 1)
        for idx, angle in enumerate([np.pi / 2, np.pi, 3 * np.pi / 2]):
            c, s = float(np.cos(angle)), float(np.sin(angle))
            dR = torch.tensor([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]], device=TCV_O.device, dtype=TCV_O.dtype)
            TCV_O[:, :, idx + 1, :3, :3] = dR @ TCV_O[:, :, idx + 1, :3, :3]
        TCV_O = TCV_O.flatten(1, 2)
    return TCV_O
