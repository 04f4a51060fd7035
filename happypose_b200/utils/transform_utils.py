"""SO(3) grid loading (happypose/toolbox/utils/transform_utils.py:24-48)."""
from __future__ import annotations

import os

import numpy as np
import torch

_DATA_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")


def unitquat_to_rotmat(quat: torch.Tensor) -> torch.Tensor:
    """(x, y, z, w) unit quaternions -> rotation matrices, the formula of roma.unitquat_to_rotmat (roma 1.5.0)."""
    x, y, z, w = quat.unbind(-1)
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    R = torch.stack(
        [1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)],
        dim=-1,
    )
    return R.reshape(quat.shape[:-1] + (3, 3))


def load_SO3_grid(resolution: int) -> torch.Tensor:
    """The reference reads megapose/data/data_{resolution}.qua (x y z w rows, generated with
    http://lavalle.pl/software/so3/so3.html); the same numbers ship here as data/so3_grid_{resolution}.npy.
    Returns rotmats [N,3,3] float32."""
    path = os.path.join(_DATA_DIR, f"so3_grid_{resolution}.npy")
    assert os.path.isfile(path), f"File {path} not found"
    quats = torch.tensor(np.load(path), dtype=torch.float32)
    return unitquat_to_rotmat(quats)
