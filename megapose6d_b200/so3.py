"""SO(3) hypothesis grid (576 rotations by default).

Mirrors load_SO3_grid of the reference (src/megapose/utils/transform_utils.py:27-50): the same
xyzw unit quaternions (data assets converted by tools/make_so3_grid.py) turned into rotation
matrices with the unit-quaternion formula the reference obtains from `roma`.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

_DATA = Path(__file__).resolve().parent / "data"


def load_SO3_grid(resolution: int) -> torch.Tensor:
    path = _DATA / f"so3_grid_{resolution}.npy"
    assert path.is_file(), f"File {path} not found"
    q = torch.tensor(np.load(path).tolist())  # float32, as torch.tensor(list of python floats)
    x, y, z, w = q.unbind(-1)
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    rows = [
        torch.stack((1 - (tyy + tzz), txy - twz, txz + twy), -1),
        torch.stack((txy + twz, 1 - (txx + tzz), tyz - twx), -1),
        torch.stack((txz - twy, tyz + twx, 1 - (txx + tyy)), -1),
    ]
    return torch.stack(rows, -2)
