"""Views resident on the device (torch only: importable without libCudaRasterizer.so, so that harnesses timing other
implementations can share the camera plumbing without mapping the library)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
import torch

from .synthetic import Camera


@dataclass
class DeviceCamera:
    """A view with its three small matrices resident on the device."""
    image_width: int
    image_height: int
    FoVx: float
    FoVy: float
    world_view_transform: torch.Tensor
    full_proj_transform: torch.Tensor
    camera_center: torch.Tensor

    @staticmethod
    def from_packed(cam: Camera, packed: torch.Tensor) -> "DeviceCamera":
        """`packed` is a [35] device tensor: viewmatrix | projmatrix | campos (Camera.packed())."""
        return DeviceCamera(cam.image_width, cam.image_height, cam.FoVx, cam.FoVy,
                            packed[0:16].view(4, 4), packed[16:32].view(4, 4), packed[32:35])

    @staticmethod
    def upload(cam: Camera, device) -> "DeviceCamera":
        return DeviceCamera.from_packed(cam, torch.from_numpy(cam.packed()).to(device))


def upload_cameras(cams: Sequence[Camera], device) -> List[DeviceCamera]:
    """One H2D copy for the whole camera set."""
    if not cams:
        return []
    packed = torch.from_numpy(np.stack([c.packed() for c in cams])).to(device)
    return [DeviceCamera.from_packed(c, packed[i]) for i, c in enumerate(cams)]
