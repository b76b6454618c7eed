"""Ray generation with the reference's signatures (mega_nerf/ray_utils.py:6-84) on libmn_b200.so."""
from __future__ import annotations

from typing import List, Optional

import torch

from . import _cabi as K


def get_ray_directions(W: int, H: int, fx: float, fy: float, cx: float, cy: float, center_pixels: bool,
                       device: torch.device) -> torch.Tensor:
    device = torch.device(device)
    h = K.ctx(device)
    out = torch.empty(H, W, 3, device=device, dtype=torch.float32)
    K.check(K.lib().mn_ray_directions(h, W, H, float(fx), float(fy), float(cx), float(cy), int(center_pixels), K.ptr(out),
                                      K.stream_of(device)), h)
    return out


def _rays(directions: torch.Tensor, c2w: torch.Tensor, near: float, far: float, ray_altitude_range: Optional[List[float]],
          batched: bool) -> torch.Tensor:
    dev = directions.device
    h = K.ctx(dev)
    d = K.f32c(directions)
    m = K.f32c(c2w.to(dev))
    n_poses = m.shape[0] if batched else 1
    P = d.numel() // 3 // (n_poses if batched else 1)
    out = torch.empty(*d.shape[:-1], 8, device=dev, dtype=torch.float32)
    has_alt = ray_altitude_range is not None
    K.check(K.lib().mn_rays(h, K.ptr(d), int(batched), K.ptr(m), n_poses, P, float(near), float(far), int(has_alt),
                            float(ray_altitude_range[0]) if has_alt else 0.0,
                            float(ray_altitude_range[1]) if has_alt else 0.0, K.ptr(out), K.stream_of(dev)), h)
    return out


def get_rays(directions: torch.Tensor, c2w: torch.Tensor, near: float, far: float,
             ray_altitude_range: List[float]) -> torch.Tensor:
    """[H,W,3] x [3,4] -> [H,W,8]  (ray_utils.py:21-30)."""
    return _rays(directions, c2w, near, far, ray_altitude_range, False)


def get_rays_batch(directions: torch.Tensor, c2w: torch.Tensor, near: float, far: float,
                   ray_altitude_range: List[float]) -> torch.Tensor:
    """[n,P,3] x [n,3,4] -> [n,P,8]  (ray_utils.py:33-41)."""
    return _rays(directions, c2w, near, far, ray_altitude_range, True)
