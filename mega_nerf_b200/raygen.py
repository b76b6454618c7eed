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
    # batched poses with ONE shared direction table [P,3] (the only shape the reference's loader passes,
    # filesystem_dataset.py:118): `directions @ c2w[:, :, :3].transpose(1, 2)` broadcasts to [n,P,3]
    shared_dirs = batched and d.dim() == 2
    if batched and not shared_dirs and d.shape[0] != n_poses:
        if d.shape[0] != 1:
            raise RuntimeError(f'get_rays_batch: directions {tuple(d.shape)} do not broadcast against {n_poses} poses')
        d, shared_dirs = d[0], True
    if shared_dirs:
        P = d.shape[0]
        out = torch.empty(n_poses, P, 8, device=dev, dtype=torch.float32)
    else:
        P = d.numel() // 3 // n_poses
        out = torch.empty(*d.shape[:-1], 8, device=dev, dtype=torch.float32)
    has_alt = ray_altitude_range is not None
    K.check(K.lib().mn_rays(h, K.ptr(d), int(batched and not shared_dirs), K.ptr(m), n_poses, P, float(near), float(far),
                            int(has_alt), float(ray_altitude_range[0]) if has_alt else 0.0,
                            float(ray_altitude_range[1]) if has_alt else 0.0, K.ptr(out), K.stream_of(dev)), h)
    return out


def get_rays(directions: torch.Tensor, c2w: torch.Tensor, near: float, far: float,
             ray_altitude_range: List[float]) -> torch.Tensor:
    """[H,W,3] x [3,4] -> [H,W,8]  (ray_utils.py:21-30)."""
    return _rays(directions, c2w, near, far, ray_altitude_range, False)


def get_rays_batch(directions: torch.Tensor, c2w: torch.Tensor, near: float, far: float,
                   ray_altitude_range: List[float]) -> torch.Tensor:
    """[n,P,3] or [P,3] x [n,3,4] -> [n,P,8]  (ray_utils.py:33-41; the loader passes [P,3], filesystem_dataset.py:118)."""
    return _rays(directions, c2w, near, far, ray_altitude_range, True)


def get_rays_pairs(directions: torch.Tensor, c2ws: torch.Tensor, image_index: torch.Tensor, pixel_index: torch.Tensor,
                   near: float, far: float, ray_altitude_range: Optional[List[float]]) -> torch.Tensor:
    """[P,3] directions x [n,3,4] poses, evaluated for M (image, pixel) pairs only -> [M,8].
    Equals `get_rays_batch(directions, c2ws, ...)[image_index, pixel_index]` (what filesystem_dataset.py:109-125 computes
    through the full [n,P,8] product and a host-side gather) without the product."""
    dev = directions.device
    h = K.ctx(dev)
    d = K.f32c(directions).view(-1, 3)
    m = K.f32c(c2ws.to(dev)).view(-1, 3, 4)
    ii = image_index.to(dev, torch.int32).contiguous().view(-1)
    pi = pixel_index.to(dev, torch.int32).contiguous().view(-1)
    if ii.shape != pi.shape:
        raise ValueError('image_index and pixel_index must have the same length')
    M = ii.shape[0]
    out = torch.empty(M, 8, device=dev, dtype=torch.float32)
    has_alt = ray_altitude_range is not None
    K.check(K.lib().mn_rays_pairs(h, K.ptr(d), d.shape[0], K.ptr(m), m.shape[0], K.ptr(ii), K.ptr(pi), M, float(near), float(far),
                                  int(has_alt), float(ray_altitude_range[0]) if has_alt else 0.0,
                                  float(ray_altitude_range[1]) if has_alt else 0.0, K.ptr(out), K.stream_of(dev)), h)
    return out
