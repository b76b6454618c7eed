"""Per-image cluster masks of scripts/create_cluster_masks.py (the hot loop at :139-201) on libmn_b200.so:
ray generation + `ray_samples` depths per ray + distance-ratio minimum per centroid in one kernel, instead of
the reference's [rays, samples, clusters] distance tensors."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _cabi as K
from .raygen import get_ray_directions, get_rays


def min_dist_ratios(rays: torch.Tensor, z_steps: torch.Tensor, centroids: torch.Tensor, cluster_2d: bool,
                    boundary_margin: Optional[float] = None):
    """rays [N,8] -> ratios [N,K] (create_cluster_masks.py:155-185); with `boundary_margin` also the uint8 masks
    [K,N] = ratio <= margin (:199-201)."""
    dev = rays.device
    h = K.ctx(dev)
    r = K.f32c(rays).view(-1, 8)
    t = K.f32c(z_steps.to(dev))
    c = K.f32c(centroids.to(dev))
    N, Kc = r.shape[0], c.shape[0]
    ratios = torch.empty(N, Kc, device=dev, dtype=torch.float32)
    mask = torch.empty(Kc, N, device=dev, dtype=torch.uint8) if boundary_margin is not None else None
    K.check(K.lib().mn_cluster_min_dist_ratios(h, K.ptr(r), N, K.ptr(t), t.numel(), K.ptr(c), Kc, int(cluster_2d),
                                               float(boundary_margin if boundary_margin is not None else 0.0),
                                               K.ptr(ratios), K.ptr(mask), K.stream_of(dev)), h)
    return (ratios, mask) if boundary_margin is not None else ratios


def image_cluster_masks(W: int, H: int, intrinsics: Sequence[float], c2w: torch.Tensor, near: float, far: float,
                        ray_altitude_range: List[float], center_pixels: bool, z_steps: torch.Tensor, centroids: torch.Tensor,
                        cluster_2d: bool, boundary_margin: float, device: torch.device) -> torch.Tensor:
    """One image's [K,H,W] bool masks (create_cluster_masks.py:139-201)."""
    device = torch.device(device)
    dirs = get_ray_directions(W, H, intrinsics[0], intrinsics[1], intrinsics[2], intrinsics[3], center_pixels, device)
    rays = get_rays(dirs, c2w.to(device), near, far, ray_altitude_range).view(-1, 8)
    _, mask = min_dist_ratios(rays, z_steps, centroids, cluster_2d, boundary_margin)
    return mask.view(centroids.shape[0], H, W).bool()
