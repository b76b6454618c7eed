"""eval_sh with the reference's signature (mega_nerf/spherical_harmonics.py:55-106)."""
from __future__ import annotations

import torch

from . import _cabi as K


def eval_sh(deg: int, sh: torch.Tensor, dirs: torch.Tensor) -> torch.Tensor:
    """sh [..., 3, (deg+1)^2], dirs [..., 3] -> [..., 3] (no sigmoid, as the reference function)."""
    assert 0 <= deg <= 4 and (deg + 1) ** 2 == sh.shape[-1]
    if sh.shape[-2] != 3:
        raise NotImplementedError('the kernel evaluates the 3-channel (rgb) head used by the hot path')
    dev = sh.device
    h = K.ctx(dev)
    nc = (deg + 1) ** 2
    B = sh.numel() // (3 * nc)
    coef = torch.empty(B, 3 * nc + 1, device=dev, dtype=torch.float32)
    coef[:, :3 * nc] = sh.reshape(B, 3 * nc)
    coef[:, 3 * nc] = 0
    d = K.f32c(dirs).reshape(B, 3)
    out = torch.empty(B, 4, device=dev, dtype=torch.float32)
    K.check(K.lib().mn_sh_to_rgb(h, deg, K.ptr(coef), coef.shape[1], K.ptr(d), 3, 1, B, 0, K.ptr(out), K.stream_of(dev)), h)
    return out[:, :3].reshape(*sh.shape[:-2], 3)
