"""Builds mega_nerf_b200 networks from a plain description (spec attributes + per-sub-module state dicts).

Used by bench.py, __graft_entry__.smoke() and the GPU tests to instantiate the SAME seeded random-init weights the CPU checker /
the reference were given (there are no datasets or checkpoints on the GPU box).  Nothing here computes:
it is `NeRF(...)` / `MegaNeRF(...)` / `Cascade(...)` + `load_state_dict`.  `net` is duck-typed: any object with
`kind` ('nerf' | 'cascade' | 'mega'), `spec` (the NeRF constructor arguments of models/nerf.py:46-49 as attributes),
`weights` (list of state dicts with the reference's parameter names) and, for 'mega', `centroids`, `boundary_margin`,
`xyz_real`, `cluster_2d` (models/mega_nerf.py:8-17)."""
from __future__ import annotations

import torch
from torch import nn

from .modules import NeRF, MegaNeRF, Cascade, ShiftedSoftplus


def nerf_from_spec(spec, state_dict) -> NeRF:
    net = NeRF(spec.pos_xyz_dim, spec.pos_dir_dim, spec.layers, list(spec.skip_layers), spec.layer_dim,
               spec.appearance_dim, spec.affine_appearance, spec.appearance_count, spec.rgb_dim, spec.xyz_dim,
               ShiftedSoftplus() if spec.shifted_softplus else nn.ReLU())
    net.load_state_dict(state_dict)
    return net


def build_net(net, device=None, trainable: bool = False) -> nn.Module:
    """-> NeRF / Cascade / MegaNeRF on `device`, in eval mode.  `trainable=False` freezes the parameters so that calls
    outside `no_grad` are not recorded for backward (a recording call always runs the fp32 training kernels)."""
    subs = [nerf_from_spec(net.spec, w) for w in net.weights]
    if net.kind == 'nerf':
        out = subs[0]
    elif net.kind == 'cascade':
        out = Cascade(subs[0], subs[1])
    elif net.kind == 'mega':
        out = MegaNeRF(subs, net.centroids.clone(), net.boundary_margin, net.xyz_real, net.cluster_2d)
    else:
        raise ValueError(f'unknown network kind {net.kind!r}')
    if device is not None:
        out = out.to(device)
    return out.eval().requires_grad_(trainable)
