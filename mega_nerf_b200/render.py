"""render_rays() with the reference's signature and result dictionary (mega_nerf/rendering.py:15-173),
orchestrating the sm_100a kernels of libmn_b200.so.  Host syncs happen only where the reference has
them too (sphere check :412, background ray selection :37).

When autograd is recording and a network parameter requires grad (the reference's training step,
runner.py:346-378), the model queries and the compositing go through mega_nerf_b200/autograd.py and the
returned rgb_* / bg_lambda_* carry a graph whose backward runs mn_composite_backward / mn_model_backward
(SURVEY.md §8f-1).  Otherwise nothing is recorded.
"""
from __future__ import annotations

import ctypes as C
import os
from argparse import Namespace
from typing import Callable, Dict, Optional, Tuple

import torch
from torch import nn

from . import _cabi as K
from . import autograd as AG
from .modules import NeRF, MegaNeRF, Cascade, RayRows

TO_COMPOSITE = ('rgb', 'depth')


def _unwrap(m: Optional[nn.Module]):
    if m is None:
        return None
    return m.module if hasattr(m, 'module') and not isinstance(m, (NeRF, MegaNeRF, Cascade)) else m


class _Stage:
    """Thin typed wrappers over the stage entry points, bound to one device/stream."""

    def __init__(self, device: torch.device):
        self.dev = device
        self.L = K.lib()
        self.h = K.ctx(device)

    @property
    def st(self):
        return K.stream_of(self.dev)

    def new(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, device=self.dev, dtype=dtype)

    def sample_coarse(self, rays, far, steps, rand, perturb, N, S):
        z, xyz = self.new(N, S), self.new(N, S, 3)
        K.check(self.L.mn_sample_coarse(self.h, K.ptr(rays), K.ptr(far), K.ptr(steps), K.ptr(rand), float(perturb), N, S,
                                        K.ptr(z), K.ptr(xyz), self.st), self.h)
        return z, xyz

    def stratify(self, z1d, rand, perturb, N, S):
        out = self.new(N, S)
        K.check(self.L.mn_stratify(self.h, K.ptr(z1d), 0, K.ptr(rand), float(perturb), N, S, K.ptr(out), self.st), self.h)
        return out

    def points_from_z(self, rays, z):
        N, S = z.shape
        xyz = self.new(N, S, 3)
        K.check(self.L.mn_points_from_z(self.h, K.ptr(rays), K.ptr(z), N, S, K.ptr(xyz), self.st), self.h)
        return xyz

    def sample_pdf(self, z_coarse, weights, u, F):
        N, S = z_coarse.shape
        out = self.new(N, F)
        ustride = 0 if u.dim() == 1 else F
        K.check(self.L.mn_sample_pdf(self.h, K.ptr(z_coarse), K.ptr(weights), weights.shape[1], None, K.ptr(u), ustride,
                                     N, S, F, K.ptr(out), None, None, self.st), self.h)
        return out

    def sort_cat(self, a, b, descending=False):
        N = a.shape[0]
        out = self.new(N, a.shape[1] + b.shape[1])
        K.check(self.L.mn_sort_cat(self.h, K.ptr(a), a.shape[1], K.ptr(b), b.shape[1], N, int(descending), K.ptr(out),
                                   self.st), self.h)
        return out

    def composite(self, raw, z, dreal, raw2, z2, dreal2, last_delta, flip, want_w, want_rgb, want_depth, want_var,
                  want_lambda):
        N, S = z.shape
        S2 = 0 if z2 is None else z2.shape[1]
        w = self.new(N, S + S2) if want_w else None
        rgb = self.new(N, 3) if want_rgb else None
        depth = self.new(N) if want_depth else None
        var = self.new(N) if want_var else None
        lam = self.new(N) if want_lambda else None
        K.check(self.L.mn_composite(self.h, K.ptr(raw), K.ptr(z), K.ptr(dreal), S, K.ptr(raw2), K.ptr(z2), K.ptr(dreal2), S2,
                                    K.ptr(last_delta), N, int(flip), K.ptr(w), K.ptr(rgb), K.ptr(depth), K.ptr(var),
                                    K.ptr(lam), self.st), self.h)
        return w, rgb, depth, var, lam

    def intersect_sphere(self, rays, center, radius):
        N = rays.shape[0]
        out = self.new(N)
        K.check(self.L.mn_intersect_sphere(self.h, K.ptr(rays), K.ptr(center), K.ptr(radius), N, K.ptr(out), self.st), self.h)
        # the reference raises from a host-side `.any()` here (rendering.py:412-414)
        K.check(self.L.mn_check_status(self.h, self.st), self.h)
        return out

    def points_outside(self, rays, ids, depth, center, radius, real, c2d):
        n, S = depth.shape
        pts = self.new(n, S, 7 if real else 4)
        dreal = self.new(n, S)
        K.check(self.L.mn_points_outside(self.h, K.ptr(rays), K.ptr(ids), K.ptr(depth), K.ptr(center), K.ptr(radius), n, S,
                                         int(real), int(c2d), K.ptr(pts), K.ptr(dreal), self.st), self.h)
        return pts, dreal

    def sh_to_rgb(self, deg, coef, dirs, S):
        B = coef.shape[0]
        out = self.new(B, 4)
        K.check(self.L.mn_sh_to_rgb(self.h, deg, K.ptr(coef), coef.shape[1], K.ptr(dirs), dirs.stride(0), S, B, 1,
                                    K.ptr(out), self.st), self.h)
        return out


def _query(sg: _Stage, net: nn.Module, hparams: Namespace, typ: str, xyz: torch.Tensor, dirs: torch.Tensor,
           idx: Optional[torch.Tensor], call: Optional[nn.Module] = None) -> torch.Tensor:
    """Model query for [n,S,C] points -> raw [n,S,4] = (rgb, sigma).  rendering.py:275-334.
    `call` is the module as the caller handed it in (e.g. DistributedDataParallel around `net`)."""
    n, S, Cc = xyz.shape
    B = n * S
    native = net._native()
    first = native.subs[0]
    use_dirs = hparams.pos_dir_dim != 0
    noise = None
    if net.training:
        # same draw order / shapes as the reference's per-chunk torch.rand (rendering.py:294,321)
        ch = hparams.model_chunk_size
        noise = torch.cat([torch.rand(min(ch, B - a), 1, device=xyz.device) for a in range(0, B, ch)], 0)
    rr = RayRows(xyz, S, dirs if use_dirs else None, idx)
    target = call if call is not None else net
    ep = getattr(net, '_ep', None)
    if ep is not None:
        # owner-computes execution over the process group (mega_nerf_b200/expert_parallel.py): the rows travel, so
        # they are materialised like the reference does (rendering.py:275-292,311-319)
        cols = [xyz.reshape(B, Cc)]
        if use_dirs:
            cols.append(dirs.unsqueeze(1).expand(n, S, 3).reshape(B, 3))
        if idx is not None:
            cols.append(idx.view(n, 1, 1).expand(n, S, 1).reshape(B, 1))
        out = ep.forward(torch.cat(cols, 1) if len(cols) > 1 else cols[0], noise)
    elif native.needs_grad():
        # through the wrapper's __call__, like `nerf(x)` in the reference (rendering.py:296-299)
        out = target(typ == 'coarse', rr, sigma_noise=noise) if isinstance(net, Cascade) else target(rr, sigma_noise=noise)
    else:
        rows, keep = rr.rows()
        out = native.forward(rows, B, xyz.device, typ == 'coarse', False, noise, first.rgb_dim + 1, keep)
    if hparams.pos_dir_dim == 0 and hparams.sh_deg is not None:
        out = AG.sh_apply(sg, hparams.sh_deg, out, dirs, S) if out.requires_grad else sg.sh_to_rgb(hparams.sh_deg, out, dirs, S)
    return out.view(n, S, 4)


def _two_pass(sg: _Stage, net: nn.Module, hparams: Namespace, dirs: torch.Tensor, idx: Optional[torch.Tensor],
              xyz_coarse: torch.Tensor, z: torch.Tensor, last_delta: torch.Tensor, get_depth: bool,
              get_depth_variance: bool, get_bg_lambda: bool, flip: bool, depth_real: Optional[torch.Tensor],
              xyz_fine_fn: Callable, call: Optional[nn.Module] = None) -> Dict[str, torch.Tensor]:
    """coarse -> resample -> fine  (rendering.py:176-248 with _inference :251-393 inlined)."""
    res: Dict[str, torch.Tensor] = {}
    fine = hparams.fine_samples > 0
    cascade = hparams.use_cascade
    training = net.training

    # ---- coarse pass
    xyz_c, z_c = xyz_coarse, z
    if flip:
        xyz_c = torch.flip(xyz_coarse, dims=[-2]).contiguous()
        z_c = torch.flip(z, dims=[-1]).contiguous()
    raw_c = _query(sg, net, hparams, 'coarse', xyz_c, dirs, idx, call)
    grad = raw_c.requires_grad

    def composite(raw, zz, dreal, raw2, z2, dreal2, want_depth, want_var, want_lambda):
        """rgb (+ depth, variance, bg_lambda) of one pass; recorded for backward iff the queries were."""
        if grad:
            return AG.composite_apply(sg, raw, zz, dreal, raw2, z2, dreal2, last_delta, flip, want_depth, want_var, want_lambda)
        return sg.composite(raw, zz, dreal, raw2, z2, dreal2, last_delta, flip, False, True, want_depth, want_var,
                            want_lambda)[1:]

    if not grad:
        w, rgb, depth, var, lam = sg.composite(raw_c, z_c, depth_real, None, None, None, last_delta, flip,
                                               want_w=fine, want_rgb=cascade,
                                               want_depth=(not fine) and (get_depth or get_depth_variance),
                                               want_var=(not fine) and get_depth_variance,
                                               want_lambda=get_bg_lambda and cascade)
    else:
        w = rgb = depth = var = lam = None
        if cascade:
            rgb, depth, var, lam = composite(raw_c, z_c, depth_real, None, None, None,
                                             (not fine) and (get_depth or get_depth_variance),
                                             (not fine) and get_depth_variance, get_bg_lambda)
        if fine or not cascade:
            # resampling weights (detached in the reference, rendering.py:215) and, for a coarse-only non-cascade
            # call, the depth terms (no_grad, rendering.py:381): nothing here carries a gradient
            with torch.no_grad():
                w, _, d2, v2, _ = sg.composite(raw_c.detach(), z_c, depth_real, None, None, None, last_delta, flip,
                                               want_w=fine, want_rgb=False,
                                               want_depth=(not cascade) and (not fine) and (get_depth or get_depth_variance),
                                               want_var=(not cascade) and (not fine) and get_depth_variance,
                                               want_lambda=False)
            if not cascade:
                depth, var = d2, v2
    if lam is not None:
        res['bg_lambda_coarse'] = lam
    if rgb is not None:
        res['rgb_coarse'] = rgb
    if (not fine) and get_depth:
        res['depth_coarse'] = depth
    if var is not None:
        res['depth_variance_coarse'] = var
    if not fine:
        return res

    # ---- resample (bins from the unflipped depths, weights as computed: quirk Q7)
    perturb = hparams.perturb if training else 0
    F = hparams.fine_samples // 2 if flip else hparams.fine_samples
    n = z.shape[0]
    if perturb == 0:
        u = torch.linspace(0, 1, F, device=z.device)
    else:
        u = torch.rand(n, F, device=z.device)
    z_f = sg.sample_pdf(z, w, u, F)
    if cascade:
        z_f = sg.sort_cat(z, z_f)
    xyz_f, dreal_f = xyz_fine_fn(z_f)

    # ---- fine pass
    if cascade:
        if flip:
            xyz_f = torch.flip(xyz_f, dims=[-2]).contiguous()
            z_f = torch.flip(z_f, dims=[-1]).contiguous()
        raw_f = _query(sg, net, hparams, 'fine', xyz_f, dirs, idx, call)
        rgb, depth, var, lam = composite(raw_f, z_f, dreal_f, None, None, None, get_depth or get_depth_variance,
                                         get_depth_variance, get_bg_lambda)
    else:
        raw_f = _query(sg, net, hparams, 'fine', xyz_f, dirs, idx, call)
        rgb, depth, var, lam = composite(raw_f, z_f, dreal_f, raw_c, z_c, depth_real if dreal_f is not None else None,
                                         get_depth or get_depth_variance, get_depth_variance, get_bg_lambda)
    res['rgb_fine'] = rgb
    if lam is not None:
        res['bg_lambda_fine'] = lam
    if get_depth:
        res['depth_fine'] = depth
    if var is not None:
        res['depth_variance_fine'] = var
    return res


def render_rays(nerf: nn.Module,
                bg_nerf: Optional[nn.Module],
                rays: torch.Tensor,
                image_indices: Optional[torch.Tensor],
                hparams: Namespace,
                sphere_center: Optional[torch.Tensor],
                sphere_radius: Optional[torch.Tensor],
                get_depth: bool,
                get_depth_variance: bool,
                get_bg_fg_rgb: bool) -> Tuple[Dict[str, torch.Tensor], bool]:
    net = _unwrap(nerf)
    bg = _unwrap(bg_nerf)
    if not isinstance(net, (NeRF, MegaNeRF, Cascade)) or (bg is not None and not isinstance(bg, (NeRF, MegaNeRF, Cascade))):
        raise TypeError('mega_nerf_b200.render_rays needs mega_nerf_b200 modules (use get_nerf / install())')
    recording = net._native().needs_grad() or (bg is not None and bg._native().needs_grad())
    if recording:
        # training step (runner.py:346-358): queries and compositing are recorded, see mega_nerf_b200/autograd.py
        return _render(net, bg, rays.detach(), image_indices, hparams, sphere_center, sphere_radius, get_depth,
                       get_depth_variance, get_bg_fg_rgb, nerf, bg_nerf)
    with torch.no_grad():
        return _render(net, bg, rays, image_indices, hparams, sphere_center, sphere_radius, get_depth,
                       get_depth_variance, get_bg_fg_rgb)


def _render(net, bg, rays, image_indices, hparams, sphere_center, sphere_radius, get_depth, get_depth_variance,
            get_bg_fg_rgb, call_net=None, call_bg=None):
    dev = rays.device
    sg = _Stage(dev)
    rays = K.f32c(rays)
    N = rays.shape[0]
    idx = None
    if image_indices is not None:
        idx = K.f32c(image_indices.to(dev)).view(-1)        # int32 in training, float in eval (runner.py:246,554)
    dirs = rays[:, 3:6]                                      # strided view, [N,3] with row stride 8
    perturb = hparams.perturb if net.training else 0
    S = hparams.coarse_samples
    last_delta = torch.full((N,), 1e10, device=dev, dtype=torch.float32)
    far_override = None
    with_bg = None
    bg_res = None
    center = K.f32c(sphere_center.to(dev)) if sphere_center is not None else None
    radius = K.f32c(sphere_radius.to(dev)) if sphere_radius is not None else None

    if bg is not None:
        fg_far = sg.intersect_sphere(rays, center, radius)
        fg_far = torch.maximum(fg_far, rays[:, 6])
        with_bg = torch.arange(N, device=dev)[rays[:, 7] > fg_far]          # host sync, as in the reference (:37)
        nb = with_bg.shape[0]
        if nb > 0:
            last_delta[with_bg] = fg_far[with_bg]
            far_override = torch.minimum(rays[:, 7], fg_far)
            half = S // 2
            bz1 = torch.linspace(0, 1, half, device=dev)
            rnd = torch.rand(nb, half, device=dev) if perturb > 0 else None
            bz = sg.stratify(bz1, rnd, perturb, nb, half)
            real = hparams.container_path is not None or hparams.train_mega_nerf is not None
            c2d = real and net.cluster_dim_start == 1
            mk = lambda zz: sg.points_outside(rays, with_bg, zz, center, radius, real, c2d)
            bpts, breal = mk(bz)
            bg_dirs = rays[with_bg][:, 3:6]
            bg_idx = idx[with_bg].contiguous() if idx is not None else None
            bg_res = _two_pass(sg, bg, hparams, bg_dirs, bg_idx, bpts, bz,
                               torch.full((nb,), 1e10, device=dev, dtype=torch.float32), get_depth,
                               get_depth_variance, False, True, breal, mk, call_bg)

    steps = torch.linspace(0, 1, S, device=dev)
    rnd = torch.rand(N, S, device=dev) if perturb > 0 else None
    z, xyz = sg.sample_coarse(rays, far_override, steps, rnd, perturb, N, S)
    res = _two_pass(sg, net, hparams, dirs, idx, xyz, z, last_delta, get_depth, get_depth_variance, bg is not None,
                    False, None, lambda zz: (sg.points_from_z(rays, zz), None), call_net)

    if bg is not None:
        types = ['fine' if hparams.fine_samples > 0 else 'coarse']
        if hparams.use_cascade and hparams.fine_samples > 0:
            types.append('coarse')
        for typ in types:
            for key in TO_COMPOSITE:
                name = f'{key}_{typ}'
                if name not in res:
                    continue
                val = res[name]
                if with_bg.shape[0] > 0:
                    lam = res[f'bg_lambda_{typ}'][with_bg]
                    add = torch.zeros_like(val)
                    add[with_bg] = bg_res[name] * (lam.unsqueeze(-1) if val.dim() > 1 else lam)
                    if get_bg_fg_rgb:
                        res[f'fg_{name}'] = val
                        res[f'bg_{name}'] = add
                    res[name] = val + add
                elif get_bg_fg_rgb:
                    res[f'fg_{name}'] = val
                    res[f'bg_{name}'] = torch.zeros_like(val)
    present = bool(bg is not None and with_bg.shape[0] > 0)
    if bg is not None and not present and 'RANK' in os.environ and net.training:
        # Distributed training with no background ray in this batch (rendering.py:143-171): the reference renders ONE
        # dummy background ray through bg_nerf - i.e. through its DistributedDataParallel wrapper, whose forward is
        # what arms the gradient reducer for this iteration - and adds 0 x its colour, so that this rank joins the
        # bg all-reduce with all-zero gradients and the optimiser steps on every rank.  Same here, through `call_bg`;
        # the random draws (jitter, density noise, resampling) are consumed in the reference's order.
        half = S // 2
        bz1 = torch.linspace(0, 1, half, device=dev)
        rnd = torch.rand(1, half, device=dev) if perturb > 0 else None
        bz = sg.stratify(bz1, rnd, perturb, 1, half)
        real = hparams.train_mega_nerf is not None                       # rendering.py:147 (no container_path here)
        c2d = real and net.cluster_dim_start == 1
        first = torch.zeros(1, device=dev, dtype=with_bg.dtype)
        mk = lambda zz: sg.points_outside(rays, first, zz, center, radius, real, c2d)
        bpts, breal = mk(bz)
        dummy = _two_pass(sg, bg, hparams, rays[:1, 3:6], idx[:1].contiguous() if idx is not None else None, bpts, bz,
                          torch.ones(1, device=dev, dtype=torch.float32), get_depth, get_depth_variance, False, True,
                          breal, mk, call_bg)
        key = f'rgb_{"fine" if hparams.fine_samples > 0 else "coarse"}'
        # `results[key][:0] += 0 * grad_results[key]`: an EMPTY slice - the values never mix (a non-finite dummy colour
        # cannot poison the batch), only the graph edge to the bg parameters is added
        res[key] = torch.cat([res[key], (0 * dummy[key])[:0]], 0)
        present = True
    return res, present


def render_rays_fused(nerf: nn.Module, rays: torch.Tensor, image_indices: Optional[torch.Tensor], hparams: Namespace,
                      get_depth: bool, get_depth_variance: bool) -> Dict[str, torch.Tensor]:
    """The foreground inference path of `render_rays(nerf, None, ...)` as ONE library call (`mn_render_rays`): the same
    kernels in the same order, sequenced in C on the current stream instead of from Python.  Eval mode, no background
    network; returns the same keys and values as `render_rays(...)[0]` for that case."""
    net = _unwrap(nerf)
    if not isinstance(net, (NeRF, MegaNeRF, Cascade)):
        raise TypeError('mega_nerf_b200.render_rays_fused needs mega_nerf_b200 modules (use get_nerf / install())')
    if net.training:
        raise ValueError('render_rays_fused is the inference path; call nerf.eval() first')
    if bool(hparams.use_cascade) != isinstance(net, Cascade):
        raise ValueError('hparams.use_cascade does not match the network')
    from .modules import get_precision
    dev = rays.device
    L, h = K.lib(), K.ctx(dev)
    native = net._native()
    native.sync(dev)
    rays = K.f32c(rays)
    N = rays.shape[0]
    idx = K.f32c(image_indices.to(dev)).view(-1) if image_indices is not None else None
    Sc, Sf = hparams.coarse_samples, hparams.fine_samples
    sh_deg = hparams.sh_deg if (hparams.pos_dir_dim == 0 and hparams.sh_deg is not None) else -1
    prec = K.PRECISIONS[get_precision()]
    steps = torch.linspace(0, 1, Sc, device=dev)
    u = torch.linspace(0, 1, Sf, device=dev) if Sf > 0 else None
    typ = 'fine' if Sf > 0 else 'coarse'
    rgb = torch.empty(N, 3, device=dev, dtype=torch.float32)
    depth = torch.empty(N, device=dev, dtype=torch.float32) if get_depth else None
    var = torch.empty(N, device=dev, dtype=torch.float32) if get_depth_variance else None
    rgb_c = torch.empty(N, 3, device=dev, dtype=torch.float32) if (hparams.use_cascade and Sf > 0) else None
    nbytes = int(L.mn_render_rays_workspace_bytes(native.handle, N, Sc, Sf, int(hparams.use_cascade), sh_deg, prec))
    ws = torch.empty(max(nbytes, 256), device=dev, dtype=torch.uint8)
    K.check(L.mn_render_rays(h, native.handle, K.ptr(rays), K.ptr(idx), N, K.ptr(steps), Sc, K.ptr(u), Sf, int(hparams.use_cascade),
                             sh_deg, prec, K.ptr(rgb), K.ptr(depth), K.ptr(var), K.ptr(rgb_c), K.ptr(ws), ws.numel(),
                             K.stream_of(dev)), h)
    res = {f'rgb_{typ}': rgb}
    if depth is not None:
        res[f'depth_{typ}'] = depth
    if var is not None:
        res[f'depth_variance_{typ}'] = var
    if rgb_c is not None:
        res['rgb_coarse'] = rgb_c
    return res
