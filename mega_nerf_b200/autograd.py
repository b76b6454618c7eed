"""torch.autograd bindings of the backward entry points of libmn_b200.so: what
`loss.backward()` runs through the hot path in the reference's training step (runner.py:346-378, :265).

Gradient flow is the reference's: per-sample (rgb, sigma) receive gradients from the composited colour
and from bg_lambda; depth / depth_variance / weights are outputs without gradient (rendering.py:381,
:215); sample positions, directions and image indices are inputs without gradient.  Only fp32 (CUDA-core)
kernels exist for the backward pass in this round, so a recording forward always runs in fp32 whatever
`set_precision` says.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _cabi as K


class _ModelFn(torch.autograd.Function):
    """nn.Module.__call__ of NeRF / Cascade / MegaNeRF on rows (nerf.py:115-160, mega_nerf.py:19-61)."""

    @staticmethod
    def forward(ctx, native, rows, B, device, use_coarse, sigma_noise, out_cols, keep, *params):
        out, tape = native.forward_train(rows, B, device, use_coarse, sigma_noise, out_cols)
        ctx.native, ctx.B, ctx.device, ctx.use_coarse = native, B, device, use_coarse
        ctx.tape = tape
        ctx.plist = native.param_list()
        assert len(ctx.plist) == len(params)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        grads = ctx.native.backward(ctx.B, ctx.device, ctx.use_coarse, grad_out, ctx.tape, ctx.plist)
        ctx.tape = None
        need = ctx.needs_input_grad[8:]
        return (None,) * 8 + tuple(g if n else None for g, n in zip(grads, need))


def model_apply(native, rows, B, device, use_coarse, sigma_noise, out_cols, keep) -> torch.Tensor:
    params = [p for _, _, p in native.param_list()]
    return _ModelFn.apply(native, rows, B, device, use_coarse, sigma_noise, out_cols, keep, *params)


class _CompositeFn(torch.autograd.Function):
    """Merge + volume rendering (rendering.py:336-393): differentiable outputs rgb and bg_lambda."""

    @staticmethod
    def forward(ctx, sg, raw, z, dreal, raw2, z2, dreal2, last_delta, flip, want_depth, want_var, want_lambda):
        _, rgb, depth, var, lam = sg.composite(raw, z, dreal, raw2, z2, dreal2, last_delta, flip, False, True,
                                               want_depth, want_var, want_lambda)
        ctx.sg, ctx.flip = sg, flip
        ctx.has2 = raw2 is not None
        ctx.save_for_backward(raw, z, raw2, z2, last_delta)
        outs = [rgb]
        nd = []
        for t in (depth, var):
            if t is not None:
                nd.append(t)
        if nd:
            ctx.mark_non_differentiable(*nd)
        ctx.slots = (depth is not None, var is not None, lam is not None)
        return rgb, depth, var, lam

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_var, g_lam):
        raw, z, raw2, z2, last_delta = ctx.saved_tensors
        sg = ctx.sg
        N, S = z.shape
        S2 = z2.shape[1] if ctx.has2 else 0
        g_raw = torch.empty_like(raw)
        g_raw2 = torch.empty_like(raw2) if ctx.has2 else None
        if g_rgb is None:
            g_rgb = torch.zeros(N, 3, device=z.device, dtype=torch.float32)
        g_rgb = K.f32c(g_rgb)
        g_lam = K.f32c(g_lam) if g_lam is not None else None
        K.check(sg.L.mn_composite_backward(sg.h, K.ptr(raw), K.ptr(z), S, K.ptr(raw2), K.ptr(z2), S2, K.ptr(last_delta), N,
                                           int(ctx.flip), K.ptr(g_rgb), K.ptr(g_lam), K.ptr(g_raw), K.ptr(g_raw2), sg.st), sg.h)
        return None, g_raw, None, None, g_raw2, None, None, None, None, None, None, None


def composite_apply(sg, raw, z, dreal, raw2, z2, dreal2, last_delta, flip, want_depth, want_var, want_lambda):
    """-> (rgb, depth, var, lam) like _Stage.composite(..., want_w=False, want_rgb=True, ...)."""
    return _CompositeFn.apply(sg, raw.contiguous(), z, dreal, raw2.contiguous() if raw2 is not None else None, z2, dreal2,
                              last_delta, flip, want_depth, want_var, want_lambda)


class _ShFn(torch.autograd.Function):
    """eval_sh + sigmoid on the MLP's raw coefficients (spherical_harmonics.py:55-106, rendering.py:301-306)."""

    @staticmethod
    def forward(ctx, sg, deg, coef, dirs, S):
        out = sg.sh_to_rgb(deg, coef, dirs, S)
        ctx.sg, ctx.deg, ctx.S = sg, deg, S
        ctx.save_for_backward(coef, dirs)
        return out

    @staticmethod
    def backward(ctx, g_out):
        coef, dirs = ctx.saved_tensors
        sg = ctx.sg
        B = coef.shape[0]
        g = K.f32c(g_out)
        g_coef = torch.zeros_like(coef)
        K.check(sg.L.mn_sh_to_rgb_backward(sg.h, ctx.deg, K.ptr(coef), coef.shape[1], K.ptr(dirs), dirs.stride(0), ctx.S, B, 1,
                                           K.ptr(g), K.ptr(g_coef), sg.st), sg.h)
        return None, None, g_coef, None, None


def sh_apply(sg, deg, coef, dirs, S) -> torch.Tensor:
    return _ShFn.apply(sg, deg, coef, dirs, S)
