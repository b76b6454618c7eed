"""GPU: size-independent properties of the path at BASELINE.json's FULL sizes, where the CPU oracle is too slow to be
the checker (C2: 4096 rays x (64 + 128) samples x 8 sub-modules; C3's per-GPU shard: 8192 rays; C4: 25 x 512; C5: SH head).

 * rays are independent units (rendering.py reduces along the sample axis only): rendering a batch in two pieces, or in a
   permuted order, gives the same per-ray results;
 * volume rendering (rendering.py:352-393): weights >= 0, sum of weights <= 1, rgb in [0,1] (convex combination of
   sigmoid outputs), depth inside [near, far];
 * inverse-CDF resampling (rendering.py:505-536, deterministic u): fine depths are non-decreasing along a ray and stay
   inside the coarse depth range;
 * compositing is linear in the sample colours, its backward is linear in the upstream gradient;
 * an empty ray batch raises (the reference raises too: torch.cat of an empty chunk list, rendering.py:330).
"""
from argparse import Namespace

import pytest
import torch

import cases as C
from oracle import mn_oracle as O
from test_gpu_parity import DEV, M, product_net, stage

pytestmark = pytest.mark.gpu


def _full_case(n_rays: int, grid=(2, 4), margin=1.15, **spec):
    sp = O.NerfSpec(**spec)
    cents = O.grid_centroids(*grid)
    net = O.make_net('mega', sp, seed=0, n_sub=cents.shape[0], centroids=cents, boundary_margin=margin, cluster_2d=True)
    rays = O.synthetic_rays(n_rays, seed=0)
    idx = O.synthetic_indices(n_rays, sp.appearance_count)
    opts = O.RenderOpts(coarse_samples=64, fine_samples=128, pos_dir_dim=sp.pos_dir_dim,
                        sh_deg=2 if sp.rgb_dim == 27 else None)
    return product_net(net), rays.to(DEV), idx.to(DEV), Namespace(**vars(opts))


def _render(pn, rays, idx, hp):
    with torch.no_grad():
        res, _ = M().render_rays(pn, None, rays, idx, hp, None, None, True, True, False)
    return res


def _close(a, b, tol):
    scale = float(b.abs().max())
    assert float((a - b).abs().max()) <= tol * max(scale, 1e-30), float((a - b).abs().max()) / max(scale, 1e-30)


@pytest.mark.parametrize('prec,tol', [('tc_f16', 1e-5), ('fp32', 1e-6)])
def test_c2_full_size_rays_are_independent(prec, tol):
    M().set_precision(prec)
    pn, rays, idx, hp = _full_case(4096)
    whole = _render(pn, rays, idx, hp)
    assert set(whole) == {'rgb_fine', 'depth_fine', 'depth_variance_fine'}
    # two ragged pieces (1001 is neither a multiple of the 4 rays per CTA nor of a 128-row tile)
    a, b = _render(pn, rays[:1001], idx[:1001], hp), _render(pn, rays[1001:], idx[1001:], hp)
    perm = torch.randperm(4096, generator=torch.Generator().manual_seed(1)).to(DEV)
    shuffled = _render(pn, rays[perm], idx[perm], hp)
    for k, v in whole.items():
        assert torch.isfinite(v).all(), k
        t = 5 * tol if 'variance' in k else tol
        _close(torch.cat([a[k], b[k]], 0), v, t)
        _close(shuffled[k], v[perm], t)


@pytest.mark.parametrize('name,kw', [('c2', dict(n_rays=4096)), ('c3_shard', dict(n_rays=8192)),
                                     ('c4', dict(n_rays=4096, grid=(5, 5), layer_dim=512)),
                                     ('c5', dict(n_rays=8192, pos_dir_dim=0, rgb_dim=27)),
                                     ('c2_hard', dict(n_rays=4096, margin=1.0))])
def test_full_size_volume_rendering_bounds(name, kw):
    M().set_precision('tc_f16')
    pn, rays, idx, hp = _full_case(**kw)
    res = _render(pn, rays, idx, hp)
    rgb, depth, var = res['rgb_fine'], res['depth_fine'], res['depth_variance_fine']
    assert rgb.shape == (kw['n_rays'], 3) and depth.shape == (kw['n_rays'],)
    assert torch.isfinite(rgb).all() and torch.isfinite(depth).all() and torch.isfinite(var).all()
    assert float(rgb.min()) >= 0.0 and float(rgb.max()) <= 1.0 + 1e-5           # convex combination of sigmoid outputs
    near, far = rays[:, 6], rays[:, 7]
    assert bool((depth >= near * (1 - 1e-4)).all()) and bool((depth <= far * (1 + 1e-4)).all())
    assert float(var.min()) >= 0.0 and float(var.max()) <= 1.001 * float(((far - near) ** 2).max())
    slots, tiles = pn._native().stats(DEV)
    rows = kw['n_rays'] * 128
    assert rows <= slots <= 4 * rows and tiles * 128 >= slots                  # every fine row reached >= 1 sub-module


def test_full_size_weights_and_resampling():
    """Stage level at C2 size: the coarse weights are a sub-probability distribution, the fine depths drawn from them
    are sorted and stay inside the coarse range."""
    sg = stage()
    n, S, F = 4096, 64, 128
    rays = O.synthetic_rays(n, seed=0).to(DEV)
    g = torch.Generator().manual_seed(2)
    raw = torch.rand(n, S, 4, generator=g)
    raw[..., 3] = raw[..., 3] * 40 * (torch.rand(n, S, generator=g) > 0.5)          # density with empty stretches
    raw = raw.to(DEV)
    steps = torch.linspace(0, 1, S, device=DEV)
    z, _ = sg.sample_coarse(rays, None, steps, None, 0.0, n, S)
    ld = torch.full((n,), 1e10, device=DEV)
    w, rgb, depth, var, lam = sg.composite(raw, z, None, None, None, None, ld, False, True, True, True, True, True)
    assert float(w.min()) >= 0.0 and float(w.sum(-1).max()) <= 1.0 + 1e-5
    assert float(lam.min()) >= 0.0 and float(lam.max()) <= 1.0 + 1e-6
    assert bool((w.sum(-1) + lam <= 1.0 + 1e-4).all())                              # rendered + transmitted <= 1
    u = torch.linspace(0, 1, F, device=DEV)
    zf = sg.sample_pdf(z, w, u, F)
    assert bool((zf[:, 1:] >= zf[:, :-1]).all()), 'fine depths must be non-decreasing for deterministic u'
    assert bool((zf >= z[:, :1]).all()) and bool((zf <= z[:, -1:]).all())
    merged = sg.sort_cat(z, zf)
    assert bool((merged[:, 1:] >= merged[:, :-1]).all()) and merged.shape == (n, S + F)
    # linearity in the colours
    raw2 = raw.clone()
    raw2[..., :3] *= 2
    _, rgb2, _, _, _ = sg.composite(raw2, z, None, None, None, None, ld, False, False, True, False, False, False)
    assert float((rgb2 - 2 * rgb).abs().max()) <= 1e-6


def test_composite_backward_is_linear_in_the_upstream_gradient():
    from mega_nerf_b200 import _cabi as K
    sg = stage()
    n, S = 2048, 192
    g = torch.Generator().manual_seed(3)
    z = torch.sort(torch.rand(n, S, generator=g), -1)[0].to(DEV)
    raw = torch.rand(n, S, 4, generator=g)
    raw[..., 3] *= 20
    raw = raw.to(DEV)
    ld = torch.full((n,), 1e10, device=DEV)

    def bwd(g_rgb):
        out = torch.empty_like(raw)
        K.check(sg.L.mn_composite_backward(sg.h, K.ptr(raw), K.ptr(z), S, None, None, 0, K.ptr(ld), n, 0, K.ptr(g_rgb), None,
                                           K.ptr(out), None, sg.st), sg.h)
        return out
    g1 = torch.randn(n, 3, generator=g).to(DEV)
    g2 = torch.randn(n, 3, generator=g).to(DEV)
    a, b, ab = bwd(g1), bwd(g2), bwd(g1 + 3 * g2)
    assert torch.isfinite(ab).all()
    scale = float(ab.abs().max())
    assert float((ab - (a + 3 * b)).abs().max()) <= 2e-5 * scale


def test_empty_ray_batch_raises_like_the_reference():
    pn, rays, idx, hp = _full_case(8)
    with pytest.raises(Exception):
        M().render_rays(pn, None, rays[:0], idx[:0], hp, None, None, True, True, False)
