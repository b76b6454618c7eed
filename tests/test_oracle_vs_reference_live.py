"""CPU-only, build container only (skipped where /root/reference is absent, e.g. on the GPU box): randomised
configurations - widths, depths, skip layers, SH degree, appearance / affine, cascade, background, routing margin,
2-D / 3-D clustering, train / eval mode - rendered by the UNMODIFIED reference (imported read-only) and by the oracle
with the same seeds.  Results and parameter gradients must agree bit for bit.  This is the live form of the pinning that
the committed fixtures freeze (tests/golden/*.pt)."""
import dataclasses
import os
import random
import sys

import pytest
import torch

import cases as C
from oracle import mn_oracle as O

REF = os.environ.get('MEGA_NERF_REFERENCE', '/root/reference')
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'mega_nerf')), reason='reference checkout not present')


@pytest.fixture(scope='module')
def MG():
    sys.path.insert(0, os.path.join(C.ROOT, 'tests', 'golden'))
    import make_golden
    return make_golden


def random_case(seed: int):
    rnd = random.Random(seed)
    sh = rnd.random() < 0.25
    app = rnd.choice([0, 16, 48])
    affine = app > 0 and not sh and rnd.random() < 0.25
    layers = rnd.choice([2, 4, 8])
    spec = O.NerfSpec(pos_xyz_dim=rnd.choice([6, 12]), pos_dir_dim=0 if sh else rnd.choice([2, 4]), layers=layers,
                      skip_layers=(rnd.randrange(1, layers),) if rnd.random() < 0.8 else (), layer_dim=rnd.choice([32, 64, 96]),
                      appearance_dim=app, affine_appearance=affine, appearance_count=9, rgb_dim=27 if sh else 3,
                      shifted_softplus=rnd.random() < 0.8)
    kind = rnd.choice(['nerf', 'cascade', 'mega'])
    cascade = kind == 'cascade'
    grid = rnd.choice([(2, 2), (1, 3), (2, 4)])
    cents = O.grid_centroids(*grid) if kind == 'mega' else None
    c2d = rnd.random() < 0.6
    if cents is not None and not c2d:
        cents = cents.clone()
        cents[:, 0] = torch.rand(cents.shape[0], generator=torch.Generator().manual_seed(seed)) * 0.4 - 0.2
    margin = rnd.choice([1.0, 1.15, 1.4]) if kind == 'mega' else 1.0
    net = O.make_net(kind, spec, seed=seed, n_sub=0 if cents is None else cents.shape[0], centroids=cents,
                     boundary_margin=margin, cluster_2d=c2d)
    has_bg = rnd.random() < 0.3
    bg = None
    center = radius = None
    n_rays = rnd.choice([7, 33, 64])
    rays = O.synthetic_rays(n_rays, seed=seed, far=1e5 if has_bg else 0.6)
    if has_bg:
        bg = O.make_net('cascade' if cascade else 'nerf', dataclasses.replace(spec, xyz_dim=4), seed=seed + 1)
        center, radius = torch.tensor([0.05, -0.02, 0.03]), torch.tensor([0.8, 0.9, 1.0])
        rays[::2, 7] = 0.4
    idx = O.synthetic_indices(n_rays, 9, seed=seed) if app > 0 else None
    fine = rnd.choice([0, 8, 24]) if cascade else rnd.choice([8, 24])
    opts = O.RenderOpts(coarse_samples=rnd.choice([8, 16, 31]), fine_samples=fine, use_cascade=cascade, perturb=1.0,
                        pos_dir_dim=spec.pos_dir_dim, sh_deg=2 if sh else None, model_chunk_size=rnd.choice([64, 1000, 32768]))
    return net, bg, rays, idx, opts, center, radius, rnd.random() < 0.5


@pytest.mark.parametrize('seed', list(range(16)))
def test_random_configuration_bit_exact(MG, seed):
    net, bg, rays, idx, opts, c, r, training = random_case(seed)
    rn = MG.ref_net(net)
    rb = MG.ref_net(bg) if bg is not None else None
    for mod in (rn, rb):
        if mod is not None:
            mod.train(training)
            for p in mod.parameters():
                p.requires_grad_(True)
    nt = dataclasses.replace(net, training=training)
    bt = dataclasses.replace(bg, training=training) if bg is not None else None
    key = f'rgb_{"fine" if opts.fine_samples > 0 else "coarse"}'
    cot = torch.randn(rays.shape[0], 3, generator=torch.Generator().manual_seed(seed))
    torch.manual_seed(seed)
    ref, rp = MG.R_render.render_rays(rn, rb, rays, idx, MG.hparams_of(opts), c, r, False, True, False)
    (ref[key] * cot).sum().backward()
    torch.manual_seed(seed)
    got, gn, gb = O.render_grads(nt, bt, rays, idx, opts, c, r, {key: cot})
    assert set(got) == set(ref)
    for k in ref:
        assert torch.equal(ref[k].detach(), got[k]), (seed, k, float((ref[k].detach() - got[k]).abs().max()))
    sys.path.insert(0, os.path.join(C.ROOT, 'tests', 'golden'))
    import make_golden_backward as MB
    for mod, n_, g_ in ((rn, net, gn), (rb, bg, gb)):
        if mod is None:
            continue
        for a, b in zip(MB.ref_grads(mod, n_), g_):
            for k in a:
                assert torch.equal(a[k], b[k]), (seed, k, float((a[k] - b[k]).abs().max()))


def test_call_surface_signatures_match_reference(MG):
    """Drop-in boundary: every replaced symbol takes the reference's parameters, in order, with the
    reference's defaults."""
    import inspect
    import mega_nerf_b200 as M
    from mega_nerf import ray_utils as R_rays, rendering as R_rendering
    from mega_nerf.spherical_harmonics import eval_sh as R_eval_sh
    from mega_nerf.models import nerf as R_nerf, mega_nerf as R_mega, cascade as R_cascade
    pairs = [(M.render_rays, R_rendering.render_rays), (M.get_rays, R_rays.get_rays), (M.get_rays_batch, R_rays.get_rays_batch),
             (M.get_ray_directions, R_rays.get_ray_directions), (M.eval_sh, R_eval_sh),
             (M.NeRF.__init__, R_nerf.NeRF.__init__), (M.NeRF.forward, R_nerf.NeRF.forward),
             (M.MegaNeRF.__init__, R_mega.MegaNeRF.__init__), (M.MegaNeRF.forward, R_mega.MegaNeRF.forward),
             (M.Cascade.__init__, R_cascade.Cascade.__init__), (M.Cascade.forward, R_cascade.Cascade.forward),
             (M.Embedding.__init__, R_nerf.Embedding.__init__), (M.ShiftedSoftplus.__init__, R_nerf.ShiftedSoftplus.__init__)]
    for mine, ref in pairs:
        a, b = inspect.signature(mine), inspect.signature(ref)
        pa = [(p.name, p.default, p.kind) for p in a.parameters.values()]
        pb = [(p.name, p.default, p.kind) for p in b.parameters.values()]
        assert [x[0] for x in pa] == [x[0] for x in pb], (ref.__qualname__, pa, pb)
        assert [x[1:] for x in pa] == [x[1:] for x in pb], (ref.__qualname__, pa, pb)
    # model_utils needs configargparse-free import: compare by source inspection of the two factory signatures
    import importlib
    mu = importlib.import_module('mega_nerf.models.model_utils')
    for name in ('get_nerf', 'get_bg_nerf'):
        assert list(inspect.signature(getattr(M, name)).parameters) == list(inspect.signature(getattr(mu, name)).parameters)
    # state-dict layout of every model family
    spec = O.NerfSpec(layer_dim=32, appearance_count=5)
    for kind in ('nerf', 'cascade', 'mega'):
        cents = O.grid_centroids(2, 2) if kind == 'mega' else None
        net = O.make_net(kind, spec, seed=1, n_sub=4 if kind == 'mega' else 1, centroids=cents, cluster_2d=True)
        ref = MG.ref_net(net)
        from test_host_factories import M as _M  # noqa: F401
        if kind == 'nerf':
            mine = M.NeRF(spec.pos_xyz_dim, spec.pos_dir_dim, spec.layers, list(spec.skip_layers), spec.layer_dim, spec.appearance_dim,
                          spec.affine_appearance, spec.appearance_count, spec.rgb_dim, spec.xyz_dim, M.ShiftedSoftplus())
        elif kind == 'cascade':
            mk = lambda: M.NeRF(spec.pos_xyz_dim, spec.pos_dir_dim, spec.layers, list(spec.skip_layers), spec.layer_dim,  # noqa: E731
                                spec.appearance_dim, spec.affine_appearance, spec.appearance_count, spec.rgb_dim, spec.xyz_dim,
                                M.ShiftedSoftplus())
            mine = M.Cascade(mk(), mk())
        else:
            mk = lambda: M.NeRF(spec.pos_xyz_dim, spec.pos_dir_dim, spec.layers, list(spec.skip_layers), spec.layer_dim,  # noqa: E731
                                spec.appearance_dim, spec.affine_appearance, spec.appearance_count, spec.rgb_dim, spec.xyz_dim,
                                M.ShiftedSoftplus())
            mine = M.MegaNeRF([mk() for _ in range(4)], cents, 1.15, False, True)
        a, b = mine.state_dict(), ref.state_dict()
        assert list(a) == list(b), (kind, set(a) ^ set(b))
        assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in b)
        mine.load_state_dict(b)                      # reference checkpoints load
