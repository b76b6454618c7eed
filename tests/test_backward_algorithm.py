"""CPU-only: the chain rule exactly as the backward kernels apply it (csrc/mn_backward.cu: which saved
activation masks which gradient, which weight sub-matrix each data-gradient step streams, where the sigma
head joins, the affine-appearance and embedding gathers, the compositing suffix sums) restated step by step
with plain tensor algebra and checked against the oracle's autograd.  Guards the derivation; the CUDA
kernels themselves are checked on the GPU by tests/test_gpu_zc_backward.py."""
import pytest
import torch
import torch.nn.functional as F

import cases as C
from oracle import mn_oracle as O


def kernel_chain_rule(spec: O.NerfSpec, w, x, cot, noise=None):
    """-> gradient dict in state-dict layout, following mlp_simt_kernel<SAVE> + mlp_bwd_data_kernel + weight op table."""
    L, layers = spec.layer_dim, spec.layers
    in_xyz = spec.in_xyz
    # ---- training forward: the activation tape
    pe = O.embed(x[:, :spec.xyz_dim], spec.pos_xyz_dim)
    aux = []
    if spec.pos_dir_dim > 0:
        aux.append(O.embed(x[:, -4:-1], spec.pos_dir_dim))
    ids = x[:, -1].long() if spec.appearance_dim > 0 else None
    app_in_dira = spec.appearance_dim > 0 and not spec.affine_appearance
    if app_in_dira:
        aux.append(w['embedding_a.weight'][ids])
    aux = torch.cat(aux, -1) if aux else torch.zeros(x.shape[0], 0)
    h, xin = [], []
    cur = pe
    for i in range(layers):
        inp = torch.cat([pe, cur], -1) if i in spec.skip_layers else cur
        xin.append(inp)
        cur = torch.relu(F.linear(inp, w[f'xyz_encodings.{i}.0.weight'], w[f'xyz_encodings.{i}.0.bias']))
        h.append(cur)
    sig_pre = F.linear(h[-1], w['sigma.weight'], w['sigma.bias'])[:, 0]
    if noise is not None:
        sig_pre = sig_pre + noise.view(-1)
    if spec.has_dir_a:
        f = F.linear(h[-1], w['xyz_encoding_final.weight'], w['xyz_encoding_final.bias'])
        g = torch.relu(F.linear(torch.cat([f, aux], -1), w['dir_a_encoding.0.weight'], w['dir_a_encoding.0.bias']))
        rgb_src = g
    else:
        rgb_src = h[-1]
    lin = F.linear(rgb_src, w['rgb.weight'], w['rgb.bias'])
    affine = spec.affine_appearance and spec.appearance_dim > 0
    if affine:
        e = w['embedding_a.weight'][ids]
        A = F.linear(e, w['affine.weight'], w['affine.bias']).view(-1, 3, 4)
        pre = (A[:, :, :3] @ lin.unsqueeze(-1)).squeeze(-1) + A[:, :, 3]
    else:
        pre = lin
    s = torch.sigmoid(pre) if spec.rgb_dim == 3 else pre

    # ---- heads
    G = {k: torch.zeros_like(v) for k, v in w.items()}
    go_rgb, go_sig = cot[:, :spec.rgb_dim], cot[:, spec.rgb_dim]
    if spec.shifted_softplus:
        y = sig_pre - 1
        d = torch.where(y > 20, torch.ones_like(y), 1 / (1 + torch.exp(-y)))
    else:
        d = (sig_pre > 0).float()
    ds = go_sig * d
    dv = go_rgb * (1 - s) * s if spec.rgb_dim == 3 else go_rgb
    if affine:
        dA = torch.zeros(x.shape[0], 12)
        for c in range(3):
            for q in range(3):
                dA[:, c * 4 + q] = dv[:, c] * lin[:, q]
            dA[:, c * 4 + 3] = dv[:, c]
        dl = torch.einsum('bcq,bc->bq', A[:, :, :3], dv)
        G['affine.bias'] += dA.sum(0)
        G['affine.weight'] += dA.t() @ e                                  # [12][app]
        G['embedding_a.weight'].index_add_(0, ids, dA @ w['affine.weight'])
        dr = dl
    else:
        dr = dv
    # ---- weight ops: dW = dZ^T X, db = sum dZ
    def wop(name, dz, xx):
        G[name + '.weight'] += dz.t() @ xx
        G[name + '.bias'] += dz.sum(0)

    wop('rgb', dr, rgb_src)
    G['sigma.weight'] += (ds.unsqueeze(-1) * h[-1]).sum(0, keepdim=True)
    G['sigma.bias'] += ds.sum().view(1)
    d_src = dr @ w['rgb.weight']                                           # [B, rgb_in]
    if spec.has_dir_a:
        dz_dira = d_src * (g > 0)
        wop('dir_a_encoding.0', dz_dira, torch.cat([f, aux], -1))
        Wd = w['dir_a_encoding.0.weight']
        dz_final = dz_dira @ Wd[:, :L]                                     # BwdLayout.dira_f
        if app_in_dira:
            d_e = dz_dira @ Wd[:, L + spec.in_dir:]                        # BwdLayout.dira_e
            G['embedding_a.weight'].index_add_(0, ids, d_e)
        wop('xyz_encoding_final', dz_final, h[-1])
        d_h = dz_final @ w['xyz_encoding_final.weight']
    else:
        d_h = d_src
    d_h = d_h + ds.unsqueeze(-1) * w['sigma.weight']                       # sigma head joins before the last ReLU mask
    dz = d_h * (h[-1] > 0)
    for i in range(layers - 1, -1, -1):
        wop(f'xyz_encodings.{i}.0', dz, xin[i])
        if i == 0:
            break
        Wi = w[f'xyz_encodings.{i}.0.weight']
        Wh = Wi[:, in_xyz:] if i in spec.skip_layers else Wi               # BwdLayout.w[i]: hidden columns only
        dz = (dz @ Wh) * (h[i - 1] > 0)
    return G


@pytest.mark.parametrize('vname', list(C.NERF_VARIANTS))
def test_mlp_chain_rule(vname):
    spec = C.NERF_VARIANTS[vname]['spec']
    if spec.layer_dim > 256:
        spec = O.NerfSpec(**{**spec.__dict__, 'layer_dim': 128})
    net = O.make_net('nerf', spec, seed=21)
    x = C.nerf_rows(spec, 120, 31)
    g = torch.Generator().manual_seed(5)
    cot = torch.randn(120, spec.rgb_dim + 1, generator=g)
    noise = torch.rand(120, 1, generator=g)
    _, want = O.net_forward_grads(net, x, cot, sigma_noise=noise)
    with torch.no_grad():
        got = kernel_chain_rule(spec, net.weights[0], x, cot, noise)
    assert set(got) == set(want[0])
    for k, v in want[0].items():
        scale = float(v.abs().max())
        err = float((got[k] - v).abs().max())
        assert err <= 2e-5 * max(scale, 1e-12), (k, err, scale)


def test_blend_weight_scales_upstream_gradient():
    """mega_nerf.py:46-49: out[row] = sum_k w_k head_k(row)  =>  each slot's upstream gradient is w_k * dL/dout[row]."""
    net = C.mega_net('blend2d')
    x = C.mega_rows(net, 300, 51)
    cot = torch.randn(300, 4, generator=torch.Generator().manual_seed(6))
    _, want = O.net_forward_grads(net, x, cot)
    _, wts = O.route(net, x)
    with torch.no_grad():
        for i, w in enumerate(net.weights):
            mask = wts[:, i] > 0
            got = kernel_chain_rule(net.spec, w, x[mask], cot[mask] * wts[mask, i].unsqueeze(-1))
            for k, v in want[i].items():
                scale = float(v.abs().max())
                assert float((got[k] - v).abs().max()) <= 2e-5 * max(scale, 1e-12), (i, k)
