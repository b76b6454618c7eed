"""GPU: the alternatives of the default 256-wide tensor-core MLP kernel (the TMEM ping-pong kernel: activations in tensor memory,
weight stages shared by the two tiles of a pair) - the shared-memory ping-pong kernel (MN_TC_TP=0; it also serves the training
modes) and the single-tile kernel (MN_TC_PINGPONG=0), environment switches libmn_b200.so reads once per process - each in its
own subprocess, compared with the default kernel on the same 2048-ray C2 batch and with the reference fixture.  (Round 2 measured and then deleted the other variants: single-tile A-from-TMEM, the cta_group::2 CTA pair
with three handshakes and shared weight stages, biases folded into the GEMMs - DESIGN.md §7.)
Every mbarrier wait in these kernels is bounded (a protocol bug traps after ~2 s instead of hanging), and the subprocess has its
own timeout."""
import os
import subprocess
import sys

import pytest
import torch

# (first run on a B200 in round 2: all green, see profiles/r2_staging_tests.log)
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + '/tests')
from argparse import Namespace
import cases as C
from oracle import mn_oracle as O
from test_gpu_parity import DEV, M, product_net, relerr
m = M(); m.set_precision('tc_f16')
golden = torch.load(C.GOLDEN_PATH, map_location='cpu', weights_only=False)
net, _, rays, idx, opts, _, _ = C.render_case('c2_mega8_blend')
pn = product_net(net); hp = Namespace(**vars(opts))
with torch.no_grad():
    res, _ = m.render_rays(pn, None, rays.to(DEV), idx.to(DEV), hp, None, None, True, True, False)
    for k in ('rgb_fine', 'depth_fine'):
        e = relerr(res[k], golden['render_c2_mega8_blend']['out'][k]); assert e <= 2e-4, (k, e)
    big = O.synthetic_rays(2048, seed=3).to(DEV); bidx = O.synthetic_indices(2048, 100, seed=4).to(DEV)
    out, _ = m.render_rays(pn, None, big, bidx, hp, None, None, True, False, False)
    torch.cuda.synchronize()
torch.save({{k: v.cpu() for k, v in out.items()}}, sys.argv[1])
print('VARIANT_OK')
'''


def run_variant(tmp_path, name, env):
    out = tmp_path / f'{name}.pt'
    e = dict(os.environ)
    for k in ('MN_TC_PINGPONG', 'MN_TC_TP'):
        e.pop(k, None)
    e.update(env)
    r = subprocess.run([sys.executable, '-c', CHILD.format(root=ROOT), str(out)], env=e, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and 'VARIANT_OK' in r.stdout, f'{name}: rc={r.returncode}\n{r.stdout[-1500:]}\n{r.stderr[-3000:]}'
    return torch.load(out, map_location='cpu', weights_only=False)


@pytest.fixture(scope='module')
def default_out(tmp_path_factory):
    return run_variant(tmp_path_factory.mktemp('variants'), 'default', {})


@pytest.mark.parametrize('name,env', [('smem_pingpong', {'MN_TC_TP': '0'}), ('single_tile', {'MN_TC_PINGPONG': '0'})])
def test_variant_matches_default_kernel(tmp_path, default_out, name, env):
    got = run_variant(tmp_path, name, env)
    for k, v in default_out.items():
        scale = float(v.abs().max())
        err = float((got[k] - v).abs().max())
        assert err <= 2e-5 * scale, (name, k, err / scale)
