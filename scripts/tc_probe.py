"""GPU diagnostic for the tcgen05 MLP kernel: compares it with the fp32 CUDA-core kernel on small nets,
under both UMMA descriptor conventions.  Usage (on the GPU box): python scripts/tc_probe.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

CONFIGS = [
    dict(name='L64_l1', layer_dim=64, layers=1, skip_layers=(), n=256),
    dict(name='L256_l1', layer_dim=256, layers=1, skip_layers=(), n=256),
    dict(name='L256_l2', layer_dim=256, layers=2, skip_layers=(), n=300),
    dict(name='L256_l8', layer_dim=256, layers=8, skip_layers=(4,), n=1000),
]


def child():
    import torch
    import mega_nerf_b200 as M
    from oracle import mn_oracle as O
    import cases as C
    from mega_nerf_b200.synthetic import build_net
    dev = torch.device('cuda:0')
    for cfg in CONFIGS:
        spec = O.NerfSpec(layer_dim=cfg['layer_dim'], layers=cfg['layers'], skip_layers=cfg['skip_layers'])
        net = O.make_net('nerf', spec, seed=3)
        x = C.nerf_rows(spec, cfg['n'], 9)
        p = build_net(net, dev)
        for so in (True, False):
            xin = (C.nerf_rows(spec, cfg['n'], 9, sigma_only=True) if so else x).to(dev)
            M.set_precision('fp32')
            ref = p(xin, sigma_only=so)
            M.set_precision('tc_f16')
            try:
                out = p(xin, sigma_only=so)
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                print(f'{cfg["name"]} sigma_only={so}: EXC {e}')
                return
            err = float((out - ref).abs().max() / ref.abs().max())
            bad = int((~torch.isfinite(out)).sum())
            print(f'{cfg["name"]} sigma_only={so}: relerr {err:.3e} nonfinite {bad} '
                  f'out[0]={out[0].tolist()} ref[0]={ref[0].tolist()}', flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'child':
        child()
    else:
        for swap in ('0', '1'):
            print(f'=== MN_TC_DESC_SWAP={swap}', flush=True)
            env = dict(os.environ, MN_TC_DESC_SWAP=swap)
            try:
                r = subprocess.run([sys.executable, __file__, 'child'], env=env, timeout=120, capture_output=True, text=True)
                print(r.stdout[-3000:])
                if r.returncode != 0:
                    print('rc', r.returncode, r.stderr[-1500:])
            except subprocess.TimeoutExpired:
                print('TIMEOUT')
