"""CPU-only, world_size 2 over gloo: host logic of the ray-sharded multi-GPU path (shard bounds, the single
padded all-gather, result reassembly).  The render function is a stand-in with render_rays' signature — the
CUDA path itself has no CPU fallback."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fake_render(rays, idx, scale):
    rgb = rays[:, 0:3] * scale + (idx.unsqueeze(-1) if idx is not None else 0)
    return {'rgb_fine': rgb, 'depth_fine': rays[:, 6] + rays[:, 7]}, bool((rays[:, 7] > 0.55).any())


def worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    from mega_nerf_b200 import dist as D
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    rays = torch.rand(n, 8, generator=g)
    idx = torch.randint(0, 10, (n,), generator=g).float()
    res, present = D.render_rays_sharded(fake_render, rays, idx, 2.0)
    ref, ref_present = fake_render(rays, idx, 2.0)
    ok = all(torch.equal(res[k], ref[k]) for k in ref) and set(res) == set(ref) and present == ref_present
    q.put((rank, ok, D.shard_bounds(n, world, rank)))
    dist.barrier()
    dist.destroy_process_group()


_port = [0]


def free_port():
    import socket
    with socket.socket() as s_:
        s_.bind(('127.0.0.1', 0))
        return s_.getsockname()[1]


def run(n):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = free_port()
    ps = [ctx.Process(target=worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(out)


def test_sharded_render_matches_single_rank_even():
    out = run(64)
    assert all(ok for _, ok, _ in out)
    assert out[0][2] == (0, 32) and out[1][2] == (32, 64)


def test_sharded_render_ragged_and_tiny():
    out = run(33)                      # ragged: 17 + 16
    assert all(ok for _, ok, _ in out)
    assert out[0][2] == (0, 17) and out[1][2] == (17, 33)
    out = run(1)                       # one rank gets nothing
    assert all(ok for _, ok, _ in out)


def test_shard_bounds_cover():
    from mega_nerf_b200 import dist as D
    for n in (0, 1, 7, 4096, 65536 + 3):
        for w in (1, 2, 4, 8):
            spans = [D.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


# ------------------------------------------------------------------------------------------------
# training: the recording path under DistributedDataParallel (runner.py:121, rendering.py:296-299)
# ------------------------------------------------------------------------------------------------
class _StandInFn(torch.autograd.Function):
    """Same contract as mega_nerf_b200.autograd._ModelFn: non-tensor arguments first, parameters as tensor inputs,
    gradients returned positionally."""

    @staticmethod
    def forward(ctx, rr, weight):
        ctx.save_for_backward(rr.xyz, weight)
        return rr.xyz @ weight.t()

    @staticmethod
    def backward(ctx, g):
        xyz, weight = ctx.saved_tensors
        return None, g.t() @ xyz


class _StandIn(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.lin = torch.nn.Linear(3, 4, bias=False)

    def forward(self, x, sigma_noise=None):
        from mega_nerf_b200.modules import RayRows
        assert isinstance(x, RayRows), type(x)          # DDP.__call__ must hand the object through untouched
        return _StandInFn.apply(x, self.lin.weight)


def ddp_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from mega_nerf_b200.modules import RayRows
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    net = _StandIn()
    ddp = torch.nn.parallel.DistributedDataParallel(net)
    xyz = torch.rand(50, 3, generator=torch.Generator().manual_seed(10 + rank))
    out = ddp(RayRows(xyz, 5, None, None), sigma_noise=None)
    out.sum().backward()
    q.put((rank, net.lin.weight.grad.tolist(), xyz.sum(0).tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_hands_rayrows_through_and_reduces_gradients():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = free_port()
    ps = [ctx.Process(target=ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted([q.get(timeout=120) for _ in ps], key=lambda t: t[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    mean = (torch.tensor(out[0][2]) + torch.tensor(out[1][2])) / 2    # d(sum(xyz W^T))/dW[n] = sum_rows xyz, averaged over ranks
    for _, g, _ in out:
        assert torch.allclose(torch.tensor(g), mean.expand(4, 3), atol=1e-5)


# ------------------------------------------------------------------------------------------------
# owner-computes ("one centroid per GPU") execution of a MegaNeRF: dispatch / return / accumulate
# ------------------------------------------------------------------------------------------------
def ep_worker(rank, world, port, mname, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import types
    import cases as C
    from oracle import mn_oracle as O
    from mega_nerf_b200.expert_parallel import ExpertParallel, plan_dispatch, owner_of
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    net = C.mega_net(mname)
    K_ = len(net.weights)
    x = C.mega_rows(net, 300 + 37 * rank, 51 + rank)            # every rank routes its own (differently sized) rows
    calls = []

    def sub_fn(k, rows, noise):
        assert owner_of(k, world) == rank, (k, rank)             # only owned sub-modules are ever evaluated here
        calls.append((k, rows.shape[0]))
        return O.nerf_forward(net.spec, net.weights[k], rows, sigma_noise=noise)

    mega = types.SimpleNamespace(sub_modules=[types.SimpleNamespace(rgb_dim=net.spec.rgb_dim)] * K_, xyz_real=net.xyz_real,
                                 parameters=lambda: iter(()))
    ep = ExpertParallel(mega, None, route_fn=lambda xx: O.route(net, xx), sub_fn=sub_fn)
    g = torch.Generator().manual_seed(5 + rank)
    noise = torch.rand(x.shape[0], 1, generator=g)
    with torch.inference_mode():
        got = ep.forward(x)
        got_n = ep.forward(x, noise)
        want = O.mega_forward(net, x)
        want_n = O.mega_forward(net, x, sigma_noise=noise)
    err = float((got - want).abs().max() / want.abs().max())
    err_n = float((got_n - want_n).abs().max() / want_n.abs().max())
    total = torch.tensor([ep.last_pairs, ep.last_owned])
    dist.all_reduce(total)
    q.put((rank, err, err_n, sorted(set(k for k, _ in calls)), int(total[0]), int(total[1])))
    dist.barrier()
    dist.destroy_process_group()


def run_ep(mname):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = free_port()
    ps = [ctx.Process(target=ep_worker, args=(r, 2, port, mname, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted([q.get(timeout=180) for _ in ps])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def test_expert_parallel_matches_single_device_mixture():
    for mname in ('hard2d', 'blend2d', 'hard3d_bgreal', 'blend25'):
        out = run_ep(mname)
        for rank, err, err_n, ks, sent, owned in out:
            assert err <= 1e-6 and err_n <= 1e-6, (mname, rank, err, err_n)
            assert all(k % 2 == rank for k in ks), (mname, rank, ks)
            assert sent == owned                                   # every dispatched pair was computed exactly once


def test_plan_dispatch_order_and_counts():
    from mega_nerf_b200.expert_parallel import plan_dispatch
    w = torch.tensor([[0.5, 0.0, 0.5, 0.0], [0.0, 1.0, 0.0, 0.0], [0.2, 0.3, 0.0, 0.5]])
    rows, subs, ww, counts = plan_dispatch(None, w, 4, 2)
    assert subs.tolist() == [0, 0, 2, 1, 1, 3] and rows.tolist() == [0, 2, 0, 1, 2, 2] and counts.tolist() == [3, 3]
    assert torch.allclose(ww, torch.tensor([0.5, 0.2, 0.5, 1.0, 0.3, 0.5]))
    rows, subs, ww, counts = plan_dispatch(torch.tensor([3, 0, 1, 1]), None, 4, 3)
    assert subs.tolist() == [0, 3, 1, 1] and rows.tolist() == [1, 0, 2, 3] and ww is None and counts.tolist() == [2, 2, 0]
