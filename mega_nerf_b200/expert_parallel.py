"""Owner-computes ("one centroid per GPU") execution of a MegaNeRF over a process group - SURVEY.md §8f-5.

`models/mega_nerf.py:19-61` loops over the sub-modules on one device.  Here sub-module k lives on rank k % G only
(round-robin, BASELINE.json configs[2] / [3]): every rank routes ITS OWN sample rows (the centroids are tiny and
replicated), ships each (row, sub-module) pair to the owner with one all-to-all, the owners run their sub-modules on
what they received, a second all-to-all returns the (rgb, sigma) rows, and the home rank accumulates them in ascending
sub-module order with the blend weights - the same arithmetic as `results[mask] += sub_result * weights[mask, i]`
(`mega_nerf.py:46-49`).  Compositing stays on the home rank: it is non-linear along a ray, which is why a per-ray
all-gather alone cannot express this partitioning.

Payload per routed pair: the child's input row (xyz, dir, image index: 28 B) + sub-module id (+ density noise) out,
16 B back.  The exchange is `torch.distributed.all_to_all_single` (NCCL over NVLink on GPUs, gloo in the CPU tests of
this host logic); the split sizes are exchanged first, which costs one host sync per query - the reference itself syncs
once per sub-module (`x[cluster_mask]`).  Inference only; every rank must issue the same sequence of queries (true for
`render_rays` on the foreground network with the same sampling configuration on every rank).

The two device-side steps are injectable (`route_fn`, `sub_fn`) so that the dispatch / return / accumulation logic is
testable without a GPU; the defaults call libmn_b200.so (`mn_model_route`, and `NeRF.forward` of the owned sub-module).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _cabi as K


class _RouteOnly:
    """A native MegaNeRF model used for routing only: centroids are set, no weights are packed."""

    def __init__(self, mega):
        self.mega = mega
        self.handle = None
        self.device = None
        self.stamp = None

    def __del__(self):
        try:
            if self.handle is not None:
                K.lib().mn_model_destroy(self.handle)
        except Exception:
            pass

    def sync(self, device: torch.device):
        from .modules import model_desc
        L = K.lib()
        h = K.ctx(device)
        m = self.mega
        if self.handle is None or self.device != device:
            if self.handle is not None:
                L.mn_model_destroy(self.handle)
            d = model_desc(m.sub_modules[0], 2, len(m.sub_modules), m.boundary_margin, m.xyz_real, m.cluster_dim_start)
            out = C.c_void_p()
            K.check(L.mn_model_create(h, C.byref(d), C.byref(out)), h)
            self.handle, self.device, self.stamp = out.value, device, None
        stamp = (m.centroids.data_ptr(), m.centroids._version)
        if stamp != self.stamp:
            c = K.f32c(m.centroids.to(device))
            K.check(L.mn_model_set_centroids(self.handle, K.ptr(c), K.stream_of(device)), h)
            self.keep, self.stamp = c, stamp
        return h

    def route(self, x: torch.Tensor) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        """-> (assign int64 [B] or None, weights [B,K] or None), like models/mega_nerf.py:21-30."""
        dev = x.device
        h = self.sync(dev)
        xin = K.f32c(x)
        rows = K.Rows()
        rows.mode, rows.x_d, rows.cols = 0, xin.data_ptr(), xin.shape[1]
        B, Kn = xin.shape[0], len(self.mega.sub_modules)
        if self.mega.boundary_margin > 1:
            w = torch.empty(B, Kn, device=dev, dtype=torch.float32)
            K.check(K.lib().mn_model_route(h, self.handle, C.byref(rows), B, None, K.ptr(w), K.stream_of(dev)), h)
            return None, w
        a = torch.empty(B, device=dev, dtype=torch.int32)
        K.check(K.lib().mn_model_route(h, self.handle, C.byref(rows), B, K.ptr(a), None, K.stream_of(dev)), h)
        return a.long(), None


def owner_of(k: int, world: int) -> int:
    """Round-robin sub-module -> rank (BASELINE.json configs[3])."""
    return k % world


def plan_dispatch(assign: Optional[torch.Tensor], weights: Optional[torch.Tensor], n_sub: int, world: int):
    """(row, sub-module) pairs of one query, ordered by (destination rank, sub-module, row).
    -> rows [P] int64, subs [P] int64, blend weights [P] or None, send counts [world] int64."""
    if weights is None:
        rows = torch.arange(assign.shape[0], device=assign.device)
        subs = assign
        w = None
    else:
        nz = (weights > 0).nonzero()
        rows, subs = nz[:, 0], nz[:, 1]
        w = weights[rows, subs]
    dest = subs % world
    order = torch.argsort(dest * n_sub + subs, stable=True)
    rows, subs, dest = rows[order], subs[order], dest[order]
    if w is not None:
        w = w[order]
    counts = torch.bincount(dest, minlength=world)
    return rows, subs, w, counts


class ExpertParallel:
    def __init__(self, mega, group=None, route_fn: Optional[Callable] = None, sub_fn: Optional[Callable] = None):
        self.mega = mega
        self.group = group
        self.n_sub = len(mega.sub_modules)
        self._router = None
        self.route_fn = route_fn or self._route_native
        self.sub_fn = sub_fn or self._sub_native
        self.last_pairs = 0          # routed pairs of the last query that originated on this rank
        self.last_owned = 0          # pairs this rank computed for everybody

    # ---- device-side defaults
    def _route_native(self, x):
        if self._router is None:
            self._router = _RouteOnly(self.mega)
        return self._router.route(x)

    def _sub_native(self, k: int, rows: torch.Tensor, sigma_noise: Optional[torch.Tensor]) -> torch.Tensor:
        return self.mega.sub_modules[k](rows, sigma_noise=sigma_noise)

    def owned(self) -> List[int]:
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        return [k for k in range(self.n_sub) if owner_of(k, world) == rank]

    # ---- nn.Module.__call__ of MegaNeRF on rows, distributed
    def forward(self, x: torch.Tensor, sigma_noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.mega.parameters()):
            raise RuntimeError('expert-parallel execution is inference-only (wrap the call in torch.no_grad())')
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        B = x.shape[0]
        assign, weights = self.route_fn(x)
        rows, subs, w, counts = plan_dispatch(assign, weights, self.n_sub, world)
        child = x[:, 3:] if self.mega.xyz_real else x                        # mega_nerf.py:36
        cols = [child[rows], subs.to(x.dtype).unsqueeze(1)]
        if sigma_noise is not None:
            cols.append(sigma_noise.reshape(B, 1)[rows])
        payload = torch.cat(cols, 1).contiguous()
        width = payload.shape[1]

        # split sizes (one host sync), then the rows
        recv_counts = torch.empty_like(counts)
        dist.all_to_all_single(recv_counts, counts, group=self.group)
        send_l, recv_l = counts.tolist(), recv_counts.tolist()
        recv = payload.new_empty(sum(recv_l), width)
        dist.all_to_all_single(recv, payload, recv_l, send_l, group=self.group)

        # owners compute
        c_in = child.shape[1]
        rk = recv[:, c_in].long()
        out_cols = self.mega.sub_modules[0].rgb_dim + 1
        res = recv.new_zeros(recv.shape[0], out_cols)
        for k in self.owned():
            m = rk == k
            n = int(m.sum())
            if n == 0:
                continue
            nz = recv[m, c_in + 1] if sigma_noise is not None else None
            res[m] = self.sub_fn(k, recv[m, :c_in].contiguous(), nz.unsqueeze(1) if nz is not None else None).to(res.dtype)
        self.last_pairs, self.last_owned = int(rows.shape[0]), int(recv.shape[0])

        # results travel back along the same routes
        back = res.new_empty(rows.shape[0], out_cols)
        dist.all_to_all_single(back, res, send_l, recv_l, group=self.group)

        # accumulate at home, ascending sub-module order (mega_nerf.py:34,46-49)
        out = back.new_zeros(B, out_cols)
        if w is None:
            out[rows] = back
        else:
            for k in range(self.n_sub):
                m = subs == k
                if bool(m.any()):
                    out[rows[m]] += back[m] * w[m].unsqueeze(-1)
        return out


def enable(mega, group=None, **kw) -> ExpertParallel:
    """Attach owner-computes execution to a MegaNeRF: `render_rays` then queries it through the process group.
    Sub-modules this rank does not own are never evaluated here (their parameters may stay on the CPU)."""
    ep = ExpertParallel(mega, group, **kw)
    object.__setattr__(mega, '_ep', ep)
    return ep


def disable(mega) -> None:
    object.__setattr__(mega, '_ep', None)
