"""CUDA-graph replay of `render_rays` for a fixed chunk shape.

One call of `render_rays` (mega_nerf/rendering.py:15-173) is ~17 kernel launches plus the Python / ctypes work that
issues them; for the reference's chunk sizes (`image_pixel_batch_size` rays per call, runner.py:567-578) the host side
costs about as much as the GPU work itself.  Every kernel of the foreground path takes its sizes from device-side
counters (slot counts, tile counts), so a chunk of a given ray count is a static launch sequence: capture it once, then
replay it with no host work besides the copy of the inputs.

    g = GraphedRenderRays(nerf, hparams, n_rays=4096, device=dev)
    results = g(rays, image_indices)        # same dict as render_rays(...)[0]; tensors are reused by the next call

The background (NeRF++) path is not captured: the reference synchronises with the host there (the bounds check that
raises `Exception`, rendering.py:37,42,412-414) and so does ours; use `render_rays` for it.
"""
from argparse import Namespace
from typing import Dict, Optional

import torch
from torch import nn

from .render import render_rays


class GraphedRenderRays:
    def __init__(self, nerf: nn.Module, hparams: Namespace, n_rays: int, device: torch.device, with_indices: bool = True,
                 get_depth: bool = True, get_depth_variance: bool = False, warmup: int = 2, post=None):
        """`post(results)`, if given, runs right after render_rays INSIDE the captured region - e.g. the per-chunk
        exchange of a multi-GPU render (`torch.distributed.all_gather_into_tensor` on NCCL is capturable), so that a
        step stays one graph launch; whatever it returns is kept in `self.post_result`."""
        if nerf.training:
            raise ValueError('GraphedRenderRays replays the inference path; call nerf.eval() first')
        self.nerf, self.hparams = nerf, hparams
        self.flags = (get_depth, get_depth_variance, False)
        self.post = post
        self.post_result = None
        self.rays = torch.zeros(n_rays, 8, device=device, dtype=torch.float32)
        self.indices = torch.zeros(n_rays, device=device, dtype=torch.float32) if with_indices else None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.results: Optional[Dict[str, torch.Tensor]] = None
        self.warmup = warmup
        net = nerf.module if hasattr(nerf, 'module') and not hasattr(nerf, '_native') else nerf
        self._native = net._native()
        self._params = [p for sub in self._native.subs for p in sub.parameters()]
        self._versions = -1

    def _weights_version(self) -> int:
        return sum(p._version for p in self._params)

    def refresh_weights(self) -> None:
        """Re-pack the native weight images (in place: the captured graph reads the same device buffers) if a parameter
        changed since the last pack - e.g. an optimiser step between two validation renders.  Called by every replay."""
        v = self._weights_version()
        if v != self._versions:
            self._native.sync(self.rays.device)
            self._versions = v

    def _run(self) -> Dict[str, torch.Tensor]:
        with torch.no_grad():      # inference path only (a recording call would switch to the fp32 training kernels)
            res, _ = render_rays(self.nerf, None, self.rays, self.indices, self.hparams, None, None, *self.flags)
            if self.post is not None:
                self.post_result = self.post(res)
        return res

    def _load(self, rays: torch.Tensor, image_indices: Optional[torch.Tensor]) -> None:
        if rays.shape != self.rays.shape:
            raise ValueError(f'captured for rays of shape {tuple(self.rays.shape)}, got {tuple(rays.shape)}')
        self.rays.copy_(rays, non_blocking=True)
        if self.indices is not None:
            self.indices.copy_(image_indices.view(-1), non_blocking=True)   # int32 (training loaders) or float, as render_rays

    def capture(self, rays: torch.Tensor, image_indices: Optional[torch.Tensor]) -> None:
        """Warm up on a side stream (weight packing, cudaFuncSetAttribute, allocator pools), then record the graph."""
        self._load(rays, image_indices)
        cur = torch.cuda.current_stream(self.rays.device)
        side = torch.cuda.Stream(self.rays.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self._run()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.rays.device)
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: other threads of the process (e.g. the NCCL watchdog polling its events) must not invalidate the capture
        with torch.cuda.graph(self.graph, capture_error_mode='thread_local'):
            self.results = self._run()
        self._versions = self._weights_version()

    def __call__(self, rays: torch.Tensor, image_indices: Optional[torch.Tensor]) -> Dict[str, torch.Tensor]:
        if self.graph is None:
            self.capture(rays, image_indices)
        self.refresh_weights()
        self._load(rays, image_indices)
        self.graph.replay()
        return self.results
