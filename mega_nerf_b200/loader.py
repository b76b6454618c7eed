"""Loader-side ray generation: a replacement for `FilesystemDataset._load_chunk_inner`
(mega_nerf/datasets/filesystem_dataset.py:95-131) that turns a parquet chunk's (image index, pixel index) columns into
rays with ONE launch of `mn_rays_pairs` per chunk instead of, per 64k rows, the full [#unique images, #unique pixels, 8]
`get_rays_batch` product on the device, its `.cpu()` copy and a host-side gather.  Same return value, same side effects
(`torch.cuda.set_device` under torchrun, the chunk iterator); `install()` binds it over the reference's method.

Only the branch with a shared direction table (`self._directions is not None`, i.e. identical intrinsics - the case the
reference generates rays in) changes; chunks that store rays explicitly are read as before."""
from __future__ import annotations

import os
from typing import Tuple

import torch

from .raygen import get_rays_pairs


def chunk_rays(directions: torch.Tensor, c2ws: torch.Tensor, img_indices: torch.Tensor, pixel_indices: torch.Tensor, near: float,
               far: float, ray_altitude_range, device: torch.device) -> torch.Tensor:
    """rays [M,8] (host tensor, like the reference's `loaded_rays`) of a chunk's (image, pixel) pairs."""
    pairs = get_rays_pairs(directions.to(device), c2ws.to(device), img_indices, pixel_indices, near, far, ray_altitude_range)
    return pairs.cpu()


def _load_chunk_inner(self) -> Tuple[str, torch.FloatTensor, torch.FloatTensor, torch.Tensor]:
    """Drop-in for FilesystemDataset._load_chunk_inner (same attribute names as the reference's dataset object)."""
    import pyarrow.parquet as pq
    if 'RANK' in os.environ:
        torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    next_index = next(self._chunk_index)
    chosen = self._parquet_paths[next_index]
    loaded_chunk = pq.read_table(chosen)
    loaded_img_indices = torch.IntTensor(loaded_chunk['img_indices'].to_numpy().astype('int32'))
    if self._directions is not None:
        loaded_pixel_indices = torch.IntTensor(loaded_chunk['pixel_indices'].to_numpy())
        loaded_rays = chunk_rays(self._directions, self._c2ws, loaded_img_indices, loaded_pixel_indices, self._near, self._far,
                                 self._ray_altitude_range, self._device)
    else:
        loaded_rays = torch.FloatTensor(loaded_chunk.to_pandas()[['rays_{}'.format(i) for i in range(8)]].to_numpy())
    rgbs = torch.FloatTensor(loaded_chunk.to_pandas()[['rgbs_{}'.format(i) for i in range(3)]].to_numpy()) / 255.
    return str(chosen), rgbs, loaded_rays, loaded_img_indices
