"""GPU: loader-side ray generation.  `get_rays_pairs` against the oracle's get_rays_batch product gathered at
the same (image, pixel) pairs, and `mega_nerf_b200.loader._load_chunk_inner` against the reference's own
`FilesystemDataset._load_chunk_inner` (baseline/_ref, unbound method on the same stand-in dataset object and the same parquet
chunk written with pyarrow in the reference's column layout, filesystem_dataset.py:95-131,222-260)."""
import os
import sys
import types
from itertools import cycle
from pathlib import Path

import pytest
import torch

import cases  # noqa: F401
from oracle import mn_oracle as O
from test_gpu_parity import DEV, M

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')


def scene(n_img=7, W=13, H=9):
    g = torch.Generator().manual_seed(11)
    dirs = O.ray_directions(W, H, 9.5, 9.1, 6.2, 3.4, True).view(-1, 3)
    q, _ = torch.linalg.qr(torch.randn(n_img, 3, 3, generator=g))
    c2w = torch.cat([q, torch.cat([-0.3 - 0.2 * torch.rand(n_img, 1, generator=g), torch.rand(n_img, 2, generator=g) - 0.5], 1).unsqueeze(-1)], -1)
    return dirs, c2w, g


@pytest.mark.parametrize('alt', [None, [-0.35, 0.05]])
def test_rays_pairs_match_batch_product(alt):
    dirs, c2w, g = scene()
    Mp = 5000
    ii = torch.randint(0, c2w.shape[0], (Mp,), generator=g, dtype=torch.int32)
    pi = torch.randint(0, dirs.shape[0], (Mp,), generator=g, dtype=torch.int32)
    want = O.rays_from_pose_batch(dirs.view(1, -1, 3).expand(c2w.shape[0], -1, -1).contiguous(), c2w, 0.1, 3.0, alt)[ii.long(), pi.long()]
    got = M().raygen.get_rays_pairs(dirs.to(DEV), c2w.to(DEV), ii.to(DEV), pi.to(DEV), 0.1, 3.0, alt)
    assert got.shape == (Mp, 8)
    assert float((got.cpu() - want).abs().max()) <= 5e-7
    # and the product entry itself at those pairs (the patched get_rays_batch with the loader's [P,3] call shape)
    prod = M().get_rays_batch(dirs.to(DEV), c2w.to(DEV), 0.1, 3.0, alt)[ii.long().to(DEV), pi.long().to(DEV)]
    assert torch.equal(prod, got)


def test_rays_pairs_bad_index_raises():
    dirs, c2w, g = scene()
    m = M()
    ii = torch.tensor([0, c2w.shape[0]], dtype=torch.int32)          # second image index is out of range
    pi = torch.tensor([0, 1], dtype=torch.int32)
    out = m.raygen.get_rays_pairs(dirs.to(DEV), c2w.to(DEV), ii.to(DEV), pi.to(DEV), 0.1, 3.0, None)
    assert torch.isnan(out[1]).all() and torch.isfinite(out[0]).all()
    from mega_nerf_b200 import _cabi as K
    with pytest.raises(RuntimeError, match='index out of range'):
        K.check(K.lib().mn_check_status(K.ctx(DEV), K.stream_of(DEV)), K.ctx(DEV))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'mega_nerf')), reason='baseline/_ref not present')
def test_chunk_loader_matches_reference_method(tmp_path):
    import numpy as np
    import pyarrow as pa
    import pyarrow.parquet as pq
    from ref_shims import install_shims
    install_shims()
    sys.path.insert(0, REF)
    try:
        from mega_nerf.datasets.filesystem_dataset import FilesystemDataset     # unmodified reference class
    finally:
        sys.path.remove(REF)
    assert FilesystemDataset._load_chunk_inner.__module__ == 'mega_nerf.datasets.filesystem_dataset'
    dirs, c2w, g = scene(n_img=5, W=16, H=10)
    rows = 70000                                                          # > RAY_CHUNK_SIZE: the reference loops twice
    img = torch.randint(0, c2w.shape[0], (rows,), generator=g, dtype=torch.int32)
    pix = torch.randint(0, dirs.shape[0], (rows,), generator=g, dtype=torch.int32)
    rgb = torch.randint(0, 256, (rows, 3), generator=g, dtype=torch.uint8)
    path = tmp_path / 'chunk0.parquet'
    cols = {'img_indices': pa.array(img.numpy()), 'pixel_indices': pa.array(pix.numpy())}
    for c in range(3):
        cols[f'rgbs_{c}'] = pa.array(rgb[:, c].numpy())
    pq.write_table(pa.table(cols), path)

    def dataset():
        return types.SimpleNamespace(_chunk_index=cycle(range(1)), _parquet_paths=[Path(path)], _directions=dirs.to(DEV), _c2ws=c2w,
                                     _device=DEV, _near=0.1, _far=3.0, _ray_altitude_range=[-0.35, 0.05])
    want = FilesystemDataset._load_chunk_inner(dataset())                    # reference: get_rays_batch product + .cpu() + gather
    from mega_nerf_b200 import loader
    got = loader._load_chunk_inner(dataset())
    assert got[0] == want[0]
    assert torch.equal(got[1], want[1]) and torch.equal(got[3], want[3])
    assert got[2].shape == want[2].shape and got[2].device.type == 'cpu'
    assert float((got[2] - want[2]).abs().max()) <= 2e-6                  # torch-CUDA matmul vs the FMA chain of mn_rays
