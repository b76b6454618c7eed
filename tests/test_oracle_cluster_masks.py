"""CPU-only: oracle cluster masks (oracle/mn_oracle.py::image_cluster_masks) == the masks the reference's
scripts/create_cluster_masks.py wrote for the synthetic dataset of tests/cases.py::cluster_mask_case
(tests/golden/cluster_masks_v1.pt, tests/golden/make_cluster_masks.py).  Bit-exact: the output is boolean."""
import torch

import cases as C
from oracle import mn_oracle as O


def test_cluster_masks_match_reference():
    gd = torch.load(C.CLUSTER_GOLDEN_PATH, map_location='cpu', weights_only=False)
    case = C.cluster_mask_case()
    cams = torch.stack([im['c2w'][:3, 3] for im in case['images']])
    cent, _, _ = O.grid_centroids_from_cameras(cams, case['grid_dim'])
    assert torch.equal(cent, gd['centroids'])
    zs = torch.linspace(0, 1, case['ray_samples'])
    for im, want in zip(case['images'], gd['masks']):
        got = O.image_cluster_masks(im['W'], im['H'], im['intrinsics'], im['c2w'], gd['near'], gd['far'],
                                    gd['ray_altitude_range'], case['center_pixels'], zs, cent, case['cluster_2d'],
                                    case['boundary_margin'], case['ray_chunk_size'])
        assert got.dtype == torch.bool and torch.equal(got, want)
        assert 0 < int(want.sum()) < want.numel()


def test_every_ray_is_in_its_nearest_cluster():
    """min over samples of d_k / (d_min + 1e-8) is < 1 + eps for the cluster that is nearest at some sample."""
    rays = O.synthetic_rays(200, seed=9)
    cent = O.grid_centroids(2, 4)
    r = O.cluster_min_dist_ratios(rays, torch.linspace(0, 1, 32), cent, True)
    assert r.shape == (200, 8) and float(r.min(dim=1)[0].max()) <= 1.0 and float(r.min()) > 0.99
