"""CPU: invariants of the role tables of the default inference MLP kernel (csrc/mn_mlp_tp.cuh), read through the host-only
test hook mn_debug_tp_program: the TMA producer's stage list and the MMA issuers' block list must describe the same ring
traffic, every accumulator half must be opened and closed exactly once, and a block must fit the ring with room to prefetch.
The kernel follows models/nerf.py:115-160 (the Linear chain of one NeRF); the shapes below are the reference's configs
(configs/mega-nerf/*.yaml: 8 x 256 with a skip at 4) plus the narrower / SH / no-direction variants the parity tests use."""
import ctypes as C

import pytest

from mega_nerf_b200 import _cabi as K

TF_FIRST, TF_LAST, TF_FROM_X, TF_WAIT_A = 2, 4, 8, 16


def desc(layer_dim=256, layers=8, skips=(4,), pos_dir_dim=4, appearance_dim=48, rgb_dim=3, affine=0, pos_xyz_dim=12):
    d = K.ModelDesc()
    d.kind, d.n_sub = 0, 1
    d.pos_xyz_dim, d.pos_dir_dim = pos_xyz_dim, pos_dir_dim
    d.layers, d.layer_dim = layers, layer_dim
    d.appearance_dim, d.affine_appearance, d.appearance_count = appearance_dim, affine, 100
    d.rgb_dim, d.xyz_dim, d.shifted_softplus = rgb_dim, 3, 1
    d.n_skip = len(skips)
    for i, s in enumerate(skips):
        d.skip_layers[i] = s
    d.boundary_margin, d.xyz_real, d.cluster_dim_start = 1.0, 1, 1
    return d


def program(d):
    cap = 1024
    tab = (C.c_uint * (4 * cap))()
    info = (C.c_int * 8)()
    rc = K.lib().mn_debug_tp_program(C.byref(d), tab, cap, info)
    if rc != 0:
        return rc, None, None, None
    n_prog, n_prog_t, n_loads, n_loads_t = info[0], info[1], info[2], info[3]
    ent = [tuple(tab[4 * i + j] for j in range(4)) for i in range(n_prog + n_loads)]
    return rc, ent[:n_prog], ent[n_prog:], list(info)


SHAPES = {
    'c2_8x256': dict(),
    'sh_head': dict(pos_dir_dim=0, appearance_dim=0, rgb_dim=27),
    'narrow_64': dict(layer_dim=64, layers=4, skips=(2,)),
    'narrow_128': dict(layer_dim=128, layers=4, skips=()),
    'w192': dict(layer_dim=192, layers=3, skips=(1,)),
    'no_appearance': dict(appearance_dim=0),
    'pe16': dict(pos_xyz_dim=16),                 # 99 encoding columns -> 7-stage feature blocks
    'deep_12': dict(layers=12, skips=(4, 8)),
}


@pytest.mark.parametrize('name', sorted(SHAPES))
def test_tables_describe_the_same_ring_traffic(name):
    d = desc(**SHAPES[name])
    rc, prog, loads, info = program(d)
    assert rc == 0
    n_prog, n_prog_t, n_loads, n_loads_t, tp_bytes, stages, smem, x_tile = info
    assert 0 < n_prog_t <= n_prog and 0 < n_loads_t <= n_loads
    assert stages >= 8 and smem <= 227 * 1024
    # the blocks consume exactly the stages the producer loads - for a full call and for a sigma_only call (trunk prefix)
    assert sum((z >> 12) & 0xF for _, _, z, _ in prog) == n_loads
    assert sum((z >> 12) & 0xF for _, _, z, _ in prog[:n_prog_t]) == n_loads_t
    # weight slices: inside the image, non-overlapping, covering it completely
    spans = sorted((x, x + y) for x, y, _, _ in loads)
    assert spans[0][0] == 0 and spans[-1][1] == tp_bytes
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 == b0, 'gap or overlap in the weight image'
    pos = 0
    open_half = None
    for x, idesc, z, code in prog:
        nw, ns, nk_last, fl = z & 0xFFF, (z >> 12) & 0xF, (z >> 16) & 0xF, z >> 20
        assert 1 <= ns <= stages - 2, 'a block must leave room in the ring for the next block to be prefetched'
        assert nw % 16 == 0 and 16 <= nw <= 128 and (idesc >> 17) & 0x3F == nw >> 3 and (idesc >> 24) & 0x1F == 8
        blk = loads[pos:pos + ns]
        pos += ns
        if fl & TF_FROM_X:
            assert nk_last == 1 and not (fl & TF_WAIT_A)
            for _, y, xo, _ in blk:
                assert y == 16 * nw * 2 and xo != 0xFFFFFFFF and (xo << 4) + 16 * 128 * 2 <= x_tile
        else:
            assert 1 <= nk_last <= 4
            for i, (_, y, xo, _) in enumerate(blk):
                assert xo == 0xFFFFFFFF and y == (nk_last if i == ns - 1 else 4) * 16 * nw * 2
        # every (GEMM, N-half) accumulator is opened by exactly one FIRST block and closed by one LAST block
        if fl & TF_FIRST:
            assert open_half is None
            open_half = code
        assert open_half == code
        if fl & TF_LAST:
            open_half = None
    assert open_half is None and pos == n_loads
    # the A operand is awaited once per GEMM that reads activations: first activation block of half 0
    waits = [code for _, _, z, code in prog if (z >> 20) & TF_WAIT_A]
    assert len(waits) == len(set(waits)) and all(c % 2 == 0 for c in waits)
    assert len(waits) == len({c >> 1 for _, _, z, c in prog if not ((z >> 20) & TF_FROM_X)})


def test_c2_counts():
    """8 x 256 with a skip at 4: 21 accumulator halves, 23 blocks (the skip layer and the view layer have a feature and an
    activation block per half), 95 ring stages per tile pair."""
    rc, prog, loads, info = program(desc())
    assert rc == 0 and len(loads) == 95 and len(prog) == 23
    assert info[4] == 2 * (80 * 256 + 3 * 256 * 256 + (80 + 256) * 256 + 3 * 256 * 256 + 256 * 256 + (256 + 80) * 128 + 128 * 32)


def test_unsupported_shapes_are_refused():
    assert program(desc(layer_dim=512))[0] != 0          # the 512-wide kernel serves this width
    assert program(desc(layer_dim=96))[0] != 0           # not a multiple of 64: fp32 kernels only
