"""CPU-only: discrete-event check of the cross-CTA barrier protocols of the CTA-pair MLP kernel (csrc/mn_mlp_c2.cuh;
MN_TC_C2 = 1, 2, 3) under random interleavings: no deadlock, the accumulator is never overwritten before all 32 epilogue
warps have read it, an activation slab is never read before all of them have written it (scripts/c2_protocol_sim.py)."""
import importlib.util
import os
import random

import cases as C


def _sim():
    spec = importlib.util.spec_from_file_location('c2_protocol_sim', os.path.join(C.ROOT, 'scripts', 'c2_protocol_sim.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_pair_kernel_handshake_protocols():
    sim = _sim()
    plan = [(4, 0)] + [(4, 4)] * 7 + [(4, 4), (2, 4), (0, 2)]          # 8 x 256 network, see the script
    for variant in (1, 2, 3):
        for s in range(6):
            assert sim.simulate(variant, n_quads=2, gemms=plan, rng=random.Random(17 * variant + s)) > 0
