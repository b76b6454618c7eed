"""CPU: the barrier protocol of the default inference MLP kernel (csrc/mn_mlp_tp.cuh) in a discrete-event model driven by the
kernel's real role tables (tests/tp_protocol_sim.py): no deadlock and no stale read - ring stages, accumulators, A operands -
under randomised timing, for every network shape the parity tests use, full and sigma_only calls, odd tile counts and the
smallest ring the launcher accepts.  A ring that cannot hold a block must be reported as a deadlock (the model can fail)."""
import pytest

import tp_protocol_sim as S
from test_tp_program import SHAPES, desc, program


@pytest.mark.parametrize('name', sorted(SHAPES))
@pytest.mark.parametrize('odd_tail', [False, True])
def test_no_deadlock_no_stale_read(name, odd_tail):
    rc, prog, loads, info = program(desc(**SHAPES[name]))
    assert rc == 0
    n_prog_t, n_loads_t, stages = info[1], info[3], info[5]
    for seed in range(4):
        S.simulate(prog, loads, stages, n_pairs=3, odd_tail=odd_tail, seed=seed)
    S.simulate(prog, loads, 8, n_pairs=3, odd_tail=odd_tail, seed=11)                       # smallest ring the launcher accepts
    S.simulate(prog[:n_prog_t], loads[:n_loads_t], stages, n_pairs=3, odd_tail=odd_tail, seed=5)      # sigma_only: trunk prefix


def test_model_reports_a_ring_smaller_than_a_block():
    rc, prog, loads, info = program(desc())
    assert rc == 0 and max((z >> 12) & 0xF for _, _, z, _ in prog) == 5
    with pytest.raises(S.Deadlock):
        S.simulate(prog, loads, 4, n_pairs=2, odd_tail=False, seed=0)


def test_model_reports_a_missing_release():
    """Dropping the second release of a ring stage when the pair has one tile (what the odd-tail path of the issuer adds) starves
    the producer: the model must notice."""
    rc, prog, loads, info = program(desc())
    orig = S.MBar.arrive
    try:
        S.MBar.arrive = lambda self, n=1: orig(self, 1)
        with pytest.raises(S.Deadlock):
            S.simulate(prog, loads, info[5], n_pairs=2, odd_tail=True, seed=0)
    finally:
        S.MBar.arrive = orig
