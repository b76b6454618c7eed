#!/usr/bin/env python
"""Discrete-event check of the cross-CTA handshake of tc_mlp_c2_kernel (csrc/mn_mlp_c2.cuh), all three variants:
MN_TC_C2=1 (32 arrivals on the leader's epi_done), =2 (relay), =3 (relay + trailing epilogue with per-slab barriers).

Actors of one cluster: the leader's MMA warp, 16 epilogue warps per CTA, the peer's relay lane.  mbarriers are modelled
with their phase-parity semantics (a wait on parity P succeeds once the phase with parity P has completed - so a
skipped phase would alias; that is exactly the bug class this looks for).  Threads are interleaved at random; the script
asserts (a) no deadlock, (b) the MMAs of a slot's next GEMM start only after all 32 warps have READ the accumulator of the
previous one, (c) a K-step reading activation slab j starts only after all 32 warps have WRITTEN slab j, (d) an epilogue
starts only after its GEMM's commit.  Not a performance model.

    python scripts/c2_protocol_sim.py [--schedules 300]
"""
import argparse
import random


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, 'more arrivals than the barrier expects in one phase'
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def done(self, parity):
        return (self.phase & 1) != parity


def simulate(variant, n_quads, gemms, rng):
    """gemms: list of (n_out_slabs_written, k_slabs_read_from_H) per GEMM of the plan."""
    W = 16
    trail = variant == 3
    relay = variant >= 2
    acc_full = [[Bar(1), Bar(1)] for _ in range(2)]          # [cta][slot]
    epi_done = [Bar(W + 1 if relay else 2 * W) for _ in range(2)]
    epi_local = [Bar(W) for _ in range(2)]
    h_ready = [[Bar(W + 1) for _ in range(4)] for _ in range(2)]
    h_local = [[Bar(W) for _ in range(4)] for _ in range(2)]
    loaded = {}     # (slot, seq) -> warps that finished reading the accumulator
    stored = {}     # (slot, seq, slab) -> warps that stored the slab
    committed = set()
    seq_of = [0, 0]  # GEMMs committed so far per slot (MMA side)

    def mma():
        ph = [0, 0]
        started = [False, False]
        for q in range(n_quads):
            for gi, (n_out, k_in) in enumerate(gemms):
                for sl in range(2):
                    prev = started[sl]
                    waited = 0
                    seq = seq_of[sl]
                    if trail:
                        if prev:
                            yield ('wait', h_ready[sl][0], ph[sl]); waited = 1
                    elif prev:
                        yield ('wait', epi_done[sl], ph[sl]); ph[sl] ^= 1
                    started[sl] = True
                    if prev:
                        assert len(loaded.get((sl, seq - 1), ())) == 2 * W, ('accumulator overwritten before it was read', variant, sl, seq)
                    for j in range(k_in):        # K-steps over H slabs (feature stages need no handshake)
                        if trail and prev:
                            while waited <= j:
                                yield ('wait', h_ready[sl][waited], ph[sl]); waited += 1
                        if prev:
                            assert len(stored.get((sl, seq - 1, j), ())) == 2 * W, ('activation slab read before it was written', variant, sl, seq, j)
                        yield ('step',)
                    if trail and prev:
                        while waited < 4:
                            yield ('wait', h_ready[sl][waited], ph[sl]); waited += 1
                        ph[sl] ^= 1
                    committed.add((sl, seq))
                    acc_full[0][sl].arrive(); acc_full[1][sl].arrive()     # multicast commit
                    seq_of[sl] += 1
                    yield ('step',)

    def epilogue(cta, w):
        aph = [0, 0]
        seq = [0, 0]
        for q in range(n_quads):
            for gi, (n_out, k_in) in enumerate(gemms):
                for sl in range(2):
                    yield ('wait', acc_full[cta][sl], aph[sl]); aph[sl] ^= 1
                    s = seq[sl]
                    assert (sl, s) in committed, 'epilogue before commit'
                    yield ('step',)                                   # TMEM loads
                    loaded.setdefault((sl, s), set()).add((cta, w))
                    for j in range(4):
                        yield ('step',)                               # convert / store slab j (or nothing)
                        stored.setdefault((sl, s, j), set()).add((cta, w))
                        if trail:
                            (h_ready if cta == 0 else h_local)[sl][j].arrive()
                    if not trail:
                        if relay and cta == 1:
                            epi_local[sl].arrive()
                        else:
                            epi_done[sl].arrive()
                    seq[sl] += 1

    def relay_lane():
        lp = [0, 0]
        for q in range(n_quads):
            for gi in range(len(gemms)):
                for sl in range(2):
                    if trail:
                        for j in range(4):
                            yield ('wait', h_local[sl][j], lp[sl])
                            h_ready[sl][j].arrive()
                    else:
                        yield ('wait', epi_local[sl], lp[sl])
                        epi_done[sl].arrive()
                    lp[sl] ^= 1

    actors = [mma()] + [epilogue(c, w) for c in range(2) for w in range(W)] + ([relay_lane()] if relay else [])
    pending = {i: None for i in range(len(actors))}      # i -> blocked wait op or None
    alive = set(pending)
    steps = 0
    while alive:
        runnable = [i for i in alive if pending[i] is None or pending[i][1].done(pending[i][2])]
        assert runnable, f'deadlock (variant {variant}): {len(alive)} actors blocked'
        i = rng.choice(runnable)
        pending[i] = None
        try:
            op = next(actors[i])
        except StopIteration:
            alive.discard(i)
            continue
        if op[0] == 'wait' and not op[1].done(op[2]):
            pending[i] = op
        steps += 1
    return steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--schedules', type=int, default=200)
    args = ap.parse_args()
    # (slabs written, H slabs read) per GEMM of the 8 x 256 network: layer 0 reads features only; the skip layer reads features + H;
    # dir_a writes 2 slabs; the rgb head reads 2 and writes none
    plan = [(4, 0)] + [(4, 4)] * 7 + [(4, 4), (2, 4), (0, 2)]
    for variant in (1, 2, 3):
        total = 0
        for s in range(args.schedules):
            total += simulate(variant, n_quads=3, gemms=plan, rng=random.Random(1000 * variant + s))
        print(f'MN_TC_C2={variant}: {args.schedules} random schedules, {total} events, no deadlock, no hazard')


if __name__ == '__main__':
    main()
