"""Discrete-event model of the barrier protocol of the default inference MLP kernel (csrc/mn_mlp_tp.cuh), driven by the REAL role
tables (mn_debug_tp_program).  Four actors - TMA producer, the two MMA issuers, the epilogue warps (in lockstep) - exchange the
kernel's mbarriers (full / empty per ring stage, acc_full / d_free / a_ready per tile slot, the issue token); every action takes
a random time, so many seeds explore many interleavings.  Independent of the barriers, the model tracks WHAT each resource holds
(which load a ring stage contains, which accumulator half a tile slot's TMEM columns hold and whether it was drained, which
GEMM's output the A operand is) and asserts that every read sees what the kernel's arithmetic needs.  A deadlock (all actors
blocked, nothing in flight) or a stale read fails the test.  Test infrastructure only.

With a `timing` dict (clock cycles, see scripts/tp_pipeline_model.py) the same model runs deterministically with a shared tensor
pipe and a wake-up latency per barrier hand-off, and reports where the time goes - a what-if tool for the kernel's dependency ring,
calibrated against the ncu captures under profiles/."""
import collections
import heapq
import random

TF_FIRST, TF_LAST, TF_FROM_X, TF_WAIT_A = 2, 4, 8, 16


class MBar:
    """mbarrier: `count` arrivals complete a phase; wait(parity) passes once the phase with that parity has completed."""

    def __init__(self, count, name=''):
        self.count, self.pending, self.phase, self.name = count, count, 0, name

    def arrive(self, n=1):
        for _ in range(n):
            self.pending -= 1
            if self.pending == 0:
                self.pending, self.phase = self.count, self.phase ^ 1

    def passed(self, parity):
        return self.phase != parity


class Deadlock(AssertionError):
    pass


def simulate(prog, loads, n_stages, n_pairs, odd_tail, seed=0, max_events=2_000_000, timing=None, stats=None):
    """prog / loads: the tables (issuer entries (x, idesc, z, code), producer entries).  odd_tail: the last pair has one tile."""
    rnd = random.Random(seed)
    full = [MBar(1, 'full') for _ in range(n_stages)]
    empty = [MBar(2, 'empty') for _ in range(n_stages)]
    acc_full, d_free, a_ready = [MBar(1, 'acc_full'), MBar(1, 'acc_full')], [MBar(16, 'd_free'), MBar(16, 'd_free')], [MBar(16, 'a_ready'), MBar(16, 'a_ready')]
    turn = [MBar(1, 'token'), MBar(1, 'token')]
    T = timing
    pipe_free = [0.0]                                # timing mode: the tensor pipe executes blocks in issue order
    if stats is not None:
        stats.update(tensor_busy=0.0, wait=collections.defaultdict(float))
    stage_holds = [None] * n_stages                  # (pair, load index) a ring stage contains
    acc = [dict(state='free', what=None), dict(state='free', what=None)]      # accumulator of a tile slot
    a_op = [None, None]                              # (pair, gemm) whose output the A operand of a slot holds
    a_readers = [0, 0]                               # issued, not yet completed blocks that read a slot's A operand
    stage_readers = [0] * n_stages                   # ... that read a ring stage
    n_gemm = max(code for _, _, _, code in prog) // 2 + 1
    halves = {gi: max(code for _, _, _, code in prog if code // 2 == gi) % 2 + 1 for gi in range(n_gemm)}
    reads_a = {code // 2 for _, _, z, code in prog if not ((z >> 20) & TF_FROM_X)}
    events, now, seq = [], [0.0], [0]
    in_flight = [0]

    def later(dt, fn):
        seq[0] += 1
        in_flight[0] += 1
        heapq.heappush(events, (now[0] + dt, seq[0], fn))

    def valid1(pr):
        return not (odd_tail and pr == n_pairs - 1)

    # ---- actors are generators yielding either ('wait', barrier, parity) or ('sleep', dt)
    def producer():
        stage, phase = 0, 0
        for pr in range(n_pairs):
            for li in range(len(loads)):
                yield ('wait', empty[stage], phase ^ 1)
                s = stage

                def land(s=s, pr=pr, li=li):
                    assert stage_readers[s] == 0, f'TMA overwrites ring stage {s} under {stage_readers[s]} unfinished block(s)'
                    stage_holds[s] = (pr, li)
                    full[s].arrive()
                later(T['tma'] if T else rnd.uniform(0.2, 3.0), land)           # TMA in flight
                stage += 1
                if stage == n_stages:
                    stage, phase = 0, phase ^ 1
                yield ('sleep', T['prod_stage'] if T else rnd.uniform(0.05, 0.3))

    def issuer(sl):
        stage, phase, dph, aph = 0, 0, 0, 0
        tph = 0 if sl else 1
        done_at = [0.0]                                   # completion time of this issuer's last MMA (commits are in order)
        for pr in range(n_pairs):
            v1 = valid1(pr)
            if sl and not v1:
                return
            li = 0
            for x, idesc, z, code in prog:
                ns, fl = (z >> 12) & 0xF, z >> 20
                gi, h = code // 2, code % 2
                n_mma = ns if (fl & TF_FROM_X) else 4 * (ns - 1) + ((z >> 16) & 0xF)
                if T:
                    yield ('sleep', T['decode'])            # table entry, flags, operand arithmetic
                if fl & TF_FIRST:
                    yield ('wait', d_free[sl], dph ^ 1)
                    dph ^= 1
                if fl & TF_WAIT_A:
                    yield ('wait', a_ready[sl], aph)
                    aph ^= 1
                blk = []
                for _ in range(ns):
                    yield ('wait', full[stage], phase)
                    blk.append(stage)
                    stage += 1
                    if stage == n_stages:
                        stage, phase = 0, phase ^ 1
                if v1:
                    yield ('wait', turn[sl], tph)
                    tph ^= 1
                # ---- issue: what the MMAs are about to read / write must be what the arithmetic needs
                for k, s in enumerate(blk):
                    assert stage_holds[s] == (pr, li + k), f'slot {sl} pair {pr} block {code}: stage {s} holds {stage_holds[s]}, wants load {li + k}'
                if fl & TF_FIRST:
                    assert acc[sl]['state'] == 'free', f'slot {sl} pair {pr} block {code}: accumulator not drained ({acc[sl]})'
                    acc[sl].update(state='accumulating', what=(pr, gi, h))
                else:
                    assert acc[sl] == dict(state='accumulating', what=(pr, gi, h)), f'slot {sl}: accumulates into {acc[sl]}'
                if not (fl & TF_FROM_X):
                    assert a_op[sl] == (pr, gi - 1), f'slot {sl} pair {pr} gemm {gi}: A operand holds {a_op[sl]}'
                li += ns
                reads_a_op = not (fl & TF_FROM_X)
                for s in blk:
                    stage_readers[s] += 1
                a_readers[sl] += reads_a_op
                t_issue0 = now[0]
                yield ('sleep', T['issue_fixed'] + T['issue_mma'] * n_mma if T else rnd.uniform(0.1, 0.6))          # the issue sequence itself
                if T:
                    # the pipe starts on the block's first MMA soon after the sequence begins and cannot finish before it ends
                    start = max(pipe_free[0], t_issue0 + T['first_mma'])
                    pipe_free[0] = max(start + n_mma * T['mma'], now[0])
                    done_at[0] = pipe_free[0] + T['commit']
                    if stats is not None:
                        stats['tensor_busy'] += n_mma * T['mma']
                else:
                    done_at[0] = max(done_at[0], now[0]) + rnd.uniform(0.5, 2.0) * ns     # the tensor pipe executes the block

                def complete(blk=tuple(blk), last=bool(fl & TF_LAST), sl=sl, what=(pr, gi, h), twice=not v1, reads_a_op=reads_a_op):
                    a_readers[sl] -= reads_a_op
                    for s in blk:
                        stage_readers[s] -= 1
                        empty[s].arrive(2 if twice else 1)
                    if last:
                        assert acc[sl]['what'] == what
                        acc[sl]['state'] = 'complete'
                        acc_full[sl].arrive()
                later(done_at[0] - now[0], complete)
                if v1:
                    turn[sl ^ 1].arrive()

    def epilogue():
        aph = [0, 0]
        for pr in range(n_pairs):
            slots = [0, 1] if valid1(pr) else [0]
            for gi in range(n_gemm):
                for h in range(halves[gi]):
                    for sl in slots:
                        yield ('wait', acc_full[sl], aph[sl])
                        aph[sl] ^= 1
                        assert acc[sl] == dict(state='complete', what=(pr, gi, h)), f'epilogue pair {pr} gemm {gi}.{h} slot {sl}: {acc[sl]}'
                        last_h = h == halves[gi] - 1
                        yield ('sleep', T['ld'] if T else rnd.uniform(0.1, 0.5))      # tcgen05.ld
                        if T and (last_h or T.get('d_free_late')):
                            yield ('sleep', T['math'] / 2)          # the second load completes under the first piece's arithmetic
                        acc[sl].update(state='free', what=None)
                        d_free[sl].arrive(16)
                        if T:
                            yield ('sleep', T['math'] / 2 if (last_h or T.get('d_free_late')) else T['math'])
                            if last_h and gi in _publishers:
                                yield ('sleep', T['st'])            # tcgen05.st + wait::st + fence
                        else:
                            yield ('sleep', rnd.uniform(0.1, 1.0))      # arithmetic
                        if h == halves[gi] - 1 and gi in _publishers:
                            assert a_readers[sl] == 0, f'epilogue overwrites the A operand of slot {sl} under unfinished MMAs'
                            a_op[sl] = (pr, gi)
                            a_ready[sl].arrive(16)

    # GEMMs whose output is some later GEMM's A operand (everything but the colour head / a sigma_only call's last trunk layer)
    _publishers = {g - 1 for g in reads_a if g - 1 >= 0}
    actors = {'producer': producer(), 'issuer0': issuer(0), 'issuer1': issuer(1), 'epilogue': epilogue()}
    blocked = {}

    def step(name):
        gen = actors.get(name)
        if gen is None:
            return
        while True:
            try:
                req = next(gen)
            except StopIteration:
                del actors[name]
                return
            if req[0] == 'sleep':
                later(req[1], lambda name=name: step(name))
                return
            _, bar, parity = req
            if not bar.passed(parity):
                blocked[name] = (bar, parity, now[0])
                return

    for name in list(actors):
        step(name)
    n_ev = 0
    while actors:
        # wake whoever can proceed
        progressed = False
        for name, (bar, parity, t0) in list(blocked.items()):
            if bar.passed(parity):
                del blocked[name]
                hop = (T.get('hop_' + name.rstrip('01'), T['hop']) if T else 0.0)      # per-role override: hop_issuer / hop_epilogue / hop_producer
                if stats is not None:
                    stats['wait'][(name, bar.name)] += now[0] - t0 + hop
                if T:
                    later(hop, lambda name=name: step(name))        # wake-up latency of a barrier hand-off
                else:
                    step(name)
                progressed = True
        if progressed:
            continue
        if not events:
            raise Deadlock(f'deadlock at t={now[0]:.1f}: blocked {sorted(blocked)}, running {sorted(set(actors) - set(blocked))}')
        t, _, fn = heapq.heappop(events)
        now[0] = t
        in_flight[0] -= 1
        fn()
        n_ev += 1
        assert n_ev < max_events, 'simulation does not terminate'
    # drain what is still in flight (commits of the last blocks)
    while events:
        t, _, fn = heapq.heappop(events)
        now[0] = t
        fn()
    assert all(a['state'] == 'free' for a in acc)
    return now[0]
