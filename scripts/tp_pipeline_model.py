"""What-if model of the TMEM ping-pong kernel's dependency ring (csrc/mn_mlp_tp.cuh): the CPU protocol model of
tests/tp_protocol_sim.py in its deterministic timing mode, driven by the kernel's real role tables for the 8 x 256 network.
Durations are clock cycles of the SM under load (~1.7 GHz), taken from the ncu captures under profiles/ (instruction counts x
~4.5-5 clk per issuer instruction, epilogue busy time per accumulator half) except `hop` - the wake-up latency of a barrier
hand-off - which is fitted to the measured rate.  Prints clk per tile pair, tensor-pipe utilisation and the waits per role, then a
sensitivity table.  No GPU needed:  python scripts/tp_pipeline_model.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import tp_protocol_sim as S
from test_tp_program import desc, program

BASE = dict(
    mma=64,            # 128 x 128 x 16 MMA at the pipe's floor (scripts/probes/mma_chain_probe.cu)
    issue_fixed=100,   # issue sequence of a block: ~135 SASS instructions for 16 MMAs + 4 commits, ~60 for 5 + 5
    issue_mma=36,
    first_mma=150,     # sequence start -> first MMA in the queue
    decode=350,        # table entry, flags, barrier addresses, operand arithmetic of the NEXT block (~80 instructions)
    commit=50,         # last MMA done -> mbarrier arrival
    hop=700,           # arrival -> the waiting thread's next instruction: FITTED so that the model reproduces both measured rates
    ld=220,            # two tcgen05.ld x16 + wait
    math=620,          # bias / ReLU / pack of 2 x 16 columns in 16 warps (4 per scheduler), incl. their internal stalls
    st=160,            # tcgen05.st x2 + wait::st + fence
    tma=800, prod_stage=260,
)


def run(T, n_pairs=8):
    rc, prog, loads, info = program(desc())
    assert rc == 0
    st_a, st_b = {}, {}
    t_a = S.simulate(prog, loads, info[5], n_pairs=2, odd_tail=False, timing=T, stats=st_a)
    t_b = S.simulate(prog, loads, info[5], n_pairs=n_pairs, odd_tail=False, timing=T, stats=st_b)
    per_pair = (t_b - t_a) / (n_pairs - 2)
    busy = (st_b['tensor_busy'] - st_a['tensor_busy']) / (n_pairs - 2)
    waits = {k: (st_b['wait'][k] - st_a['wait'].get(k, 0.0)) / (n_pairs - 2) for k in st_b['wait']}
    return per_pair, busy / per_pair, waits


def main():
    per_pair, util, waits = run(dict(BASE, d_free_late=True))
    print(f'token-alternating issuers, d_free after the first piece (measured: 1197 TFLOP/s = 64.9k clk per pair, tensor pipe 54 %):')
    print(f'  model {per_pair / 1e3:.1f}k clk per pair, tensor pipe {100 * util:.0f} %')
    per_pair, util, waits = run(BASE)
    print(f'final build, d_free before the arithmetic (measured: 1220 TFLOP/s = 63.7k clk per pair):')
    print(f'  model {per_pair / 1e3:.1f}k clk per pair, tensor pipe {100 * util:.0f} %')
    print('  waits per pair (clk, incl. the wake-up latency):')
    for (who, bar), v in sorted(waits.items(), key=lambda kv: -kv[1]):
        if v > 200:
            print(f'    {who:9s} on {bar:8s} {v / 1e3:6.1f}k')
    base = per_pair
    print('sensitivity (final build): parameter -> clk per pair, change')
    for k, f in [('hop', 0.5), ('hop', 0.15), ('hop_issuer', 0.15), ('hop_epilogue', 0.15), ('math', 0.5), ('ld', 0.5), ('st', 0.0), ('decode', 0.5),
                 ('issue_mma', 0.5), ('commit', 0.0)]:
        T = dict(BASE)
        T[k] = BASE.get(k, BASE['hop']) * f
        pp, u, _ = run(T)
        print(f'  {k:10s} x {f:<4} -> {pp / 1e3:6.1f}k  ({100 * (base / pp - 1):+5.1f} % throughput, tensor pipe {100 * u:.0f} %)')
    ideal = sum((((z >> 12) & 0xF) if ((z >> 20) & 8) else 4 * (((z >> 12) & 0xF) - 1) + ((z >> 16) & 0xF)) for _, _, z, _ in program(desc())[1]) * 64 * 2
    print(f'ideal (tensor pipe never idle): {ideal / 1e3:.1f}k clk per pair')


if __name__ == '__main__':
    main()
