"""In-kernel timeline of CTA 0 of the ping-pong MLP kernel (MN_TC_TRACE=1): where do the GEMM periods go?"""
import ctypes as C
import os
import sys

os.environ['MN_TC_TRACE'] = '1'
os.environ.setdefault('MN_TC_TP', '0')      # the issuer / epilogue timeline events live in the shared-memory ping-pong kernel
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import mega_nerf_b200 as M
from mega_nerf_b200 import _cabi as K
from oracle import mn_oracle as O
import cases as Cs
from mega_nerf_b200.synthetic import build_net

dev = torch.device('cuda:0')
WIDTH = int(os.environ.get('TRACE_WIDTH', '256'))
spec = O.NerfSpec(layer_dim=WIDTH)
net = O.make_net('nerf', spec, seed=3)
n = 148 * 128 * (8 if WIDTH == 256 else 4)
x = Cs.nerf_rows(spec, n, 9).to(dev)
p = build_net(net, dev)
M.set_precision('tc_f16')
lib = K.lib()
lib.mn_debug_read_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
for it in range(3):
    lib.mn_debug_read_trace(None, None, 1)
    p(x)
buf = (C.c_ulonglong * (4 * 4096))()
cnt = (C.c_uint * 2)()
lib.mn_debug_read_trace(buf, cnt, 0)
clk = (C.c_ulonglong * 4)()
lib.mn_debug_read_clock.argtypes = [C.c_void_p]
lib.mn_debug_read_clock(clk)
if clk[3] > clk[1]:
    print(f'SM clock during the kernel (CTA 0): {(clk[2] - clk[0]) / (clk[3] - clk[1]):.3f} GHz over {(clk[3] - clk[1]) / 1e3:.1f} us')
ev = []
for who in range(2):
    for i in range(min(cnt[who], 2048)):
        tag, t = buf[(who * 2048 + i) * 2], buf[(who * 2048 + i) * 2 + 1]
        ev.append((t, who, tag >> 32, (tag >> 16) & 0xffff, tag & 0xffff))
ev.sort()
t0 = ev[0][0]
names = {1: 'mma_start', 2: 'mma_issued', 3: 'epi_accready', 4: 'epi_done'}
print('counts', list(cnt))
W0 = int(os.environ.get('TRACE_SKIP', '0'))
for t, who, e, sl, gi in ev[W0:W0 + int(os.environ.get('TRACE_SHOW', '140'))]:
    print(f'{t - t0:9d} ns  {names[e]:13s} slot {sl} gemm {gi}')
# per-GEMM period statistics for slot 0 mma_start events
st = [t for t, who, e, sl, gi in ev if e == 1 and sl == 0]
if len(st) > 20:
    d = [b - a for a, b in zip(st[5:], st[6:])]
    print('median ns between consecutive slot-0 GEMM starts:', sorted(d)[len(d) // 2])

# per (gemm, slot/half) medians over all recorded tiles: issue time, dependency stall before the next start, epilogue time
import collections
mma = [(t, e, sl, gi) for t, who, e, sl, gi in ev if who == 0]
epi = [(t, e, sl, gi) for t, who, e, sl, gi in ev if who == 1]
issue, stall, epit = collections.defaultdict(list), collections.defaultdict(list), collections.defaultdict(list)
for (ta, ea, sa, ga), (tb, eb, sb, gb) in zip(mma, mma[1:]):
    if ea == 1 and eb == 2 and (sa, ga) == (sb, gb):
        issue[(ga, sa)].append(tb - ta)
    if ea == 2 and eb == 1:
        stall[(gb, sb)].append(tb - ta)
for (ta, ea, sa, ga), (tb, eb, sb, gb) in zip(epi, epi[1:]):
    if ea == 3 and eb == 4 and (sa, ga) == (sb, gb):
        epit[(ga, sa)].append(tb - ta)
med = lambda v: sorted(v)[len(v) // 2] if v else -1
print('gemm slot/half | issue ns | stall-before ns | epilogue ns')
tot_i = tot_s = 0
for k in sorted(issue):
    print(f'{k[0]:4d} {k[1]:2d} | {med(issue[k]):6d} | {med(stall[k]):6d} | {med(epit[k]):6d}')
    tot_i += med(issue[k]); tot_s += max(med(stall[k]), 0)
print('sum of medians: issue', tot_i, 'stall', tot_s)
