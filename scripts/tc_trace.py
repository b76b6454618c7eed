"""In-kernel timeline of CTA 0 of the ping-pong MLP kernel (MN_TC_TRACE=1): where do the GEMM periods go?"""
import ctypes as C
import os
import sys

os.environ['MN_TC_TRACE'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import mega_nerf_b200 as M
from mega_nerf_b200 import _cabi as K
from oracle import mn_oracle as O
import cases as Cs
from test_gpu_parity import product_net

dev = torch.device('cuda:0')
spec = O.NerfSpec()
net = O.make_net('nerf', spec, seed=3)
n = 148 * 128 * 8
x = Cs.nerf_rows(spec, n, 9).to(dev)
p = product_net(net)
M.set_precision('tc_f16')
lib = K.lib()
lib.mn_debug_read_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
for it in range(3):
    lib.mn_debug_read_trace(None, None, 1)
    p(x)
buf = (C.c_ulonglong * (4 * 4096))()
cnt = (C.c_uint * 2)()
lib.mn_debug_read_trace(buf, cnt, 0)
ev = []
for who in range(2):
    for i in range(min(cnt[who], 2048)):
        tag, t = buf[(who * 2048 + i) * 2], buf[(who * 2048 + i) * 2 + 1]
        ev.append((t, who, tag >> 32, (tag >> 16) & 0xffff, tag & 0xffff))
ev.sort()
t0 = ev[0][0]
names = {1: 'mma_start', 2: 'mma_issued', 3: 'epi_accready', 4: 'epi_done'}
print('counts', list(cnt))
for t, who, e, sl, gi in ev[:140]:
    print(f'{t - t0:9d} ns  {names[e]:13s} slot {sl} gemm {gi}')
# per-GEMM period statistics for slot 0 mma_start events
st = [t for t, who, e, sl, gi in ev if e == 1 and sl == 0]
if len(st) > 20:
    d = [b - a for a, b in zip(st[5:], st[6:])]
    print('median ns between consecutive slot-0 GEMM starts:', sorted(d)[len(d) // 2])
