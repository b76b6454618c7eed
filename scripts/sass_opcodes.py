"""Per-kernel SASS opcode histogram of libmn_b200.so (no GPU needed): python scripts/sass_opcodes.py > profiles/r2_sass_opcodes.txt
Counts the mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,
UBLKCP = cp.async.bulk (1-D TMA), UTMALDG = cp.async.bulk.tensor (tensor-map TMA), SYNCS = mbarrier ops, UTCBAR = tcgen05.commit,
plus the legacy tensor path (HMMA) which must be absent, and the fences that must not sit in inner loops (MEMBAR.ALL.GPU, CCTL.IVALL)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'mega_nerf_b200', 'libmn_b200.so')
KEYS = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'UTCCP', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'SYNCS', 'HMMA', 'FFMA', 'LDS', 'STS', 'LDG', 'STG',
        'REDG', 'ATOMG', 'RED', 'MEMBAR.ALL.GPU', 'CCTL.IVALL', 'MUFU', 'DFMA', 'DMUL']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', name)
            cur = per.setdefault(re.sub(r'\(.*$', '', name), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r'/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', line)
        if m:
            op = m.group(1)
            cur['_total'] += 1
            for k in KEYS:
                if op == k or op.startswith(k + '.') or (k in ('MEMBAR.ALL.GPU', 'CCTL.IVALL') and op.startswith(k)):
                    cur[k] += 1
    print(f'# SASS opcode counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; static counts, not executions)')
    print(f'# {"kernel":58s} {"instrs":>7s} ' + ' '.join(f'{k:>8s}' for k in KEYS if k not in ('MEMBAR.ALL.GPU', 'CCTL.IVALL')) + '  MEMBAR.GPU CCTL.IVALL')
    for name, c in per.items():
        print(f'{name[:60]:60s} {c["_total"]:7d} ' + ' '.join(f'{c[k]:8d}' for k in KEYS if k not in ('MEMBAR.ALL.GPU', 'CCTL.IVALL')) +
              f'  {c["MEMBAR.ALL.GPU"]:10d} {c["CCTL.IVALL"]:10d}')
    tc = [n for n, c in per.items() if c['UTCHMMA'] or c['UTCQMMA']]
    print(f'# tensor-core (tcgen05) kernels: {", ".join(tc)}')
    print(f'# legacy HMMA anywhere: {sum(c["HMMA"] for c in per.values())}')


if __name__ == '__main__':
    main()
