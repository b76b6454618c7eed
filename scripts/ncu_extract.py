"""Reads committed `ncu --set full` reports (no GPU needed: `ncu -i`) and writes the per-kernel summary bench.py and
profiles/README.md quote:  python scripts/ncu_extract.py
  profiles/kernel_traffic.json   one entry per (report, kernel): DRAM bytes per launch, duration, tensor-pipe %, L2 -> SM bytes,
                                 shared-memory wavefronts by client, SM clock during the capture.
Entries are keyed by (kernel, workload, precision) - the workload / precision of a report come from REPORTS below."""
import csv
import io
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, 'profiles')
# report file -> (workload, precision, what was captured)
REPORTS = {
    'r1_tc_mlp_pp_kernel.ncu-rep': ('c2', 'tc_f16', 'fine pass of bench.py (4980 tiles x 128 rows), round-1 ping-pong kernel'),
    'r1_tc_mlp_wide_kernel.ncu-rep': ('c4', 'tc_f16', 'fine pass of bench.py --workload c4 (5120 tiles x 128 rows)'),
}
for fn in sorted(os.listdir(PROF)):
    if fn.endswith('.ncu-rep') and fn not in REPORTS and os.path.exists(os.path.join(PROF, fn[:-8] + '.meta.json')):
        m = json.load(open(os.path.join(PROF, fn[:-8] + '.meta.json')))
        REPORTS[fn] = (m['workload'], m['precision'], m['what'])

WANT = {
    'gpu__time_duration.sum': 'duration',
    'dram__bytes_read.sum': 'dram_read',
    'dram__bytes_write.sum': 'dram_write',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed': 'tensor_pipe_active_pct',
    'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed': 'tensor_pipe_active_pct_alt',
    'l1tex__m_xbar2l1tex_read_bytes.sum': 'l2_to_sm_bytes',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum': 'smem_wavefronts_lsu',
    'l1tex__data_pipe_tc_wavefronts_mem_shared.sum': 'smem_wavefronts_tensor_core',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum': 'smem_lsu_arbitration_replays',
    'sm__cycles_elapsed.max': 'sm_cycles_elapsed',
    'sm__cycles_elapsed.max.per_second': 'sm_clock_hz',
    'lts__t_sector_hit_rate.pct': 'l2_hit_rate_pct',
    'launch__registers_per_thread': 'registers_per_thread',
    'launch__grid_size': 'grid',
    'launch__block_size': 'block',
}
UNIT = {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12, 'byte': 1.0, 'us': 1e-6, 'ms': 1e-3, 'ns': 1e-9, 's': 1.0,
        'Ghz': 1e9, 'Mhz': 1e6, 'hz': 1.0, 'GHz': 1e9, 'MHz': 1e6, 'cycle/nsecond': 1e9, 'cycle/usecond': 1e6, 'cycle/second': 1.0}


def num(v, unit):
    try:
        x = float(v.replace(',', ''))
    except ValueError:
        return v
    return x * UNIT.get(unit, 1.0)


def main():
    out = []
    for fn, (wl, prec, what) in sorted(REPORTS.items(), reverse=True):      # latest round first: bench.py takes the first match
        path = os.path.join(PROF, fn)
        if not os.path.exists(path):
            continue
        r = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True)
        rows = list(csv.reader(io.StringIO(r.stdout)))
        if len(rows) < 3:
            print(f'{fn}: no rows ({r.stderr[-200:]})')
            continue
        hdr, units = rows[0], rows[1]
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            u = dict(zip(hdr, units))
            name = d.get('Kernel Name', '?')
            mt = re.search(r'(\w+_kernel)', name)
            short = mt.group(1) if mt else name
            e = {'kernel': short, 'kernel_full': name, 'workload': wl, 'precision': prec, 'what': what, 'source': f'profiles/{fn}'}
            for k, key in WANT.items():
                if k in d and d[k] != '':
                    e[key] = num(d[k], u.get(k, ''))
            if 'dram_read' in e and 'dram_write' in e:
                e['dram_bytes'] = e['dram_read'] + e['dram_write']
            out.append(e)
            print(f"{fn}: {short}: {e.get('duration', 0) * 1e6:.1f} us, dram {e.get('dram_bytes', 0) / 1e6:.1f} MB, tensor pipe "
                  f"{e.get('tensor_pipe_active_pct', '-')} %, L2->SM {e.get('l2_to_sm_bytes', 0) / 1e9:.2f} GB, SM clock {e.get('sm_clock_hz', 0) / 1e9:.3f} GHz")
    json.dump(out, open(os.path.join(PROF, 'kernel_traffic.json'), 'w'), indent=1)
    print(f'wrote profiles/kernel_traffic.json ({len(out)} entries)')


if __name__ == '__main__':
    main()
