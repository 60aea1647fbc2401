#!/usr/bin/env python
"""Turn ncu outputs into the committed summaries.

    python profiles/summarize_ncu.py launches <launches.csv> <out.md> "<command>"
    python profiles/summarize_ncu.py full <raw.csv (ncu -i rep --page raw --csv)> <out.md> <traffic.json> "<command>"
"""
import csv
import io
import json
import sys
from collections import OrderedDict


def read_csv(path):
    txt = open(path, errors='replace').read()
    start = txt.find('"ID"')
    return list(csv.DictReader(io.StringIO(txt[start:])))


def launches(src, out, cmd):
    rows = read_csv(src)
    per = OrderedDict()
    for r in rows:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        v *= {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'nsecond': 1e-6, 'ms': 1.0, 'msecond': 1.0}.get(unit, 1e-6)
        per.setdefault(r['Kernel Name'], []).append(v)
    with open(out, 'w') as f:
        f.write('# ncu launch list\n\nCommand: `%s`\n(per-launch times under ncu are cold-cache and serialised - compare shares)\n\n' % cmd)
        f.write('| kernel | launches | mean ms | min | max |\n|---|---|---|---|---|\n')
        for k, v in per.items():
            f.write('| `%s` | %d | %.3f | %.3f | %.3f |\n' % (k[:70], len(v), sum(v)/len(v), min(v), max(v)))
        it = [k for k in per if any(t in k for t in ('k_fwd_last', 'k_c2c_fast', 'k_mid_green', 'k_inv_last', 'k_cg_update', 'k_cg_scal'))]
        tot = 0.0
        lines = []
        for k in it:
            m = sum(per[k])/len(per[k])
            mult = 2 if 'k_cg_scal' in k else 1
            tot += mult*m
            lines.append((k, mult*m))
        f.write('\nOne CG iteration (`fh_cg_steps`): S1 + S2 + S3 + S4 + S5 + update + 2 scalar kernels:\n\n| kernel | ms under ncu | share |\n|---|---|---|\n')
        for k, m in lines:
            f.write('| `%s` | %.3f | %.1f %% |\n' % (k[:70], m, 100*m/tot))
        f.write('| sum | %.3f | |\n' % tot)


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__block_size', 'launch__grid_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct']


def full(src, out, traffic_out, cmd):
    txt = open(src, errors='replace').read()
    start = txt.find('"ID"')
    rd = csv.reader(io.StringIO(txt[start:]))
    head = next(rd)
    units = next(rd)
    traffic = {}
    with open(out, 'w') as f:
        f.write('# ncu --set full of the CG-iteration kernels\n\nCommand: `%s`\n\n`traffic` = dram__bytes_read.sum + '
                'dram__bytes_write.sum per launch; algorithmic bytes per launch in DESIGN.md section 4.\n\n' % cmd)
        for row in rd:
            d = dict(zip(head, row))
            u = dict(zip(head, units))
            name = d.get('Kernel Name', '?')
            f.write('## `%s`\n\n| metric | value |\n|---|---|\n' % name[:90])
            tr = 0.0
            for m in WANT:
                if m in d:
                    f.write('| %s | %s %s |\n' % (m, d[m], u.get(m, '')))
                    if m.startswith('dram__bytes'):
                        val = float(d[m].replace(',', ''))
                        val *= {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}.get(u.get(m, 'byte'), 1.0)
                        tr += val
            stalls = sorted(((float(d[k].replace(',', '') or 0), k) for k in head
                             if k.startswith('smsp__pcsamp_warps_issue_stalled_') and not k.endswith('_not_issued') and d.get(k)),
                            reverse=True)[:6]
            if stalls:
                f.write('| top stall reasons (pc samples) | %s |\n' % ', '.join(
                    '%s %d' % (k.replace('smsp__pcsamp_warps_issue_stalled_', ''), v) for v, k in stalls))
            f.write('\n')
            traffic[name] = tr
    json.dump({'source': out, 'dram_bytes_per_launch': traffic}, open(traffic_out, 'w'), indent=1)


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5])
