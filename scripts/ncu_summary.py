#!/usr/bin/env python
"""Summarise ncu captures into profiles/ (run here, no GPU needed).

    python scripts/ncu_summary.py launches gpurun_out/launches_TAG.csv        > profiles/launches_TAG.md
    python scripts/ncu_summary.py full     gpurun_out/prof_X.ncu-rep [...]    > profiles/full_TAG.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'smsp__cycles_active.avg',
        'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fma.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active']


def launches(path):
    txt = open(path).read()
    r = csv.DictReader(io.StringIO(txt[txt.index('"ID"'):]))
    agg = collections.OrderedDict()
    for row in r:
        if row['Metric Name'] != 'gpu__time_duration.sum':
            continue
        k = row['Kernel Name'].split('(')[0]
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(row['Metric Unit'], 1.0)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print('| kernel | launches | total ms | avg us | share |')
    print('|---|---:|---:|---:|---:|')
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print('| `%s` | %d | %.3f | %.1f | %.1f%% |' % (k[:90], a[0], a[1] / 1e6, a[1] / a[0] / 1e3, 100 * a[1] / tot))
    print('\ntotal %.3f ms over %d launches (cold-cache, serialised: compare shares)' % (tot / 1e6, sum(a[0] for a in agg.values())))


def full(paths):
    for p in paths:
        out = subprocess.run(['ncu', '-i', p, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            print('### `%s`  (%s, id %s)\n' % (d.get('Kernel Name', '?')[:100], p.split('/')[-1], d.get('ID')))
            print('| metric | value | unit |')
            print('|---|---:|---|')
            for k in KEYS:
                if k in d:
                    print('| %s | %s | %s |' % (k, d[k], units[hdr.index(k)]))
            print()


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2])
    else:
        full(sys.argv[2:])
