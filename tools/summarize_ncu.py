#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ into profiles/ (tracked).

  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r1_launches.md
  python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/r1_full.md
"""
import csv
import subprocess
import sys
from collections import defaultdict

FULL_METRICS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    d = defaultdict(list)
    for r in rows[1:]:
        try:
            d[r[ki].split('(')[0][-70:]].append(float(r[vi].replace(',', '')))
        except ValueError:
            pass
    tot = sum(sum(v) for v in d.values())
    with open(dst, 'w') as f:
        f.write(f'# ncu launch list ({src}): gpu__time_duration.sum, --clock-control none\n\n')
        f.write('Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n')
        f.write('| kernel | launches | avg us | share |\n|---|---|---|---|\n')
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write(f'| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / tot:.3f} |\n')


def full(src, dst):
    if src.endswith('.csv'):       # already exported on the GPU box: ncu -i X.ncu-rep --page raw --csv > X.raw.csv
        out = open(src).read()
    else:
        out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, 'w') as f:
        f.write(f'# ncu --set full ({src})\n\n')
        for r in rows[2:]:
            if len(r) < len(hdr):
                continue
            f.write(f'## `{r[hdr.index("Kernel Name")][:110]}`\n\n| metric | value | unit |\n|---|---|---|\n')
            for m in FULL_METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f'| {m} | {r[i]} | {units[i]} |\n')
            f.write('\n')


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])
