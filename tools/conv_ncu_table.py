"""Markdown table + traffic JSON from an ncu CSV (--csv --page raw style log) of the convolution launches of one step.
    python tools/conv_ncu_table.py gpurun_out/r2_conv_metrics.csv > profiles/r2_conv_ncu_summary.md
Also rewrites profiles/conv_tc_traffic.json (mean DRAM bytes per convolution launch: bench.py reports it as roofline.traffic)."""
import collections
import csv
import json
import os
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
by_id = collections.OrderedDict()
for r in rows:
    d = by_id.setdefault(r['ID'], dict(kernel=re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('shgan::', ''), grid=r['Grid Size']))
    try:
        d[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
    except ValueError:
        pass
print('| # | kernel | grid | time us | tensor pipe active % | DRAM rd+wr MB | L2 hit % |')
print('|---|---|---|---|---|---|---|')
tot_bytes, n, tot_us = 0.0, 0, 0.0
for i, (k, d) in enumerate(by_id.items()):
    t = d.get('gpu__time_duration.sum', 0) / 1e3
    byts = d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
    print(f"| {i} | {d['kernel'][:40]} | {d['grid']} | {t:.1f} | {d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 0):.1f} | "
          f"{byts / 1e6:.1f} | {d.get('lts__t_sector_hit_rate.pct', 0):.0f} |")
    tot_bytes += byts; n += 1; tot_us += t
print(f'\n{n} launches, {tot_us / 1e3:.3f} ms serialized, mean DRAM traffic per launch {tot_bytes / max(n, 1) / 1e6:.1f} MB')
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'conv_tc_traffic.json')
json.dump(dict(source=f'profiles/r2_conv_ncu_summary.md (ncu metrics pass over the {n} convolution launches of one batch-16 512^2 step; '
                      'a cited ncu figure, not re-measured by bench.py)', launches=n, dram_bytes_per_launch_avg=tot_bytes / max(n, 1)),
          open(out, 'w'), indent=1)
