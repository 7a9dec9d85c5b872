"""Summarise ncu outputs into profiles/: launch list shares and key raw metrics per kernel.
usage: python scripts/summarize_ncu.py launches.csv report.ncu-rep out_prefix"""
import csv
import subprocess
import sys
from collections import defaultdict

launches, rep, prefix = sys.argv[1:4]
rows = list(csv.reader(open(launches)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]
ki, vi, ui, mi = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit'), h.index('Metric Name')
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[hdr + 1:]:
    if len(r) <= vi or r[mi] != 'gpu__time_duration.sum':
        continue
    v = float(r[vi].replace(',', ''))
    u = r[ui]
    ms = v / 1e6 if u in ('ns', 'nsecond') else (v / 1e3 if u.startswith('us') else (v if u.startswith('ms') else v * 1e3))
    tot[r[ki][:90]] += ms
    cnt[r[ki][:90]] += 1
T = sum(tot.values())
with open(prefix + '_launches.md', 'w') as f:
    f.write('| kernel | launches | total ms | share |\n|---|---|---|---|\n')
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        f.write(f'| `{k}` | {cnt[k]} | {v:.3f} | {100 * v / T:.1f}% |\n')
print(open(prefix + '_launches.md').read())

raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
names, units = rr[0], rr[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.avg.per_second', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'launch__shared_mem_per_block_dynamic']
idx = [(w, [k for k, nm in enumerate(names) if nm == w or nm.endswith('.' + w)][0]) for w in want
       if any(nm == w or nm.endswith('.' + w) for nm in names)]
with open(prefix + '_kernels.md', 'w') as f:
    for row in rr[2:]:
        f.write('\n'.join(f'- {w}: {row[i]} {units[i]}' for w, i in idx) + '\n\n')
print(open(prefix + '_kernels.md').read())
