#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) and/or an ncu launch-list csv into a small text file for profiles/."""
import collections
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_static', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
        'smsp__cycles_active.avg', 'smsp__inst_executed.sum']


def rep_summary(path, out):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        out.write('no kernels in %s\n' % path)
        return
    head, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(head)}
    for r in rows[2:]:
        out.write('kernel: %s\n' % r[idx['Kernel Name']])
        for w in WANT:
            if w in idx:
                out.write('  %-62s %s %s\n' % (w, r[idx[w]], units[idx[w]]))
        out.write('\n')


def source_hotspots(path, out, kernel_regex, top=25):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kernel_regex],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return
    head = rows[1]
    si, ai, ii = head.index('# Samples'), head.index('Source'), head.index('Instructions Executed')
    seen, data = set(), []
    for r in rows[2:]:
        if len(r) > si and r[si].isdigit() and r[0] not in seen:
            seen.add(r[0])
            data.append(r)
    total = sum(int(r[si]) for r in data) or 1
    out.write('hot instructions of %s (warp stall samples, %d total):\n' % (kernel_regex, total))
    for r in sorted(data, key=lambda r: -int(r[si]))[:top]:
        out.write('  %5.1f%%  exec=%10s  %s\n' % (100.0 * int(r[si]) / total, r[ii], r[ai].strip()[:100]))
    out.write('\n')


def launches_summary(path, out):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    head = rows[hdr]
    ki, vi, mi = head.index('Kernel Name'), head.index('Metric Value'), head.index('Metric Name')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        if len(r) <= vi or r[mi] != 'gpu__time_duration.sum':
            continue
        agg[r[ki]][0] += 1
        agg[r[ki]][1] += float(r[vi].replace(',', ''))
    total = sum(v[1] for v in agg.values())
    out.write('launch list %s: %d launches, %.3f ms of kernel time\n' % (path, sum(v[0] for v in agg.values()), total / 1e6))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write('  %6.2f%%  n=%4d  avg=%10.3f us  %s\n' % (100 * v[1] / total, v[0], v[1] / v[0] / 1e3, k[:110]))
    out.write('\n')


if __name__ == '__main__':
    out = sys.stdout
    args = sys.argv[1:]
    hot = None
    if '--hot' in args:
        i = args.index('--hot')
        hot = args[i + 1]
        del args[i:i + 2]
    for p in args:
        if p.endswith('.ncu-rep'):
            rep_summary(p, out)
            if hot:
                source_hotspots(p, out, hot)
        else:
            launches_summary(p, out)
