#!/usr/bin/env python
"""Per-source-line view of an ncu capture: instructions executed and warp-stall samples of one kernel, attributed to the
lines of petastorm_b200/csrc/kernels_decode.cu (ncu's CSV export of the source page is SASS-only; the line table comes from
nvdisasm of a cubin rebuilt from the same source, joined instruction by instruction).

    python tools/ncu_lines.py gpurun_out/prof_x.ncu-rep k_snappy_pages [--src petastorm_b200/csrc/kernels_decode.cu] [--top 40]
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

STALLS = ['stall_barrier', 'stall_wait', 'stall_short_sb', 'stall_long_sb', 'stall_selected', 'stall_branch_resolving',
          'stall_no_inst', 'stall_not_selected', 'stall_math', 'stall_mio', 'stall_lg', 'stall_membar', 'stall_dispatch']


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('kernel')
    ap.add_argument('--src', default='petastorm_b200/csrc/kernels_decode.cu')
    ap.add_argument('--top', type=int, default=40)
    ap.add_argument('--flags', default='')
    a = ap.parse_args()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, a.src)
    tmp = tempfile.mkdtemp()
    cubin = os.path.join(tmp, 'k.cubin')
    subprocess.check_call(['nvcc', '-std=c++17', '-O3', '-lineinfo', '-gencode', 'arch=compute_100a,code=sm_100a',
                           '-I' + os.path.join(root, 'include'), '-I' + os.path.dirname(src), '-cubin', '-o', cubin, src]
                          + a.flags.split())
    sass = subprocess.check_output(['nvdisasm', '--print-line-info', cubin]).decode()
    dis, cur, infn, fname = [], None, False, None
    for l in sass.split('\n'):
        if l.startswith('.text.') and a.kernel in l and '$' not in l:
            infn = True
            continue
        if infn and l.startswith('\t.section'):
            break
        if not infn:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            fname, cur = os.path.basename(m.group(1)), int(m.group(2))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            dis.append((fname, cur, m.group(2).strip()))
    out = subprocess.check_output(['ncu', '-i', a.report, '--page', 'source', '--csv', '--kernel-name',
                                   'regex:' + a.kernel, '--launch-count', '1'], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if r and r[0] == 'Address')
    col = {c: i for i, c in enumerate(hdr)}
    ncu = [r for r in rows if r and r[0].startswith('0x') and len(r) == len(hdr)]
    if len(ncu) != len(dis):
        sys.exit('SASS of the capture (%d instructions) and of the rebuilt source (%d) differ: rebuild the capture\'s '
                 'commit' % (len(ncu), len(dis)))
    bad = sum(1 for (f, l, t), r in zip(dis, ncu) if t.split()[0] != r[1].split()[0])
    if bad:
        sys.exit('%d instructions differ between the capture and the rebuilt source' % bad)
    lines = {}
    per = collections.defaultdict(lambda: [0, 0])
    stalls = collections.Counter()
    tot_i = tot_s = 0
    for (f, ln, t), r in zip(dis, ncu):
        n, s = int(r[col['Instructions Executed']]), int(r[col['# Samples']])
        per[(f, ln)][0] += n
        per[(f, ln)][1] += s
        tot_i += n
        tot_s += s
        for c in STALLS:
            if c in col:
                stalls[c] += int(r[col[c]])
    print('%s: %d SASS instructions, %d executed (warp level), %d stall samples' % (a.kernel, len(dis), tot_i, tot_s))
    ts = sum(stalls.values()) or 1
    print('stall reasons: ' + ', '.join('%s %.1f%%' % (k.replace('stall_', ''), 100.0 * v / ts)
                                        for k, v in stalls.most_common(8)))
    for (f, ln), (n, s) in sorted(per.items(), key=lambda x: -x[1][1])[:a.top]:
        if f not in lines:
            pth = os.path.join(os.path.dirname(src), f or '')
            lines[f] = open(pth).read().split('\n') if f and os.path.exists(pth) else []
        text = lines[f][ln - 1].strip() if ln and ln <= len(lines[f]) else ''
        print('%-18s %5s  inst %5.1f%%  samples %5.1f%%  %s' % (f, ln, 100.0 * n / tot_i, 100.0 * s / tot_s, text[:100]))


if __name__ == '__main__':
    main()
