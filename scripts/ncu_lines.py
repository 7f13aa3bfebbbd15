"""Aggregate an `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda` export by source file / line.

usage: python scripts/ncu_lines.py src.csv [top_n]
Prints, per file, samples and executed instructions, then the top_n lines by stall samples with their dominant stall reasons.
"""
import csv
import sys
from collections import defaultdict


def main(path, top=40):
    cur = None
    hdr = None
    files = defaultdict(lambda: [0, 0])
    lines = []
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            hdr = None
            continue
        if r[0] == 'Function Name':
            continue
        if r[0] == 'Line No':
            hdr = r
            continue
        if hdr is None or r[2] != '-':      # keep only the per-CUDA-line summary rows (address column is '-')
            continue
        d = dict(zip(hdr[4:], r[4:]))
        try:
            samp = int(d['# Samples'])
            inst = int(d['Instructions Executed'])
        except (KeyError, ValueError):
            continue
        files[cur][0] += samp
        files[cur][1] += inst
        stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith('stall_') and '(Not Issued)' not in k and v.isdigit() and int(v)}
        lines.append((samp, inst, cur, r[0], r[1].strip()[:90], stalls))
    ts = sum(v[0] for v in files.values()) or 1
    ti = sum(v[1] for v in files.values()) or 1
    print('file                      samples   %     inst      %')
    for f, (s, i) in sorted(files.items(), key=lambda kv: -kv[1][0]):
        print('%-24s %8d %5.1f %10d %5.1f' % (f, s, 100. * s / ts, i, 100. * i / ti))
    print()
    for samp, inst, f, ln, src, st in sorted(lines, key=lambda t: -t[0])[:top]:
        tops = ' '.join('%s=%d' % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print('%5.1f%% s %5.1f%% i  %s:%s  %s   [%s]' % (100. * samp / ts, 100. * inst / ti, f, ln, src, tops))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
