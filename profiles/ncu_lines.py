#!/usr/bin/env python
"""Per CUDA source line: instructions executed and stall samples (ncu source page, needs
-lineinfo and --import-source on).  usage: python profiles/ncu_lines.py rep.ncu-rep [topN]"""
import csv
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source',
                          'cuda,sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    path = '?'
    hdr = None
    agg = defaultdict(lambda: [0, 0, ''])
    tot = tots = 0
    for r in rows:
        if not r:
            continue
        if r[0] in ('File Path', 'File Name'):
            path = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            hdr = r
            col = {}
            for i, k in enumerate(hdr):
                col.setdefault(k, i)
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        try:
            ex = int(r[col['Instructions Executed']]); smp = int(r[col['# Samples']])
        except ValueError:
            continue
        key = (path, r[0])
        agg[key][0] += ex; agg[key][1] += smp
        if r[1].strip():
            agg[key][2] = r[1].strip()[:100]
        tot += ex; tots += smp
    print('total inst %d, samples %d' % (tot, tots))
    for (fn, ln), (ex, smp, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
        print('%5.1f%% smp %5.1f%% inst  %-12s:%-4s %s' % (100.0 * smp / max(tots, 1), 100.0 * ex / max(tot, 1), fn, ln, src))


if __name__ == '__main__':
    main()
