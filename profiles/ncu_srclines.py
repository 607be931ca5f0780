#!/usr/bin/env python
"""Per CUDA source line of an ncu --set full --import-source on capture: warp instructions executed,
stall samples, shared-memory wavefronts.  usage: python profiles/ncu_srclines.py rep.ncu-rep [topN]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    path, col, agg = '?', None, []
    for r in rows:
        if not r:
            continue
        if r[0] in ('File Path', 'File Name'):
            path = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            col = {}
            for i, k in enumerate(r):
                col.setdefault(k, i)
            continue
        if col is None or not r[0].isdigit() or 'Instructions Executed' not in col or len(r) <= col['Instructions Executed']:
            continue
        try:
            ex = int(r[col['Instructions Executed']]); smp = int(r[col['# Samples']])
            wf = int(r[col['L1 Wavefronts Shared']] or 0)
        except (ValueError, KeyError):
            continue
        agg.append((ex, smp, wf, path, r[0], r[1].strip()[:110]))
    tot = sum(a[0] for a in agg); tots = sum(a[1] for a in agg); totw = sum(a[2] for a in agg)
    print('total warp instructions %d, samples %d, shared wavefronts %d' % (tot, tots, totw))
    for ex, smp, wf, path, ln, src in sorted(agg, reverse=True)[:topn]:
        print('%5.1f%% inst %5.1f%% smp %5.1f%% wf  %s:%s  %s' % (100.0 * ex / max(tot, 1), 100.0 * smp / max(tots, 1),
                                                               100.0 * wf / max(totw, 1), path, ln, src))


if __name__ == '__main__':
    main()
