#!/usr/bin/env python
"""Aggregate the ncu source page (needs -lineinfo + --import-source on) by CUDA source line and
by SASS opcode.  usage: python profiles/ncu_source.py rep.ncu-rep [topN]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] in ('Address', 'Line No'))
    hdr = rows[hi]
    col = {k: i for i, k in enumerate(hdr)}
    isrc = col['Source']
    iex = col['Instructions Executed']
    ismp = col['# Samples']
    by_op = defaultdict(lambda: [0, 0])
    tot = 0
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            ex = int(r[iex]); smp = int(r[ismp])
        except ValueError:
            continue
        sass = r[isrc].strip()
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', sass)
        op = m.group(2) if m else sass[:12]
        op = op.split('.')[0] + ('.' + op.split('.')[1] if '.' in op and op.split('.')[0] in ('LDS', 'STS', 'LDG', 'STG', 'BAR') else '')
        by_op[op][0] += ex; by_op[op][1] += smp
        tot += ex
    print('total warp-instructions executed: %d' % tot)
    for op, (ex, smp) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:topn]:
        print('  %-14s %12d  %5.1f%%   samples %7d' % (op, ex, 100.0 * ex / tot, smp))


if __name__ == '__main__':
    main()
