#!/usr/bin/env python
"""Dynamic SASS opcode mix of a kernel from an ncu report (source page, sass view).
usage: python profiles/ncu_opmix.py rep.ncu-rep [topN]   -> executed warp instructions per opcode,
plus the shared-memory wavefront totals (actual vs ideal)."""
import csv
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    ops = defaultdict(int)
    smp = defaultdict(int)
    tot = 0
    wf = wfi = 0
    for r in rows:
        if r and r[0] == 'Address':
            hdr = {k: i for i, k in enumerate(r)}
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        src = r[hdr['Source']].strip()
        parts = src.split()
        if not parts:
            continue
        op = parts[1] if parts[0].startswith('@') and len(parts) > 1 else parts[0]
        op = op.split('.')[0]
        try:
            ex = int(r[hdr['Instructions Executed']])
            s = int(r[hdr['# Samples']])
            wf += int(r[hdr['L1 Wavefronts Shared']]); wfi += int(r[hdr['L1 Wavefronts Shared Ideal']])
        except ValueError:
            continue
        ops[op] += ex; smp[op] += s; tot += ex
    print('total executed warp instructions %d; shared wavefronts %d (ideal %d)' % (tot, wf, wfi))
    ts = sum(smp.values())
    for op, n in sorted(ops.items(), key=lambda kv: -kv[1])[:topn]:
        print('%-10s %12d  %5.1f%% inst  %5.1f%% samples' % (op, n, 100.0 * n / tot, 100.0 * smp[op] / max(ts, 1)))


if __name__ == '__main__':
    main()
