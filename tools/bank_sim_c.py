#!/usr/bin/env python
"""Shared-memory bank model for stage3c.cuh (order 3, two elements per warp): wavefronts of every
access pattern for a padded DOF-block layout (row stride RS, plane stride SZ, element stride EL,
in doubles).  64-bit accesses are served per half-warp over 16 bank pairs, 128-bit accesses per
quarter-warp over 8 bank quads.  usage: python tools/bank_sim_c.py [RS SZ EL]"""
import itertools
import sys


def wf64(addrs):          # addrs: 32 double offsets (None = inactive)
    tot = 0
    for h in range(2):
        banks = {}
        for a in addrs[16 * h:16 * h + 16]:
            if a is None:
                continue
            banks.setdefault(a % 16, set()).add(a)
        tot += max([len(v) for v in banks.values()], default=0)
    return tot


def wf128(addrs):         # addrs: 32 double offsets (even), each lane reads 2 doubles
    tot = 0
    for q in range(4):
        banks = {}
        for a in addrs[8 * q:8 * q + 8]:
            if a is None:
                continue
            banks.setdefault((a // 2) % 8, set()).add(a)
        tot += max([len(v) for v in banks.values()], default=0)
    return tot


def model(RS, SZ, EL, D1=4):
    lanes = range(32)
    el = [L // 16 for L in lanes]
    b = [(L // 4) % 4 for L in lanes]
    a = [L % 4 for L in lanes]
    res = {}
    # rows (x-round, fill, tail XO reads): lane = (el, iz=b, iy=a), two 128-bit halves
    res['row128'] = sum(wf128([el[L] * EL + b[L] * SZ + a[L] * RS + 2 * h for L in lanes]) for h in range(2))
    # y-lines: lane = (el, iz=b, ix=a): 64-bit at k
    res['yline64'] = sum(wf64([el[L] * EL + b[L] * SZ + k * RS + a[L] for L in lanes]) for k in range(D1))
    # y-lines with the lane map of stage3c.cuh (order 3): half-warp h owns planes {h, h + 2} of both
    # elements: ely = (lane >> 3) & 1, ya = lane & 3, yb = 2 * ((lane >> 2) & 1) + (lane >> 4)
    ely = [(L >> 3) & 1 for L in lanes]
    ya = [L & 3 for L in lanes]
    yb = [2 * ((L >> 2) & 1) + (L >> 4) for L in lanes]
    res['yline64_remapped'] = sum(wf64([ely[L] * EL + yb[L] * SZ + k * RS + ya[L] for L in lanes])
                                  for k in range(D1))
    # z-lines: lane = (el, iy=b, ix=a)
    res['zline64'] = sum(wf64([el[L] * EL + k * SZ + b[L] * RS + a[L] for L in lanes]) for k in range(D1))
    return res


if __name__ == '__main__':
    if len(sys.argv) == 4:
        print(model(*[int(x) for x in sys.argv[1:]]))
    else:
        best = []
        for RS, SZp, ELp in itertools.product([4, 6, 8], range(0, 17, 2), range(0, 17, 2)):
            SZ = 4 * RS + SZp
            EL = 4 * SZ + ELp
            r = model(RS, SZ, EL)
            best.append((r['row128'] + min(r['yline64'], r['yline64_remapped']) + r['zline64'], EL, RS, SZ, r))
        best.sort(key=lambda t: (t[0], t[1]))
        for t in best[:12]:
            print(t)
