#!/usr/bin/env python
"""Shared-memory wavefront model of k_stage3w (warp per element, DMMA fragments): lane = 4g + c.
64-bit accesses are served per half-warp (16 lanes, 16 bank pairs), 128-bit per quarter-warp
(8 lanes, 8 bank quads).  Compares the layouts of the first version ('old') with the swizzled
ones ('new'); run it when changing SmemW.   usage: python tools/bank_sim_w.py [D1 Q]"""
import sys
from collections import defaultdict

D1, Q = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4, 6)
RQ = (Q + 1) & ~1
QQ, NL, NY, NC, NT1, NT2 = Q * Q, D1 * D1, D1 * Q, Q * Q, 6 * D1, 6 * Q
KF, KB = (D1 + 3) // 4, (Q + 3) // 4
LANES = [(l >> 2, l & 3) for l in range(32)]


def wf64(addrs):
    tot = 0
    for h in (0, 16):
        banks = defaultdict(set)
        for a in addrs[h:h + 16]:
            if a is not None:
                banks[a % 16].add(a)
        tot += max([len(v) for v in banks.values()] + [0])
    return tot


def wf128(addrs):
    tot = 0
    for h in range(0, 32, 8):
        banks = defaultdict(set)
        for a in addrs[h:h + 8]:
            if a is not None:
                assert a % 2 == 0
                banks[(a // 2) % 8].add(a)
        tot += max([len(v) for v in banks.values()] + [0])
    return tot


def face_geom(f):
    axis = 2 if f in (0, 5) else (1 if f in (1, 3) else 0)
    side = 1 if f in (2, 3, 5) else 0
    return axis, side


class Old:
    PZ = ((QQ + 3) // 8) * 8 + 4
    PA = D1 * PZ
    name = 'old'

    def U(s, z, y, x): return (z * D1 + y) * D1 + x
    def BU(s, arr, z, y, qx): return arr * NL * RQ + (z * D1 + y) * RQ + qx
    def G3(s, arr, z, qy, qx): return arr * s.PA + z * s.PZ + qy * Q + qx
    def S2(s, iz, iy, qx): return (iz * D1 + iy) * RQ + qx
    def F1(s, f, j, q): return (f * D1 + j) * RQ + q
    def lineB(s, line): return divmod(line, Q)            # -> (z, qx)
    def bu_pair(s): return True
    def g3_pair_b(s): return False
    def g3_pair_c(s): return False
    def s2_pair(s): return False


class New(Old):
    name = 'new'
    SZU = D1 * D1 + 4

    def U(s, z, y, x): return z * s.SZU + y * D1 + (x ^ (z & (D1 - 1)))
    def BU(s, arr, z, y, qx):
        r = z * D1 + y
        return arr * NL * RQ + r * RQ + (qx ^ ((r >> 3) & 1))
    def G3(s, arr, z, qy, qx): return arr * (QQ * D1) + qy * Q * D1 + ((D1 * qx + z) ^ (D1 * ((qy >> 1) & 1)))
    def S2(s, iz, iy, qx): return (D1 * D1 + 4) * qx + D1 * iz + iy
    def lineB(s, line):
        qx, z = divmod(line, D1)
        return z, qx
    def g3_pair_b(s): return True
    def g3_pair_c(s): return True


def run(L):
    res = {}
    def add(name, n):
        res[name] = res.get(name, 0) + n
    # ---- A: fwd-x loads U as A operand (row = line (z,y), k = ix)
    for t in range((NL + 7) // 8):
        for ks in range(KF):
            ad = []
            for g, c in LANES:
                line, k = t * 8 + g, ks * 4 + c
                ad.append(L.U(line // D1, line % D1, k) if line < NL and k < D1 else None)
            add('A load U', wf64(ad))
        for arr in range(2):      # double2 store of (qx = 2c, 2c+1)
            ad = []
            for g, c in LANES:
                line = t * 8 + g
                ad.append(min(L.BU(arr, line // D1, line % D1, 2 * c), L.BU(arr, line // D1, line % D1, 2 * c + 1))
                          if line < NL and 2 * c < RQ else None)
            add('A store BU/GU (128)', wf128(ad))
    # ---- A face: own loads (rows = (f,jb), k = ja)
    for t in range((NT1 + 7) // 8):
        for ks in range(KF):
            ad = []
            for g, c in LANES:
                line, k = t * 8 + g, ks * 4 + c
                if line < NT1 and k < D1:
                    f, jb = divmod(line, D1)
                    axis, side = face_geom(f)
                    idx = [0, 0, 0]
                    idx[axis] = side * (D1 - 1)
                    rem = [a for a in range(3) if a != axis]
                    idx[rem[0]] = k; idx[rem[1]] = jb
                    ad.append(L.U(idx[2], idx[1], idx[0]))
                else:
                    ad.append(None)
            add('A face load own', wf64(ad))
            add('A face load NB', wf64([(t * 8 + g) * D1 + ks * 4 + c if t * 8 + g < NT1 and ks * 4 + c < D1 else None
                                        for g, c in LANES]))
        add('A face store F1 (128)', wf128([L.F1((t * 8 + g) // D1, (t * 8 + g) % D1, 2 * c)
                                            if t * 8 + g < NT1 and 2 * c < RQ else None for g, c in LANES]))
    # ---- B: fwd-y, B operand data[k = iy][n = line]; D rows = qy, cols = lines 2c, 2c+1
    for t in range((NY + 7) // 8):
        for ks in range(KF):
            for arr in range(2):
                ad = []
                for g, c in LANES:
                    line, k = t * 8 + g, ks * 4 + c
                    if line < NY and k < D1:
                        z, qx = L.lineB(line)
                        ad.append(L.BU(arr, z, k, qx))
                    else:
                        ad.append(None)
                add('B load BU/GU', wf64(ad))
        for arr in range(3):
            if L.g3_pair_b():
                ad = []
                for g, c in LANES:
                    ls = t * 8 + 2 * c
                    if g < Q and ls < NY:
                        z, qx = L.lineB(ls)
                        z1, qx1 = L.lineB(ls + 1)
                        a0, a1 = L.G3(arr, z, g, qx), L.G3(arr, z1, g, qx1)
                        assert abs(a0 - a1) == 1 and min(a0, a1) % 2 == 0
                        ad.append(min(a0, a1))
                    else:
                        ad.append(None)
                add('B store G3 (128)', wf128(ad))
            else:
                for h in range(2):
                    ad = []
                    for g, c in LANES:
                        ls = t * 8 + 2 * c + h
                        if g < Q and ls < NY:
                            z, qx = L.lineB(ls)
                            ad.append(L.G3(arr, z, g, qx))
                        else:
                            ad.append(None)
                    add('B store G3', wf64(ad))
    # ---- B face fused: loads F1[(f, kk), qa], stores same positions (ib = 2c, 2c+1)
    for t in range((NT2 + 7) // 8):
        for ks in range(KF):
            ad = []
            for g, c in LANES:
                line, kk = t * 8 + g, ks * 4 + c
                ad.append(L.F1(line // Q, kk, line % Q) if line < NT2 and kk < D1 else None)
            add('B face load F1', wf64(ad))
        for h in range(2):
            ad = []
            for g, c in LANES:
                line = t * 8 + g
                ad.append(L.F1(line // Q, 2 * c + h, line % Q) if line < NT2 and 2 * c + h < D1 else None)
            add('B face store F1', wf64(ad))
    # ---- C: z-stage, A operand rows = col (qy,qx), k = iz; stores (col, iz = 2c, 2c+1)
    for t in range((NC + 7) // 8):
        for ks in range(KF):
            for arr in range(3):
                ad = []
                for g, c in LANES:
                    col, kk = t * 8 + g, ks * 4 + c
                    ad.append(L.G3(arr, kk, col // Q, col % Q) if col < NC and kk < D1 else None)
                add('C load G3', wf64(ad))
        if L.g3_pair_c():
            ad = []
            for g, c in LANES:
                col = t * 8 + g
                ad.append(min(L.G3(0, 2 * c, col // Q, col % Q), L.G3(0, 2 * c + 1, col // Q, col % Q))
                          if col < NC and 2 * c + 1 < D1 else None)
            add('C store T4 (128)', wf128(ad))
        else:
            for h in range(2):
                ad = []
                for g, c in LANES:
                    col = t * 8 + g
                    ad.append(L.G3(0, 2 * c + h, col // Q, col % Q) if col < NC and 2 * c + h < D1 else None)
                add('C store T4', wf64(ad))
    # ---- C face back-a: loads F1[line, kk], stores FD[line*D1 + 2c(+1)]
    for t in range((NT1 + 7) // 8):
        for ks in range(KB):
            add('C face load F1', wf64([L.F1((t * 8 + g) // D1, (t * 8 + g) % D1, ks * 4 + c)
                                        if t * 8 + g < NT1 and ks * 4 + c < Q else None for g, c in LANES]))
        for h in range(2):
            add('C face store FD', wf64([(t * 8 + g) * D1 + 2 * c + h if t * 8 + g < NT1 and 2 * c + h < D1 else None
                                         for g, c in LANES]))
    # ---- D: bwd-y, B operand data[k = qy][n = line (iz,qx)]; D rows = iy, cols = lines 2c, 2c+1
    for t in range((NY + 7) // 8):
        for ks in range(KB):
            ad = []
            for g, c in LANES:
                line, kk = t * 8 + g, ks * 4 + c
                if line < NY and kk < Q:
                    iz, qx = L.lineB(line)
                    ad.append(L.G3(0, iz, kk, qx))
                else:
                    ad.append(None)
            add('D load T4', wf64(ad))
        for h in range(2):
            ad = []
            for g, c in LANES:
                ls = t * 8 + 2 * c + h
                if g < D1 and ls < NY:
                    iz, qx = L.lineB(ls)
                    ad.append(L.S2(iz, g, qx))
                else:
                    ad.append(None)
            add('D store S2', wf64(ad))
    # ---- E: bwd-x, A operand rows = lines (iz,iy), k = qx
    for t in range((NL + 7) // 8):
        for ks in range(KB):
            ad = []
            for g, c in LANES:
                line, kk = t * 8 + g, ks * 4 + c
                ad.append(L.S2(line // D1, line % D1, kk) if line < NL and kk < Q else None)
            add('E load S2', wf64(ad))
    # ---- tail: u[k] = U[j], j = lane + 32 k
    for k in range((D1 ** 3 + 31) // 32):
        ad = []
        for l in range(32):
            j = l + 32 * k
            ad.append(L.U(j // (D1 * D1), (j // D1) % D1, j % D1) if j < D1 ** 3 else None)
        add('tail load U', wf64(ad))
    return res


def check_injective(L):
    for name, gen in (('U', [(z, y, x) for z in range(D1) for y in range(D1) for x in range(D1)]),
                      ('BU', [(a, z, y, q) for a in range(2) for z in range(D1) for y in range(D1) for q in range(Q)]),
                      ('G3', [(a, z, qy, qx) for a in range(3) for z in range(D1) for qy in range(Q) for qx in range(Q)]),
                      ('S2', [(iz, iy, q) for iz in range(D1) for iy in range(D1) for q in range(Q)])):
        s = [getattr(L, name)(*i) for i in gen]
        assert len(set(s)) == len(s), name + ' layout is not injective'
        print('   %s: %d values in [0, %d)' % (name, len(s), max(s) + 1))


if __name__ == '__main__':
    for L in (Old(), New()):
        print('layout', L.name)
        check_injective(L)
        r = run(L)
        for k, v in r.items():
            print('  %-26s %4d' % (k, v))
        print('  total %d' % sum(r.values()))


def search_f1():
    import itertools
    best = None
    for o1 in range(6, 10):
        for o2 in range(o1 + 6, o1 + 10):
            for o3 in range(o2 + 6, o2 + 12):
                for FS in range(o3 + 6, o3 + 14, 2):
                    if o1 % 2 or o2 % 2 or o3 % 2:
                        continue
                    off = [0, o1, o2, o3]

                    class T(New):
                        def F1(s, f, j, q, off=off, FS=FS): return FS * f + off[j] + q
                    r = run(T())
                    tot = sum(v for k, v in r.items() if 'F1' in k)
                    if best is None or tot < best[0]:
                        best = (tot, off, FS, {k: v for k, v in r.items() if 'F1' in k})
    print('best F1', best)


if __name__ == '__main__' and len(sys.argv) > 3:
    search_f1()
