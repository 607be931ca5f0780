#!/usr/bin/env python
"""Shared-memory wavefront model for the access patterns of k_stage3p (64-bit accesses: a warp is
served in half-warps of 16 lanes, one 128-byte wavefront per half-warp when the 16 addresses fall
in distinct 8-byte bank pairs).  Prints wavefronts per phase against the conflict-free count, for
a candidate set of strides.  usage: python tools/bank_sim.py [D1 Q E]"""
import sys
from collections import defaultdict


def wavefronts(addr_fn, ntasks, T):
    """addr_fn(task) -> list of double-indices accessed by successive instructions of that task."""
    total = ideal = 0
    for base in range(0, ntasks, T):
        n = min(T, ntasks - base)
        for w0 in range(0, n, 32):
            lanes = list(range(w0, min(w0 + 32, n)))
            seqs = [addr_fn(base + t) for t in lanes]
            for k in range(len(seqs[0])):
                for h0 in range(0, len(lanes), 16):
                    banks = defaultdict(set)
                    for s in seqs[h0:h0 + 16]:
                        banks[s[k] % 16].add(s[k])
                    total += max(len(v) for v in banks.values())
                    ideal += 1
    return total, ideal


def main():
    D1, Q, E = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (4, 6, 8)
    ND, QQ, NF, NFD = D1 ** 3, Q * Q, 6, D1 * D1
    TPE = max(QQ, 32)
    T = ((TPE * E + 31) // 32) * 32
    NL, NY, NT1, NT2 = E * D1 * D1, E * D1 * Q, E * NF * D1, E * NF * Q
    import itertools
    # strides (doubles): RU row of U, RC row of BU/GU/S2 (length Q), ZB plane of GB (length QQ),
    # EB element of GB (3*D1 planes), RF row of F1 (length Q)
    def report(RC, ZB, EB, RF, RX, verbose):
        res = {}
        offGU = NL * RC
        def A_vol_r(id):
            arr, l = divmod(id, NL)
            return [l * D1 + i for i in range(D1)]
        def A_vol_w(id):
            arr, l = divmod(id, NL)
            return [arr * offGU + l * RC + q for q in range(Q)]
        res['A vol read U'] = wavefronts(A_vol_r, 2 * NL, T)
        res['A vol write BU/GU'] = wavefronts(A_vol_w, 2 * NL, T)
        def A_face_w(l):
            return [l * RF + q for q in range(Q)]
        def A_face_nb(l):
            return [l * D1 + i for i in range(D1)]
        def A_face_own(l):
            ef, jb = divmod(l, D1)
            e, f = divmod(ef, NF)
            axis = 2 if f in (0, 5) else (1 if f in (1, 3) else 0)
            side = 1 if f in (2, 3, 5) else 0
            s1 = D1 if axis == 0 else 1
            s2 = D1 if axis == 2 else D1 * D1
            sa = (1, D1, D1 * D1)[axis]
            return [e * ND + side * (D1 - 1) * sa + jb * s2 + i * s1 for i in range(D1)]
        res['A face read NB'] = wavefronts(A_face_nb, NT1, T)
        res['A face read own'] = wavefronts(A_face_own, NT1, T)
        res['A face write F1'] = wavefronts(A_face_w, NT1, T)
        def B_r(id):
            arr, l = divmod(id, NY)
            ez, qx = divmod(l, Q)
            return [(offGU if arr == 0 else 0) + (ez * D1 + i) * RC + qx for i in range(D1)]
        def B_w(id):
            arr, l = divmod(id, NY)
            ez, qx = divmod(l, Q)
            e, z = divmod(ez, D1)
            return [e * EB + (arr * D1 + z) * ZB + q * Q + qx for q in range(Q)]
        res['B fwd-y read'] = wavefronts(B_r, 3 * NY, T)
        res['B fwd-y write GB'] = wavefronts(B_w, 3 * NY, T)
        def F2(id):
            ef, qa = divmod(id, Q)
            return [(ef * D1 + i) * RF + qa for i in range(D1)]
        res['B face read/write F1'] = wavefronts(F2, NT2, T)
        def C_r(id):
            ze, zr = divmod(id, QQ)
            return [ze * EB + k * ZB + zr for k in range(3 * D1)]
        def C_w(id):
            ze, zr = divmod(id, QQ)
            return [ze * EB + k * ZB + zr for k in range(D1)]
        res['C z read GB'] = wavefronts(C_r, E * QQ, T)
        res['C z write'] = wavefronts(C_w, E * QQ, T)
        res['C face read F1'] = wavefronts(lambda id: [id * RF + q for q in range(Q)], NT1, T)
        res['C face write FD'] = wavefronts(lambda id: [id * D1 + i for i in range(D1)], NT1, T)
        def D_r(id):
            eiz, qx = divmod(id, Q)
            e, iz = divmod(eiz, D1)
            return [e * EB + iz * ZB + q * Q + qx for q in range(Q)]
        def D_w(id):
            eiz, qx = divmod(id, Q)
            return [(eiz * D1 + i) * RC + qx for i in range(D1)]
        res['D read GB'] = wavefronts(D_r, NY, T)
        res['D write S2'] = wavefronts(D_w, NY, T)
        res['E read S2'] = wavefronts(lambda id: [id * RC + q for q in range(Q)], NL, T)
        res['E write X'] = wavefronts(lambda id: [id * RX + i for i in range(D1)], NL, T)
        def E_fc(id):
            e, r = divmod(id, D1 * D1)
            b, a = divmod(r, D1)
            base = e * NF * NFD
            out = [base + 4 * NFD + b * D1 + a, base + 2 * NFD + b * D1 + a]
            for i in range(D1):
                out += [base + 1 * NFD + b * D1 + i, base + 3 * NFD + b * D1 + i,
                        base + 0 * NFD + a * D1 + i, base + 5 * NFD + a * D1 + i]
            return out
        res['E read FD'] = wavefronts(E_fc, NL, T)
        tot = sum(v[0] for v in res.values()); idl = sum(v[1] for v in res.values())
        if verbose:
            for k, (t, i) in res.items():
                print('  %-24s %6d wavefronts (ideal %6d)  x%.2f' % (k, t, i, t / i))
            print('  total %d vs ideal %d  (x%.2f)' % (tot, idl, tot / idl))
        return tot, idl
    print('baseline strides RC=%d ZB=%d EB=%d RF=%d RX=%d' % (Q, QQ, 3 * D1 * QQ, Q, D1))
    report(Q, QQ, 3 * D1 * QQ, Q, D1, True)
    best = None
    for RC in range(Q, Q + 3):
        for ZB in range(QQ, QQ + 9):
            for EBp in range(0, 17):
                EB = 3 * D1 * ZB + EBp
                for RF in range(Q, Q + 3):
                    t, i = report(RC, ZB, EB, RF, D1, False)
                    if best is None or t < best[0]:
                        best = (t, RC, ZB, EB, RF)
    print('best', best)
    report(best[1], best[2], best[3], best[4], D1, True)


if __name__ == '__main__':
    main()
