"""Generates the small input meshes used by the tests (MFEM mesh v1.0 text format).

The GPU box has no /root/reference, so the meshes the reference's known answers are quoted on
are re-generated here from their definitions (Cartesian 3x3 / 3x3x3 periodic grids on [-1,1]^d
with coordinates rounded to the digits the reference files carry, the 2x2x2 unit cube, and the
4x4 INLINE unit square).  tests/test_oracle_golden.py checks, when /root/reference is present,
that they are numerically identical to the reference's data/*.mesh.

Run:  python tests/data/make_meshes.py
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..', 'oracle'))
from remhos_oracle import mesh as om  # noqa: E402

LEX2MFEM = {2: [0, 1, 3, 2], 3: [0, 1, 3, 2, 4, 5, 7, 6]}


def write_v1(path, m, periodic, digits):
    dim = m.dim
    geom = 3 if dim == 2 else 5
    with open(path, 'w') as f:
        f.write('MFEM mesh v1.0\n\ndimension\n%d\n\nelements\n%d\n' % (dim, m.ne))
        for e in range(m.ne):
            v = m.ev[e][LEX2MFEM[dim]]
            f.write('1 %d %s\n' % (geom, ' '.join(str(int(x)) for x in v)))
        f.write('\nboundary\n0\n\nvertices\n%d\n' % m.nv)
        fmt = '%.' + str(digits) + 'g'
        if periodic:
            f.write('\nnodes\nFiniteElementSpace\nFiniteElementCollection: L2_T1_%dD_P1\n'
                    'VDim: %d\nOrdering: 1\n\n' % (dim, dim))
            for e in range(m.ne):
                for n in range(2 ** dim):
                    f.write(' '.join(fmt % (round(x, digits) + 0.0) for x in m.X[e, n]) + '\n')
                f.write('\n')
        else:
            coords = np.zeros((m.nv, dim))
            coords[m.ev.reshape(-1)] = m.X.reshape(-1, dim)
            f.write('%d\n' % dim)
            for v in range(m.nv):
                f.write(' '.join(fmt % x for x in coords[v]) + '\n')


def main():
    sq = om.cartesian_mesh([3, 3], [2.0, 2.0], origin=[-1.0, -1.0], periodic=True)
    sq.X = np.round(sq.X, 9)
    write_v1(os.path.join(HERE, 'periodic-square.mesh'), sq, True, 9)
    cu = om.cartesian_mesh([3, 3, 3], [2.0, 2.0, 2.0], origin=[-1.0, -1.0, -1.0], periodic=True)
    cu.X = np.round(cu.X, 6)
    write_v1(os.path.join(HERE, 'periodic-cube.mesh'), cu, True, 6)
    hx = om.cartesian_mesh([2, 2, 2], [1.0, 1.0, 1.0])
    write_v1(os.path.join(HERE, 'cube01_hex.mesh'), hx, False, 17)
    with open(os.path.join(HERE, 'inline-quad.mesh'), 'w') as f:
        f.write('MFEM INLINE mesh v1.0\n\ntype = quad\nnx = 4\nny = 4\nsx = 1.0\nsy = 1.0\n')


if __name__ == '__main__':
    main()
