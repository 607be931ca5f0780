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


# periodic hexagon: three rhombi (120-degree rotations of one another), 2 x 2 quadrilaterals each,
# opposite sides of the hexagon identified.  Element -> vertex table (MFEM counter-clockwise
# order) of the 12-vertex periodic topology.
HEXAGON_EV = [[0, 2, 8, 5], [2, 1, 3, 8], [5, 8, 6, 11], [8, 3, 0, 6],
              [0, 4, 9, 6], [4, 1, 2, 9], [6, 9, 7, 11], [9, 2, 0, 7],
              [0, 3, 10, 7], [3, 1, 4, 10], [7, 10, 5, 11], [10, 4, 0, 5]]


def write_hexagon(path):
    with open(path, 'w') as f:
        f.write('MFEM mesh v1.0\n\ndimension\n2\n\nelements\n12\n')
        for k, ev in enumerate(HEXAGON_EV):
            f.write('%d 3 %s\n' % (k // 4 + 1, ' '.join(str(v) for v in ev)))
        f.write('\nboundary\n0\n\nvertices\n12\n\nnodes\nFiniteElementSpace\n'
                'FiniteElementCollection: L2_T1_2D_P1\nVDim: 2\nOrdering: 1\n\n')
        h = 0.8660254037844386          # sqrt(3)/2 as the reference file carries it
        # (origin, s-direction, t-direction) of the three rhombi: 120-degree rotations
        rh = [((-0.5, -h), (1.0, 0.0), (0.5, h)), ((1.0, 0.0), (-0.5, h), (-1.0, 0.0)),
              ((-0.5, h), (-0.5, -h), (0.5, -h))]
        for org, ds, dt in rh:
            org, ds, dt = np.array(org), np.array(ds), np.array(dt)
            for j in range(2):
                for i in range(2):
                    for b in range(2):
                        for aa in range(2):
                            x = org + 0.5 * (i + aa) * ds + 0.5 * (j + b) * dt + 0.0
                            f.write('%.16g %.16g\n' % (x[0], x[1]))
                    f.write('\n')


def main():
    sq = om.cartesian_mesh([3, 3], [2.0, 2.0], origin=[-1.0, -1.0], periodic=True)
    sq.X = np.round(sq.X, 9)
    write_v1(os.path.join(HERE, 'periodic-square.mesh'), sq, True, 9)
    cu = om.cartesian_mesh([3, 3, 3], [2.0, 2.0, 2.0], origin=[-1.0, -1.0, -1.0], periodic=True)
    cu.X = np.round(cu.X, 6)
    write_v1(os.path.join(HERE, 'periodic-cube.mesh'), cu, True, 6)
    hx = om.cartesian_mesh([2, 2, 2], [1.0, 1.0, 1.0])
    write_v1(os.path.join(HERE, 'cube01_hex.mesh'), hx, False, 17)
    write_hexagon(os.path.join(HERE, 'periodic-hexagon.mesh'))
    with open(os.path.join(HERE, 'inline-quad.mesh'), 'w') as f:
        f.write('MFEM INLINE mesh v1.0\n\ntype = quad\nnx = 4\nny = 4\nsx = 1.0\nsy = 1.0\n')


if __name__ == '__main__':
    main()
