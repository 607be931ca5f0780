"""Host mesh module (C++) against the oracle's mesh code: bit-exact integer maps
(BdrDofs, NbrDof, Sub2Ind, neighbour elements, overlap-bounds entity partition) and nodes."""
import os

import numpy as np
import pytest

from helpers import DATA
import remhos_b200 as rb
from remhos_oracle import mesh as om, dg

CASES = [('periodic-square.mesh', 2, 3), ('periodic-square.mesh', 1, 1), ('inline-quad.mesh', 2, 2),
         ('inline-quad.mesh', 1, 4), ('cube01_hex.mesh', 1, 2), ('cube01_hex.mesh', 2, 3),
         ('periodic-cube.mesh', 1, 3), ('periodic-cube.mesh', 0, 4), ('periodic-cube.mesh', 1, 1),
         ('periodic-hexagon.mesh', 2, 3), ('periodic-hexagon.mesh', 1, 2)]


@pytest.mark.parametrize('mesh,rs,p', CASES)
def test_index_maps_bit_exact(mesh, rs, p):
    m = om.read_mesh(os.path.join(DATA, mesh))
    for _ in range(rs):
        m = om.refine_uniform(m)
    m = om.set_curvature(m, 2)
    topo = om.Topology(m)
    pm = rb.Mesh.load(os.path.join(DATA, mesh)).refine(rs).set_curvature(2)
    maps = pm.dof_maps(p)
    assert pm.ne == m.ne and pm.dim == m.dim
    assert np.abs(pm.nodes() - m.X).max() < 1e-15
    assert np.array_equal(maps['bdr_dofs'], dg.bdr_dofs(p, m.dim))
    assert np.array_equal(maps['nbr_dof'], dg.nbr_dof_map(topo, p))
    assert np.array_equal(maps['nbr_elem'], topo.nbr_elem)
    if p > 1:
        assert np.array_equal(maps['sub2ind'], dg.sub2ind(p, m.dim))
    # same partition of (element, lattice position) pairs into shared entities
    a = maps['lat'].reshape(-1); b = topo.lat.reshape(-1)
    pairs = np.unique(np.stack([a, b], 1), axis=0)
    assert len(pairs) == len(np.unique(a)) == len(np.unique(b))
    assert maps['n_ent'] >= a.max() + 1


def test_nbr_dof_is_an_involution_and_matches_geometry():
    """NbrDof(NbrDof) = identity on interior faces, and matched DOFs coincide physically."""
    pm = rb.Mesh.load(os.path.join(DATA, 'cube01_hex.mesh')).refine(1).set_curvature(2)
    p = 3
    maps = pm.dof_maps(p)
    from remhos_b200.setup_problem import mesh_eval
    x = mesh_eval(pm, np.arange(p + 1) / p).reshape(-1, 3)
    bd, nb = maps['bdr_dofs'], maps['nbr_dof']
    nd = maps['nd']
    ne, nf, nfd = nb.shape
    own = (np.arange(ne)[:, None, None] * nd + bd.T[None, :, :])
    inner = nb >= 0
    assert np.abs(x[own[inner]] - x[nb[inner]]).max() < 1e-14
    lut = -np.ones(ne * nd * nf, dtype=np.int64)
    # map (global dof, face) -> neighbour dof, then apply twice
    back = {}
    for e in range(ne):
        for f in range(nf):
            for j in range(nfd):
                if nb[e, f, j] >= 0:
                    back[(own[e, f, j], nb[e, f, j] // nd)] = nb[e, f, j]
    for (g, ne2), g2 in back.items():
        assert back[(g2, g // nd)] == g


def test_cartesian_generator_and_refinement_counts():
    m = rb.Mesh.cartesian([3, 3, 3], [2.0] * 3, origin=[-1.0] * 3, periodic=True).refine(2)
    assert m.ne == 27 * 64
    lo, hi = m.bounding_box()
    assert np.allclose(lo, -1) and np.allclose(hi, 1)
    maps = m.dof_maps(2)
    assert (maps['nbr_dof'] >= 0).all()          # periodic: no domain boundary
    m2 = rb.Mesh.cartesian([4, 2], [1.0, 1.0])
    maps2 = m2.dof_maps(1)
    assert (maps2['nbr_elem'] < 0).sum() == 2 * (4 + 2)


def test_problem_functions_match_oracle():
    from remhos_oracle import problems
    from remhos_b200.setup_problem import velocity, u0
    rng = np.random.default_rng(3)
    for dim in (2, 3):
        x = rng.uniform(-1, 1, size=(500, dim))
        lo, hi = -np.ones(dim), np.ones(dim)
        for prob in (0, 1, 3, 4, 5, 6, 7, 10, 11, 14):
            if prob == 11 and dim == 3:
                continue
            assert np.allclose(velocity(prob, x, lo, hi), problems.velocity(prob, x, lo, hi),
                               rtol=0, atol=1e-15)
        for prob in (0, 1, 2, 3, 4, 5, 6, 7):
            if prob in (2, 3, 4) and dim == 3:
                continue
            assert np.allclose(u0(prob, x, lo, hi), problems.u0(prob, x, lo, hi), rtol=0, atol=2e-15)


def _brute_structured(lat, nbr, dim):
    ne, n3 = lat.shape
    ent = {}
    for e in range(ne):
        for t in range(n3):
            ent.setdefault(int(lat[e, t]), set()).add(e)
    for e in range(ne):
        for t in range(n3):
            c = [(t // 3 ** a) % 3 for a in range(dim)]
            T = set()
            for d in range(n3):
                da = [(d // 3 ** a) % 3 - 1 for a in range(dim)]
                if all((c[a] == 0 and da[a] <= 0) or (c[a] == 1 and da[a] == 0) or (c[a] == 2 and da[a] >= 0)
                       for a in range(dim)) and nbr[e, d] >= 0:
                    T.add(int(nbr[e, d]))
            if T != ent[int(lat[e, t])]:
                return False
    return True


@pytest.mark.parametrize('name,structured', [('cart3p', True), ('cart3', True), ('cart2p', True),
                                             ('hexagon', False), ('cube01', True)])
def test_nbr_lattice(name, structured):
    """rmh_nbr_lattice: the 3^dim neighbourhood of every element derived from the lattice-entity map;
    `structured` iff it reproduces every entity's element set (so overlap bounds can be formed from
    neighbour values alone); face directions agree with the face-neighbour map"""
    import remhos_b200 as rb
    m = {'cart3p': lambda: rb.Mesh.cartesian([3, 3, 3], [2.] * 3, origin=[-1.] * 3, periodic=True).refine(1),
         'cart3': lambda: rb.Mesh.cartesian([3, 2, 4], [1.] * 3, periodic=False),
         'cart2p': lambda: rb.Mesh.cartesian([4, 5], [1.] * 2, periodic=True),
         'hexagon': lambda: rb.Mesh.load(os.path.join(DATA, 'periodic-hexagon.mesh')).refine(1),
         'cube01': lambda: rb.Mesh.load(os.path.join(DATA, 'cube01_hex.mesh')).refine(1)}[name]()
    nbr, ok = m.nbr_lattice()
    maps = m.dof_maps(1)
    dim, n3 = m.dim, 3 ** m.dim
    assert ok == structured
    assert ok == _brute_structured(maps['lat'], nbr, dim)
    assert (nbr[:, n3 // 2] == np.arange(nbr.shape[0])).all()
    # faces: quad S E N W, hex bottom south east north west top -> (axis, side)
    faces = {2: [(1, 0), (0, 1), (1, 1), (0, 0)], 3: [(2, 0), (1, 0), (0, 1), (1, 1), (0, 0), (2, 1)]}[dim]
    for f, (axis, side) in enumerate(faces):
        d = n3 // 2 + (1 if side else -1) * 3 ** axis
        assert np.array_equal(nbr[:, d], maps['nbr_elem'][:, f]), f


@pytest.mark.parametrize('name,rs,g', [('periodic-cube.mesh', 1, 2), ('inline-quad.mesh', 1, 2), ('cube01_hex.mesh', 1, 3),
                                       ('periodic-hexagon.mesh', 0, 2), ('periodic-square.mesh', 1, 1)])
def test_mesh_save_round_trip(name, rs, g, tmp_path):
    """rmh_mesh_save (Mesh::Print in "MFEM mesh v1.0", nodes as L2_T1 Gauss-Lobatto field; -save / -visit,
    remhos.cpp:1016-1043): the file reads back to the same nodes, vertices and DofInfo maps through the
    product's reader and through the oracle's, and lists exactly the faces without a neighbour as boundary"""
    import ctypes as C
    from remhos_b200.capi import lib, check
    from remhos_oracle import mesh as om
    m = rb.Mesh.load(os.path.join(DATA, name)).refine(rs)
    m.set_curvature(g)
    p = str(tmp_path / 'out.mesh')
    check(lib().rmh_mesh_save(m.h, p.encode(), None, 17))
    m2 = rb.Mesh.load(p)
    assert m2.ne == m.ne and m2.geom_order == g
    assert np.array_equal(m.nodes(), m2.nodes()) and np.array_equal(m.elem_vertices(), m2.elem_vertices())
    a, b = m.dof_maps(2), m2.dof_maps(2)
    for k in ('nbr_dof', 'lat', 'nbr_elem', 'bdr_dofs'):
        assert np.array_equal(a[k], b[k]), k
    mo = om.read_mesh(p)
    assert mo.ne == m.ne and np.array_equal(mo.X, m.nodes())
    txt = open(p).read()
    nb = int(txt.split('boundary\n')[1].split('\n')[0])
    assert nb == int((a['nbr_elem'] < 0).sum())
    # moved nodes (remap) and the GridFunction writer
    x = m.nodes() + 0.01
    check(lib().rmh_mesh_save(m.h, p.encode(), x.ctypes.data_as(C.c_void_p), 17))
    assert np.array_equal(rb.Mesh.load(p).nodes(), x)
    vals = np.linspace(0.0, 1.0, 37)
    g_path = str(tmp_path / 'u.gf')
    check(lib().rmh_gf_save(g_path.encode(), m.dim, 3, 2, C.c_int64(vals.size), vals.ctypes.data_as(C.c_void_p), 17))
    lines = open(g_path).read().split('\n')
    assert lines[0] == 'FiniteElementSpace' and lines[1] == 'FiniteElementCollection: L2_T2_%dD_P3' % m.dim
    assert lines[2] == 'VDim: 1' and lines[3] == 'Ordering: 0'
    assert np.array_equal(np.array([float(v) for v in lines[5:5 + vals.size]]), vals)


@pytest.mark.parametrize('name,rs,g,periodic', [('periodic-square.mesh', 1, 2, True), ('periodic-hexagon.mesh', 0, 2, True),
                                                ('periodic-cube.mesh', 0, 2, True), ('inline-quad.mesh', 1, 2, False),
                                                ('cube01_hex.mesh', 1, 3, False)])
@pytest.mark.parametrize('factor', [1, 2, 3])
def test_make_refined_is_the_subcell_mesh(name, rs, g, periodic, factor, tmp_path):
    """rmh_mesh_make_refined = Mesh::MakeRefined(mesh, order, ClosedUniform) (remhos.cpp:801), the mesh -save
    writes as meshLO_*.mesh: factor^dim linear sub-elements per element on the uniform lattice, shared lattice
    points are shared vertices (periodic identification kept), positions = the geometry evaluated at i / factor,
    and the result is a valid mesh for every other function of the module (topology, writer, reader)."""
    from remhos_b200.capi import lib, check
    m = rb.Mesh.load(os.path.join(DATA, name)).refine(rs)
    m.set_curvature(g)
    dim, ne = m.dim, m.ne
    r = m.make_refined(factor)
    assert r.dim == dim and r.ne == ne * factor ** dim and r.geom_order == 1
    # vertex count: on a torus (all-quad / all-hex periodic mesh) #vertices = #elements; on the box meshes
    # (n sub-elements per direction) (n + 1)^dim
    if periodic:
        assert r.nv == r.ne
    else:
        n1 = round(ne ** (1.0 / dim)) * factor
        assert r.nv == (n1 + 1) ** dim
    # positions: corner k of sub-element (a, b, c) is the lattice point (a + k_x, b + k_y, c + k_z) / factor
    pts = np.arange(factor + 1) / factor
    from remhos_b200.setup_problem import mesh_eval
    xl = mesh_eval(m, pts).reshape(ne, *([factor + 1] * dim)[::-1], dim)       # [e][z][y][x][dim]
    X = r.nodes().reshape(ne, factor ** dim, 2 ** dim, dim)
    for sc in range(factor ** dim):
        c = [(sc // factor ** a) % factor for a in range(dim)]
        for k in range(2 ** dim):
            idx = tuple(c[a] + ((k >> a) & 1) for a in range(dim))[::-1]
            assert np.abs(X[:, sc, k] - xl[(slice(None),) + idx]).max() < 1e-14
    # the same vertex id <=> the same point (up to the periodic shift); every vertex id is used
    ev = r.elem_vertices()
    assert np.array_equal(np.unique(ev), np.arange(r.nv))
    if not periodic:
        pos = np.full((r.nv, dim), np.nan)
        flat_v, flat_x = ev.reshape(-1), r.nodes().reshape(-1, dim)
        pos[flat_v] = flat_x
        assert np.abs(pos[flat_v] - flat_x).max() < 1e-13
        assert np.unique(np.round(pos, 9), axis=0).shape[0] == r.nv
    # a valid mesh: neighbour maps, boundary faces, round trip through the writer and the reader
    maps = r.dof_maps(1)
    nb = int((maps['nbr_elem'] < 0).sum())
    nb0 = int((m.dof_maps(1)['nbr_elem'] < 0).sum())
    assert nb == nb0 * factor ** (dim - 1)
    p = str(tmp_path / 'lo.mesh')
    check(lib().rmh_mesh_save(r.h, p.encode(), None, 17))
    r2 = rb.Mesh.load(p)
    assert r2.ne == r.ne and np.array_equal(r2.nodes(), r.nodes()) and np.array_equal(r2.elem_vertices(), ev)
    # moved nodes (remap): linear in the nodes
    x = m.nodes() * 1.5 + 0.25
    rm = m.make_refined(factor, x)
    assert np.abs(rm.nodes() - (r.nodes() * 1.5 + 0.25)).max() < 1e-13
    # total volume is kept when the geometry is (multi)linear
    if g == 1 or name in ('periodic-square.mesh', 'periodic-cube.mesh', 'inline-quad.mesh', 'cube01_hex.mesh'):
        assert abs((r.elem_sizes() ** dim).sum() - (m.elem_sizes() ** dim).sum()) < 1e-12 * ne
