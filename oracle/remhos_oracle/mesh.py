"""TEST INFRASTRUCTURE ONLY -- CPU oracle (numpy) for the Remhos RK-stage hot path.

Mesh reader (MFEM mesh v1.0 / MFEM INLINE mesh v1.0 text formats, quads and hexes),
uniform refinement, element-wise nodal geometry, face-neighbour topology and the
H1-style "macro lattice" entity numbering used for overlap bounds.

Restates the MFEM mesh semantics Remhos relies on (remhos.cpp:448-463 load + refine,
:510-513 SetCurvature, :457 GetBoundingBox); MFEM itself is not under /root/reference,
see SURVEY.md Appendix C.  Conventions:

  * Element vertices are stored in LEXICOGRAPHIC corner order (x fastest), converted from
    MFEM's counter-clockwise order on input.
  * Geometry is an element-wise (L2-style) nodal field X[e, node, comp] on the tensor
    Gauss-Lobatto lattice of degree `gorder`; this represents both H1 and periodic (L2) nodes.
  * Local faces follow MFEM: quad edges 0 S(y=0) 1 E(x=1) 2 N(y=1) 3 W(x=0)
    (remhos_tools.cpp:1367-1376); hex faces 0 bottom(z=0) 1 south(y=0) 2 east(x=1)
    3 north(y=1) 4 west(x=0) 5 top(z=1) (remhos_tools.cpp:1086,1126,1166,1206,1246,1286).
"""
import numpy as np
from . import fe

# MFEM corner order -> lexicographic corner order
_MFEM2LEX = {2: np.array([0, 1, 3, 2]), 3: np.array([0, 1, 3, 2, 4, 5, 7, 6])}

# local face -> (fixed axis, side) ; the remaining axes (ascending) parametrise the face
FACE_AXIS = {
    2: [(1, 0), (0, 1), (1, 1), (0, 0)],
    3: [(2, 0), (1, 0), (0, 1), (1, 1), (0, 0), (2, 1)],
}


def face_corner_lex(dim):
    """[nf, 2^(dim-1)] lexicographic corner indices of each local face, ordered in the
    face's natural (ascending remaining axes) parametrisation."""
    out = []
    for axis, side in FACE_AXIS[dim]:
        rem = [a for a in range(dim) if a != axis]
        corners = []
        for t in range(2 ** (dim - 1)):
            c = [0] * dim
            c[axis] = side
            for m, a in enumerate(rem):
                c[a] = (t >> m) & 1
            corners.append(sum(c[a] << a for a in range(dim)))
        out.append(corners)
    return np.array(out)


class Mesh:
    def __init__(self, dim, ev, X, gorder, nv):
        self.dim = dim
        self.ev = np.ascontiguousarray(ev, dtype=np.int64)      # [NE, 2^d] lexicographic
        self.X = np.ascontiguousarray(X, dtype=np.float64)      # [NE, (g+1)^d, dim]
        self.gorder = gorder
        self.nv = nv

    @property
    def ne(self):
        return self.ev.shape[0]


# --------------------------------------------------------------------------- reading
def _tokens(path):
    toks = []
    with open(path) as f:
        for line in f:
            line = line.split('#')[0]
            toks.extend(line.split())
    return toks


def read_mesh(path):
    with open(path) as f:
        first = f.readline().strip()
    if first.startswith('MFEM INLINE mesh'):
        return _read_inline(path)
    if not first.startswith('MFEM mesh v1.0'):
        raise ValueError('unsupported mesh format: ' + first)
    with open(path) as f:
        f.readline()
        toks = []
        for line in f:
            line = line.split('#')[0]
            toks.extend(line.split())
    pos = 0

    def expect(word):
        nonlocal pos
        if toks[pos] != word:
            raise ValueError('expected %s, got %s' % (word, toks[pos]))
        pos += 1

    expect('dimension')
    dim = int(toks[pos]); pos += 1
    expect('elements')
    ne = int(toks[pos]); pos += 1
    nvert = 2 ** dim
    ev = np.empty((ne, nvert), dtype=np.int64)
    ev_file = np.empty((ne, nvert), dtype=np.int64)          # MFEM's own vertex order
    for e in range(ne):
        geom = int(toks[pos + 1])
        if (dim, geom) not in ((2, 3), (3, 5)):
            raise ValueError('only quadrilateral / hexahedral meshes are supported')
        v = np.array([int(t) for t in toks[pos + 2: pos + 2 + nvert]])
        ev_file[e] = v
        ev[e] = v[_MFEM2LEX[dim]]
        pos += 2 + nvert
    expect('boundary')
    nb = int(toks[pos]); pos += 1
    for _ in range(nb):
        geom = int(toks[pos + 1])
        nbv = {1: 2, 3: 4, 0: 1}[geom]
        pos += 2 + nbv
    expect('vertices')
    nv = int(toks[pos]); pos += 1
    if pos < len(toks) and toks[pos] == 'nodes':
        pos += 1
        expect('FiniteElementSpace')
        assert toks[pos] == 'FiniteElementCollection:'
        fec = toks[pos + 1]; pos += 2
        assert toks[pos] == 'VDim:'
        vdim = int(toks[pos + 1]); pos += 2
        assert toks[pos] == 'Ordering:'
        ordering = int(toks[pos + 1]); pos += 2
        data = np.array([float(t) for t in toks[pos:]])
        if fec.startswith('L2_T1_'):
            g = int(fec.split('_P')[1])
            npe = (g + 1) ** dim
            nd = ne * npe
            assert data.size == nd * vdim
            vals = data.reshape(nd, vdim) if ordering == 1 else data.reshape(vdim, nd).T
            X = vals.reshape(ne, npe, vdim)
            return Mesh(dim, ev, X, g, nv)
        if fec in ('Linear',) or (fec.startswith('H1_') and fec.endswith('_P1')):
            assert data.size == nv * vdim
            coords = data.reshape(nv, vdim) if ordering == 1 else data.reshape(vdim, nv).T
            return Mesh(dim, ev, coords[ev], 1, nv)
        if fec in ('Quadratic', 'H1_2D_P2') and dim == 2:
            return Mesh(dim, ev, _quadratic_nodes(ev_file, nv, data, vdim, ordering), 2, nv)
        if fec in ('Cubic', 'H1_2D_P3') and dim == 2:
            return Mesh(dim, ev, _cubic_nodes(ev_file, nv, data, vdim, ordering), 3, nv)
        raise ValueError('unsupported nodal collection ' + fec)
    sdim = int(toks[pos]); pos += 1
    coords = np.array([float(t) for t in toks[pos: pos + nv * sdim]]).reshape(nv, sdim)
    return Mesh(dim, ev, coords[ev], 1, nv)


def _quadratic_nodes(ev_file, nv, data, vdim, ordering):
    """Element-wise 3 x 3 nodes of an H1 order-2 nodal field on quadrilaterals (legacy `Quadratic`
    collection): global dofs are [vertices | edges | elements], edges numbered in the order of their
    first appearance over the elements and their local edges (0,1), (1,2), (2,3), (3,0) -- MFEM's
    vertex-to-vertex table [MFEM-K; validated by remhos_tests.cpp:88-91 through the star-q2 run]."""
    ne = ev_file.shape[0]
    edge_id = {}
    el_edges = np.empty((ne, 4), dtype=np.int64)
    for e in range(ne):
        v = ev_file[e]
        for j, (a, b) in enumerate(((0, 1), (1, 2), (2, 3), (3, 0))):
            key = (min(v[a], v[b]), max(v[a], v[b]))
            if key not in edge_id:
                edge_id[key] = len(edge_id)
            el_edges[e, j] = edge_id[key]
    nedge = len(edge_id)
    nd = nv + nedge + ne
    assert data.size == nd * vdim, (data.size, nd, vdim)
    vals = data.reshape(nd, vdim) if ordering == 1 else data.reshape(vdim, nd).T
    X = np.empty((ne, 9, vdim))
    v = ev_file
    X[:, 0] = vals[v[:, 0]]; X[:, 2] = vals[v[:, 1]]; X[:, 8] = vals[v[:, 2]]; X[:, 6] = vals[v[:, 3]]
    X[:, 1] = vals[nv + el_edges[:, 0]]; X[:, 5] = vals[nv + el_edges[:, 1]]
    X[:, 7] = vals[nv + el_edges[:, 2]]; X[:, 3] = vals[nv + el_edges[:, 3]]
    X[:, 4] = vals[nv + nedge + np.arange(ne)]
    return X


def _cubic_nodes(ev_file, nv, data, vdim, ordering):
    """Element-wise 4 x 4 Gauss-Lobatto nodes of an H1 order-3 nodal field on quadrilaterals given in
    the legacy `Cubic` collection (equispaced nodes): global dofs are [vertices | 2 per edge | 4 per
    element]; the two edge dofs run from the edge's lower-numbered vertex to the higher one, the
    element dofs are (1/3,1/3), (2/3,1/3), (1/3,2/3), (2/3,2/3) [MFEM-K; checked geometrically on
    data/star-q3.mesh: edge nodes ordered along their edges, positive Jacobians, interior nodes
    closest to the transfinite interpolant, area equal to star-q2's to 1e-6]."""
    from . import fe
    ne = ev_file.shape[0]
    edge_id = {}
    el_edges = np.empty((ne, 4), dtype=np.int64)
    el_fwd = np.empty((ne, 4), dtype=bool)
    for e in range(ne):
        v = ev_file[e]
        for j, (a, b) in enumerate(((0, 1), (1, 2), (2, 3), (3, 0))):
            key = (min(v[a], v[b]), max(v[a], v[b]))
            if key not in edge_id:
                edge_id[key] = len(edge_id)
            el_edges[e, j] = edge_id[key]
            el_fwd[e, j] = v[a] < v[b]
    nedge = len(edge_id)
    nd = nv + 2 * nedge + 4 * ne
    assert data.size == nd * vdim, (data.size, nd, vdim)
    vals = data.reshape(nd, vdim) if ordering == 1 else data.reshape(vdim, nd).T
    X = np.empty((ne, 4, 4, vdim))                                   # [e][iy][ix]
    v = ev_file
    X[:, 0, 0] = vals[v[:, 0]]; X[:, 0, 3] = vals[v[:, 1]]
    X[:, 3, 3] = vals[v[:, 2]]; X[:, 3, 0] = vals[v[:, 3]]

    def edof(j, k):                                                  # k-th node along the local edge direction
        g = np.where(el_fwd[:, j], k, 1 - k)
        return vals[nv + 2 * el_edges[:, j] + g]
    X[:, 0, 1] = edof(0, 0); X[:, 0, 2] = edof(0, 1)                 # v0 -> v1: +x at y = 0
    X[:, 1, 3] = edof(1, 0); X[:, 2, 3] = edof(1, 1)                 # v1 -> v2: +y at x = 1
    X[:, 3, 2] = edof(2, 0); X[:, 3, 1] = edof(2, 1)                 # v2 -> v3: -x at y = 1
    X[:, 2, 0] = edof(3, 0); X[:, 1, 0] = edof(3, 1)                 # v3 -> v0: -y at x = 0
    base = nv + 2 * nedge + 4 * np.arange(ne)
    X[:, 1, 1] = vals[base]; X[:, 1, 2] = vals[base + 1]
    X[:, 2, 1] = vals[base + 2]; X[:, 2, 2] = vals[base + 3]
    # equispaced -> Gauss-Lobatto nodes of the same cubic map
    T = fe.lagrange(np.array([0.0, 1.0 / 3.0, 2.0 / 3.0, 1.0]), fe.gauss_lobatto_01(4))   # [gll][equi]
    X = np.einsum('pj,qi,ejic->epqc', T, T, X)
    return X.reshape(ne, 16, vdim)


def _read_inline(path):
    kv = {}
    with open(path) as f:
        f.readline()
        for line in f:
            line = line.split('#')[0].strip()
            if '=' in line:
                k, v = line.split('=')
                kv[k.strip()] = v.strip()
    if kv.get('type') == 'quad':
        return cartesian_mesh([int(kv['nx']), int(kv['ny'])],
                              [float(kv['sx']), float(kv['sy'])])
    if kv.get('type') == 'hex':
        return cartesian_mesh([int(kv['nx']), int(kv['ny']), int(kv['nz'])],
                              [float(kv['sx']), float(kv['sy']), float(kv['sz'])])
    raise ValueError('unsupported inline mesh type')


def cartesian_mesh(n, size, origin=None, periodic=False):
    """Cartesian n[0] x n[1] (x n[2]) mesh of [origin, origin+size]; elements x fastest.
    periodic=True identifies opposite sides topologically (needs n >= 3 per direction)."""
    dim = len(n)
    origin = np.zeros(dim) if origin is None else np.asarray(origin, dtype=np.float64)
    n = np.asarray(n)
    nvd = n if periodic else n + 1
    idx = np.stack(np.meshgrid(*[np.arange(m) for m in n], indexing='ij'), -1)
    idx = idx.reshape(-1, dim)
    # element order: x fastest
    order = np.lexsort([idx[:, a] for a in range(dim)])
    idx = idx[order]
    ne = idx.shape[0]
    ev = np.empty((ne, 2 ** dim), dtype=np.int64)
    X = np.empty((ne, 2 ** dim, dim))
    for c in range(2 ** dim):
        off = np.array([(c >> a) & 1 for a in range(dim)])
        vi = (idx + off) % nvd if periodic else idx + off
        vid = np.zeros(ne, dtype=np.int64)
        for a in reversed(range(dim)):
            vid = vid * nvd[a] + vi[:, a]
        ev[:, c] = vid
        X[:, c, :] = origin + (idx + off) * (np.asarray(size) / n)
    return Mesh(dim, ev, X, 1, int(np.prod(nvd)))


# ------------------------------------------------------------------- topology helpers
def _unique_ids(keys):
    """keys [m, w] int -> ids [m] identifying equal (row-sorted) keys, and the id count."""
    ks = np.sort(keys, axis=1)
    _, inv = np.unique(ks, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    return inv, (int(inv.max()) + 1 if inv.size else 0)


def macro_lattice(ev, nv, dim):
    """Entity ids on the 3^dim lattice of every element (vertices, edge interiors, face
    interiors, element interior), i.e. an H1 order-2 numbering: two elements get the same
    id at a lattice position iff they share that topological entity.  This is what
    DofInfo::ComputeOverlapBounds needs from its H1 space (remhos_tools.cpp:432-495;
    any numbering that identifies coincident lattice points is equivalent).

    Returns (lat [NE, 3^dim], n_entities)."""
    ne = ev.shape[0]
    lat = np.empty((ne, 3 ** dim), dtype=np.int64)
    offset = nv
    by_width = {}
    for t in range(3 ** dim):
        tt = [(t // 3 ** a) % 3 for a in range(dim)]
        corners = []
        for c in range(2 ** dim):
            cc = [(c >> a) & 1 for a in range(dim)]
            if all(tt[a] == 1 or tt[a] == 2 * cc[a] for a in range(dim)):
                corners.append(c)
        by_width.setdefault(len(corners), []).append((t, corners))
    for width in sorted(by_width):
        items = by_width[width]
        if width == 1:
            for t, corners in items:
                lat[:, t] = ev[:, corners[0]]
        elif width == 2 ** dim:
            for t, corners in items:
                lat[:, t] = offset + np.arange(ne)
            offset += ne
        else:
            keys = np.concatenate([ev[:, corners] for t, corners in items], axis=0)
            ids, cnt = _unique_ids(keys)
            for m, (t, corners) in enumerate(items):
                lat[:, t] = offset + ids[m * ne:(m + 1) * ne]
            offset += cnt
    return lat, offset


def refine_uniform(mesh):
    """One level of uniform refinement (remhos.cpp:449): each quad/hex -> 4/8 children with
    the parent's orientation, children of element e stored contiguously at 2^d*e + child,
    child index lexicographic.  Geometry nodes are interpolated from the parent (MFEM
    updates the nodal grid function through the refinement interpolation)."""
    dim, g = mesh.dim, mesh.gorder
    ne = mesh.ne
    lat, nvnew = macro_lattice(mesh.ev, mesh.nv, dim)
    nch = 2 ** dim
    ev = np.empty((ne, nch, nch), dtype=np.int64)
    for ch in range(nch):
        o = [(ch >> a) & 1 for a in range(dim)]
        for c in range(nch):
            cc = [(c >> a) & 1 for a in range(dim)]
            t = sum((o[a] + cc[a]) * 3 ** a for a in range(dim))
            ev[:, ch, c] = lat[:, t]
    gll = fe.gauss_lobatto_01(g + 1)
    R = [fe.lagrange(gll, 0.5 * gll), fe.lagrange(gll, 0.5 + 0.5 * gll)]
    n1 = g + 1
    Xp = mesh.X.reshape((ne,) + (n1,) * dim + (dim,))   # axes: e, z, y, x, comp  (x fastest)
    X = np.empty((ne, nch) + (n1,) * dim + (dim,))
    for ch in range(nch):
        o = [(ch >> a) & 1 for a in range(dim)]
        T = Xp
        # axis index of direction a in the array is (dim - a) (after the element axis)
        for a in range(dim):
            ax = dim - a
            T = np.moveaxis(np.tensordot(R[o[a]], T, axes=([1], [ax])), 0, ax)
        X[:, ch] = T
    return Mesh(dim, ev.reshape(ne * nch, nch), X.reshape(ne * nch, n1 ** dim, dim), g, nvnew)


def set_curvature(mesh, order):
    """Mesh::SetCurvature(order, discont) (remhos.cpp:513): re-express the geometry as a
    degree-`order` Gauss-Lobatto nodal field by nodal interpolation."""
    dim, g = mesh.dim, mesh.gorder
    if order == g:
        return mesh
    src = fe.gauss_lobatto_01(g + 1)
    dst = fe.gauss_lobatto_01(order + 1)
    I1 = fe.lagrange(src, dst)
    I = fe.tensor_basis([I1] * dim)
    X = np.einsum('qn,enc->eqc', I, mesh.X)
    return Mesh(dim, mesh.ev, X, order, mesh.nv)


def bounding_box(mesh):
    """Mesh::GetBoundingBox (remhos.cpp:457). The nodal lattice contains the extreme points
    for the straight-sided / Q2 meshes in scope."""
    pts = mesh.X.reshape(-1, mesh.dim)
    return pts.min(axis=0), pts.max(axis=0)


class Topology:
    """Face-neighbour topology of a conforming quad/hex mesh.

    nbr_elem[e, f]  neighbour element across local face f (-1 = domain boundary)
    nbr_face[e, f]  local face id of that face in the neighbour
    fmap[e, f, :]   for each own face corner t (natural parametrisation, see
                    face_corner_lex) the index of the coincident corner in the neighbour's
                    face corner list (defines the relative orientation)
    lat[e, 3^dim]   macro-lattice entity ids (overlap bounds), n_ent their count
    """

    def __init__(self, mesh):
        dim = mesh.dim
        ne = mesh.ne
        fcl = face_corner_lex(dim)
        nf, nfc = fcl.shape
        fv = mesh.ev[:, fcl]                               # [NE, nf, nfc] global vertex ids
        ids, nfaces = _unique_ids(fv.reshape(ne * nf, nfc))
        order = np.argsort(ids, kind='stable')
        sid = ids[order]
        cnt = np.bincount(ids, minlength=nfaces)
        if cnt.max() > 2:
            raise ValueError('non-manifold or degenerate periodic mesh (face shared by >2)')
        first = np.concatenate(([0], np.cumsum(cnt)[:-1]))
        nbr = -np.ones(ne * nf, dtype=np.int64)
        two = np.nonzero(cnt == 2)[0]
        a = order[first[two]]
        b = order[first[two] + 1]
        nbr[a] = b
        nbr[b] = a
        self.nbr_elem = np.where(nbr >= 0, nbr // nf, -1).reshape(ne, nf)
        self.nbr_face = np.where(nbr >= 0, nbr % nf, -1).reshape(ne, nf)
        fvf = fv.reshape(ne * nf, nfc)
        fmap = -np.ones((ne * nf, nfc), dtype=np.int64)
        has = np.nonzero(nbr >= 0)[0]
        own = fvf[has]                                     # [m, nfc]
        oth = fvf[nbr[has]]                                # [m, nfc]
        eq = own[:, :, None] == oth[:, None, :]            # [m, own corner, nbr corner]
        if not np.all(eq.sum(axis=2) == 1):
            raise ValueError('degenerate face (repeated vertex ids) - periodic mesh too coarse')
        fmap[has] = np.argmax(eq, axis=2)
        self.fmap = fmap.reshape(ne, nf, nfc)
        self.lat, self.n_ent = macro_lattice(mesh.ev, mesh.nv, dim)
        self.dim = dim
        self.nf = nf
