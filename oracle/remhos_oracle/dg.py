"""TEST INFRASTRUCTURE ONLY -- CPU oracle (numpy) for the Remhos RK-stage hot path.

Discretisation in *full-assembly style*: dense element / face matrices built directly from
the integrator definitions (independent of any sum-factorised kernel), plus the integer
index maps of DofInfo.

Restates:
  * DofInfo::ExtractBdrDofs / FillNeighborDofs / FillSubcell2CellDof
    (remhos_tools.cpp:1356-1431, 525-676, 678-734)
  * MassIntegrator / ConvectionIntegrator / TransposeIntegrator(DGTraceIntegrator)
    as used at remhos.cpp:640-679 (MFEM semantics)
  * Assembly::ComputeFluxTerms (remhos_tools.cpp:788-858)
DOF layout: element-major, dof = k*nd + j, j lexicographic (x fastest).
"""
import numpy as np
from . import fe
from .mesh import FACE_AXIS, face_corner_lex


def bdr_dofs(p, dim):
    """BdrDofs[nfd, nf] exactly as ExtractBdrDofs (remhos_tools.cpp:1356-1431)."""
    n = p + 1
    if dim == 2:
        d = np.empty((n, 4), dtype=np.int64)
        for i in range(n):
            d[i, 0] = i
            d[i, 1] = i * n + p
            d[i, 2] = n * n - 1 - i
            d[i, 3] = (p - i) * n
        return d
    d = np.empty((n * n, 6), dtype=np.int64)
    d[:, 0] = np.arange(n * n)
    d[:, 1] = (np.arange(n)[:, None] * n * n + np.arange(n)[None, :]).reshape(-1)
    d[:, 2] = np.arange(p, n ** 3, n)
    d[:, 3] = (np.arange(n)[:, None] * n * n + np.arange(p * n, n * n)[None, :]).reshape(-1)
    d[:, 4] = np.arange(0, n ** 3, n)
    d[:, 5] = np.arange(p * n * n, n ** 3)
    return d


def sub2ind(p, dim):
    """Sub2Ind[numSubcells, 2^dim] (remhos_tools.cpp:678-734)."""
    n = p + 1
    ns = p ** dim
    out = np.empty((ns, 2 ** dim), dtype=np.int64)
    for m in range(ns):
        if dim == 2:
            aux = m + m // p
            out[m] = [aux, aux + 1, aux + p + 1, aux + p + 2]
        else:
            aux = m + m // p + (p + 1) * (m // (p * p))
            out[m] = [aux, aux + 1, aux + p + 1, aux + p + 2,
                      aux + n * n, aux + n * n + 1, aux + n * n + p + 1, aux + n * n + p + 2]
    return out


def dof_lattice(p, dim):
    """[nd, dim] integer lattice coordinates of each local DOF (x fastest)."""
    n = p + 1
    idx = np.arange(n ** dim)
    return np.stack([(idx // n ** a) % n for a in range(dim)], axis=1)


def nbr_dof_map(topo, p):
    """NbrDof[NE, nf, nfd] as DofInfo::FillNeighborDofs (remhos_tools.cpp:525-676):
    NbrDof(k,f,j) = global index (nbr*nd + local) of the neighbour DOF coincident with the
    own DOF BdrDofs(j,f); -1 on the domain boundary.  The reference derives the local index
    from MFEM's face orientation code through hand-written tables
    (GetLocalFaceDofIndex, :1078-1352); here it is derived geometrically from the shared
    vertices, which yields the same map on conforming meshes."""
    dim = topo.dim
    n = p + 1
    nd = n ** dim
    bd = bdr_dofs(p, dim)
    nfd, nf = bd.shape
    ne = topo.nbr_elem.shape[0]
    lat = dof_lattice(p, dim)
    fcl = face_corner_lex(dim)
    out = -np.ones((ne, nf, nfd), dtype=np.int64)
    # corner coordinates (0/1 per axis) of the lexicographic element corners
    cc = np.array([[(c >> a) & 1 for a in range(dim)] for c in range(2 ** dim)])
    for f in range(nf):
        axis, side = FACE_AXIS[dim][f]
        rem = [a for a in range(dim) if a != axis]
        own_l = lat[bd[:, f]]                      # [nfd, dim] own lattice coords
        ab = own_l[:, rem]                         # [nfd, dim-1] face params (natural)
        for f2 in range(nf):
            sel = np.nonzero((topo.nbr_face[:, f] == f2) & (topo.nbr_elem[:, f] >= 0))[0]
            if sel.size == 0:
                continue
            fm = topo.fmap[sel, f, :]              # [m, nfc] own corner t -> nbr corner idx
            ncorn = cc[fcl[f2]]                    # [nfc, dim] nbr element corner coords
            o = ncorn[fm[:, 0]]                    # [m, dim] image of own corner t=0
            res = o[:, None, :] * p                # start point
            res = np.broadcast_to(res, (sel.size, nfd, dim)).copy()
            for m_ax in range(dim - 1):
                dvec = ncorn[fm[:, 1 << m_ax]] - o # [m, dim] direction of own face axis m_ax
                res += ab[None, :, m_ax, None] * dvec[:, None, :]
            loc = np.zeros((sel.size, nfd), dtype=np.int64)
            for a in reversed(range(dim)):
                loc = loc * n + res[:, :, a]
            out[sel, f, :] = topo.nbr_elem[sel, f][:, None] * nd + loc
    return out


class Space:
    """Bernstein DG space of order p on a quad/hex mesh with tensor Gauss-Legendre rule of
    Q = p + dim points per direction at mesh_order 2 ((D1D,Q1D) triples at
    remhos.cpp:405-435; integrator default orders, Appendix C-6)."""

    def __init__(self, dim, p, gorder):
        self.dim, self.p, self.g = dim, p, gorder
        self.n1 = p + 1
        self.nd = self.n1 ** dim
        order = 2 * p + dim * gorder - 1
        self.Q = order // 2 + 1
        self.xq, self.wq = fe.gauss_legendre_01(self.Q)
        self.B1 = fe.bernstein(p, self.xq)
        self.G1 = fe.bernstein_deriv(p, self.xq)
        self.gll = fe.gauss_lobatto_01(gorder + 1)
        self.L1 = fe.lagrange(self.gll, self.xq)
        self.dL1 = fe.lagrange_deriv(self.gll, self.xq)
        d = dim
        self.Bt = fe.tensor_basis([self.B1] * d)                        # [Q^d, nd]
        self.Gt = [fe.tensor_basis([self.G1 if a == b else self.B1 for b in range(d)])
                   for a in range(d)]
        self.Lt = fe.tensor_basis([self.L1] * d)
        self.dLt = [fe.tensor_basis([self.dL1 if a == b else self.L1 for b in range(d)])
                    for a in range(d)]
        self.wt = fe.tensor_basis([self.wq[:, None]] * d)[:, 0]         # [Q^d]
        # lattice (uniform) points for projection / nodal sampling
        lat = np.arange(self.n1) / max(p, 1) if p > 0 else np.array([0.5])
        self.Llat = fe.tensor_basis([fe.lagrange(self.gll, lat)] * d)   # [nd, ng]
        self.bd = bdr_dofs(p, dim)
        self.nfd, self.nf = self.bd.shape
        # face tables: reference face quadrature points in the face's natural param
        self.face = []
        wf = fe.tensor_basis([self.wq[:, None]] * (d - 1))[:, 0] if d > 1 else np.ones(1)
        for f in range(self.nf):
            axis, side = FACE_AXIS[dim][f]
            one = np.array([float(side)])
            Bs, Ls, dLs = [], [], [[] for _ in range(d)]
            for b in range(d):
                if b == axis:
                    Bs.append(fe.bernstein(p, one))
                    Ls.append(fe.lagrange(self.gll, one))
                else:
                    Bs.append(self.B1)
                    Ls.append(self.L1)
                for a in range(d):
                    if b == axis:
                        dLs[a].append(fe.lagrange_deriv(self.gll, one) if a == b
                                      else fe.lagrange(self.gll, one))
                    else:
                        dLs[a].append(self.dL1 if a == b else self.L1)
            self.face.append(dict(axis=axis, sign=(1.0 if side == 1 else -1.0),
                                  B=fe.tensor_basis(Bs), L=fe.tensor_basis(Ls),
                                  dL=[fe.tensor_basis(x) for x in dLs], w=wf))

    # ---------------------------------------------------------------- geometry
    def jacobians(self, X, dLt=None):
        """J[e,q,i,j] = d x_i / d xi_j at the given points (default: volume quad points)."""
        dLt = self.dLt if dLt is None else dLt
        J = np.stack([np.einsum('qn,eni->eqi', dLt[a], X) for a in range(self.dim)], axis=-1)
        return J

    @staticmethod
    def det_adj(J):
        dim = J.shape[-1]
        if dim == 2:
            det = J[..., 0, 0] * J[..., 1, 1] - J[..., 0, 1] * J[..., 1, 0]
            adj = np.empty_like(J)
            adj[..., 0, 0] = J[..., 1, 1]; adj[..., 0, 1] = -J[..., 0, 1]
            adj[..., 1, 0] = -J[..., 1, 0]; adj[..., 1, 1] = J[..., 0, 0]
            return det, adj
        adj = np.empty_like(J)
        for i in range(3):
            for j in range(3):
                # adj[i,j] = cofactor[j,i]
                r = [a for a in range(3) if a != j]
                c = [a for a in range(3) if a != i]
                minor = J[..., r[0], c[0]] * J[..., r[1], c[1]] - J[..., r[0], c[1]] * J[..., r[1], c[0]]
                adj[..., i, j] = ((-1) ** (i + j)) * minor
        det = (J[..., 0, :] * adj[..., :, 0]).sum(axis=-1)
        return det, adj

    def quad_points(self, X):
        return np.einsum('qn,eni->eqi', self.Lt, X)

    def dof_points(self, X):
        return np.einsum('qn,eni->eqi', self.Llat, X)

    # ---------------------------------------------------------------- element matrices
    def mass_matrices(self, X):
        """M[e,i,j] = sum_q w_q detJ_q phi_i phi_j (MassIntegrator, remhos.cpp:641-644)."""
        det, _ = self.det_adj(self.jacobians(X))
        return np.einsum('q,eq,qi,qj->eij', self.wt, det, self.Bt, self.Bt, optimize=True)

    def quad_detw(self, X):
        det, _ = self.det_adj(self.jacobians(X))
        return det * self.wt[None, :]

    def conv_matrices(self, X, vq, alpha):
        """K[e,i,j] = alpha sum_q w_q phi_i (adj(J) v)_q . grad_ref phi_j
        (ConvectionIntegrator; PA data alpha*w*adj(J)*v restated at remhos_lo.cpp:1155-1190)."""
        _, adj = self.det_adj(self.jacobians(X))
        D = alpha * self.wt[None, :, None] * np.einsum('eqij,eqj->eqi', adj, vq)
        K = np.zeros((X.shape[0], self.nd, self.nd))
        for a in range(self.dim):
            K += np.einsum('eq,qi,qj->eij', D[:, :, a], self.Bt, self.Gt[a], optimize=True)
        return K

    def face_quad_points(self, X, f):
        return np.einsum('qn,eni->eqi', self.face[f]['L'], X)

    def face_flux_matrices(self, X, vfun, remap):
        """bdrInt[e, f, i, j] = - sum_q w_q phi_i phi_j vn_q |J_F|  in BdrDofs ordering
        (Assembly::ComputeFluxTerms, remhos_tools.cpp:788-858), with vn = min(0, v.n) for
        transport and -max(0, v.n) for remap, n the outward normal.
        vfun(e_pts[e,q,dim], f) -> velocity at the face quadrature points."""
        ne = X.shape[0]
        out = np.zeros((ne, self.nf, self.nfd, self.nfd))
        for f in range(self.nf):
            ft = self.face[f]
            J = self.jacobians(X, ft['dL'])
            _, adj = self.det_adj(J)
            nds = ft['sign'] * adj[:, :, ft['axis'], :]          # unnormalised outward normal
            pts = np.einsum('qn,eni->eqi', ft['L'], X)
            v = vfun(pts, f)
            vn = (v * nds).sum(axis=-1)
            vn = np.minimum(0.0, vn) if not remap else -np.maximum(0.0, vn)
            Bf = ft['B'][:, self.bd[:, f]]                       # [qf, nfd]
            out[:, f] = -np.einsum('q,eq,qi,qj->eij', ft['w'], vn, Bf, Bf, optimize=True)
        return out
