"""TEST INFRASTRUCTURE (oracle): SmoothnessIndicator of remhos_tools.cpp:24-354 (`-si 1|2`).  Pinned for
`-o 1` by the reference's two monolithic-solver known answers (autotest/out_baseline.dat:212-220);
orders above 1 follow the same statements on the subcell mesh and have no reference number.

For order 1 the "subcell mesh" is the mesh itself (remhos.cpp:870) and the H1 space of
positive order-1 elements has one DOF per mesh vertex; ShapeEval is the identity (Bernstein
values at the lattice points 0, 1).  The indicator (ComputeSmoothnessIndicator, :153-184):

    rhs = MassMixed u          (H1 x DG mass: element Q1 mass blocks, :186-237)
    y   = two lumped-mass Jacobi sweeps for  M y = rhs        (ApproximateLaplacian, :239-318)
    r2  = LaplaceOp y,   LaplaceOp = -(grad u, grad v) + <du/dn, v> on the domain boundary
          (DiffusionIntegrator(-1) + DGDiffusionIntegrator(-1, 0, 0) on boundary faces, :44-49)
    g   = two sweeps for  M g = r2
    gmin / gmax over the sparsity pattern of M (vertices sharing an element, :320-354)
    si  = 1 - ((|gmin - gmax| + 1e-50) / (|gmin| + |gmax| + 1e-50))^5                    (type 1)
        = min(1, 3 max(0, gmin gmax) / (max(gmin^2, gmax^2) + 1e-15))                    (type 2)

Element matrices are integrated with a 3-point Gauss rule per direction on the multilinear map
through the element's corners: exact on the parallelogram / rectangular meshes the known answers
use, so MFEM's own quadrature orders do not matter there."""
import numpy as np

from . import fe


class SmoothnessIndicator:
    """Any order p >= 1.  The H1 space is the positive order-1 space on the SUBCELL mesh (every element
    split into p^dim subcells whose vertices are the lattice points i/p, remhos.cpp:797-868): one DOF
    per distinct lattice point.  The DG solution enters through its values at the lattice points
    (ShapeEval, remhos_tools.cpp:107-125) and the subcell Q1 mass blocks (ComputeVariationalMatrix,
    :192-237); for p = 1 the subcell mesh is the mesh and ShapeEval the identity."""

    def __init__(self, run, si_type):
        sp, m, topo = run.space, run.mesh, run.topo
        assert si_type in (1, 2), 'Bad smoothness indicator id!'
        self.type = si_type
        self.param = 5.0 if si_type == 1 else 3.0
        dim, ne, p, nd = sp.dim, m.ne, sp.p, sp.nd
        nv = 2 ** dim
        from . import dg
        from .dg import FACE_AXIS
        import scipy.sparse as sps
        from scipy.sparse.csgraph import connected_components
        # H1 dofs = classes of coincident lattice points: DG dof (e, BdrDofs(j, f)) coincides with
        # NbrDof(e, f, j) (remhos_tools.cpp:525-676); the classes are the connected components
        nbr = run.disc.nbr                                              # [ne, nf, nfd]
        own = (np.arange(ne)[:, None, None] * nd + sp.bd.T[None, :, :])  # [ne, nf, nfd]
        ok = nbr >= 0
        N_dg = ne * nd
        G = sps.coo_matrix((np.ones(ok.sum()), (own[ok], nbr[ok])), shape=(N_dg, N_dg))
        self.N, lab = connected_components(G, directed=False)
        cg_dof = lab.reshape(ne, nd)                                    # DG dof -> H1 dof
        s2i = dg.sub2ind(p, dim)                                        # [nsub, nv] lexicographic corners
        nsub = s2i.shape[0]
        self.cg = cg_dof[:, s2i]                                        # [ne, nsub, nv]
        xlat = sp.dof_points(m.X)                                       # [ne, nd, dim]
        Xc = xlat[:, s2i, :].reshape(ne * nsub, nv, dim)                # subcell corners
        # Q1 shape functions and gradients on a 3-point Gauss rule
        xq, wq = fe.gauss_legendre_01(3)
        L = fe.lagrange(np.array([0.0, 1.0]), xq)                        # [3, 2]
        dL = fe.lagrange_deriv(np.array([0.0, 1.0]), xq)
        Phi = fe.tensor_basis([L] * dim)                                # [nq, nv]
        dPhi = [fe.tensor_basis([dL if a == b else L for b in range(dim)]) for a in range(dim)]
        wt = fe.tensor_basis([wq[:, None]] * dim)[:, 0]
        J = np.stack([np.einsum('qn,eni->eqi', dPhi[a], Xc) for a in range(dim)], axis=3)  # [e,q,i,a]
        det = np.linalg.det(J)
        Jinv = np.linalg.inv(J)                                         # [e,q,a,i]
        grad = np.einsum('aqn,eqai->eqni', np.stack(dPhi), Jinv)        # physical gradients [e,q,n,i]
        wdet = wt[None, :] * np.abs(det)
        Me = np.einsum('eq,qi,qj->eij', wdet, Phi, Phi)
        Ke = -np.einsum('eq,eqid,eqjd->eij', wdet, grad, grad)
        # subcell faces on the domain boundary: + <dn phi_j, phi_i>
        xf, wf = fe.gauss_legendre_01(3)
        sub_lat = dg.dof_lattice(p - 1, dim) if p > 1 else np.zeros((1, dim), dtype=int)   # [nsub, dim]
        for f in range(sp.nf):
            axis, side = FACE_AXIS[dim][f]
            touches = sub_lat[:, axis] == (p - 1 if side else 0)        # subcells on element face f
            bnd = ((topo.nbr_elem[:, f] < 0)[:, None] & touches[None, :]).reshape(-1)
            if not bnd.any():
                continue
            one = np.array([float(side)])
            Ls = [fe.lagrange(np.array([0.0, 1.0]), one if b == axis else xf) for b in range(dim)]
            dLs = [[fe.lagrange_deriv(np.array([0.0, 1.0]), one if b == axis else xf) if a == b
                    else fe.lagrange(np.array([0.0, 1.0]), one if b == axis else xf)
                    for b in range(dim)] for a in range(dim)]
            Pf = fe.tensor_basis(Ls)                                    # [nqf, nv]
            dPf = [fe.tensor_basis(x) for x in dLs]
            wft = fe.tensor_basis([wf[:, None]] * (dim - 1))[:, 0] if dim > 1 else np.ones(1)
            Jf = np.stack([np.einsum('qn,eni->eqi', dPf[a], Xc[bnd]) for a in range(dim)], axis=3)
            detf = np.linalg.det(Jf)
            Jfi = np.linalg.inv(Jf)
            gradf = np.einsum('aqn,eqai->eqni', np.stack(dPf), Jfi)
            # outward normal times surface element: sign * det(J) * J^-T e_axis
            nrm = (1.0 if side else -1.0) * detf[:, :, None] * Jfi[:, :, axis, :]
            dn = np.einsum('eqni,eqi->eqn', gradf, nrm)                 # (grad phi_j . n) dS / w
            Bf = np.einsum('q,eqi,eqj->eij', wft, Pf[None, :, :].repeat(bnd.sum(), 0), dn)
            Ke[bnd] += Bf
        N = self.N
        cgs = self.cg.reshape(ne * nsub, nv)
        rows = np.repeat(cgs[:, :, None], nv, axis=2).reshape(-1)
        cols = np.repeat(cgs[:, None, :], nv, axis=1).reshape(-1)
        self.M = sps.csr_matrix((Me.reshape(-1), (rows, cols)), shape=(N, N))
        self.M.sum_duplicates()
        self.Lap = sps.csr_matrix((Ke.reshape(-1), (rows, cols)), shape=(N, N))
        self.Lap.sum_duplicates()
        self.ml = np.asarray(self.M.sum(axis=1)).reshape(-1)            # LumpedIntegrator: row sums
        # MassMixed: rows H1 dofs, columns the lattice-point values of the DG field (the "switchero" of
        # :76-92 only translates MFEM's counter-clockwise vertex order into the lexicographic DG order)
        dgcol = (np.arange(ne)[:, None, None] * nd + s2i[None, :, :]).reshape(ne * nsub, nv)
        mc = np.repeat(dgcol[:, None, :], nv, axis=1).reshape(-1)
        self.Mmix = sps.csr_matrix((Me.reshape(-1), (rows, mc)), shape=(N, ne * nd))
        self.Mmix.sum_duplicates()
        # ShapeEval: Bernstein coefficients -> values at the closed uniform points (:107-125)
        lat_pts = np.arange(p + 1) / max(p, 1)
        self.V = fe.tensor_basis([fe.bernstein(p, lat_pts)] * dim)       # [nd points, nd coefficients]
        # DG2CG: H1 dof of every DG dof, -1 on the domain boundary (:94-105)
        d2c = cg_dof.copy()
        for f in range(sp.nf):
            bnd = topo.nbr_elem[:, f] < 0
            if bnd.any():
                d2c[np.ix_(bnd, sp.bd[:, f])] = -1
        self.DG2CG = d2c                                                # [ne, nd]
        pat = (self.M != 0).tocsr()
        self.pI, self.pJ = pat.indptr, pat.indices

    def _solve2(self, rhs):
        """two sweeps of y <- y - (M y - rhs) / m_lumped from y = 0 (stop if |M y - rhs|_2 <= 1e-10)"""
        y = np.zeros_like(rhs)
        for _ in range(2):
            z = self.M @ y - rhs
            if np.sqrt((z * z).sum()) <= 1e-10:
                break
            y = y - z / self.ml
        return y

    def compute(self, u):
        """u: [ne, nd] DG coefficients -> si per H1 dof [N]"""
        rhs = self.Mmix @ (u @ self.V.T).reshape(-1)                    # xEval = ShapeEval u, element by element
        y = self._solve2(rhs)
        g = self._solve2(self.Lap @ y)
        gmin = np.minimum.reduceat(g[self.pJ], self.pI[:-1])
        gmax = np.maximum.reduceat(g[self.pJ], self.pI[:-1])
        if self.type == 1:
            eps = 1.0e-50
            return 1.0 - ((np.abs(gmin - gmax) + eps) / (np.abs(gmin) + np.abs(gmax) + eps)) ** self.param
        eps = 1.0e-15
        return np.minimum(1.0, self.param * np.maximum(0.0, gmin * gmax) /
                          (np.maximum(gmin * gmin, gmax * gmax) + eps))

    def dof_values(self, u):
        """tmp of remhos_mono.cpp:134-135: si at the DOF's vertex, 1 on the domain boundary"""
        si = self.compute(u)
        return np.where(self.DG2CG < 0, 1.0, si[np.maximum(self.DG2CG, 0)])
