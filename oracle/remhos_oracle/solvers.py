"""TEST INFRASTRUCTURE ONLY -- CPU oracle (numpy) for the Remhos RK-stage hot path.

The stage operator and its solver components, restated from the reference:
  LocalInverseHOSolver::CalcHOSolution          remhos_ho.cpp:84-129 (FA-exact semantics)
  DiscreteUpwind::CalcLOSolution                remhos_lo.cpp:43-100
  MassBasedAvg::CalcLOSolution                  remhos_lo.cpp:247-324
  ResidualDistribution::CalcLOSolution          remhos_lo.cpp:111-245
  Assembly::LinearFluxLumping                   remhos_tools.cpp:876-913
  DofInfo::ComputeElementsMinMax / ComputeBounds remhos_tools.cpp:497-523, 381-495
  ClipScaleSolver::CalcFCTSolution              remhos_fct.cpp:449-541
  FluxBasedFCT::CalcFCTSolution                 remhos_fct.cpp:155-181, 295-446
  AdvectionOperator::MultUnlimited / LimitMult  remhos.cpp:1596-1739, 1798-1916
All vectors are element-major [NE, nd] views of length-N arrays.
"""
import numpy as np
from . import dg


class Assembled:
    """Everything that depends on the mesh position (re-assembled per stage in remap,
    remhos.cpp:1598-1677)."""
    pass


class Discretization:
    def __init__(self, space, topo, X0, exec_mode, vel_fun=None, Vnodes=None,
                 inflow_vals=None, need_sparse=False):
        """vel_fun(points[e,q,dim]) -> v for transport; Vnodes[e,ng,dim] nodal mesh velocity
        for remap (VectorGridFunctionCoefficient v_mesh_coeff, remhos.cpp:561,655)."""
        self.sp = space
        self.topo = topo
        self.X0 = X0
        self.exec_mode = exec_mode
        self.vel_fun = vel_fun
        self.Vnodes = Vnodes
        self.ne = X0.shape[0]
        self.nd = space.nd
        self.N = self.ne * self.nd
        self.nbr = dg.nbr_dof_map(topo, space.p)                      # [NE, nf, nfd]
        self.inflow = (np.zeros((self.ne, self.nd)) if inflow_vals is None else inflow_vals)
        self.need_sparse = need_sparse
        self.cur = None

    # ------------------------------------------------------------------ assembly
    def assemble(self, t=0.0):
        sp = self.sp
        A = Assembled()
        if self.exec_mode == 1:
            X = self.X0 + t * self.Vnodes
            vq = np.einsum('qn,eni->eqi', sp.Lt, self.Vnodes)
            alpha = 1.0

            def vface(pts, f):
                return np.einsum('qn,eni->eqi', sp.face[f]['L'], self.Vnodes)
        else:
            X = self.X0
            vq = self.vel_fun(sp.quad_points(X))
            alpha = -1.0

            def vface(pts, f):
                return self.vel_fun(pts)
        A.X = X
        A.M = sp.mass_matrices(X)
        A.ml = A.M.sum(axis=2)                                        # lumped mass (row sums)
        A.wdet = sp.quad_detw(X)
        A.K = sp.conv_matrices(X, vq, alpha)
        A.bdrInt = sp.face_flux_matrices(X, vface, remap=(self.exec_mode == 1))
        self.cur = A
        return A

    # ------------------------------------------------------------------ helpers
    def face_diffs(self, u, boundary_vals):
        """xDiff[e,f,j] = u_nbr - u_own at the face DOFs (BdrDofs order).  boundary_vals:
        None -> exterior state 0 (HO operator), else [NE,nd] exterior values (inflow_gf)."""
        sp = self.sp
        uf = u.reshape(-1)
        own = u[:, sp.bd.T]                                           # [NE, nf, nfd]
        nb = self.nbr
        un = np.where(nb >= 0, uf[np.maximum(nb, 0)], 0.0)
        if boundary_vals is not None:
            un = np.where(nb >= 0, un, boundary_vals[:, sp.bd.T])
        return un - own

    def apply_K_HO(self, u):
        """rhs = K_HO u: volume convection + upwinded face terms (transposed DG trace),
        exterior state 0 on domain-boundary faces."""
        A, sp = self.cur, self.sp
        rhs = np.einsum('eij,ej->ei', A.K, u)
        diff = self.face_diffs(u, None)
        contrib = np.einsum('efij,efj->efi', A.bdrInt, diff)
        for f in range(sp.nf):
            np.add.at(rhs, (slice(None), sp.bd[:, f]), contrib[:, f, :])
        return rhs

    def ho_local_inverse(self, u):
        rhs = self.apply_K_HO(u)
        return np.linalg.solve(self.cur.M, rhs[:, :, None])[:, :, 0]

    # ------------------------------------------------------------------ LO solvers
    def lo_mass_based_avg(self, u, du_ho, dt):
        A, sp = self.cur, self.sp
        u_new = u + dt * du_ho
        uq = np.einsum('qi,ei->eq', sp.Bt, u_new)
        mass = (A.wdet * uq).sum(axis=1)
        vol = A.wdet.sum(axis=1)
        return ((mass / vol)[:, None] - u) / dt

    def du_matrix(self, K):
        """ComputeDiscreteUpwindMatrix on the element-block (volume-only) K
        (remhos_lo.cpp:76-100)."""
        Kt = np.swapaxes(K, 1, 2)
        d = np.maximum(np.maximum(0.0, -K), -Kt)
        D = K + d
        idx = np.arange(K.shape[1])
        dsum = d.sum(axis=2) - d[:, idx, idx]
        D[:, idx, idx] = K[:, idx, idx] - dsum
        return D

    def lumped_face_terms(self, u):
        """sum over faces of LinearFluxLumping with alpha = 0: y_i += (sum_j bdrInt_ij)
        (u_nbr_i - u_own_i), inflow_gf as exterior state on the boundary."""
        A, sp = self.cur, self.sp
        y = np.zeros_like(u)
        diff = self.face_diffs(u, self.inflow)
        contrib = A.bdrInt.sum(axis=3) * diff
        for f in range(sp.nf):
            np.add.at(y, (slice(None), sp.bd[:, f]), contrib[:, f, :])
        return y

    def ho_neumann(self, u):
        """NeumannHOSolver::CalcHOSolution (remhos_ho.cpp:136-187): rhs = k u + Galerkin face terms
        (LinearFluxLumping with alpha = 1, inflow exterior state), then the Neumann iteration
        du <- du - (M du - rhs)/m_L, at most 20 sweeps, stop at |res|_2 <= 1e-4 (global norm)."""
        A, sp = self.cur, self.sp
        rhs = np.einsum('eij,ej->ei', A.K, u)
        diff = self.face_diffs(u, self.inflow)
        contrib = np.einsum('efij,efj->efi', A.bdrInt, diff)
        for f in range(sp.nf):
            np.add.at(rhs, (slice(None), sp.bd[:, f]), contrib[:, f, :])
        du = np.zeros_like(u)
        for _ in range(20):
            res = np.einsum('eij,ej->ei', A.M, du) - rhs
            if np.sqrt((res * res).sum()) <= 1.0e-4:
                break
            du = du - res / A.ml
        return du

    def precond_conv(self):
        """PrecondConvectionIntegrator (remhos_tools.cpp:975-1031): M_L M^-1 K per element."""
        A = self.cur
        return A.ml[:, :, None] * np.linalg.solve(A.M, A.K)

    def lo_discrete_upwind(self, u, prec=False):
        A = self.cur
        D = self.du_matrix(self.precond_conv() if prec else A.K)
        y = np.einsum('eij,ej->ei', D, u) + self.lumped_face_terms(u)
        return y / A.ml

    def lo_residual_distribution(self, u, subcell_weights=None):
        """ResidualDistribution::CalcLOSolution (remhos_lo.cpp:111-245), gamma = 1.
        subcell_weights[e, m, c] enables the subcell variant."""
        A, sp = self.cur, self.sp
        eps = 1e-15
        nd = sp.nd
        z = np.einsum('eij,ej->ei', A.K, u)
        du = self.lumped_face_terms(u)
        xmax = u.max(axis=1); xmin = u.min(axis=1); xsum = u.sum(axis=1)
        rhoP = np.maximum(0.0, z).sum(axis=1)
        rhoN = np.minimum(0.0, z).sum(axis=1)
        sumWP = nd * xmax - xsum + eps
        sumWN = nd * xmin - xsum - eps
        wP = (xmax[:, None] - u) / sumWP[:, None]
        wN = (xmin[:, None] - u) / sumWN[:, None]
        if subcell_weights is not None:
            gamma = 1.0
            s2i = dg.sub2ind(sp.p, sp.dim)                           # [ns, nc]
            us = u[:, s2i]                                           # [NE, ns, nc]
            fluct = (subcell_weights * us).sum(axis=2)
            smax = us.max(axis=2); smin = us.min(axis=2); ssum = us.sum(axis=2)
            nc = s2i.shape[1]
            swP = nc * smax - ssum + eps
            swN = nc * smin - ssum - eps
            fP = np.maximum(0.0, fluct); fN = np.minimum(0.0, fluct)
            sfP = fP.sum(axis=1); sfN = fN.sum(axis=1)
            nwP = np.zeros_like(u); nwN = np.zeros_like(u)
            cP = fP[:, :, None] * ((smax[:, :, None] - us) / swP[:, :, None])
            cN = fN[:, :, None] * ((smin[:, :, None] - us) / swN[:, :, None])
            for m in range(s2i.shape[0]):
                for c in range(nc):
                    nwP[:, s2i[m, c]] += cP[:, m, c]
                    nwN[:, s2i[m, c]] += cN[:, m, c]
            aux = gamma / (rhoP + eps)
            wP = wP * (1.0 - np.minimum(aux * sfP, 1.0))[:, None] \
                + np.minimum(aux, 1.0 / (sfP + eps))[:, None] * nwP
            aux = gamma / (rhoN - eps)
            wN = wN * (1.0 - np.minimum(aux * sfN, 1.0))[:, None] \
                + np.maximum(aux, 1.0 / (sfN - eps))[:, None] * nwN
        return (du + wP * rhoP[:, None] + wN * rhoN[:, None]) / A.ml

    # ------------------------------------------------------------------ monolithic solver
    def nonlin_flux_lumping(self, u, alpha):
        """sum over the faces (ascending) of Assembly::NonlinFluxLumping (remhos_tools.cpp:915-973):
        the lumped face term plus the alpha-weighted anti-diffusive correction, rescaled per face so
        that the corrections of one face sum to zero whenever both signs occur.
        alpha: [NE, nd] (or None for alpha = 1).  Returns the increment y[NE, nd]."""
        A, sp = self.cur, self.sp
        eps = 1e-15
        y = np.zeros_like(u)
        diff = self.face_diffs(u, self.inflow)                        # [NE, nf, nfd]
        for f in range(sp.nf):
            B = A.bdrInt[:, f]                                        # [NE, nfd, nfd]
            d = diff[:, f]
            lump = np.zeros_like(d)
            corr = np.zeros_like(d)
            for j in range(sp.nfd):                                   # the reference's summation order
                lump = lump + B[:, :, j] * d
                corr = corr + B[:, :, j] * (d[:, j][:, None] - d)
            if alpha is not None:
                corr = corr * alpha[:, sp.bd[:, f]]
            sP = np.maximum(0.0, corr).sum(axis=1)
            sN = np.minimum(0.0, corr).sum(axis=1)
            with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
                cP = np.minimum(0.0, corr) - np.maximum(0.0, corr) * (sN / sP)[:, None]
                cN = np.maximum(0.0, corr) - np.minimum(0.0, corr) * (sP / sN)[:, None]
            tot = sP + sN
            corr = np.where((tot > eps)[:, None], cP, np.where((tot < -eps)[:, None], cN, corr))
            np.add.at(y, (slice(None), sp.bd[:, f]), lump + corr)
        return y

    def mono_rd(self, u, bounds_type, scale, subcell_weights=None, mass_lim=True, si_tmp=None):
        """MonoRDSolver::CalcSolution (remhos_mono.cpp:60-356); si_tmp[NE, nd] = the smoothness
        indicator value at each DOF (1 on the domain boundary, remhos_mono.cpp:134-135) or None:
        convex-limited residual distribution (beta = gamma = 10) with the element-local fixed-point
        mass correction (at most 101 sweeps, |res|_2 <= 1e-8).  scale[NE]: remhos_mono.cpp:40-57."""
        A, sp = self.cur, self.sp
        eps, beta, gamma, tol = 1e-15, 10.0, 10.0, 1e-8
        nd = sp.nd
        xi_min, xi_max = self.bounds(u, bounds_type)
        xe_min = u.min(axis=1); xe_max = u.max(axis=1)
        z = np.einsum('eij,ej->ei', A.K, u)
        d = z.copy()
        lo_gap = np.minimum(xi_max - u, u - xi_min)
        alpha = np.minimum(1.0, beta * lo_gap / (np.maximum(xi_max - u, u - xi_min) + eps))
        if si_tmp is not None:                                        # remhos_mono.cpp:132-153
            tmp = si_tmp
            bndN = np.maximum(0.0, tmp * (2.0 * u - xi_max) + (1.0 - tmp) * xi_min)
            bndP = np.minimum(1.0, tmp * (2.0 * u - xi_min) + (1.0 - tmp) * xi_max)
            aN = np.minimum(1.0, beta * (u - bndN) / (xi_max - u + eps))
            aP = np.minimum(1.0, beta * (bndP - u) / (u - xi_min + eps))
            ssum = xi_min + xi_max
            alpha = np.where(ssum > 2.0 * u + eps, aN, np.where(ssum < 2.0 * u - eps, aP, alpha))
        du = alpha * z
        z = z - alpha * z
        du = du + self.nonlin_flux_lumping(u, alpha)
        d = d + self.nonlin_flux_lumping(u, None)
        xsum = u.sum(axis=1)
        rhoP = np.maximum(0.0, z).sum(axis=1)
        rhoN = np.minimum(0.0, z).sum(axis=1)
        sumWP = nd * xe_max - xsum + eps
        sumWN = nd * xe_min - xsum - eps
        wP = (xe_max[:, None] - u) / sumWP[:, None]
        wN = (xe_min[:, None] - u) / sumWN[:, None]
        if subcell_weights is not None:
            s2i = dg.sub2ind(sp.p, sp.dim)
            us = u[:, s2i]
            fluct = (subcell_weights * us).sum(axis=2)
            smax = us.max(axis=2); smin = us.min(axis=2); ssum = us.sum(axis=2)
            nc = s2i.shape[1]
            swP = nc * smax - ssum + eps
            swN = nc * smin - ssum - eps
            fP = np.maximum(0.0, fluct); fN = np.minimum(0.0, fluct)
            sfP = fP.sum(axis=1); sfN = fN.sum(axis=1)
            nwP = np.zeros_like(u); nwN = np.zeros_like(u)
            cP = fP[:, :, None] * ((smax[:, :, None] - us) / swP[:, :, None])
            cN = fN[:, :, None] * ((smin[:, :, None] - us) / swN[:, :, None])
            for m in range(s2i.shape[0]):
                for c in range(nc):
                    nwP[:, s2i[m, c]] += cP[:, m, c]
                    nwN[:, s2i[m, c]] += cN[:, m, c]
            aux = gamma / (rhoP + eps)
            wP = wP * (1.0 - np.minimum(aux * sfP, 1.0))[:, None] \
                + np.minimum(aux, 1.0 / (sfP + eps))[:, None] * nwP
            aux = gamma / (rhoN - eps)
            wN = wN * (1.0 - np.minimum(aux * sfN, 1.0))[:, None] \
                + np.maximum(aux, 1.0 / (sfN - eps))[:, None] * nwN
        du = du + wP * rhoP[:, None] + wN * rhoN[:, None]
        # time derivative and mass matrix: per-element fixed point (eq. 27-29)
        m_it = np.zeros_like(u)
        if mass_lim:
            active = np.ones(u.shape[0], dtype=bool)
            msum = A.M.sum(axis=2)
            for it in range(101):
                uDot = (du + m_it) / A.ml
                uDotMin = uDot.min(axis=1)[:, None]; uDotMax = uDot.max(axis=1)[:, None]
                m_new = msum * uDot - np.einsum('eij,ej->ei', A.M, uDot)   # sum_j M_ij (uDot_i - uDot_j)
                diff = d - du
                ratio = np.abs(m_new) / (np.abs(diff) + eps)
                if si_tmp is not None:
                    ratio = np.maximum(si_tmp, ratio)
                m_new = m_new + np.minimum(1.0, ratio) * diff
                den = np.maximum(uDotMax - uDot, uDot - uDotMin) + eps
                al = np.minimum(1.0, beta * scale[:, None] * lo_gap / den)
                if si_tmp is not None:                                # remhos_mono.cpp:316-324
                    aglob = np.minimum(1.0, beta * scale[:, None] * np.minimum(1.0 - u, u - 0.0) / den)
                    al = np.minimum(np.maximum(si_tmp, al), aglob)
                m_new = m_new * al
                MP = np.maximum(0.0, m_new).sum(axis=1); MN = np.minimum(0.0, m_new).sum(axis=1)
                with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
                    cP = np.minimum(0.0, m_new) - np.maximum(0.0, m_new) * (MN / MP)[:, None]
                    cN = np.maximum(0.0, m_new) - np.minimum(0.0, m_new) * (MP / MN)[:, None]
                tot = MP + MN
                m_new = np.where((tot > eps)[:, None], cP, np.where((tot < -eps)[:, None], cN, m_new))
                res = m_new + du - A.ml * uDot
                m_it = np.where(active[:, None], m_new, m_it)
                active = active & ~(np.sqrt((res * res).sum(axis=1)) <= tol)
                if not active.any():
                    break
        return (du + m_it) / A.ml

    # ------------------------------------------------------------------ bounds
    def bounds(self, u, bounds_type, active_el=None, active_dof=None):
        """ComputeElementsMinMax + ComputeBounds (remhos_tools.cpp:381-523).  active_el / active_dof:
        the masked variants used for the product field (inactive elements and dofs do not
        contribute; dofs no active element touches get +-inf)"""
        sp, topo = self.sp, self.topo
        if active_dof is not None:
            xe_min = np.where(active_dof, u, np.inf).min(axis=1)
            xe_max = np.where(active_dof, u, -np.inf).max(axis=1)
        else:
            xe_min = u.min(axis=1); xe_max = u.max(axis=1)
        if active_el is not None:
            xe_min = np.where(active_el, xe_min, np.inf)
            xe_max = np.where(active_el, xe_max, -np.inf)
        if bounds_type == 0:
            emin = np.full(topo.n_ent, np.inf); emax = np.full(topo.n_ent, -np.inf)
            np.minimum.at(emin, topo.lat, xe_min[:, None])
            np.maximum.at(emax, topo.lat, xe_max[:, None])
            lat = dg.dof_lattice(sp.p, sp.dim)
            cls = np.where(lat == 0, 0, np.where(lat == sp.p, 2, 1))
            t = sum(cls[:, a] * 3 ** a for a in range(sp.dim))       # [nd] macro position
            ent = topo.lat[:, t]                                      # [NE, nd]
            return emin[ent], emax[ent]
        nb = topo.nbr_elem
        mn = np.where(nb >= 0, xe_min[np.maximum(nb, 0)], np.inf).min(axis=1)
        mx = np.where(nb >= 0, xe_max[np.maximum(nb, 0)], -np.inf).max(axis=1)
        mn = np.minimum(mn, xe_min); mx = np.maximum(mx, xe_max)
        return (np.repeat(mn[:, None], sp.nd, axis=1), np.repeat(mx[:, None], sp.nd, axis=1))

    # ------------------------------------------------------------------ product fields (-ps)
    @staticmethod
    def bool_indicators(u, tol=1e-12):
        """ComputeBoolIndicators (remhos_sync.cpp:24-47), EMPTY_ZONE_TOL = 1e-12"""
        dofs = u > tol
        return dofs.any(axis=1), dofs

    @staticmethod
    def compute_ratio(us, u):
        """ComputeRatio (remhos_sync.cpp:50-94): s = us / u on active dofs, the average of those
        ratios on the other dofs of an active element, 0 in inactive elements"""
        el, dofs = Discretization.bool_indicators(u)
        with np.errstate(divide='ignore', invalid='ignore'):
            r = np.where(dofs, us / u, 0.0)
        n = dofs.sum(axis=1)
        avg = np.where(n > 0, r.sum(axis=1) / np.maximum(n, 1), 0.0)
        s = np.where(dofs, r, avg[:, None])
        s = np.where(el[:, None], s, 0.0)
        return s, el, dofs

    def compatible_lo_product(self, us, m, d_us_ho, s_min, s_max, u_new, act_el, act_dof, dt):
        """FCTSolver::CalcCompatibleLOProduct (remhos_fct.cpp:26-118); returns (d_us_LO, s_min, s_max)
        with the bounds the reference adjusts in place"""
        eps = 1e-12
        s_min = s_min.copy(); s_max = s_max.copy()
        mass_us = ((us + dt * d_us_ho) * m).sum(axis=1)
        mass_u = (u_new * m).sum(axis=1)
        with np.errstate(divide='ignore', invalid='ignore'):
            s_avg = mass_us / mass_u
        smin = np.where(act_dof, s_min, np.inf).min(axis=1)
        smax = np.where(act_dof, s_max, -np.inf).max(axis=1)
        any_act = act_dof.any(axis=1)
        with np.errstate(invalid='ignore'):
            fix_lo = any_act & (s_avg < smin) & (mass_us + eps > smin * mass_u)
            s_avg = np.where(fix_lo, smin, s_avg)
            fix_hi = any_act & (s_avg > smax) & (mass_us - eps < smax * mass_u)
            s_avg = np.where(fix_hi, smax, s_avg)
        sa = s_avg[:, None]
        upd = act_dof & act_el[:, None]
        s_min = np.where(upd & (sa + eps < s_min), sa, s_min)
        s_max = np.where(upd & (sa - eps > s_max), sa, s_max)
        d_lo = np.where(act_el[:, None], (u_new * sa - us) / dt, 0.0)
        return d_lo, s_min, s_max

    @staticmethod
    def scale_product_bounds(s_min, s_max, u_new, act_el, act_dof):
        """FCTSolver::ScaleProductBounds (remhos_fct.cpp:120-153)"""
        on = act_dof & act_el[:, None]
        with np.errstate(invalid='ignore'):
            return np.where(on, s_min * u_new, 0.0), np.where(on, s_max * u_new, 0.0)

    # ------------------------------------------------------------------ FCT
    def fct_clip_scale(self, u, m, du_ho, du_lo, umin, umax, dt):
        eps = 1.0e-15
        u_new_lo = u + dt * du_lo
        fmin = m / dt * (umin - u_new_lo)
        fmax = m / dt * (umax - u_new_lo)
        f = m * (du_ho - du_lo)
        f = np.minimum(fmax, np.maximum(fmin, f))
        # sequential sums in DOF order, as the reference loop
        sumNeg = np.zeros(u.shape[0]); sumPos = np.zeros(u.shape[0])
        for j in range(u.shape[1]):
            sumNeg = sumNeg + np.minimum(f[:, j], 0.0)
            sumPos = sumPos + np.maximum(f[:, j], 0.0)
        new_mass = sumNeg + sumPos
        with np.errstate(divide='ignore', invalid='ignore'):
            fpos = np.minimum(0.0, f) - np.maximum(0.0, f) * sumNeg[:, None] / sumPos[:, None]
            f = np.where((new_mass > eps)[:, None], fpos, f)
            fneg = np.maximum(0.0, f) - np.minimum(0.0, f) * sumPos[:, None] / sumNeg[:, None]
            f = np.where((new_mass < -eps)[:, None], fneg, f)
        return du_lo + f / m

    @staticmethod
    def si_update_bounds(u_ho, si_tmp, umin, umax):
        """SmoothnessIndicator::UpdateBounds (remhos_tools.cpp:183-190), all dofs at once: si_tmp is the
        indicator at the dof (1 on the domain boundary); u_ho = u + dt du_HO"""
        return (np.maximum(0.0, si_tmp * u_ho + (1.0 - si_tmp) * umin),
                np.minimum(1.0, si_tmp * u_ho + (1.0 - si_tmp) * umax))

    @staticmethod
    def _penalty_get_z(lam, w, flux):
        return np.where(np.abs(flux) >= lam * np.abs(w), lam * w, flux)

    @classmethod
    def _penalty_get_lambda(cls, delta, w, flux, max_iter=200):
        """get_lambda (remhos_fct.cpp:843-926), statement by statement; sums run in DOF order.  The
        reference loops until abs(F) <= 1e-15 with no cap; here both loops stop after max_iter rounds (the
        bracket has collapsed to one double long before), stated in DESIGN.md."""
        tol = 1e-15

        def lsz(lam):                                          # get_lambda_times_sum_z
            z = cls._penalty_get_z(lam, w, flux)
            acc = 0.0
            for v in z:
                acc += v
            return acc
        lam = 1.0
        F = delta - lsz(lam)
        factor = 1.0
        for _ in range(max_iter):
            factor *= 2.0
            lo, hi = lam / factor, factor * lam
            FL, FU = delta - lsz(lo), delta - lsz(hi)
            if not (F * FL > 0 and F * FU > 0):
                break
        if F * FL < 0:
            hi = lam
        else:
            lo = lam
        FL, FU = delta - lsz(lo), delta - lsz(hi)
        for _ in range(max_iter):
            lam = 0.5 * (lo + hi)
            F = delta - lsz(lam)
            if F * FL < 0:
                hi, FU = lam, F
            else:
                lo, FL = lam, F
            if not abs(F) > tol:
                break
        lam = 0.5 * (lo + hi)
        return cls._penalty_get_z(lam, w, flux)

    def fct_nonlinear_penalty(self, u, m, du_ho, du_lo, umin, umax, dt, eps_w, si_tmp=None):
        """NonlinearPenaltySolver::CalcFCTSolution + CorrectFlux (remhos_fct.cpp:760-996).  eps_w =
        GetElementSize(0, 0) / order (:961)."""
        if si_tmp is not None:
            umin, umax = self.si_update_bounds(u + dt * du_ho, si_tmp, umin, umax)
        star = np.minimum((umax - u) / dt, np.maximum(du_ho, (umin - u) / dt))     # uses u at the old time
        fL = m * (star - du_lo)
        fH = m * (star - du_ho)
        corr = np.zeros_like(fL)
        for e in range(u.shape[0]):
            fl, fh = fL[e], fH[e]
            fp = 0.0; fn = 0.0
            for v in fl:
                if v >= 0.0:
                    fp += v
                else:
                    fn += v
            delta = fp + fn
            if delta == 0.0:
                continue
            mx = max(np.abs(fh).max(), -1.0)
            if delta > 0.0:
                w = np.where(fl > 0.0, eps_w * np.abs(fl) + abs(mx), 0.0)
            else:
                w = np.where(fl < 0.0, -eps_w * np.abs(fl) - abs(mx), 0.0)
            corr[e] = -self._penalty_get_lambda(delta, w, fl)
        return du_lo + (fL + corr) / m

    def fct_project(self, u, du_ho, du_lo, umin, umax, dt):
        """ElementFCTProjection::CalcFCTSolution (remhos_fct.cpp:613-733): element-local Zalesak
        limiter on the fluxes F_ij = M_ij (du_i - du_j) + (beta_j z_i - beta_i z_j), beta = M_L / sum M_L,
        z = M du_HO - M_L du_LO, started from the LO rate."""
        A = self.cur
        M = A.M
        ML = M.sum(axis=2)
        rhs = np.einsum('eij,ej->ei', M, du_ho)
        beta = ML / ML.sum(axis=1)[:, None]
        z = rhs - ML * du_lo
        F = M * (du_ho[:, :, None] - du_ho[:, None, :]) + \
            (beta[:, None, :] * z[:, :, None] - beta[:, :, None] * z[:, None, :])
        idx = np.arange(M.shape[1])
        F[:, idx, idx] = 0.0
        sp_ = np.maximum(0.0, F).sum(axis=2)
        sm_ = np.minimum(0.0, F).sum(axis=2)
        du_max = (umax - u) / dt
        du_min = (umin - u) / dt
        rp = np.maximum(ML * (du_max - du_lo), 0.0)
        rm = np.minimum(ML * (du_min - du_lo), 0.0)
        with np.errstate(divide='ignore', invalid='ignore'):
            gp = np.where(rp < sp_, rp / sp_, 1.0)
            gm = np.where(rm > sm_, rm / sm_, 1.0)
        a = np.where(F >= 0.0, np.minimum(gp[:, :, None], gm[:, None, :]),
                     np.minimum(gm[:, :, None], gp[:, None, :]))
        return du_lo + (a * F).sum(axis=2) / ML

    def build_sparse_K_HO(self):
        """Upper-triangular coupling list of K_HO (volume blocks + face blocks) for the
        flux-based FCT: arrays (I, J, kij, kji, same_elem, Mij)."""
        A, sp = self.cur, self.sp
        if getattr(A, '_sparse_K_HO', None) is not None:          # same assembled operators: reuse
            return A._sparse_K_HO
        ne, nd = self.ne, self.nd
        import scipy.sparse as sps
        base = (np.arange(ne) * nd)[:, None, None]
        ii = np.broadcast_to(base + np.arange(nd)[None, :, None], (ne, nd, nd))
        jj = np.broadcast_to(base + np.arange(nd)[None, None, :], (ne, nd, nd))
        rows = [ii.reshape(-1)]; cols = [jj.reshape(-1)]; vals = [A.K.reshape(-1)]
        for f in range(sp.nf):
            gi = (np.arange(ne) * nd)[:, None] + sp.bd[None, :, f]    # [NE, nfd] own globals
            nb = self.nbr[:, f, :]
            r = np.broadcast_to(gi[:, :, None], (ne, sp.nfd, sp.nfd))
            c_own = np.broadcast_to(gi[:, None, :], (ne, sp.nfd, sp.nfd))
            rows.append(r.reshape(-1)); cols.append(c_own.reshape(-1))
            vals.append((-A.bdrInt[:, f]).reshape(-1))
            has = nb[:, 0] >= 0
            c_n = np.broadcast_to(nb[:, None, :], (ne, sp.nfd, sp.nfd))
            rows.append(r[has].reshape(-1)); cols.append(c_n[has].reshape(-1))
            vals.append(A.bdrInt[has, f].reshape(-1))
        rows = np.concatenate(rows); cols = np.concatenate(cols)
        Ksp = sps.coo_matrix((np.concatenate(vals), (rows, cols)), shape=(self.N, self.N)).tocsr()
        Ksp.sum_duplicates()
        # structural pattern of the assembled form (every in-element pair and every face
        # coupling, whatever the values: MFEM keeps exact zeros in the CSR pattern); take the
        # upper triangle
        P = sps.coo_matrix((np.ones(rows.size), (rows, cols)), shape=(self.N, self.N)).tocsr()
        P = (P + P.T).tocoo()
        mask = P.col > P.row
        I = P.row[mask]; J = P.col[mask]
        kij = np.asarray(Ksp[I, J]).reshape(-1)
        kji = np.asarray(Ksp[J, I]).reshape(-1)
        same = (I // nd) == (J // nd)
        Mij = np.zeros(I.size)
        Mij[same] = A.M[I[same] // nd, I[same] % nd, J[same] % nd]
        A._sparse_K_HO = (I, J, kij, kji, same, Mij)
        return A._sparse_K_HO

    def _flux_matrix(self, u, du_ho, dt):
        """FluxBasedFCT::ComputeFluxMatrix (remhos_fct.cpp:295-341) on the upper-triangular coupling
        list: f_ij = dt d_ij (u_i - u_j) + dt M_ij (du_i - du_j)"""
        I, J, kij, kji, same, Mij = self.build_sparse_K_HO()
        uf = u.reshape(-1); dho = du_ho.reshape(-1)
        dij = np.maximum(np.maximum(0.0, -kij), -kji)
        flux = dt * dij * (uf[I] - uf[J])
        flux = flux + np.where(same, Mij * dt * (dho[I] - dho[J]), 0.0)
        return I, J, same, flux

    def _flux_iterate(self, I, J, flux, uf, mf, du_lo_fct, umn, umx, dt, iter_cnt, zero_mask=None):
        """AddFluxesAtDofs / ComputeFluxCoefficients / UpdateSolutionAndFlux (remhos_fct.cpp:344-446)"""
        du = du_lo_fct.copy()
        for _ in range(iter_cnt):
            gp = np.zeros(self.N); gm = np.zeros(self.N)
            pos = flux >= 0.0
            np.add.at(gp, I[pos], flux[pos]); np.add.at(gm, J[pos], -flux[pos])
            np.add.at(gm, I[~pos], flux[~pos]); np.add.at(gp, J[~pos], -flux[~pos])
            u_lo = uf + dt * du_lo_fct
            max_pos = np.maximum((umx - u_lo) * mf, 0.0)
            min_neg = np.minimum((umn - u_lo) * mf, 0.0)
            with np.errstate(divide='ignore', invalid='ignore'):
                cp = np.where(gp > max_pos, max_pos / gp, 1.0)
                cn = np.where(gm < min_neg, min_neg / gm, 1.0)
            a = np.where(pos, np.minimum(cp[I], cn[J]), np.minimum(cn[I], cp[J]))
            fa = flux * a
            du = du_lo_fct.copy()
            np.add.at(du, I, fa / mf[I] / dt)
            np.add.at(du, J, -fa / mf[J] / dt)
            flux = flux - fa
            if zero_mask is not None:
                du = np.where(zero_mask, 0.0, du)                     # ZeroOutEmptyDofs
            du_lo_fct = du.copy()
        return du

    def fct_flux_based(self, u, m, du_ho, du_lo, umin, umax, dt, iter_cnt=1):
        I, J, same, flux = self._flux_matrix(u, du_ho, dt)
        du = self._flux_iterate(I, J, flux, u.reshape(-1), m.reshape(-1), du_lo.reshape(-1).copy(),
                                umin.reshape(-1), umax.reshape(-1), dt, iter_cnt)
        return du.reshape(u.shape)

    def fct_flux_based_product(self, us, m, d_us_ho, d_us_lo, s_min, s_max, u_new, act_el, act_dof, dt,
                               iter_cnt=1):
        """FluxBasedFCT::CalcFCTProduct (remhos_fct.cpp:183-294): the flux matrix of (us, d_us_HO),
        plus the element-local fluxes that turn the LO product into the compatible one, limited
        against the scaled bounds"""
        nd = self.nd
        I, J, same, flux = self._flux_matrix(us, d_us_ho, dt)
        d_lo_c, s_min, s_max = self.compatible_lo_product(us, m, d_us_ho, s_min, s_max, u_new, act_el,
                                                          act_dof, dt)
        us_min, us_max = self.scale_product_bounds(s_min, s_max, u_new, act_el, act_dof)
        flux_el = m * dt * (d_us_lo - d_lo_c)
        beta = m * u_new
        with np.errstate(divide='ignore', invalid='ignore'):
            beta = beta / beta.sum(axis=1)[:, None]
        e = I // nd
        add = same & act_el[e]
        i_loc, j_loc = I % nd, J % nd
        fij = beta[e, j_loc] * flux_el[e, i_loc] - beta[e, i_loc] * flux_el[e, j_loc]
        flux = flux + np.where(add, fij, 0.0)
        zero = (~act_el[:, None] & ~act_dof).reshape(-1)
        du = self._flux_iterate(I, J, flux, us.reshape(-1), m.reshape(-1), d_lo_c.reshape(-1).copy(),
                                us_min.reshape(-1), us_max.reshape(-1), dt, iter_cnt, zero_mask=zero)
        return du.reshape(us.shape)
