"""TEST INFRASTRUCTURE ONLY -- CPU oracle (numpy) for the Remhos RK-stage hot path.

End-to-end restatement of remhos() (remhos.cpp:210-1523) for the solver combinations on
the hot path, used to pin the oracle against the reference's own known answers
(autotest/out_baseline.dat, remhos_tests.cpp:38-107, README.md:221-259).

ODE solvers restate MFEM's ForwardEuler / RK2Solver(1.0) / RK3SSPSolver / RK4Solver
(remhos.cpp:488-492; formulas in SURVEY.md 3.2 / Appendix C-10).
"""
from dataclasses import dataclass
import numpy as np
from . import fe, mesh as meshmod, dg, problems
from .solvers import Discretization


# MFEM RK6Solver tables (Verner's 8-stage, 6th-order method; remhos.cpp:492).  MFEM is not in the
# reference tree, so the coefficients are restated and checked against the 6th-order conditions
# in tests/test_oracle_golden.py (test_rk6_tableau).
RK6_A = [
    .6e-1,
    .1923996296296296296296296296296296296296e-1, .7669337037037037037037037037037037037037e-1,
    .35975e-1, 0., .107925,
    1.318683415233148260919747276431735612861, 0., -5.042058063628562225427761634715637693344,
    4.220674648395413964508014358283902080483,
    -41.87259166432751461803757780644346812905, 0., 159.4325621631374917700365669070346830453,
    -122.1192135650100309202516203389242140663, 5.531743066200053768252631238332999150076,
    -54.43015693531650433250642051294142461271, 0., 207.0672513650184644273657173866509835987,
    -158.6108137845899991828742424365058599469, 6.991816585950242321992597280791793907096,
    -.1859723106220323397765171799549294623692e-1,
    -54.66374178728197680241215648050386959351, 0., 207.9528062553893734515824816699834244238,
    -159.2889574744995071508959805871426654216, 7.018743740796944434698170760964252490817,
    -.1833878590504572306472782005141738268361e-1, -.5119484997882099077875432497245168395840e-3]
RK6_B = [
    .3438957868357036009278820124728322386520e-1, 0., 0.,
    .2582624555633503404659558098586120858767, .4209371189673537150642551514069801967032,
    4.405396469669310170148836816197095664891, -176.4831190242986576151740942499002125029,
    172.3641334014150730294022582711902413315]
# RKIDPSolver tables (remhos_solvers.cpp:252-279)
IDP_TABLES = {
    12: ([.5], [0., 1.], [.5]),
    13: ([1. / 3., 0., 2. / 3.], [.25, 0., .75], [1. / 3., 2. / 3.]),
    14: ([1. / 3., -1. / 3., 1., 1., -1., 1.], [1. / 8., 3. / 8., 3. / 8., 1. / 8.],
         [1. / 3., 2. / 3., 1.]),
    16: ([.25, 1. / 8., 1. / 8., 0., -.5, 1., 3. / 16., 0., 0., 9. / 16., -3. / 7., 2. / 7.,
          12. / 7., -12. / 7., 8. / 7.], [7. / 90., 0., 32. / 90., 12. / 90., 32. / 90., 7. / 90.],
         [.25, .25, .5, .75, 1.]),
}
RK6_C = [.6e-1, .9593333333333333333333333333333333333333e-1, .1439, .4973, .9725, .9995, 1.]


@dataclass
class Options:                       # defaults: remhos.cpp:216-244
    mesh_file: str = 'default'
    problem: int = 0
    rs_levels: int = 2
    order: int = 3
    mesh_order: int = 2
    ode_solver: int = 3
    ho_type: int = 3
    lo_type: int = 0
    fct_type: int = 0
    mono_type: int = 0            # 1: MonoRDSolver, 2: with subcells (remhos.cpp:285-289)
    si_type: int = 0              # smoothness indicator (remhos.cpp:302; order 1 only here)
    dt_control: int = 0           # -dtc 1: LO bounds error time step control (remhos.cpp:312-316)
    product_sync: bool = False    # -ps: remap the product field us along with u (remhos.cpp:886-903)
    bounds_type: int = 0
    t_final: float = 4.0
    dt: float = 0.005
    max_steps: int = -1
    verify_bounds: bool = False


class Run:
    """Set-up phase of remhos() (remhos.cpp:438-1121)."""

    def __init__(self, opt, mesh=None):
        self.opt = opt
        self.exec_mode = 0 if opt.problem < 10 else 1          # :438-440
        m = meshmod.read_mesh(opt.mesh_file) if mesh is None else mesh
        for _ in range(opt.rs_levels):
            m = meshmod.refine_uniform(m)
        self.bb_min, self.bb_max = meshmod.bounding_box(m)     # :457
        m = meshmod.set_curvature(m, opt.mesh_order)           # :513
        self.mesh = m
        self.topo = meshmod.Topology(m)
        dim = m.dim
        self.space = sp = dg.Space(dim, opt.order, opt.mesh_order)
        X0 = m.X.copy()
        prob = opt.problem

        def vel(pts):
            shp = pts.shape
            return problems.velocity(prob, pts.reshape(-1, dim), self.bb_min,
                                     self.bb_max).reshape(shp)
        self.vel = vel
        dt = opt.dt
        if dt < 0.0:                                           # CFL estimate, :538-553
            gll = sp.gll
            c = np.array([0.5])
            Lc = fe.tensor_basis([fe.lagrange(gll, c)] * dim)
            dLc = [fe.tensor_basis([fe.lagrange_deriv(gll, c) if a == b else fe.lagrange(gll, c)
                                    for b in range(dim)]) for a in range(dim)]
            J = sp.jacobians(X0, dLc)
            det, _ = sp.det_adj(J)
            length = np.abs(det[:, 0]) ** (1.0 / dim)          # Mesh::GetElementSize
            xc = np.einsum('qn,eni->eqi', Lc, X0)
            vc = vel(xc)[:, 0, :]
            speed = np.sqrt((vc * vc).sum(axis=1) + 1e-14)
            dt = float(np.min(0.25 * length / speed))
        self.dt = dt
        Vnodes = None
        t_final = opt.t_final
        if self.exec_mode == 1:                                # mesh velocity, :562-584
            x = X0.copy()
            v = vel(x)
            t = 0.0
            while t < t_final:
                t += dt
                x = x + min(dt, t_final - t) * v
                v = vel(x)
            Vnodes = x - X0
            t_final = 1.0                                      # :1128-1134
        self.t_final = t_final
        infl = problems.inflow(prob, sp.dof_points(X0).reshape(-1, dim)).reshape(m.ne, sp.nd)
        if prob == 7:
            # "Convergence test: use high order projection" (remhos.cpp:628-635): interpolate the
            # inflow function at the tensor Gauss-Legendre points (L2_FECollection's nodal basis),
            # then take the Bernstein coefficients of that very polynomial (ProjectGridFunction
            # between two bases of the same space) -- not the nodal samples at the lattice points
            xg, _ = fe.gauss_legendre_01(sp.p + 1)
            Lg = fe.tensor_basis([fe.lagrange(sp.gll, xg)] * dim)
            pts = np.einsum('qn,eni->eqi', Lg, X0)
            vals = problems.inflow(prob, pts.reshape(-1, dim)).reshape(m.ne, sp.nd)
            Vinv = np.linalg.inv(fe.bernstein(sp.p, xg))               # [p+1 coeffs, p+1 points]
            T = fe.tensor_basis([Vinv] * dim)
            infl = vals @ T.T
        self.disc = Discretization(sp, self.topo, X0, self.exec_mode, vel_fun=vel,
                                   Vnodes=Vnodes, inflow_vals=infl)
        self.disc.assemble(0.0)
        pts = sp.dof_points(X0)                                # ProjectCoefficient, :883
        self.u = problems.u0(prob, pts.reshape(-1, dim), self.bb_min,
                             self.bb_max).reshape(m.ne, sp.nd)
        self.u0_min, self.u0_max = float(self.u.min()), float(self.u.max())
        self.mass0 = float((self.disc.cur.ml * self.u).sum())
        self.masses0 = self.disc.cur.ml.copy()
        self.us = None
        if opt.product_sync:                                   # remhos.cpp:886-903
            assert self.exec_mode == 1, 'Products are processed only in remap mode.'
            el, _ = Discretization.bool_indicators(self.u)
            P = pts.reshape(-1, dim)
            s0 = (2.0 + np.sin(2 * np.pi * P[:, 0]) * np.sin(2 * np.pi * P[:, 1])).reshape(m.ne, sp.nd)
            self.us = self.u * np.where(el[:, None], s0, 0.0)
            self.mass0_us = float((self.disc.cur.ml * self.us).sum())
        self.subcell_weights = None
        self.mono_scale = None
        if opt.mono_type:
            self.mono_scale = self._mono_scale()
        self.si = None
        if opt.si_type:
            from .smoothness import SmoothnessIndicator
            self.si = SmoothnessIndicator(self, opt.si_type)

    # ---------------------------------------------------------------- stage operator
    def _mono_scale(self):
        """MonoRDSolver constructor (remhos_mono.cpp:40-57): scale_e = vmax / (2 sqrt(dim) h_e / p),
        vmax over the Gauss-Legendre rule of order OrderW + 2p + 2 max(OrderGrad, 0) [MFEM-K: for
        tensor elements OrderW = dim*mo - 1, OrderGrad = mo*(dim-1) + p - 1], h_e = GetElementSize."""
        sp, m = self.space, self.mesh
        dim, p, mo = sp.dim, sp.p, sp.g
        q_ord = (dim * mo - 1) + 2 * p + 2 * max(mo * (dim - 1) + p - 1, 0)
        n = q_ord // 2 + 1
        xq, _ = np.polynomial.legendre.leggauss(n)
        xq = 0.5 * (xq + 1.0)
        L = fe.tensor_basis([fe.lagrange(sp.gll, xq)] * dim)
        pts = np.einsum('qn,eni->eqi', L, m.X)
        v = self.vel(pts)
        vmax = np.sqrt((v * v).sum(axis=2)).max(axis=1)
        c = np.array([0.5])
        dLc = [fe.tensor_basis([fe.lagrange_deriv(sp.gll, c) if a == b else fe.lagrange(sp.gll, c)
                                for b in range(dim)]) for a in range(dim)]
        det, _ = sp.det_adj(sp.jacobians(m.X, dLc))
        h = np.abs(det[:, 0]) ** (1.0 / dim)
        return vmax / (2.0 * (np.sqrt(dim) * h / p))

    def mult(self, u, t, dt):
        """LimitedTimeDependentOperator::Mult = MultUnlimited + LimitMult
        (remhos_solvers.hpp:46-50; remhos.cpp:1596-1739, 1798-1916)."""
        if u.ndim == 3:                                        # (u, us): product remap
            return self.limit_mult(u, self.mult_unlimited(u, t, dt), dt)
        o, d = self.opt, self.disc
        self._t = t
        if self.exec_mode == 1:
            d.assemble(t)
        A = d.cur
        if o.mono_type:                                        # remhos.cpp:1687
            sw = self.get_subcell_weights() if o.mono_type == 2 else None
            mass_lim = o.problem not in (6, 7)                 # remhos.cpp:999
            si_tmp = self.si.dof_values(u) if self.si is not None else None
            return d.mono_rd(u, o.bounds_type, self.mono_scale, sw, mass_lim, si_tmp)
        if o.fct_type:
            du_ho = self.calc_ho(u)
            du_lo = self.calc_lo(u, du_ho, dt)
            umin, umax = d.bounds(u, o.bounds_type)
            if o.verify_bounds:
                self.check(u, dt, du_lo, umin, umax, 'LO')
            si_tmp = self.si.dof_values(u) if (self.si is not None and o.fct_type in (2, 3)) else None
            if o.fct_type == 2:
                bmin, bmax = umin, umax
                if si_tmp is not None:      # the bound relaxation the reference's kernel intends (remhos_fct.cpp:498-504)
                    bmin, bmax = d.si_update_bounds(u + dt * du_ho, si_tmp, umin, umax)
                du = d.fct_clip_scale(u, A.ml, du_ho, du_lo, bmin, bmax, dt)
            elif o.fct_type == 1:
                du = d.fct_flux_based(u, A.ml, du_ho, du_lo, umin, umax, dt)
            elif o.fct_type == 3:
                du = d.fct_nonlinear_penalty(u, A.ml, du_ho, du_lo, umin, umax, dt, self.penalty_eps(), si_tmp)
            elif o.fct_type == 4:
                du = d.fct_project(u, du_ho, du_lo, umin, umax, dt)
            else:
                raise NotImplementedError('fct type %d' % o.fct_type)
            if o.dt_control:                                   # remhos.cpp:1839-1842
                self.update_dt_estimate(u, du_lo, umin, umax, dt)
            if o.verify_bounds:
                self.check(u, dt, du, umin, umax, 'FCT')
            return du
        if o.lo_type:
            return self.calc_lo(u, None, dt)
        return self.calc_ho(u)

    def penalty_eps(self):
        """GetElementSize(0, 0) / GetOrder(0) of NonlinearPenaltySolver::CorrectFlux (remhos_fct.cpp:961):
        size of element 0 = |det J(centre)|^(1/dim) at the current mesh position"""
        sp = self.space
        dim = sp.dim
        c = np.array([0.5])
        dLc = [fe.tensor_basis([fe.lagrange_deriv(sp.gll, c) if a == b else fe.lagrange(sp.gll, c)
                                for b in range(dim)]) for a in range(dim)]
        det, _ = sp.det_adj(sp.jacobians(self.disc.cur.X[:1], dLc))
        return float(np.abs(det[0, 0]) ** (1.0 / dim)) / sp.p

    def calc_ho(self, u):
        if self.opt.ho_type == 1:
            return self.disc.ho_neumann(u)
        return self.disc.ho_local_inverse(u)                  # -ho 3, and -ho 2 (CG to 1e-12)

    def update_dt_estimate(self, x, dx, x_min, x_max, dt_cur):
        """AdvectionOperator::UpdateTimeStepEstimate (remhos.cpp:1968-1998)"""
        eps = 1e-12
        with np.errstate(divide='ignore', invalid='ignore'):
            up = np.where(dx > eps, (x_max - x) / dx, np.inf)
            dn = np.where(dx < -eps, (x_min - x) / dx, np.inf)
        dt = float(min(up.min(), dn.min()))
        self.dt_est = min(getattr(self, 'dt_est', np.inf), dt)
        self.dt_ratio = min(getattr(self, 'dt_ratio', np.inf), dt / dt_cur if dt_cur != 0.0 else 0.0)

    def mult_unlimited(self, u, t, dt):
        """AdvectionOperator::MultUnlimited (remhos.cpp:1596-1739).  A state of shape [2, NE, nd]
        is (u, us): the product field is remapped with the same operator (remhos.cpp:1714-1738)."""
        if u.ndim == 3:
            ku = self.mult_unlimited(u[0], t, dt)
            assert self.opt.fct_type, 'product remap is restated for the FCT path'
            return np.stack([ku, self.calc_ho(u[1])])
        o, d = self.opt, self.disc
        self._t = t
        if self.exec_mode == 1:
            d.assemble(t)
        if o.fct_type:
            return self.calc_ho(u)
        if o.lo_type:
            return self.calc_lo(u, None, dt)
        return self.calc_ho(u)

    def limit_mult(self, u, du_ho, dt):
        """AdvectionOperator::LimitMult (remhos.cpp:1798-1916): du_ho is the (combined) HO rate."""
        o, d = self.opt, self.disc
        if u.ndim == 3:                                        # remhos.cpp:1848-1915
            du = self.limit_mult(u[0], du_ho[0], dt)
            return np.stack([du, self.limit_product(u[0], du, u[1], du_ho[1], dt)])
        if not o.fct_type:
            return du_ho
        A = d.cur
        du_lo = self.calc_lo(u, du_ho, dt)
        umin, umax = d.bounds(u, o.bounds_type)
        si_tmp = self.si.dof_values(u) if (self.si is not None and o.fct_type in (2, 3)) else None
        if o.fct_type == 2:
            if si_tmp is not None:          # the bound relaxation the reference's kernel intends (remhos_fct.cpp:498-504)
                umin, umax = d.si_update_bounds(u + dt * du_ho, si_tmp, umin, umax)
            return d.fct_clip_scale(u, A.ml, du_ho, du_lo, umin, umax, dt)
        if o.fct_type == 3:
            return d.fct_nonlinear_penalty(u, A.ml, du_ho, du_lo, umin, umax, dt, self.penalty_eps(), si_tmp)
        if o.fct_type == 4:
            return d.fct_project(u, du_ho, du_lo, umin, umax, dt)
        return d.fct_flux_based(u, A.ml, du_ho, du_lo, umin, umax, dt)

    def limit_product(self, u, du, us, d_us_ho, dt):
        """second pass of LimitMult (remhos.cpp:1848-1915) + CalcFCTProduct of ClipScaleSolver /
        ElementFCTProjection (remhos_fct.cpp:543-563, 735-758): bounds on s = us / u from the old
        active dofs, compatible LO product, limiter on us, empty dofs zeroed"""
        o, d = self.opt, self.disc
        assert o.fct_type in (1, 2, 4), 'product remap is restated for -fct 1, 2 and 4'
        A = d.cur
        s, s_el, s_dof = d.compute_ratio(us, u)
        s_min, s_max = d.bounds(s, o.bounds_type, active_el=s_el, active_dof=s_dof)
        s_min0, s_max0 = s_min, s_max
        u_new = u + dt * du
        el_new, dof_new = d.bool_indicators(u_new)
        d_lo, s_min, s_max = d.compatible_lo_product(us, A.ml, d_us_ho, s_min, s_max, u_new, el_new,
                                                     dof_new, dt)
        us_min, us_max = d.scale_product_bounds(s_min, s_max, u_new, el_new, dof_new)
        if o.fct_type == 1:        # NeedsLOProductInput (remhos.cpp:1865-1869)
            d_us_lo = self.calc_lo(us, d_us_ho, dt)
            return d.fct_flux_based_product(us, A.ml, d_us_ho, d_us_lo, s_min0, s_max0, u_new, el_new,
                                            dof_new, dt)
        if o.fct_type == 2:
            d_us = d.fct_clip_scale(us, A.ml, d_us_ho, d_lo, us_min, us_max, dt)
        else:
            d_us = d.fct_project(us, d_us_ho, d_lo, us_min, us_max, dt)
        # ZeroOutEmptyDofs (remhos_sync.cpp:96-114)
        return np.where(~el_new[:, None] & ~dof_new, 0.0, d_us)

    @staticmethod
    def compute_mask(x):
        """AdvectionOperator::ComputeMask (remhos.cpp:1741-1796): only a product state (u, us) is
        masked; an element is on iff all of its u dofs are active; the same mask for both fields"""
        if x.ndim != 3:
            return np.ones(x.shape, dtype=bool)
        full = (x[0] > 1e-12).all(axis=1)
        return np.broadcast_to(full[None, :, None], x.shape).copy()

    def idp_step(self, u, t, dt, use_mask=False):
        """ForwardEulerIDPSolver / RKIDPSolver::Step (remhos_solvers.cpp:29-38, 40-95, 171-249; tables
        :252-279).  The driver switches the masks off (remhos.cpp:502-507); use_mask=True restates
        the masked variant (ComputeMask / UpdateMask / AddMasked, remhos_solvers.cpp:97-147)."""
        s = self.opt.ode_solver
        if s == 11:
            k = self.limit_mult(u, self.mult_unlimited(u, t, dt), dt)
            return u + dt * k
        a, b, c = IDP_TABLES[s]
        ns = len(b)
        # ConstructD
        d = np.zeros(ns * (ns + 1) // 2)
        an, ao, i_o, c_o = 0, 0, -1, 0.0           # offsets into `a` (an == -1: use b)
        row = lambda off: (b if off < 0 else a[off:])
        for i in range(ns):
            c_n = c[i] if i < ns - 1 else 1.0
            dc = c_n - c_o
            di = i * (i + 1) // 2
            for j in range(i):
                a_oj = row(ao)[j] if j <= i_o else 0.0
                m = (row(an)[j] - a_oj) / dc
                if m == 0.0:
                    d[di + j] = 0.0
                    continue
                dj = j * (j + 1) // 2
                dij = m / d[dj + j]
                for k in range(j):
                    d[di + k] -= d[dj + k] * dij
                d[di + j] = dij
            d[di + i] = row(an)[i] / dc
            c_next = c[i + 1] if i < ns - 2 else 1.0
            if c_next > c_n:
                i_o, c_o, ao = i, c_n, an
            if i < ns - 2:
                an += i + 1
            else:
                an = -1
        # Step
        x = u.copy()
        ks = [None] * ns
        c_o = 0.0
        tcur = t
        ks[0] = self.limit_mult(x, self.mult_unlimited(x, tcur, c[0] * dt), c[0] * dt)
        c_next = c[1] if ns > 2 else 1.0
        mask = None
        if c_next > c[0]:
            x = x + c[0] * dt * ks[0]
            if use_mask:
                mask = self.compute_mask(x)
            tcur = t + c[0] * dt
            c_o = c[0]
        elif use_mask:
            mask = self.compute_mask(x + c[0] * dt * ks[0])
        for i in range(1, ns):
            c_n = c[i] if i < ns - 1 else 1.0
            dct = (c_n - c_o) * dt
            di = i * (i + 1) // 2
            k = self.mult_unlimited(x, tcur, dct)
            if use_mask:
                mask = mask & self.compute_mask(x + dct * k if dct != 0.0 else x)       # UpdateMask
                k = k + np.where(mask, (d[di + i] - 1.0) * k, 0.0)                       # AddMasked
                for j in range(i):
                    k = k + np.where(mask, d[di + j] * ks[j], 0.0)
            else:
                k = k * d[di + i]
                for j in range(i):
                    k = k + d[di + j] * ks[j]
            ks[i] = self.limit_mult(x, k, dct)
            c_next = c[i + 1] if i < ns - 2 else 1.0
            if i == ns - 1 or c_next > c_n:
                tcur = t + c_n * dt
                x = x + dct * ks[i]
                c_o = c_n
        return x

    def calc_lo(self, u, du_ho, dt):
        o, d = self.opt, self.disc
        if o.lo_type == 5:
            if du_ho is None:
                du_ho = self.calc_ho(u)
            return d.lo_mass_based_avg(u, du_ho, dt)
        if o.lo_type == 1:
            return d.lo_discrete_upwind(u)
        if o.lo_type == 2:
            return d.lo_discrete_upwind(u, prec=True)
        if o.lo_type == 3:
            return d.lo_residual_distribution(u)
        if o.lo_type == 4:
            return d.lo_residual_distribution(u, self.get_subcell_weights())
        raise NotImplementedError('lo type %d' % o.lo_type)

    def get_subcell_weights(self):
        from .subcell import subcell_weights
        if self.exec_mode == 1 or self.subcell_weights is None:
            self.subcell_weights = subcell_weights(self, getattr(self, '_t', 0.0))
        return self.subcell_weights

    @staticmethod
    def check(u, dt, du, umin, umax, info, tol=1e-12):         # check_violation, :1576-1594
        un = u + dt * du
        bad = (un + tol < umin) | (un > umax + tol)
        if bad.any():
            raise RuntimeError(info + ' bounds violation')

    # ---------------------------------------------------------------- time stepping
    def step(self, u, t, dt):
        s = self.opt.ode_solver
        f = self.mult
        if s in (11, 12, 13, 14, 16):
            return self.idp_step(u, t, dt)
        if s == 1:
            k = f(u, t, dt)
            return u + dt * k
        if s == 2:                                             # RK2Solver(1.0)
            k = f(u, t, dt)
            x1 = u + 0.5 * dt * k
            x = u + dt * k
            k = f(x, t + dt, dt)
            return x1 + 0.5 * dt * k
        if s == 3:                                             # RK3SSPSolver
            k = f(u, t, dt)
            y = u + dt * k
            k = f(y, t + dt, dt)
            y = y + dt * k
            y = 0.75 * u + 0.25 * y
            k = f(y, t + dt / 2, dt)
            y = y + dt * k
            return (1. / 3) * u + (2. / 3) * y
        if s == 4:
            k1 = f(u, t, dt)
            k2 = f(u + dt / 2 * k1, t + dt / 2, dt)
            k3 = f(u + dt / 2 * k2, t + dt / 2, dt)
            k4 = f(u + dt * k3, t + dt, dt)
            return u + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        if s == 6:                                             # RK6Solver (ExplicitRKSolver)
            a, b, c = RK6_A, RK6_B, RK6_C
            ks = [f(u, t, dt)]
            for i in range(1, 8):
                ai = a[i * (i - 1) // 2:i * (i + 1) // 2]
                y = u.copy()
                for j in range(i):
                    y = y + dt * ai[j] * ks[j]
                ks.append(f(y, t + c[i - 1] * dt, dt))
            y = u.copy()
            for j in range(8):
                y = y + dt * b[j] * ks[j]
            return y
        raise NotImplementedError('ode solver %d' % s)

    def run(self, callback=None):
        """Time loop (remhos.cpp:1146-1330) and final mass / max (:1382-1436)."""
        o = self.opt
        t, dt = 0.0, self.dt
        u = self.u if self.us is None else np.stack([self.u, self.us])
        ti = 0
        done = False
        steady = o.problem in (6, 7, 8)                        # remhos.cpp:1146-1330
        res = u.copy()
        ml0 = self.disc.cur.ml
        self.residual = 0.0
        self.repeats = 0
        while not done:
            dt_real = min(dt, self.t_final - t)
            self.dt_est = np.inf; self.dt_ratio = np.inf       # ResetTimeStepRatio
            u_old = u
            u = self.step(u, t, dt_real)
            t += dt_real
            ti += 1
            if o.dt_control:                                   # remhos.cpp:1178-1197
                if self.dt_ratio < 1.0:
                    ti -= 1; t -= dt_real; u = u_old
                    dt = 0.85 * dt
                    self.repeats += 1
                    assert dt >= 1e-12, 'The time step crashed!'
                    continue
                elif self.dt_ratio > 1.25:
                    dt *= 1.02
            done = t >= self.t_final - 1e-8 * dt
            if steady:                                         # :1263-1290
                self.residual = float(np.sqrt((((ml0 * u) / dt - (ml0 * res) / dt) ** 2).sum()))
                if self.residual < 1e-12 and t >= 1.0:
                    done = True
                    u = res
                else:
                    res = u.copy()
            if ti == o.max_steps:
                done = True
            if callback:
                callback(ti, t, u)
        if self.us is not None:
            u, self.us = u[0], u[1]
        self.u = u
        self.t = t
        self.steps = ti
        if self.exec_mode == 1:
            A = self.disc.assemble(t)
            ml = A.ml
        else:
            ml = self.masses0
        self.final_mass = float((ml * u).sum())
        if self.us is not None:
            self.final_mass_us = float((ml * self.us).sum())
        self.final_max = float(u.max())
        return self.final_mass, self.final_max


def run(**kw):
    r = Run(Options(**kw))
    r.run()
    return r
