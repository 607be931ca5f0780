"""GPU parity tests of the matrix-based solver entry points (DiscreteUpwind, ResidualDistribution,
FluxBasedFCT), of the general stage operator rmh_mult and of rmh_ode_step (RK1/2/3/4/6), against
the CPU oracle -- which is itself pinned on the reference's `-lo 1 -fct 1` and `-lo 3 -fct 2` rows
of autotest/out_baseline.dat (tests/test_oracle_golden.py).

Tolerances: single-kernel outputs 1e-11 of the field's max norm; whole runs 1e-12 (L1, Linf, mass).
"""
import numpy as np
import pytest

from helpers import oracle_run, ctx_from_oracle, rel_err, subcell_setup_from_oracle

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

TOL = 1e-11

CASES = [
    ('periodic-square.mesh', dict(problem=5, rs_levels=2, order=2)),
    ('periodic-square.mesh', dict(problem=1, rs_levels=1, order=3)),
    ('periodic-hexagon.mesh', dict(problem=0, rs_levels=1, order=2)),
    ('inline-quad.mesh', dict(problem=14, rs_levels=1, order=2, dt=0.002, t_final=0.75)),
    ('inline-quad.mesh', dict(problem=4, rs_levels=1, order=1)),
    ('periodic-cube.mesh', dict(problem=0, rs_levels=1, order=2)),
    ('periodic-cube.mesh', dict(problem=1, rs_levels=0, order=3)),
    ('cube01_hex.mesh', dict(problem=10, rs_levels=1, order=2, dt=0.02, t_final=0.7)),
    ('cube01_hex.mesh', dict(problem=1, rs_levels=1, order=1)),
]
IDS = ['%s-p%d-o%d-rs%d' % (m.split('.')[0], o['problem'], o['order'], o['rs_levels'])
       for m, o in CASES]


def dev(a):
    return torch.tensor(np.ascontiguousarray(a, dtype=np.float64).reshape(-1), device='cuda')


def host(t, shape):
    return t.cpu().numpy().reshape(shape)


def empty(ctx):
    return torch.empty(ctx.ndofs, dtype=torch.float64, device='cuda')


@pytest.fixture(scope='module', params=list(range(len(CASES))), ids=IDS)
def setup(request):
    mesh, opt = CASES[request.param]
    run = oracle_run(mesh, ho_type=3, lo_type=1, fct_type=1, **opt)
    ctx = ctx_from_oracle(run)
    ctx.fa_setup()
    rng = np.random.default_rng(20260102)
    u = run.u + 0.05 * rng.standard_normal(run.u.shape)
    yield run, ctx, u
    ctx.close()


def at_time(run, ctx, t):
    if run.exec_mode == 1:
        run.disc.assemble(t)
        ctx.set_time(t)


def test_discrete_upwind(setup):
    run, ctx, u = setup
    at_time(run, ctx, 0.3)
    ref = run.disc.lo_discrete_upwind(u)
    out = empty(ctx)
    ctx.lo_discrete_upwind(dev(u), out)
    assert rel_err(host(out, u.shape), ref) < TOL


def test_discrete_upwind_preconditioned(setup):
    run, ctx, u = setup
    at_time(run, ctx, 0.3)
    ref = run.disc.lo_discrete_upwind(u, prec=True)
    out = empty(ctx)
    ctx.lo_discrete_upwind_prec(dev(u), out)
    assert rel_err(host(out, u.shape), ref) < 1e-10


def test_neumann_ho(setup):
    run, ctx, u = setup
    at_time(run, ctx, 0.3)
    ref = run.disc.ho_neumann(u)
    out = empty(ctx)
    ctx.ho_neumann(dev(u), out)
    assert rel_err(host(out, u.shape), ref) < TOL


def test_residual_distribution(setup):
    run, ctx, u = setup
    at_time(run, ctx, 0.3)
    ref = run.disc.lo_residual_distribution(u)
    out = empty(ctx)
    ctx.lo_res_dist(dev(u), out)
    assert rel_err(host(out, u.shape), ref) < TOL


def test_residual_distribution_subcell(setup):
    run, ctx, u = setup
    if run.space.p < 2:
        import remhos_b200 as rb
        with pytest.raises(rb.RmhError):
            subcell_setup_from_oracle(run, ctx)          # remhos.cpp:613-616
        return
    subcell_setup_from_oracle(run, ctx)
    at_time(run, ctx, 0.3)
    run._t = 0.3 if run.exec_mode == 1 else 0.0
    run.subcell_weights = None
    ref = run.disc.lo_residual_distribution(u, run.get_subcell_weights())
    out = empty(ctx)
    ctx.lo_res_dist_subcell(dev(u), out)
    assert rel_err(host(out, u.shape), ref) < TOL


def test_lo_solutions_conserve_mass(setup):
    """sum_i m_i du_i of DU and RD equals the net lumped face flux; on periodic meshes with a
    divergence-free velocity it vanishes."""
    run, ctx, u = setup
    if 'periodic' not in run.opt.mesh_file:
        pytest.skip('needs a periodic mesh')
    ml = run.disc.cur.ml
    for fn in (ctx.lo_discrete_upwind, ctx.lo_res_dist):
        out = empty(ctx)
        fn(dev(u), out)
        du = host(out, u.shape)
        assert abs((ml * du).sum()) < 1e-12 * np.abs(ml * du).sum()


@pytest.mark.parametrize('lo', [1, 3])
def test_flux_based_fct(setup, lo):
    run, ctx, u = setup
    at_time(run, ctx, 0.3)
    d = run.disc
    dt = 0.01
    du_ho = d.ho_local_inverse(u)
    du_lo = d.lo_discrete_upwind(u) if lo == 1 else d.lo_residual_distribution(u)
    umin, umax = d.bounds(u, 0)
    ref = d.fct_flux_based(u, d.cur.ml, du_ho, du_lo, umin, umax, dt)
    out = empty(ctx)
    ctx.fct_flux_based(dt, dev(u), dev(d.cur.ml), dev(du_ho), dev(du_lo), dev(umin), dev(umax), out)
    g = host(out, u.shape)
    assert rel_err(g, ref) < TOL
    # antisymmetric fluxes: the correction carries no net mass
    res = (d.cur.ml * (g - du_lo)).sum()
    assert abs(res) < 1e-12 * np.abs(d.cur.ml * (du_ho - du_lo)).sum()
    # Zalesak limiter keeps u + dt du inside the bounds whenever the LO update is inside
    un_lo = u + dt * du_lo
    ok = (un_lo >= umin - 1e-12) & (un_lo <= umax + 1e-12)
    un = u + dt * g
    assert (un[ok] >= umin[ok] - 1e-11).all() and (un[ok] <= umax[ok] + 1e-11).all()


COMBOS = [(1, 2, 2), (1, 1, 1), (3, 2, 2), (1, 0, 0), (3, 1, 2), (3, 1, 1), (3, 3, 2), (3, 3, 1), (3, 5, 2), (3, 5, 1), (3, 0, 0), (0, 1, 0),
          (0, 3, 0), (3, 5, 0)]


@pytest.mark.parametrize('ho,lo,fct', COMBOS)
def test_mult_matches_oracle(setup, ho, lo, fct):
    run, ctx, u = setup
    o = run.opt
    o.ho_type, o.lo_type, o.fct_type = ho, lo, fct
    t = 0.3 if run.exec_mode == 1 else 0.0
    dt = 0.01
    ref = run.mult(np.clip(u, 0.0, None), t, dt)
    k = empty(ctx)
    ctx.mult(ho, lo, fct, t, dt, dev(np.clip(u, 0.0, None)), k)
    tol = TOL if run.space.p <= 3 else 1e-9
    assert rel_err(host(k, u.shape), ref) < tol


def test_mult_rejects_unsupported(setup):
    import remhos_b200 as rb
    run, ctx, u = setup
    k = empty(ctx)
    for combo in [(2, 1, 2), (3, 6, 2), (3, 1, 5), (3, 0, 2), (0, 5, 0), (0, 0, 0)]:
        with pytest.raises(rb.RmhError):
            ctx.mult(*combo, 0.0, 0.01, dev(u), k)


RUNS = [
    # configs[0] of BASELINE.json: 2D periodic-square, order 2, -rs 3, RK3, DU + ClipScale
    ('periodic-square.mesh', dict(problem=5, rs_levels=3, order=2, dt=0.004, t_final=0.8,
                                  ho_type=3, lo_type=1, fct_type=2, ode_solver=3), 12),
    ('periodic-square.mesh', dict(problem=5, rs_levels=2, order=2, dt=0.004, t_final=0.8,
                                  ho_type=3, lo_type=1, fct_type=1, ode_solver=4), 8),
    ('periodic-cube.mesh', dict(problem=0, rs_levels=1, order=2, dt=0.015, t_final=2.0,
                                ho_type=3, lo_type=1, fct_type=1, ode_solver=3), 6),
    ('periodic-cube.mesh', dict(problem=0, rs_levels=1, order=2, dt=0.015, t_final=2.0,
                                ho_type=3, lo_type=3, fct_type=2, ode_solver=6), 1),
    ('periodic-cube.mesh', dict(problem=0, rs_levels=1, order=2, dt=0.015, t_final=2.0,
                                ho_type=3, lo_type=0, fct_type=0, ode_solver=6), 5),
    ('periodic-cube.mesh', dict(problem=0, rs_levels=1, order=2, dt=0.015, t_final=2.0,
                                ho_type=3, lo_type=3, fct_type=2, ode_solver=4), 4),
    ('inline-quad.mesh', dict(problem=14, rs_levels=1, order=2, dt=0.0015, t_final=0.75,
                              ho_type=3, lo_type=1, fct_type=1, ode_solver=3), 10),
    ('inline-quad.mesh', dict(problem=14, rs_levels=1, order=2, dt=0.0015, t_final=0.75,
                              ho_type=3, lo_type=3, fct_type=2, ode_solver=2), 10),
    ('cube01_hex.mesh', dict(problem=10, rs_levels=1, order=2, dt=0.02, t_final=0.7,
                             ho_type=3, lo_type=5, fct_type=2, ode_solver=4), 5),
    ('periodic-square.mesh', dict(problem=5, rs_levels=2, order=3, dt=0.004, t_final=0.8,
                                  ho_type=3, lo_type=5, fct_type=2, ode_solver=6), 1),
    ('periodic-square.mesh', dict(problem=0, rs_levels=2, order=2, dt=0.004, t_final=0.8,
                                  ho_type=3, lo_type=0, fct_type=0, ode_solver=1), 10),
    # subcell residual distribution (-lo 4), transport and remap
    ('periodic-hexagon.mesh', dict(problem=0, rs_levels=1, order=3, dt=0.005, t_final=2.5,
                                   ho_type=3, lo_type=4, fct_type=2, ode_solver=3), 8),
    ('periodic-cube.mesh', dict(problem=0, rs_levels=1, order=2, dt=0.015, t_final=2.0,
                                ho_type=3, lo_type=4, fct_type=2, ode_solver=3), 5),
    ('inline-quad.mesh', dict(problem=14, rs_levels=1, order=3, dt=0.0015, t_final=0.75,
                              ho_type=3, lo_type=4, fct_type=2, ode_solver=3), 8),
    ('cube01_hex.mesh', dict(problem=10, rs_levels=1, order=2, dt=0.02, t_final=0.7,
                             ho_type=3, lo_type=4, fct_type=1, ode_solver=2), 4),
    # IDP Runge-Kutta solvers (-s 11/12/13/14/16, remhos_solvers.cpp)
    ('periodic-square.mesh', dict(problem=5, rs_levels=2, order=2, dt=0.004, t_final=0.8,
                                  ho_type=3, lo_type=1, fct_type=2, ode_solver=11), 6),
    ('periodic-square.mesh', dict(problem=5, rs_levels=2, order=2, dt=0.004, t_final=0.8,
                                  ho_type=3, lo_type=5, fct_type=2, ode_solver=12), 6),
    ('periodic-cube.mesh', dict(problem=0, rs_levels=1, order=2, dt=0.015, t_final=2.0,
                                ho_type=3, lo_type=5, fct_type=2, ode_solver=13), 4),
    ('inline-quad.mesh', dict(problem=14, rs_levels=1, order=2, dt=0.0015, t_final=0.75,
                              ho_type=3, lo_type=3, fct_type=2, ode_solver=13), 6),
    ('periodic-square.mesh', dict(problem=5, rs_levels=2, order=3, dt=0.004, t_final=0.8,
                                  ho_type=3, lo_type=1, fct_type=1, ode_solver=14), 4),
    ('periodic-cube.mesh', dict(problem=1, rs_levels=1, order=2, dt=0.015, t_final=2.0,
                                ho_type=3, lo_type=5, fct_type=2, ode_solver=16), 3),
]


@pytest.mark.parametrize('mesh,opt,steps', RUNS)
def test_ode_steps_match_oracle(mesh, opt, steps):
    run = oracle_run(mesh, max_steps=steps, **opt)
    ctx = ctx_from_oracle(run)
    if opt['lo_type'] == 1 or opt['fct_type'] == 1:
        ctx.fa_setup()
    if opt['lo_type'] == 4:
        subcell_setup_from_oracle(run, ctx)
    u = dev(run.u)
    t, dt = 0.0, run.dt
    for _ in range(steps):
        t = ctx.ode_step(opt['ode_solver'], opt['ho_type'], opt['lo_type'], opt['fct_type'], t,
                         min(dt, run.t_final - t), u)
    run.run()
    ug = u.cpu().numpy().reshape(run.u.shape)
    ml = run.disc.cur.ml if run.exec_mode == 1 else run.masses0
    l1 = float((ml * np.abs(ug - run.u)).sum() / (ml * np.abs(run.u)).sum())
    linf = float(np.abs(ug - run.u).max() / np.abs(run.u).max())
    # RK6's tableau has entries up to 208 and weights of -176/+172: with a limiter in F (Lipschitz
    # constant ~ 1/dt) round-off differences between two implementations are amplified by several
    # orders per step, so limited RK6 runs are compared over one step at 1e-9; the unlimited
    # (linear) RK6 run only carries the cancellation in the weights (1e-11)
    tol = 1e-12
    if opt['ode_solver'] == 6:
        tol = 1e-9 if opt['fct_type'] else 1e-11
    assert l1 < tol and linf < tol, (l1, linf)
    ctx.set_time(t)
    m = torch.empty(ctx.ndofs, dtype=torch.float64, device='cuda')
    ctx.lumped_mass(m)
    mass = ctx.reduce(0, u, m)
    assert abs(mass - run.final_mass) < 1e-12 * abs(run.final_mass)
    ctx.close()


def test_ode_step_unknown_solver_returns_3():
    import remhos_b200 as rb
    run = oracle_run('periodic-square.mesh', problem=0, rs_levels=0, order=1, ho_type=3)
    ctx = ctx_from_oracle(run)
    with pytest.raises(rb.RmhError):
        ctx.ode_step(5, 3, 0, 0, 0.0, 0.01, dev(run.u))
    ctx.close()


# ---- ElementFCTProjection (-fct 4) and -dtc 1; the oracle is pinned on out_baseline.dat:207-210
# (tests/test_oracle_mono_golden.py)
def test_fct_project_matches_oracle(setup):
    run, ctx, u = setup
    at_time(run, ctx, 0.0)
    d = run.disc
    dt = run.dt
    du_ho = d.ho_local_inverse(u)
    du_lo = d.lo_mass_based_avg(u, du_ho, dt)
    umin, umax = d.bounds(u, run.opt.bounds_type)
    ref = d.fct_project(u, du_ho, du_lo, umin, umax, dt)
    out = empty(ctx)
    ctx.fct_project(dt, dev(u), dev(du_ho), dev(du_lo), dev(umin), dev(umax), out)
    assert rel_err(host(out, u.shape), ref) < TOL
    # conservative: sum_i M_L,i (du_i - du_lo_i) = 0 per element
    ML = d.cur.M.sum(axis=2)
    corr = (ML * (host(out, u.shape) - du_lo)).sum(axis=1)
    assert np.abs(corr).max() < 1e-12 * np.abs(ML * du_lo).sum(axis=1).max()


def test_dt_ratio_matches_oracle(setup):
    run, ctx, u = setup
    if run.exec_mode == 1:
        pytest.skip('transport only')
    run.opt.fct_type = 4; run.opt.lo_type = 5; run.opt.dt_control = 1
    run.dt_ratio = np.inf; run.dt_est = np.inf
    dt = run.dt
    try:
        ref = run.mult(u, 0.0, dt)
        ctx.dt_control(1)
        ctx.dt_ratio(reset=True)
        k = empty(ctx)
        ctx.mult(3, 5, 4, 0.0, dt, dev(u), k)
        assert rel_err(host(k, u.shape), ref) < TOL
        r = ctx.dt_ratio()
        assert abs(r - run.dt_ratio) < 1e-9 * abs(run.dt_ratio)
    finally:
        ctx.dt_control(0)
        run.opt.fct_type = 1; run.opt.lo_type = 1; run.opt.dt_control = 0
