"""Product-field remap (-ps) on the device against the oracle, which reproduces all three reference
known answers for it (autotest/out_baseline.dat:187-200, tests/test_oracle_golden.py):
ComputeBoolIndicators / ComputeRatio (remhos_sync.cpp), CalcCompatibleLOProduct /
ScaleProductBounds and the three CalcFCTProduct variants (remhos_fct.cpp), the two-pass
MultUnlimited / LimitMult on the block state (remhos.cpp:1714-1738,1848-1915) and the ODE solvers
on (u, us), with and without the IDP masks.  Tolerance 1e-11 of the field's scale per evaluation
(element-local sums run in a different order), 1e-10 over runs."""
import numpy as np
import pytest

from helpers import oracle_run, ctx_from_oracle, rel_err

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu


def dev(a, dtype=torch.float64):
    return torch.tensor(np.ascontiguousarray(a).reshape(-1), device='cuda', dtype=dtype)


COMBOS = [(3, 1, 1, 1), (1, 5, 2, 12), (3, 5, 4, 13), (3, 5, 2, 11), (3, 3, 2, 2), (3, 5, 4, 14)]
IDS = ['ho%d-lo%d-fct%d-s%d' % c for c in COMBOS]


def make(ho, lo, fct, ode, steps=0, order=3, rs=2):
    run = oracle_run('inline-quad.mesh', problem=14, rs_levels=rs, order=order, dt=0.005, t_final=0.75,
                     ho_type=ho, lo_type=lo, fct_type=fct, ode_solver=ode, product_sync=True,
                     max_steps=steps if steps else -1)
    ctx = ctx_from_oracle(run)
    ctx.fa_setup()
    ctx.product_enable(True)
    return run, ctx


def evolved_state(run, n=3):
    """a few oracle steps, so that the state has partially filled elements"""
    U = np.stack([run.u, run.us])
    t = 0.0
    for _ in range(n):
        U = run.step(U, t, run.dt)
        t += run.dt
    return U, t


def test_bool_indicators_and_ratio():
    run, ctx = make(3, 5, 2, 12)
    U, _ = evolved_state(run)
    u, us = U
    ne, nd = u.shape
    d = run.disc
    s_ref, el_ref, dof_ref = d.compute_ratio(us, u)
    el = torch.zeros(ne, dtype=torch.uint8, device='cuda')
    dof = torch.zeros(ne * nd, dtype=torch.uint8, device='cuda')
    s = torch.empty(ne * nd, dtype=torch.float64, device='cuda')
    ctx.prod_compute_ratio(dev(us), dev(u), s, el, dof)
    assert np.array_equal(el.cpu().numpy().astype(bool), el_ref)
    assert np.array_equal(dof.cpu().numpy().astype(bool).reshape(ne, nd), dof_ref)
    assert 0 < el_ref.sum() < ne and (dof_ref.any(axis=1) & ~dof_ref.all(axis=1)).any(), 'state must mix full, partial and empty elements'
    assert rel_err(s.cpu().numpy().reshape(ne, nd), s_ref) < 1e-13
    ctx.prod_bool_indicators(dev(u), el, dof)
    e2, d2 = d.bool_indicators(u)
    assert np.array_equal(el.cpu().numpy().astype(bool), e2) and np.array_equal(dof.cpu().numpy().astype(bool).reshape(ne, nd), d2)
    # masked element min/max -> bounds of s
    for bt in (0, 1):
        smin_ref, smax_ref = d.bounds(s_ref, bt, active_el=el_ref, active_dof=dof_ref)
    xe_min = torch.empty(ne, dtype=torch.float64, device='cuda'); xe_max = torch.empty_like(xe_min)
    ctx.elem_min_max_masked(s, el, dof, xe_min, xe_max)
    mn = np.where(dof_ref, s_ref, np.inf).min(axis=1); mx = np.where(dof_ref, s_ref, -np.inf).max(axis=1)
    assert np.array_equal(xe_min.cpu().numpy(), mn) and np.array_equal(xe_max.cpu().numpy(), mx)
    ctx.close()


@pytest.mark.parametrize('fct', [1, 2, 4])
def test_fct_product_matches_oracle(fct):
    ho, lo = (3, 1) if fct == 1 else (3, 5)
    run, ctx = make(ho, lo, fct, 11)
    U, t = evolved_state(run)
    u, us = U
    ne, nd = u.shape
    d, dt = run.disc, run.dt
    run._t = t
    d.assemble(t); ctx.set_time(t)
    A = d.cur
    K = run.mult_unlimited(U, t, dt)
    du = run.limit_mult(u, K[0], dt)
    d_us_ho = K[1]
    s, s_el, s_dof = d.compute_ratio(us, u)
    s_min, s_max = d.bounds(s, run.opt.bounds_type, active_el=s_el, active_dof=s_dof)
    u_new = u + dt * du
    el_new, dof_new = d.bool_indicators(u_new)
    ref = run.limit_product(u, du, us, d_us_ho, dt)
    d_us_lo = run.calc_lo(us, d_us_ho, dt) if fct == 1 else None
    out = torch.empty(ne * nd, dtype=torch.float64, device='cuda')
    smin_d, smax_d = dev(s_min), dev(s_max)
    ctx.fct_product(fct, dt, dev(us), dev(A.ml), dev(d_us_ho), dev(d_us_lo) if fct == 1 else None, smin_d, smax_d,
                    dev(u_new), dev(el_new, torch.uint8), dev(dof_new, torch.uint8), out)
    scale = max(np.abs(ref).max(), np.abs(d_us_ho).max())
    assert np.abs(out.cpu().numpy().reshape(ne, nd) - ref).max() < 1e-11 * scale
    # the in-place adjustment of the bounds (CalcCompatibleLOProduct)
    _, smin2, smax2 = d.compatible_lo_product(us, A.ml, d_us_ho, s_min, s_max, u_new, el_new, dof_new, dt)
    fin = np.isfinite(smin2)
    assert np.array_equal(np.isfinite(smin_d.cpu().numpy().reshape(ne, nd)), fin)
    assert np.abs(smin_d.cpu().numpy().reshape(ne, nd)[fin] - smin2[fin]).max() < 1e-12 * max(np.abs(smin2[fin]).max(), 1.0)
    ctx.close()


@pytest.mark.parametrize('ho,lo,fct,ode', COMBOS, ids=IDS)
def test_block_mult_and_limit_match_oracle(ho, lo, fct, ode):
    """MultUnlimited + LimitMult on (u, us): rate of u and of us against the oracle"""
    run, ctx = make(ho, lo, fct, ode)
    U, t = evolved_state(run)
    dt = run.dt
    K = run.mult_unlimited(U, t, dt)
    ref = run.limit_mult(U, K, dt)
    Ud = dev(U)
    Kd = torch.empty_like(Ud)
    ctx.mult_unlimited(ho, lo, fct, t, dt, Ud, Kd)
    assert np.abs(Kd.cpu().numpy().reshape(U.shape) - K).max() < 1e-11 * np.abs(K).max()
    ctx.limit_mult(lo, fct, dt, Ud, Kd)
    got = Kd.cpu().numpy().reshape(U.shape)
    assert np.abs(got[0] - ref[0]).max() < 1e-11 * np.abs(K[0]).max()
    assert np.abs(got[1] - ref[1]).max() < 1e-11 * np.abs(K[1]).max()
    # rmh_mult = both passes
    K2 = torch.empty_like(Ud)
    ctx.mult(ho, lo, fct, t, dt, Ud, K2)
    assert torch.equal(K2, Kd)
    ctx.close()


@pytest.mark.parametrize('ho,lo,fct,ode', COMBOS, ids=IDS)
def test_product_steps_match_oracle(ho, lo, fct, ode):
    steps = 6
    run, ctx = make(ho, lo, fct, ode, steps=steps, rs=1)
    U0 = np.stack([run.u, run.us])
    Ud = dev(U0)
    t, dt = 0.0, run.dt
    for _ in range(steps):
        t = ctx.ode_step(ode, ho, lo, fct, t, dt, Ud)
    run.run()
    got = Ud.cpu().numpy().reshape(U0.shape)
    ml = run.disc.cur.ml
    for k, ref in enumerate((run.u, run.us)):
        l1 = float((ml * np.abs(got[k] - ref)).sum() / (ml * np.abs(ref)).sum())
        linf = float(np.abs(got[k] - ref).max() / np.abs(ref).max())
        assert l1 < 1e-10 and linf < 1e-10, (k, l1, linf)
    ctx.set_time(t)
    m = torch.empty(ctx.ndofs, dtype=torch.float64, device='cuda')
    ctx.lumped_mass(m)
    n = ctx.ndofs
    assert abs(ctx.reduce(0, Ud[:n], m) - run.final_mass) < 1e-11 * abs(run.final_mass)
    assert abs(ctx.reduce(0, Ud[n:], m) - run.final_mass_us) < 1e-11 * abs(run.final_mass_us)
    # s = us / u stays inside its initial range (2 + sin sin in [1, 3]) on the active dofs: the point of -ps
    u, us = got
    on = u > 1e-12
    s = us[on] / u[on]
    assert s.min() > 1.0 - 1e-6 and s.max() < 3.0 + 1e-6
    ctx.close()


def test_product_needs_remap_mode():
    import remhos_b200 as rb
    run = oracle_run('periodic-square.mesh', problem=0, rs_levels=0, order=1, ho_type=3)
    ctx = ctx_from_oracle(run)
    with pytest.raises(rb.RmhError, match='remap mode'):
        ctx.product_enable(True)
    ctx.close()


@pytest.mark.parametrize('ode', [12, 13, 14])
def test_idp_masks_match_oracle(ode):
    """RKIDPSolver with use_masks (remhos_solvers.cpp:97-147,171-249; off in the driver): elements of
    the product state that are not fully active in u advance by forward Euler stages"""
    steps = 4
    run, ctx = make(3, 5, 2, ode, rs=2, order=2)
    ctx.idp_use_mask(True)
    U = np.stack([run.u, run.us])
    Ud = dev(U)
    mask = torch.zeros(U.size, dtype=torch.uint8, device='cuda')
    ctx.compute_mask(Ud, mask)
    full = (run.u > 1e-12).all(axis=1)
    ref_mask = np.broadcast_to(full[None, :, None], U.shape)
    assert np.array_equal(mask.cpu().numpy().astype(bool).reshape(U.shape), ref_mask)
    assert 0 < full.sum() < full.size
    t, dt = 0.0, run.dt
    for _ in range(steps):
        U = run.idp_step(U, t, dt, use_mask=True)
        t = ctx.ode_step(ode, 3, 5, 2, t, dt, Ud)
    got = Ud.cpu().numpy().reshape(U.shape)
    assert np.abs(got - U).max() < 1e-10 * np.abs(U).max()
    # and the masks do change the result
    V = np.stack([run.u, run.us])
    tt = 0.0
    for _ in range(steps):
        V = run.idp_step(V, tt, dt)
        tt += dt
    assert np.abs(V - U).max() > 1e-8 * np.abs(U).max()
    ctx.close()
