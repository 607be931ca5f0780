"""GPU parity tests of the monolithic solver (MonoRDSolver, remhos_mono.cpp:60-356, `-mono 1/2`,
with and without the smoothness indicator) against the CPU oracle, whose restatement is pinned on
the reference's two known answers for -mono (tests/test_oracle_mono_golden.py; the GPU CLI
reproduces them too, tests/test_cli.py).  Configurations no reference number covers (no -si,
order > 1, subcells) are checked CUDA-vs-oracle plus the properties the scheme guarantees
(conservation on periodic meshes, local bounds).

Tolerances: one evaluation 1e-10 of the field's max norm (the fixed-point loop stops on a norm
threshold, so a round-off sized difference may add or drop one sweep whose update is below 1e-8 of
the residual scale); whole runs 1e-9.
"""
import numpy as np
import pytest

from helpers import oracle_run, ctx_from_oracle, rel_err, subcell_setup_from_oracle

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

CASES = [
    ('periodic-square.mesh', dict(problem=5, rs_levels=2, order=2, mono_type=1), 0),
    ('periodic-square.mesh', dict(problem=1, rs_levels=1, order=3, mono_type=2), 0),
    ('periodic-square.mesh', dict(problem=5, rs_levels=1, order=3, mono_type=2), 1),
    ('inline-quad.mesh', dict(problem=6, rs_levels=2, order=1, mono_type=1), 0),      # mass_lim off
    ('inline-quad.mesh', dict(problem=4, rs_levels=1, order=2, mono_type=1), 0),      # inflow boundary
    ('periodic-hexagon.mesh', dict(problem=0, rs_levels=1, order=2, mono_type=1), 0),
    ('periodic-cube.mesh', dict(problem=0, rs_levels=1, order=2, mono_type=1), 0),
    ('periodic-cube.mesh', dict(problem=1, rs_levels=0, order=3, mono_type=2), 1),
]


def dev(a):
    return torch.tensor(np.ascontiguousarray(a, dtype=np.float64).reshape(-1), device='cuda')


def make(mesh, opt, bt):
    run = oracle_run(mesh, bounds_type=bt, dt=0.002, **opt)
    ctx = ctx_from_oracle(run)
    if opt['mono_type'] == 2:
        subcell_setup_from_oracle(run, ctx)
    ctx.mono_setup(opt['mono_type'], run.opt.problem not in (6, 7), run.mono_scale)
    return run, ctx


@pytest.mark.parametrize('mesh,opt,bt', CASES)
def test_mono_rd_matches_oracle(mesh, opt, bt):
    run, ctx = make(mesh, opt, bt)
    rng = np.random.default_rng(5)
    u = np.clip(run.u + 0.05 * rng.standard_normal(run.u.shape), 0.0, 1.0)
    ref = run.mult(u, 0.0, run.dt)
    k = torch.empty(ctx.ndofs, dtype=torch.float64, device='cuda')
    ctx.mono_rd(dev(u), k)
    assert rel_err(k.cpu().numpy().reshape(u.shape), ref) < 1e-10
    # rmh_mult evaluates the monolithic solver while one is set (remhos.cpp:1687)
    k2 = torch.empty_like(k)
    ctx.mult(3, 5, 2, 0.0, run.dt, dev(u), k2)
    assert torch.equal(k, k2)
    ctx.close()


@pytest.mark.parametrize('mesh,opt,bt', [CASES[0], CASES[2], CASES[6]])
def test_mono_run_matches_oracle(mesh, opt, bt):
    run, ctx = make(mesh, dict(opt, ode_solver=3, max_steps=8), bt)
    u = dev(run.u)
    t = 0.0
    for _ in range(8):
        t = ctx.ode_step(3, 0, 0, 0, t, run.dt, u)
    run.run()
    ug = u.cpu().numpy().reshape(run.u.shape)
    assert rel_err(ug, run.u) < 1e-9
    m = torch.empty(ctx.ndofs, dtype=torch.float64, device='cuda')
    ctx.lumped_mass(m)
    mass = ctx.reduce(0, u, m)
    assert abs(mass - run.mass0) < 1e-12 * abs(run.mass0)          # periodic: conservative
    assert ug.min() > run.u0_min - 1e-10 and ug.max() < run.u0_max + 1e-10
    ctx.close()


# ---- smoothness indicator (remhos_tools.cpp:24-354), order 1; the oracle's restatement is pinned on
# both reference known answers for -mono (tests/test_oracle_mono_golden.py)
SI_CASES = [('inline-quad.mesh', 6, 2, 1, 1), ('inline-quad.mesh', 7, 2, 2, 1), ('periodic-square.mesh', 5, 2, 1, 1),
            ('cube01_hex.mesh', 1, 1, 2, 1),
            # orders above 1: H1 space on the subcell mesh (no reference number; oracle vs CUDA)
            ('inline-quad.mesh', 6, 1, 1, 2), ('periodic-square.mesh', 5, 1, 2, 3), ('cube01_hex.mesh', 1, 1, 1, 2)]


@pytest.mark.parametrize('mesh,problem,rs,si,order', SI_CASES)
def test_si_and_mono_with_si_match_oracle(mesh, problem, rs, si, order):
    run = oracle_run(mesh, mono_type=1, si_type=si, problem=problem, rs_levels=rs, order=order, dt=0.002)
    ctx = ctx_from_oracle(run)
    ctx.mono_setup(1, run.opt.problem not in (6, 7), run.mono_scale)
    ctx.si_setup(si)
    rng = np.random.default_rng(9)
    u = np.clip(run.u + 0.05 * rng.standard_normal(run.u.shape), 0.0, 1.0)
    tmp = torch.empty(ctx.ndofs, dtype=torch.float64, device='cuda')
    ctx.si_values(dev(u), tmp)
    ref_si = run.si.dof_values(u)
    # type 1 raises a ratio to the 5th power and type 2 divides two Laplacian-sized numbers:
    # compare at 1e-9 absolute (values live in [0, 1])
    assert np.abs(tmp.cpu().numpy().reshape(u.shape) - ref_si).max() < 1e-9
    k = torch.empty_like(tmp)
    ctx.mono_rd(dev(u), k)
    assert rel_err(k.cpu().numpy().reshape(u.shape), run.mult(u, 0.0, run.dt)) < 1e-9
    ctx.close()
