"""GPU parity tests of the fused RK-stage path (rmh_stage / rmh_rk_stage / rmh_rk_step /
rmh_rk_step_host) against the CPU oracle, and of whole runs against the reference's own
known answers (remhos_tests.cpp:38-107: final mass of `-ho 3 -lo 5 -fct 2` remap runs).

Tolerance (north_star): final-solution L1/Linf and total mass within 1e-12 relative.
"""
import numpy as np
import pytest

from helpers import oracle_run, ctx_from_oracle, rel_err

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu


def dev(a):
    return torch.tensor(np.ascontiguousarray(a, dtype=np.float64).reshape(-1), device='cuda')


STAGE_CASES = [
    ('periodic-square.mesh', dict(problem=5, rs_levels=2, order=3, dt=0.004), 0),
    ('periodic-square.mesh', dict(problem=1, rs_levels=2, order=2, dt=0.004), 1),
    ('inline-quad.mesh', dict(problem=14, rs_levels=1, order=3, dt=0.002, t_final=0.75), 0),
    ('periodic-hexagon.mesh', dict(problem=0, rs_levels=1, order=3, dt=0.005), 0),
    ('periodic-hexagon.mesh', dict(problem=1, rs_levels=1, order=2, dt=0.005), 1),
    ('periodic-cube.mesh', dict(problem=0, rs_levels=1, order=3, dt=0.01), 0),
    ('periodic-cube.mesh', dict(problem=1, rs_levels=1, order=2, dt=0.01), 1),
    ('periodic-cube.mesh', dict(problem=0, rs_levels=1, order=4, dt=0.005), 0),
    ('periodic-cube.mesh', dict(problem=1, rs_levels=0, order=1, dt=0.01), 0),
    ('cube01_hex.mesh', dict(problem=10, rs_levels=1, order=3, dt=0.02, t_final=0.7), 0),
]


@pytest.mark.parametrize('mesh,opt,bt', STAGE_CASES)
def test_stage_matches_oracle(mesh, opt, bt):
    run = oracle_run(mesh, ho_type=3, lo_type=5, fct_type=2, bounds_type=bt, **opt)
    ctx = ctx_from_oracle(run)
    rng = np.random.default_rng(7)
    u = np.clip(run.u + 0.02 * rng.standard_normal(run.u.shape), 0.0, None)
    t = 0.3 if run.exec_mode == 1 else 0.0
    ref = run.mult(u, t, run.dt)
    ctx.set_time(t)
    k = torch.empty(ctx.ndofs, dtype=torch.float64, device='cuda')
    ctx.stage(5, run.dt, dev(u), k)
    # order 4: the Bernstein mass matrix has cond ~ 2e6, carried by both the oracle's dense LU
    # and the Kronecker inverse
    assert rel_err(k.cpu().numpy().reshape(u.shape), ref) < (1e-10 if run.space.p <= 3 else 1e-8)
    # the limited update keeps u + dt*k inside the bounds (same verdict as the oracle's check)
    umin, umax = run.disc.bounds(u, bt)
    un = u + run.dt * k.cpu().numpy().reshape(u.shape)
    assert (un + 1e-12 >= umin).all() and (un <= umax + 1e-12).all()
    ctx.close()


RUN_CASES = [
    # (mesh, options, steps, reference final mass or None)
    ('inline-quad.mesh', dict(problem=14, rs_levels=1, order=2, dt=-1.0, t_final=0.5), 5,
     0.09711395400387984),                               # remhos_tests.cpp:40-44, 70-73
    ('cube01_hex.mesh', dict(problem=10, rs_levels=1, order=2, dt=-1.0, t_final=0.5), 5,
     0.11972857593296446),                               # remhos_tests.cpp:64-67
    ('periodic-cube.mesh', dict(problem=0, rs_levels=1, order=3, dt=0.01, t_final=1.0), 6, None),
    ('periodic-square.mesh', dict(problem=5, rs_levels=3, order=3, dt=0.004, t_final=0.8), 10, None),
]


@pytest.mark.parametrize('mesh,opt,steps,ref_mass', RUN_CASES)
@pytest.mark.parametrize('ode', [3, 2, 1])
def test_rk_steps_match_oracle(mesh, opt, steps, ref_mass, ode):
    run = oracle_run(mesh, ho_type=3, lo_type=5, fct_type=2, ode_solver=ode, max_steps=steps, **opt)
    ctx = ctx_from_oracle(run)
    u = dev(run.u)
    t, dt = 0.0, run.dt
    for _ in range(steps):
        dt_real = min(dt, run.t_final - t)
        t = ctx.rk_step(ode, 5, t, dt_real, u)
    run.run()
    ug = u.cpu().numpy().reshape(run.u.shape)
    # final solution norms (lumped-mass weighted L1, and Linf) within 1e-12 relative
    ml = run.disc.cur.ml if run.exec_mode == 1 else run.masses0
    l1 = float((ml * np.abs(ug - run.u)).sum() / (ml * np.abs(run.u)).sum())
    linf = float(np.abs(ug - run.u).max() / np.abs(run.u).max())
    assert l1 < 1e-12 and linf < 1e-12, (l1, linf)
    # total mass on the final mesh
    ctx.set_time(t)
    m = torch.empty(ctx.ndofs, dtype=torch.float64, device='cuda')
    ctx.lumped_mass(m)
    mass = ctx.reduce(0, u, m)
    assert abs(mass - run.final_mass) < 1e-12 * abs(run.final_mass)
    if ref_mass is not None and ode == 3:
        assert abs(mass - ref_mass) < 1e-12 * abs(ref_mass)
    # bound preservation verdict: global extrema do not grow (remhos.cpp:1219-1260)
    assert ug.min() > run.u0_min - 1e-10 and ug.max() < run.u0_max + 1e-10
    ctx.close()


def test_rk_step_host_equals_device():
    run = oracle_run('periodic-cube.mesh', ho_type=3, lo_type=5, fct_type=2, problem=0,
                     rs_levels=1, order=3, dt=0.01)
    ctx = ctx_from_oracle(run)
    u = dev(run.u)
    uh = torch.tensor(run.u.reshape(-1)).pin_memory()
    t1 = ctx.rk_step(3, 5, 0.0, run.dt, u)
    t2 = ctx.rk_step_host(3, 5, 0.0, run.dt, uh.data_ptr())
    assert t1 == t2
    assert torch.equal(u.cpu(), uh)
    ctx.close()


# Stage-kernel variants on affine 3D meshes (rmh_ctx_path_flags): bit 3 = velocity constant over
# every element -> constant-coefficient kernel (stage3c.cuh, default when it applies); bit 2 =
# velocity linear over every element -> DMMA kernel that rebuilds the quadrature data in-kernel
# (opt-in, RMH_LINEAR_OP=1); otherwise the DMMA kernel streams the stored quadrature data.  All
# variants must match the oracle and each other.
@pytest.mark.parametrize('problem,order,bt,const,lin', [
    (0, 3, 0, True, True), (0, 3, 1, True, True), (0, 4, 0, True, True), (0, 2, 0, True, True),
    (0, 1, 1, True, True), (1, 3, 0, False, True), (1, 4, 0, False, True), (1, 1, 0, False, True),
    (3, 3, 0, False, False), (3, 2, 1, False, False)])
def test_stage_kernel_variants(problem, order, bt, const, lin, monkeypatch):
    run = oracle_run('periodic-cube.mesh', ho_type=3, lo_type=5, fct_type=2, problem=problem,
                     rs_levels=1, order=order, dt=0.005, bounds_type=bt)
    rng = np.random.default_rng(11)
    u = np.clip(run.u + 0.02 * rng.standard_normal(run.u.shape), 0.0, None)
    ref = run.mult(u, 0.0, run.dt)
    tol = 1e-10 if order <= 3 else 1e-8
    out = []
    # bit 4: overlap bounds formed inside the constant-coefficient kernel (k_stage3c<FOLD>)
    fold = 16 if (const and bt == 0) else 0
    for env, flags in [({}, 3 | (8 if const else 0) | fold),
                       ({'RMH_LINEAR_OP': '1', 'RMH_NO_CONST_OP': '1'}, 3 | (4 if lin else 0)),
                       ({'RMH_NO_CONST_OP': '1'}, 3),
                       ({'RMH_NO_FOLD': '1'}, 3 | (8 if const else 0))]:
        for k_ in ('RMH_LINEAR_OP', 'RMH_NO_CONST_OP', 'RMH_NO_FOLD'):
            monkeypatch.delenv(k_, raising=False)
        for k_, v_ in env.items():
            monkeypatch.setenv(k_, v_)
        ctx = ctx_from_oracle(run)
        assert ctx.path_flags == flags, (env, ctx.path_flags)
        k = torch.empty(ctx.ndofs, dtype=torch.float64, device='cuda')
        ctx.stage(5, run.dt, dev(u), k)
        out.append(k.cpu().numpy())
        assert rel_err(out[-1].reshape(u.shape), ref) < tol, env
        ctx.close()
    assert rel_err(out[0], out[2]) < 1e-12 and rel_err(out[1], out[2]) < 1e-12
    assert rel_err(out[3], out[0]) < 1e-13      # bounds through the entity pass or in the kernel: same operator


def test_const_kernel_ring_depths(monkeypatch):
    """every (warps, resident blocks) configuration of k_stage3c, with the overlap bounds formed in
    the kernel or by the entity pass, gives the same RK3 step"""
    run = oracle_run('periodic-cube.mesh', ho_type=3, lo_type=5, fct_type=2, problem=0,
                     rs_levels=2, order=3, dt=0.005, max_steps=2)
    res = []
    for cfg, nofold in (('0', '0'), ('822', '0'), ('0', '1'), ('822', '1')):
        # the configuration is latched per process on first use: run each in a fresh interpreter
        import subprocess, sys, os, json
        code = ("import sys, os, numpy as np, torch; sys.path[:0] = [%r, %r, %r];"
                "from helpers import oracle_run, ctx_from_oracle;"
                "run = oracle_run('periodic-cube.mesh', ho_type=3, lo_type=5, fct_type=2, problem=0,"
                " rs_levels=2, order=3, dt=0.005, max_steps=2);"
                "ctx = ctx_from_oracle(run); assert ctx.path_flags in (11, 27);"
                "u = torch.tensor(run.u.reshape(-1), device='cuda'); t = 0.0\n"
                "for _ in range(2): t = ctx.rk_step(3, 5, t, run.dt, u)\n"
                "np.save(sys.argv[1], u.cpu().numpy())")
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        code = code % (root, os.path.join(root, 'oracle'), os.path.join(root, 'tests'))
        import tempfile
        with tempfile.TemporaryDirectory() as td:
            f = os.path.join(td, 'u.npy')
            env = dict(os.environ, RMH_C_CFG=cfg, RMH_NO_FOLD=nofold)
            subprocess.check_call([sys.executable, '-c', code, f], env=env)
            res.append(np.load(f))
    run.run()
    for r in res:
        assert np.array_equal(r, res[0])
        assert rel_err(r.reshape(run.u.shape), run.u) < 1e-12


@pytest.mark.parametrize('ode', [3, 2])
def test_trust_state_is_bit_identical(ode):
    """rmh_ctx_trust_state reuses the element min/max of the last stage's output: same bits as
    recomputing them, also after the state was touched through another entry point"""
    run = oracle_run('periodic-cube.mesh', ho_type=3, lo_type=5, fct_type=2, problem=0,
                     rs_levels=1, order=3, dt=0.01)
    ctx = ctx_from_oracle(run)
    u1, u2 = dev(run.u), dev(run.u)
    t = 0.0
    for _ in range(4):
        t = ctx.rk_step(ode, 5, t, run.dt, u1)
    ctx.trust_state(True)
    t = 0.0
    k = torch.empty_like(u2)
    for i in range(4):
        t = ctx.rk_step(ode, 5, t, run.dt, u2)
        if i == 1:
            ctx.stage(5, run.dt, u1, k)      # another vector passes through the context in between
    assert torch.equal(u1, u2)
    ctx.close()


# k_stage3c edge cases: element counts that are not a multiple of the elements a warp handles at
# once (27 elements: last group partial at every order), and domain-boundary faces (exterior state 0
# through zero-fill copies) with both bounds types
@pytest.mark.parametrize('mesh,rs,order,bt', [
    ('periodic-cube.mesh', 0, 1, 0), ('periodic-cube.mesh', 0, 2, 0), ('periodic-cube.mesh', 0, 3, 1),
    ('periodic-cube.mesh', 0, 4, 0), ('cube01_hex.mesh', 1, 3, 0), ('cube01_hex.mesh', 1, 2, 1),
    ('cube01_hex.mesh', 0, 1, 0), ('cube01_hex.mesh', 2, 4, 0)])
def test_const_kernel_edge_cases(mesh, rs, order, bt):
    run = oracle_run(mesh, ho_type=3, lo_type=5, fct_type=2, problem=0, rs_levels=rs, order=order,
                     dt=0.004, bounds_type=bt, max_steps=3)
    ctx = ctx_from_oracle(run)
    assert ctx.path_flags & 8, 'constant-coefficient kernel expected'
    rng = np.random.default_rng(13)
    u = np.clip(run.u + 0.02 * rng.standard_normal(run.u.shape), 0.0, None)
    ref = run.mult(u, 0.0, run.dt)
    k = torch.empty(ctx.ndofs, dtype=torch.float64, device='cuda')
    ctx.stage(5, run.dt, dev(u), k)
    assert rel_err(k.cpu().numpy().reshape(u.shape), ref) < (1e-10 if order <= 3 else 1e-8)
    # a few RK3 steps, state carried across steps
    ctx.trust_state(True)
    ud = dev(run.u)
    t = 0.0
    for _ in range(3):
        t = ctx.rk_step(3, 5, t, run.dt, ud)
    run.run()
    assert rel_err(ud.cpu().numpy().reshape(run.u.shape), run.u) < (1e-12 if order <= 3 else 1e-10)
    ctx.close()


@pytest.mark.parametrize('env,flags', [({'RMH_NO_TENSOR': '1'}, 1), ({'RMH_NO_PIPELINE': '1'}, None)],
                         ids=['dfma-pipelined-k_stage3p', 'one-batch-per-block-k_stage'])
@pytest.mark.parametrize('order,bt', [(3, 0), (2, 1), (4, 0)])
def test_fallback_stage_kernels(env, flags, order, bt, monkeypatch):
    """the two fall-back fused stage kernels of affine 3D meshes stay correct: k_stage3p (DFMA,
    persistent, RMH_NO_TENSOR=1: stored quadrature data in the generic order) and k_stage (the
    non-affine / 2D kernel, RMH_NO_PIPELINE=1)"""
    for k_, v_ in env.items():
        monkeypatch.setenv(k_, v_)
    run = oracle_run('periodic-cube.mesh', ho_type=3, lo_type=5, fct_type=2, problem=1, rs_levels=1, order=order,
                     dt=0.005, bounds_type=bt, max_steps=2)
    ctx = ctx_from_oracle(run)
    if flags is not None:
        assert ctx.path_flags == flags, ctx.path_flags          # affine, generic data order, no DMMA / const kernel
    rng = np.random.default_rng(3)
    u = np.clip(run.u + 0.02 * rng.standard_normal(run.u.shape), 0.0, None)
    k = torch.empty(ctx.ndofs, dtype=torch.float64, device='cuda')
    ctx.stage(5, run.dt, dev(u), k)
    assert rel_err(k.cpu().numpy().reshape(u.shape), run.mult(u, 0.0, run.dt)) < (1e-10 if order <= 3 else 1e-8)
    ud = dev(run.u)
    t = 0.0
    for _ in range(2):
        t = ctx.rk_step(3, 5, t, run.dt, ud)
    run.run()
    assert rel_err(ud.cpu().numpy().reshape(run.u.shape), run.u) < (1e-12 if order <= 3 else 1e-10)
    ctx.close()
