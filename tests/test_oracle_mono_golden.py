"""The oracle's MonoRDSolver + SmoothnessIndicator restatement against the reference's own known
answers for the monolithic solver: autotest/out_baseline.dat:212-220 (= README.md runs 12, 13),
final mass and maximum to the 10 printed digits.  These runs also pin NonlinFluxLumping, the
inflow boundary state (including the "high order projection" of remhos.cpp:628-635 for problem 7)
and the steady-state stopping rule (remhos.cpp:1276-1295)."""
import pytest

from helpers import oracle_run

ROWS = [
    # -m inline-quad.mesh -p 7 -rs 3 -o 1 -dt 0.01 -tf 20 -mono 1 -si 2     out_baseline.dat:212-215
    (dict(problem=7, rs_levels=3, si_type=2), 0.1570667907, 0.9987771164),
    # -m inline-quad.mesh -p 6 -rs 2 -o 1 -dt 0.01 -tf 20 -mono 1 -si 1     out_baseline.dat:217-220
    (dict(problem=6, rs_levels=2, si_type=1), 0.3182739921, 1.0),
]


@pytest.mark.parametrize('cfg,mass,umax', ROWS)
def test_mono_si_known_answers(cfg, mass, umax):
    run = oracle_run('inline-quad.mesh', mono_type=1, order=1, dt=0.01, t_final=20.0, ode_solver=3,
                     **cfg)
    run.run()
    assert float('%.10g' % run.final_mass) == mass
    assert float('%.10g' % run.u.max()) == umax
    assert run.residual < 1e-12 and run.t >= 1.0       # stopped by the steady-state criterion
