"""The C/OpenMP port of the stage path (oracle/c, the CPU baseline of bench.py) against the numpy
oracle -- which is pinned on the reference's known answers (test_oracle_golden.py) -- and against
remhos_tests.cpp:64-67 directly.  No GPU needed."""
import numpy as np
import pytest

from helpers import oracle_run, rel_err
from remhos_oracle.cport import Port

CASES = [
    ('periodic-cube.mesh', dict(problem=0, rs_levels=1, order=3, dt=0.01)),
    ('periodic-cube.mesh', dict(problem=1, rs_levels=1, order=2, dt=0.01)),
    ('periodic-cube.mesh', dict(problem=0, rs_levels=0, order=4, dt=0.01)),
    ('cube01_hex.mesh', dict(problem=10, rs_levels=1, order=2, dt=0.02, t_final=0.7)),
    ('cube01_hex.mesh', dict(problem=1, rs_levels=1, order=1, dt=0.01)),
]


@pytest.mark.parametrize('mesh,opt', CASES)
def test_port_stage_matches_numpy_oracle(mesh, opt):
    run = oracle_run(mesh, ho_type=3, lo_type=5, fct_type=2, **opt)
    port = Port.from_run(run)
    rng = np.random.default_rng(11)
    u = np.clip(run.u + 0.02 * rng.standard_normal(run.u.shape), 0.0, None)
    t = 0.3 if run.exec_mode == 1 else 0.0
    ref = run.mult(u, t, run.dt)
    port.set_time(t)
    assert rel_err(port.lumped_mass(), run.disc.cur.ml) < 1e-13
    k = port.stage(run.dt, u)
    tol = 1e-10 if run.space.p <= 3 else 1e-8
    assert rel_err(k, ref) < tol
    port.close()


def test_port_run_reproduces_reference_final_mass():
    """remhos_tests.cpp:64-67: cube01_hex -p 10 -rs 1 -o 2 -dt -1 -tf 0.5 -ho 3 -lo 5 -fct 2 -ms 5."""
    run = oracle_run('cube01_hex.mesh', problem=10, rs_levels=1, order=2, dt=-1.0, t_final=0.5,
                     ho_type=3, lo_type=5, fct_type=2, max_steps=5)
    port = Port.from_run(run)
    u = run.u.copy()
    t = 0.0
    for _ in range(5):
        dt = min(run.dt, run.t_final - t)
        port.rk3_step(t, dt, u)
        t += dt
    port.set_time(t)
    mass = float((port.lumped_mass() * u).sum())
    assert abs(mass - 0.11972857593296446) < 1e-13
    run.run()
    assert rel_err(u, run.u) < 1e-12
    port.close()
