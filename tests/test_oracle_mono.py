"""CPU tests of the oracle's restatement of MonoRDSolver (remhos_mono.cpp:60-356) in the
configurations no reference number covers (no smoothness indicator, order > 1, subcells): the
properties the scheme guarantees -- conservation on periodic meshes, local bounds, zero-sum rate.
The pinned configurations are in tests/test_oracle_mono_golden.py."""
import numpy as np
import pytest

from helpers import oracle_run


@pytest.mark.parametrize('mono,order,mesh,problem', [(1, 2, 'periodic-square.mesh', 5),
                                                     (2, 3, 'periodic-square.mesh', 5),
                                                     (1, 2, 'periodic-cube.mesh', 0)])
def test_mono_conservative_and_bounded(mono, order, mesh, problem):
    run = oracle_run(mesh, mono_type=mono, problem=problem, rs_levels=1, order=order, dt=0.002,
                     max_steps=10, ode_solver=3)
    run.run()
    assert abs(run.final_mass - run.mass0) < 1e-13 * abs(run.mass0)
    assert run.u.min() > run.u0_min - 1e-12 and run.u.max() < run.u0_max + 1e-12


def test_mono_rate_is_zero_sum_per_element_plus_fluxes():
    """sum_i m_i du_i over the whole periodic mesh vanishes for any state (every stage of the
    scheme -- alpha splitting, face corrections, mass correction -- is conservative)"""
    run = oracle_run('periodic-square.mesh', mono_type=1, problem=5, rs_levels=2, order=2, dt=0.002)
    rng = np.random.default_rng(3)
    u = np.clip(run.u + 0.1 * rng.standard_normal(run.u.shape), 0.0, 1.0)
    du = run.mult(u, 0.0, run.dt)
    ml = run.disc.cur.ml
    assert abs((ml * du).sum()) < 1e-12 * np.abs(ml * du).sum()


def test_mono_scale_formula():
    """scale_e = vmax / (2 sqrt(dim) h_e / p) (remhos_mono.cpp:40-57): unit-speed diagonal flow on
    the 3x3 periodic square refined twice: h = 2/12"""
    run = oracle_run('periodic-square.mesh', mono_type=1, problem=0, rs_levels=2, order=2, dt=0.002)
    h = 2.0 / 12
    assert np.allclose(run.mono_scale, 1.0 / (2 * np.sqrt(2) * h / 2), rtol=1e-12)
