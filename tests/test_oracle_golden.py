"""Pins the CPU oracle on the reference's own known answers (no GPU needed).

Sources: autotest/out_baseline.dat (final mass + max value, 10 significant digits, np=2) and
remhos_tests.cpp:38-107 (final mass of `-ho 3 -lo 5 -fct 2` remap runs, tolerance 10 eps).
`-ho 2 ... -pa` rows are reproduced with the exact local inverse (`-ho 3`): the reference's CG
solve (rel. tol 1e-12) agrees with it to the printed digits.
"""
import os

import numpy as np
import pytest

from helpers import DATA
from remhos_oracle import driver, mesh as om

REF_DATA = '/root/reference/data'


def run(mesh, **kw):
    r = driver.Run(driver.Options(mesh_file=os.path.join(DATA, mesh), **kw))
    r.run()
    return r


def digits10(x):
    return float('%.10g' % x)


# (mesh, options, mass, max) -- autotest/out_baseline.dat line numbers in the comments
BASELINE = [
    ('periodic-cube.mesh', dict(problem=0, rs_levels=1, order=2, dt=0.015, t_final=2, ho_type=3,
                                lo_type=1, fct_type=1), 0.9607429525, 0.9984668427),     # :177-180
    ('periodic-cube.mesh', dict(problem=0, rs_levels=1, order=2, dt=0.015, t_final=2, ho_type=3,
                                lo_type=3, fct_type=2), 0.9607429525, 0.9202929163),     # :103-106
    ('periodic-square.mesh', dict(problem=5, rs_levels=3, dt=0.004, t_final=0.8, ho_type=3,
                                  lo_type=3, fct_type=2), 0.1623263888, 0.6374820899),   # :98-101
    ('inline-quad.mesh', dict(problem=14, rs_levels=1, dt=0.0015, t_final=0.75, ho_type=3,
                              lo_type=1, fct_type=1), 0.08479546845, 0.905654904),       # :152-155
    ('inline-quad.mesh', dict(problem=14, rs_levels=1, dt=0.0015, t_final=0.75, ho_type=3,
                              lo_type=3, fct_type=2), 0.08479546775, 0.7779015453),      # :78-81
]


@pytest.mark.parametrize('mesh,opt,mass,umax', BASELINE,
                         ids=['cube-DU-fluxFCT', 'cube-RD-clipscale', 'square-RD-clipscale',
                              'quad-remap-DU-fluxFCT', 'quad-remap-RD-clipscale'])
def test_autotest_baseline(mesh, opt, mass, umax):
    r = run(mesh, **opt)
    assert digits10(r.final_mass) == mass
    assert digits10(r.final_max) == umax


# remhos_tests.cpp:38-91: -ho 3 -lo 5 -fct 2, -dt -1 (CFL), -tf 0.5
TESTS_CPP = [
    ('inline-quad.mesh', dict(problem=14, rs_levels=1, order=2, max_steps=5), 0.09711395400387984),
    ('inline-quad.mesh', dict(problem=14, rs_levels=4, order=2, max_steps=5), 0.09185717760402806),
    ('inline-quad.mesh', dict(problem=14, rs_levels=4, order=3, max_steps=5), 0.0930984399257905),
    ('cube01_hex.mesh', dict(problem=10, rs_levels=1, order=2, max_steps=5), 0.11972857593296446),
]


@pytest.mark.parametrize('mesh,opt,mass', TESTS_CPP,
                         ids=['quad-rs1-o2', 'quad-rs4-o2', 'quad-rs4-o3', 'hex-rs1-o2'])
def test_remhos_tests_final_mass(mesh, opt, mass):
    r = run(mesh, dt=-1.0, t_final=0.5, ho_type=3, lo_type=5, fct_type=2, **opt)
    # the reference accepts 10 eps relative to (1 + |x|) (AlmostEq, remhos_tests.cpp:13-23);
    # an independent implementation sums in a different order: hold 1e-14 relative
    assert abs(r.final_mass - mass) < 1e-14 * (1.0 + abs(mass))
    # the run really moved mass between the initial and the final mesh representation
    assert abs(r.mass0 - r.final_mass) > 1e-10


def test_known_unmatched_row_is_documented():
    """autotest/out_baseline.dat:172-175 (periodic-square, -ho 3 -lo 1 -fct 1) is the one row tried
    that the oracle does not reproduce (max 0.7879213622 vs 0.787875182); every component of that
    combination is pinned by other rows.  Keep the discrepancy visible (DESIGN.md 'Oracle')."""
    r = run('periodic-square.mesh', problem=5, rs_levels=3, dt=0.004, t_final=0.8, ho_type=3,
            lo_type=1, fct_type=1, max_steps=30)
    assert digits10(r.final_mass) == 0.1623263888


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason='reference tree not mounted')
@pytest.mark.parametrize('name', ['periodic-square', 'periodic-cube', 'cube01_hex', 'inline-quad'])
def test_generated_meshes_equal_reference_meshes(name):
    a = om.read_mesh(os.path.join(DATA, name + '.mesh'))
    b = om.read_mesh(os.path.join(REF_DATA, name + '.mesh'))
    sa = sorted(map(tuple, a.X.reshape(a.ne, -1).tolist()))
    sb = sorted(map(tuple, b.X.reshape(b.ne, -1).tolist()))
    assert sa == sb
