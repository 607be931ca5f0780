"""The oracle's ElementFCTProjection (-fct 4) and automatic time step control (-dtc 1) against the
reference's known answer autotest/out_baseline.dat:207-210 ("BLAST sharpening test"):
  -m periodic-square.mesh -p 5 -rs 3 -dt 0.01 -tf 0.8 -ho 3 -lo 5 -fct 4 -bt 1 -dtc 1
final mass and maximum to the 10 printed digits; and out_baseline.dat:202-205 ("Pacman remap auto-dt"):
  -m inline-quad.mesh -p 14 -rs 1 -dt -1 -tf 0.75 -ho 3 -lo 5 -fct 4 -bt 1 -dtc 1
final mass (10 digits) and mass loss (6 digits)."""
from helpers import oracle_run


def test_fct_project_dtc_known_answer():
    run = oracle_run('periodic-square.mesh', problem=5, rs_levels=3, order=3, dt=0.01, t_final=0.8,
                     ode_solver=3, ho_type=3, lo_type=5, fct_type=4, bounds_type=1, dt_control=1)
    run.run()
    assert float('%.10g' % run.final_mass) == 0.1623263888
    assert float('%.10g' % run.u.max()) == 0.2863317261


def test_fct_project_dtc_remap_known_answer():
    run = oracle_run('inline-quad.mesh', problem=14, rs_levels=1, order=3, dt=-1.0, t_final=0.75,
                     ode_solver=3, ho_type=3, lo_type=5, fct_type=4, bounds_type=1, dt_control=1)
    run.run()
    assert float('%.10g' % run.final_mass) == 0.08479612805
    assert float('%.6g' % abs(run.mass0 - run.final_mass)) == 6.61247e-07
