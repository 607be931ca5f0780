"""The remhos command-line driver (remhos_b200/host): flag handling without a GPU, and -- on the
GPU -- whole runs through the C++ solver-interface mirror against the reference's known answers
(remhos_tests.cpp:38-91, autotest/out_baseline.dat) and the CPU oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import DATA, oracle_run

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, 'remhos_b200', 'host', 'remhos')


def run_cli(*args):
    p = subprocess.run([EXE] + [str(a) for a in args], capture_output=True, text=True, timeout=180)
    return p.returncode, p.stdout, p.stderr


def test_driver_is_built():
    assert os.path.exists(EXE), 'run __graft_entry__.build()'


def test_bad_flag_returns_1():
    rc, out, _ = run_cli('-no-such-flag')
    assert rc == 1 and 'Usage' in out                      # remhos.cpp:335-339


def test_unknown_ode_solver_returns_3():
    rc, out, _ = run_cli('-m', 'x', '-s', '7')
    assert rc == 3 and 'Unknown ODE solver type: 7' in out  # remhos.cpp:499-500


def test_rejected_combinations_abort():
    for flags in (['-fct', '2', '-lo', '0'], ['-lo', '5', '-ho', '0'], ['-fct', '1', '-lo', '1', '-pa']):
        rc, _, err = run_cli('-m', 'x', *flags)
        assert rc == 134 and 'Verification failed' in err


def parse(out):
    g = lambda pat: float(re.search(pat, out).group(1))
    return dict(n=int(g(r'Number of unknowns: (\d+)')), mass=g(r'Final mass u:\s+(\S+)'),
                umax=g(r'Max value u:\s+(\S+)'), loss=g(r'Mass loss u:\s+(\S+)'))


def mesh(name):
    return os.path.join(DATA, name)


CLI_KNOWN = [
    # remhos_tests.cpp:40-44,64-67: final mass of -ho 3 -lo 5 -fct 2 remap runs (first 10 digits)
    (['-m', mesh('inline-quad.mesh'), '-p', 14, '-rs', 1, '-o', 2, '-dt', -1, '-tf', 0.5, '-ho', 3,
      '-lo', 5, '-fct', 2, '-ms', 5, '-no-vis'], 0.09711395400387984, None),
    (['-m', mesh('cube01_hex.mesh'), '-p', 10, '-rs', 1, '-o', 2, '-dt', -1, '-tf', 0.5, '-ho', 3,
      '-lo', 5, '-fct', 2, '-ms', 5, '-no-vis', '-pa'], 0.11972857593296446, None),
    # remhos_tests.cpp #1, #2, #5, #7 (65536 / 102400 unknowns in 2D; 3D PA order 3 -rs 3)
    (['-m', mesh('inline-quad.mesh'), '-p', 14, '-rs', 4, '-o', 3, '-dt', -1, '-tf', 0.5, '-ho', 3,
      '-lo', 5, '-fct', 2, '-ms', 5, '-no-vis', '-vs', 1], 0.0930984399257905, None),
    (['-m', mesh('inline-quad.mesh'), '-p', 14, '-rs', 4, '-o', 4, '-dt', -1, '-tf', 0.5, '-ho', 3,
      '-lo', 5, '-fct', 2, '-ms', 5, '-no-vis', '-vs', 1], 0.09237630484178257, None),
    (['-m', mesh('inline-quad.mesh'), '-pa', '-p', 14, '-rs', 4, '-o', 2, '-dt', -1, '-tf', 0.5,
      '-ho', 3, '-lo', 5, '-fct', 2, '-ms', 5, '-no-vis', '-vs', 1], 0.09185717760402806, None),
    (['-m', mesh('cube01_hex.mesh'), '-pa', '-p', 10, '-rs', 3, '-o', 3, '-dt', -1, '-tf', 0.5,
      '-ho', 3, '-lo', 5, '-fct', 2, '-ms', 1, '-no-vis', '-vs', 1], 0.11601536511552431, None),
    # autotest/out_baseline.dat:167-170, :93-96: unstructured periodic hexagon
    (['-m', mesh('periodic-hexagon.mesh'), '-p', 0, '-rs', 2, '-dt', 0.005, '-tf', 2.5, '-ho', 3,
      '-lo', 1, '-fct', 1, '-no-vis'], 0.3888354875, 0.9979069772),
    (['-m', mesh('periodic-hexagon.mesh'), '-p', 0, '-rs', 2, '-dt', 0.005, '-tf', 2.5, '-ho', 3,
      '-lo', 3, '-fct', 2, '-no-vis'], 0.3888354875, 0.9755502191),
    # -lo 4 (subcell RD) rows: out_baseline.dat:41-69; -ho 2 ... -pa rows :76-100 run verbatim
    (['-m', mesh('periodic-hexagon.mesh'), '-p', 0, '-rs', 2, '-dt', 0.005, '-tf', 2.5, '-ho', 3,
      '-lo', 4, '-fct', 2, '-no-vis'], 0.3888354875, 0.9850024108),
    (['-m', mesh('periodic-square.mesh'), '-p', 5, '-rs', 3, '-dt', 0.004, '-tf', 0.8, '-ho', 3,
      '-lo', 4, '-fct', 2, '-no-vis'], 0.1623263888, 0.7145371968),
    (['-m', mesh('periodic-cube.mesh'), '-p', 0, '-rs', 1, '-o', 2, '-dt', 0.015, '-tf', 2, '-ho', 3,
      '-lo', 4, '-fct', 2, '-no-vis'], 0.9607429525, 0.9334903111),
    (['-m', mesh('inline-quad.mesh'), '-p', 14, '-rs', 1, '-dt', 0.0015, '-tf', 0.75, '-ho', 3,
      '-lo', 4, '-fct', 2, '-no-vis'], 0.0847954729, 0.7581364675),
    (['-m', mesh('cube01_hex.mesh'), '-p', 10, '-rs', 1, '-o', 2, '-dt', 0.02, '-tf', 0.7, '-ho', 3,
      '-lo', 4, '-fct', 2, '-no-vis'], 0.1197299801, 0.9997499683),
    (['-m', mesh('periodic-hexagon.mesh'), '-p', 0, '-rs', 2, '-dt', 0.005, '-tf', 2.5, '-ho', 2,
      '-lo', 3, '-fct', 2, '-pa', '-no-vis'], 0.3888354875, 0.9755502191),
    # -ho 1 -lo 2 -fct 2 (Neumann HO + preconditioned discrete upwinding): out_baseline.dat:4-32
    (['-m', mesh('inline-quad.mesh'), '-p', 14, '-rs', 1, '-dt', 0.0015, '-tf', 0.75, '-ho', 1,
      '-lo', 2, '-fct', 2, '-no-vis'], 0.08479546635, 0.8262759545),
    (['-m', mesh('cube01_hex.mesh'), '-p', 10, '-rs', 1, '-o', 2, '-dt', 0.02, '-tf', 0.7, '-ho', 1,
      '-lo', 2, '-fct', 2, '-no-vis'], 0.1197299711, 0.9998930413),
    (['-m', mesh('periodic-hexagon.mesh'), '-p', 0, '-rs', 2, '-dt', 0.005, '-tf', 2.5, '-ho', 1,
      '-lo', 2, '-fct', 2, '-no-vis'], 0.3888354875, 0.9854644631),
    (['-m', mesh('periodic-square.mesh'), '-p', 5, '-rs', 3, '-dt', 0.004, '-tf', 0.8, '-ho', 1,
      '-lo', 2, '-fct', 2, '-no-vis'], 0.1623263888, 0.7742737139),
    (['-m', mesh('periodic-cube.mesh'), '-p', 0, '-rs', 1, '-o', 2, '-dt', 0.015, '-tf', 2, '-ho', 1,
      '-lo', 2, '-fct', 2, '-no-vis'], 0.9607429525, 0.9724537077),
    # autotest/out_baseline.dat:177-180 and :103-106 (mass and max, 10 digits)
    (['-m', mesh('periodic-cube.mesh'), '-p', 0, '-rs', 1, '-o', 2, '-dt', 0.015, '-tf', 2, '-ho', 3,
      '-lo', 1, '-fct', 1, '-no-vis'], 0.9607429525, 0.9984668427),
    (['-m', mesh('periodic-cube.mesh'), '-p', 0, '-rs', 1, '-o', 2, '-dt', 0.015, '-tf', 2, '-ho', 3,
      '-lo', 3, '-fct', 2, '-no-vis'], 0.9607429525, 0.9202929163),
    # autotest/out_baseline.dat:172-175, :98-101 (2D periodic square)
    (['-m', mesh('periodic-square.mesh'), '-p', 5, '-rs', 3, '-dt', 0.004, '-tf', 0.8, '-ho', 3,
      '-lo', 1, '-fct', 1, '-no-vis'], 0.1623263888, 0.787875182),
    (['-m', mesh('periodic-square.mesh'), '-p', 5, '-rs', 3, '-dt', 0.004, '-tf', 0.8, '-ho', 3,
      '-lo', 3, '-fct', 2, '-no-vis'], 0.1623263888, 0.6374820899),
    # remap rows :152-155, :78-81
    (['-m', mesh('inline-quad.mesh'), '-p', 14, '-rs', 1, '-dt', 0.0015, '-tf', 0.75, '-ho', 3,
      '-lo', 1, '-fct', 1, '-no-vis'], 0.08479546845, 0.905654904),
    (['-m', mesh('inline-quad.mesh'), '-p', 14, '-rs', 1, '-dt', 0.0015, '-tf', 0.75, '-ho', 3,
      '-lo', 3, '-fct', 2, '-no-vis'], 0.08479546775, 0.7779015453),
]


@pytest.mark.gpu
@pytest.mark.parametrize('args,mass,umax', CLI_KNOWN)
def test_cli_reproduces_reference_known_answers(args, mass, umax):
    rc, out, err = run_cli(*args)
    assert rc == 0, err
    r = parse(out)
    assert float('%.10g' % r['mass']) == float('%.10g' % mass)
    if umax is not None:
        assert float('%.10g' % r['umax']) == umax


@pytest.mark.gpu
def test_cli_baseline_config_c1_matches_oracle():
    """BASELINE.json configs[0]: 2D periodic-square, order 2, -rs 3, RK3, DU + ClipScale."""
    rc, out, err = run_cli('-m', mesh('periodic-square.mesh'), '-p', 5, '-rs', 3, '-o', 2, '-dt', 0.004,
                           '-tf', 0.8, '-s', 3, '-ho', 3, '-lo', 1, '-fct', 2, '-no-vis', '-vb')
    assert rc == 0, err
    r = parse(out)
    run = oracle_run('periodic-square.mesh', problem=5, rs_levels=3, order=2, dt=0.004, t_final=0.8,
                     ode_solver=3, ho_type=3, lo_type=1, fct_type=2)
    run.run()
    assert r['n'] == run.u.size
    assert abs(r['mass'] - run.final_mass) < 1e-9 * abs(run.final_mass)      # printed to 10 digits
    assert abs(r['umax'] - run.final_max) < 1e-9 * abs(run.final_max)
    assert r['loss'] < 1e-12
    assert 'time step:' in out and 'residual:' in out


@pytest.mark.gpu
@pytest.mark.parametrize('mono,order,meshname,problem', [(1, 2, 'periodic-square.mesh', 5),
                                                         (2, 3, 'periodic-square.mesh', 1),
                                                         (1, 1, 'inline-quad.mesh', 4)])
def test_cli_mono_matches_oracle(mono, order, meshname, problem):
    """-mono 1/2 (MonoRDSolver, no smoothness indicator) through the C++ MonolithicSolver mirror:
    the driver computes the scale factors itself (remhos_mono.cpp:40-57)"""
    rc, out, err = run_cli('-m', mesh(meshname), '-p', problem, '-rs', 2, '-o', order, '-dt', 0.002,
                           '-tf', 0.02, '-s', 3, '-mono', mono, '-no-vis', '-ms', 10)
    assert rc == 0, err
    r = parse(out)
    run = oracle_run(meshname, problem=problem, rs_levels=2, order=order, dt=0.002, t_final=0.02,
                     ode_solver=3, mono_type=mono, max_steps=10)
    run.run()
    assert abs(r['mass'] - run.final_mass) < 1e-9 * abs(run.final_mass)
    assert abs(r['umax'] - run.final_max) < 1e-9 * abs(run.final_max)


def test_mono_flag_validation():
    rc, _, err = run_cli('-m', 'x', '-mono', '3')
    assert rc == 134 and 'Verification failed' in err
    rc, _, err = run_cli('-m', 'x', '-mono', '2', '-o', '1')
    assert rc == 134 and 'Subcell schemes require' in err


@pytest.mark.gpu
@pytest.mark.parametrize('args,mass,umax', [
    # autotest/out_baseline.dat:212-220 (= README.md runs 12, 13): the reference's known answers
    # for the monolithic solver with smoothness indicator, steady-state stop
    (['-m', mesh('inline-quad.mesh'), '-p', 7, '-rs', 3, '-o', 1, '-dt', 0.01, '-tf', 20, '-mono', 1,
      '-si', 2, '-no-vis'], 0.1570667907, 0.9987771164),
    (['-m', mesh('inline-quad.mesh'), '-p', 6, '-rs', 2, '-o', 1, '-dt', 0.01, '-tf', 20, '-mono', 1,
      '-si', 1, '-no-vis'], 0.3182739921, 1.0)])
def test_cli_mono_si_reproduces_reference_known_answers(args, mass, umax):
    rc, out, err = run_cli(*args)
    assert rc == 0, err
    r = parse(out)
    assert float('%.10g' % r['mass']) == mass
    assert float('%.10g' % r['umax']) == umax


@pytest.mark.gpu
def test_cli_fct_project_dtc_reproduces_reference_known_answer():
    """autotest/out_baseline.dat:207-210 ("BLAST sharpening test"): -fct 4 -bt 1 -dtc 1"""
    rc, out, err = run_cli('-m', mesh('periodic-square.mesh'), '-p', 5, '-rs', 3, '-dt', 0.01, '-tf', 0.8,
                           '-ho', 3, '-lo', 5, '-fct', 4, '-bt', 1, '-dtc', 1, '-no-vis')
    assert rc == 0, err
    r = parse(out)
    assert float('%.10g' % r['mass']) == 0.1623263888
    assert float('%.10g' % r['umax']) == 0.2863317261


@pytest.mark.gpu
def test_cli_fct_project_dtc_remap_reproduces_reference_known_answer():
    """autotest/out_baseline.dat:202-205 ("Pacman remap auto-dt"): remap, CFL dt, -fct 4 -bt 1 -dtc 1"""
    rc, out, err = run_cli('-m', mesh('inline-quad.mesh'), '-p', 14, '-rs', 1, '-dt', -1, '-tf', 0.75,
                           '-ho', 3, '-lo', 5, '-fct', 4, '-bt', 1, '-dtc', 1, '-no-vis')
    assert rc == 0, err
    r = parse(out)
    assert float('%.10g' % r['mass']) == 0.08479612805
    assert float('%.6g' % r['loss']) == 6.61247e-07


@pytest.mark.gpu
def test_cli_star_q2_curved_mesh_known_answer(tmp_path):
    """remhos_tests.cpp:88-91 (#8): curved Q2 `Quadratic` mesh (tests/golden/star_q2.json), PA remap"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    from make_star_q2 import materialise
    p = materialise(str(tmp_path / 'star-q2.mesh'))
    rc, out, err = run_cli('-m', p, '-pa', '-p', 14, '-rs', 1, '-o', 3, '-dt', -1, '-tf', 0.5, '-ho', 3,
                           '-lo', 5, '-fct', 2, '-ms', 5, '-no-vis')
    assert rc == 0, err
    assert float('%.10g' % parse(out)['mass']) == float('%.10g' % 0.8069675186775516)


@pytest.mark.gpu
def test_cli_config5_star_q3_mono_subcell_matches_oracle(tmp_path):
    """BASELINE config 5: monolithic subcell residual distribution (-mono 2) on the refined star-q3
    mesh (curved, legacy `Cubic` nodes), solid-body rotation -p 4: CLI on the GPU vs the oracle, and
    the bounds verdict (no reference number exists for this configuration)"""
    import sys
    from remhos_oracle import driver
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    from make_star_q2 import materialise, HERE
    p = materialise(str(tmp_path / 'star-q3.mesh'), os.path.join(HERE, 'star_q3.json'))
    rc, out, err = run_cli('-m', p, '-p', 4, '-rs', 2, '-o', 2, '-dt', 0.002, '-tf', 0.1, '-mono', 2,
                           '-ms', 10, '-no-vis')
    assert rc == 0, err
    r = parse(out)
    run = driver.Run(driver.Options(mesh_file=p, problem=4, rs_levels=2, order=2, dt=0.002, t_final=0.1,
                                    ode_solver=3, mono_type=2, max_steps=10))
    run.run()
    assert abs(r['mass'] - run.final_mass) < 1e-9 * abs(run.final_mass)
    assert abs(r['umax'] - run.final_max) < 1e-9
    assert run.u.min() > -1e-12 and r['umax'] < 1.0 + 1e-12


def test_new_flag_combinations_validated():
    """-si acts on -mono, -fct 2 and -fct 3; -fct 4 needs assembled matrices; -fct 3 is transport only;
    -dtc 1 needs an FCT solver; -ps needs remap mode"""
    for flags, msg in ((['-si', '1', '-ho', '3', '-o', '1'], 'smoothness indicators (-si) act on'),
                       (['-si', '3', '-mono', '1', '-o', '1'], 'Bad smoothness indicator id!'),
                       (['-ho', '3', '-lo', '5', '-fct', '4', '-pa'], 'FCTProject needs the assembled'),
                       (['-ho', '3', '-lo', '5', '-fct', '3', '-p', '10'], '-fct 3 (NonlinearPenalty) is built for transport'),
                       (['-ho', '3', '-lo', '5', '-fct', '5'], 'FCT solver type must be 0 .. 4'),
                       (['-ho', '3', '-lo', '5', '-fct', '2', '-ps'], 'Products are processed only in remap mode.'),
                       (['-ho', '3', '-dtc', '1'], '-dtc 1 needs an FCT solver'),
                       (['-ho', '3', '-lo', '5', '-fct', '2', '-dtc', '2'], 'time step control must be')):
        rc, _, err = run_cli('-m', 'x', *flags)
        assert rc == 134 and msg in err, (flags, err)


@pytest.mark.gpu
@pytest.mark.parametrize('order,bt,dt', [(3, 0, '0.004'), (4, 1, '-1'), (2, 0, '-1')])
def test_cli_decomposed_matches_single_gpu(order, bt, dt):
    """`remhos -gpus 2` (one process per GPU, started by the driver itself as mpirun -np 2 starts the
    reference; halo puts, in-kernel halo wait and ncclAllReduce behind the C ABI) prints the same
    unknown count, final mass and maximum as the single-GPU run; with -dt -1 the CFL time step is the
    MPI_MIN over the ranks (remhos.cpp:551)."""
    torch = pytest.importorskip('torch')
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    args = ['-m', mesh('periodic-cube.mesh'), '-p', 0, '-rs', 2, '-o', order, '-dt', dt, '-tf', 0.05, '-ho', 3,
            '-lo', 5, '-fct', 2, '-pa', '-bt', bt, '-no-vis']
    rc1, out1, err1 = run_cli(*args)
    assert rc1 == 0, err1
    rc2, out2, err2 = run_cli(*args, '-gpus', 2)
    assert rc2 == 0, out2[-2000:] + err2[-2000:]
    a, b = parse(out1), parse(out2)
    assert a['n'] == b['n']
    assert abs(a['mass'] - b['mass']) <= 1e-10 * abs(a['mass'])     # the 10 printed digits
    assert abs(a['umax'] - b['umax']) <= 1e-10 * abs(a['umax'])
    assert out2.count('Final mass u:') == 1                          # only the root rank talks


@pytest.mark.gpu
def test_cli_default_mesh_has_epm_elements():
    """-m default: PartitionMPI builds world * elem_per_mpi elements AFTER the -rp refinements
    (verified by the reference at remhos.cpp:466-471); -rs applies to file meshes only (:448-449)."""
    for extra in ([], ['-rs', 3]):
        rc, out, err = run_cli('-m', 'default', '-dim', 3, '-epm', 64, '-rp', 1, '-o', 2, '-p', 0, '-dt', 0.002,
                               '-tf', 0.004, '-ho', 3, '-lo', 5, '-fct', 2, '-no-vis', *extra)
        assert rc == 0, err
        assert parse(out)['n'] == 64 * 27
    rc, _, err = run_cli('-m', 'default', '-dim', 3, '-epm', 12, '-rp', 1, '-o', 2, '-ho', 3, '-lo', 5, '-fct', 2)
    assert rc == 134 and 'Mesh generation error' in err     # 12 is not a multiple of 8


# Product-field remap through the GPU CLI: the three reference known answers for -ps
# (autotest/out_baseline.dat:187-200: -fct 1 -s 1, -fct 2 -s 12, -fct 4 -s 13), 10 printed digits
PS_ROWS = [
    (['-ho', 3, '-lo', 1, '-fct', 1, '-ps', '-s', 1], dict(mass_us=0.1815368098, loss_us=0.00192894)),
    (['-ho', 1, '-lo', 5, '-fct', 2, '-ps', '-s', 12], dict(mass_us=0.1796076412, loss_us=2.31348e-07)),
    (['-ho', 3, '-lo', 5, '-fct', 4, '-ps', '-s', 13], dict(mass=0.08980386855, mass_us=0.179607829)),
]


@pytest.mark.gpu
@pytest.mark.parametrize('flags,exp', PS_ROWS, ids=['fct1-s1', 'fct2-s12', 'fct4-s13'])
def test_cli_product_remap_reproduces_reference_known_answers(flags, exp):
    rc, out, err = run_cli('-no-vis', '-m', mesh('inline-quad.mesh'), '-p', 14, '-rs', 2, '-dt', 0.005, '-tf', 0.75,
                           *flags)
    assert rc == 0, err
    g = lambda pat: float(re.search(pat, out).group(1))
    got = dict(mass=g(r'Final mass u:\s+(\S+)'), mass_us=g(r'Final mass us:\s+(\S+)'),
               loss_us=g(r'Mass loss us:\s+(\S+)'), s_max=g(r'Max value s:\s+(\S+)'))
    for k, v in exp.items():
        tol = 5e-10 if k.startswith('mass') else 2e-5          # 10 digits; losses are printed with 6
        assert abs(got[k] - v) <= tol * abs(v), (k, got[k], v)
    assert 1.0 <= got['s_max'] <= 3.0 + 1e-6                   # s0 = 2 + sin sin stays in its range


@pytest.mark.gpu
def test_cli_product_remap_rejections():
    rc, _, err = run_cli('-m', mesh('periodic-square.mesh'), '-p', 0, '-ho', 3, '-lo', 5, '-fct', 2, '-ps')
    assert rc == 134 and 'Products are processed only in remap mode.' in err      # remhos.cpp:1850
    rc, _, err = run_cli('-m', mesh('inline-quad.mesh'), '-p', 14, '-ho', 3, '-lo', 5, '-fct', 4, '-ps', '-dtc', 1)
    assert rc == 134 and 'Automatic time step is not implemented for product remap.' in err


@pytest.mark.gpu
def test_cli_nonlinear_penalty_matches_oracle():
    """-fct 3 through the driver (README.md:210's combination -ho 1 -lo 4 -fct 3 on a short run): no
    reference number exists; the oracle's transcription is the checker"""
    args = dict(problem=5, rs_levels=2, order=2, dt=0.002, t_final=0.02, ho_type=3, lo_type=3, fct_type=3)
    run = oracle_run('periodic-square.mesh', **args)
    run.run()
    rc, out, err = run_cli('-no-vis', '-m', mesh('periodic-square.mesh'), '-p', 5, '-rs', 2, '-o', 2, '-dt', 0.002,
                           '-tf', 0.02, '-ho', 3, '-lo', 3, '-fct', 3)
    assert rc == 0, err
    got = parse(out)
    # (loose: the penalty solver amplifies one-ulp input differences, see tests/test_gpu_penalty.py)
    assert abs(got['mass'] - run.final_mass) < 1e-5 * abs(run.final_mass)
    assert abs(got['umax'] - run.final_max) < 1e-6


@pytest.mark.gpu
def test_cli_save_and_visit_write_mfem_files(tmp_path):
    """-save: meshHO_init/final.mesh + sltn_init/final.gf (remhos.cpp:1016-1030,1366-1380); -visit: a
    VisItDataCollection "Remhos" per visualisation step (:1034-1043,1323-1328).  Remap run, so the final
    mesh is the moved one."""
    import json
    import remhos_b200 as rb
    args = ['-no-vis', '-m', mesh('inline-quad.mesh'), '-p', 14, '-rs', 1, '-o', 2, '-dt', 0.01, '-tf', 0.1, '-ho', 3,
            '-lo', 5, '-fct', 2, '-save', '-visit', '-vs', 5]
    p = subprocess.run([EXE] + [str(a) for a in args], capture_output=True, text=True, timeout=180, cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr
    names = sorted(os.listdir(tmp_path))
    for f in ('meshHO_init.mesh', 'meshHO_final.mesh', 'meshLO_init.mesh', 'meshLO_final.mesh', 'sltn_init.gf',
              'sltn_final.gf', 'Remhos_000000.mfem_root', 'Remhos_000005.mfem_root', 'Remhos_000010.mfem_root'):
        assert f in names, names
    # no subcell scheme: subcell_mesh = &pmesh (remhos.cpp:870), the LO files repeat the HO ones
    assert open(tmp_path / 'meshLO_final.mesh').read() == open(tmp_path / 'meshHO_final.mesh').read()
    m0 = rb.Mesh.load(str(tmp_path / 'meshHO_init.mesh'))
    m1 = rb.Mesh.load(str(tmp_path / 'meshHO_final.mesh'))
    assert m0.ne == m1.ne == 64 and m0.geom_order == 2
    assert np.abs(m0.nodes() - m1.nodes()).max() > 1e-3            # the remap mesh moved
    ref = rb.Mesh.load(mesh('inline-quad.mesh')).refine(1)
    ref.set_curvature(2)
    assert np.abs(m0.nodes() - ref.nodes()).max() < 1e-7           # 8 digits written
    gf = open(tmp_path / 'sltn_final.gf').read().split('\n')
    assert gf[1] == 'FiniteElementCollection: L2_T2_2D_P2'
    vals = np.array([float(v) for v in gf[5:] if v.strip()])
    assert vals.size == 64 * 9 and abs(vals.max() - parse(p.stdout)['umax']) < 1e-7
    root = json.load(open(tmp_path / 'Remhos_000010.mfem_root'))['dsets']['main']
    assert root['cycle'] == 10 and root['domains'] == 1 and 'solution' in root['fields']
    assert os.path.exists(tmp_path / (root['mesh']['path'] % 0)) and os.path.exists(tmp_path / (root['fields']['solution']['path'] % 0))
    assert rb.Mesh.load(str(tmp_path / (root['mesh']['path'] % 0))).ne == 64


@pytest.mark.gpu
def test_cli_save_writes_the_subcell_mesh(tmp_path):
    """-save with a subcell scheme (-lo 4): meshLO_*.mesh is ParMesh::MakeRefined(pmesh, order, ClosedUniform),
    moved with the HO mesh in remap mode (remhos.cpp:801, 1021-1026, 1371-1376)"""
    import remhos_b200 as rb
    args = ['-no-vis', '-m', mesh('inline-quad.mesh'), '-p', 14, '-rs', 1, '-o', 3, '-dt', 0.01, '-tf', 0.05, '-ho', 3,
            '-lo', 4, '-fct', 2, '-save']
    p = subprocess.run([EXE] + [str(a) for a in args], capture_output=True, text=True, timeout=180, cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr
    ho0, ho1 = rb.Mesh.load(str(tmp_path / 'meshHO_init.mesh')), rb.Mesh.load(str(tmp_path / 'meshHO_final.mesh'))
    lo0, lo1 = rb.Mesh.load(str(tmp_path / 'meshLO_init.mesh')), rb.Mesh.load(str(tmp_path / 'meshLO_final.mesh'))
    assert ho0.ne == 64 and lo0.ne == lo1.ne == 64 * 9 and lo0.geom_order == 1
    assert lo0.nv == (8 * 3 + 1) ** 2
    assert np.abs(lo0.nodes() - ho0.make_refined(3).nodes()).max() < 1e-7        # 8 digits written
    assert np.abs(lo1.nodes() - ho1.make_refined(3).nodes()).max() < 1e-7
    assert np.abs(lo1.nodes() - lo0.nodes()).max() > 1e-3                       # moved


DECOMP_GENERAL = [
    ('periodic-square.mesh', ['-p', 5, '-rs', 3, '-o', 2, '-dt', 0.004, '-tf', 0.04, '-ho', 3, '-lo', 3, '-fct', 2]),
    ('periodic-square.mesh', ['-p', 5, '-rs', 3, '-o', 2, '-dt', 0.004, '-tf', 0.04, '-ho', 3, '-lo', 1, '-fct', 2, '-s', 2]),
    ('periodic-hexagon.mesh', ['-p', 0, '-rs', 2, '-o', 2, '-dt', 0.005, '-tf', 0.05, '-ho', 1, '-lo', 2, '-fct', 2]),
    ('periodic-cube.mesh', ['-p', 0, '-rs', 1, '-o', 2, '-dt', 0.01, '-tf', 0.05, '-ho', 3, '-lo', 4, '-fct', 2]),
    ('periodic-square.mesh', ['-p', 5, '-rs', 3, '-o', 3, '-dt', 0.004, '-tf', 0.04, '-ho', 3, '-lo', 5, '-fct', 4, '-s', 13]),
    ('periodic-cube.mesh', ['-p', 1, '-rs', 1, '-o', 3, '-dt', 0.01, '-tf', 0.05, '-ho', 3, '-lo', 5, '-fct', 2, '-s', 4, '-pa']),
    ('periodic-square.mesh', ['-p', 5, '-rs', 3, '-o', 2, '-dt', 0.01, '-tf', 0.08, '-ho', 3, '-lo', 5, '-fct', 4, '-bt', 1, '-dtc', 1]),
    ('periodic-square.mesh', ['-p', 5, '-rs', 2, '-o', 2, '-dt', 0.004, '-tf', 0.02, '-ho', 3, '-lo', 3, '-fct', 0]),
    ('inline-quad.mesh', ['-p', 4, '-rs', 2, '-o', 2, '-dt', 0.002, '-tf', 0.02, '-ho', 3, '-lo', 1, '-fct', 2]),
    # remap mode on a decomposed mesh: fused stage path (3D, re-assembly every stage) and solver by solver (2D)
    ('cube01_hex.mesh', ['-p', 10, '-rs', 2, '-o', 2, '-dt', -1, '-tf', 0.5, '-ho', 3, '-lo', 5, '-fct', 2, '-ms', 6, '-pa']),
    ('inline-quad.mesh', ['-p', 14, '-rs', 2, '-o', 3, '-dt', -1, '-tf', 0.5, '-ho', 3, '-lo', 3, '-fct', 2, '-ms', 8]),
    ('inline-quad.mesh', ['-p', 14, '-rs', 2, '-o', 2, '-dt', 0.004, '-tf', 0.04, '-ho', 3, '-lo', 5, '-fct', 4, '-s', 12]),
    # FluxBasedFCT across rank boundaries: ghost-face blocks formed locally, R+ / R- exchanged (remhos_fct.cpp:406-409)
    ('periodic-square.mesh', ['-p', 5, '-rs', 3, '-o', 2, '-dt', 0.004, '-tf', 0.04, '-ho', 3, '-lo', 1, '-fct', 1]),
    ('periodic-hexagon.mesh', ['-p', 0, '-rs', 2, '-o', 3, '-dt', 0.005, '-tf', 0.05, '-ho', 1, '-lo', 2, '-fct', 1, '-s', 2]),
    ('periodic-cube.mesh', ['-p', 1, '-rs', 1, '-o', 2, '-dt', 0.01, '-tf', 0.05, '-ho', 3, '-lo', 1, '-fct', 1]),
    ('inline-quad.mesh', ['-p', 4, '-rs', 2, '-o', 2, '-dt', 0.002, '-tf', 0.02, '-ho', 3, '-lo', 2, '-fct', 1]),
]


@pytest.mark.gpu
@pytest.mark.parametrize('mesh_name,flags', DECOMP_GENERAL,
                         ids=['RD-clipscale', 'DU-clipscale-rk2', 'hexagon-Neumann-DUprec', 'cube-RDsub', 'idp3-fctproject',
                              'cube-rk4-rotation', 'dtc', 'LO-only', 'inline-quad-boundary', 'remap-3d-fused',
                              'remap-2d-RD', 'remap-2d-idp2-fctproject', 'fluxfct-DU', 'fluxfct-hexagon-Neumann-rk2',
                              'fluxfct-cube-rotation', 'fluxfct-inline-quad-boundary'])
def test_cli_decomposed_solver_by_solver(mesh_name, flags):
    """Decomposed runs of the matrix-based / unfused solver combinations (DU, RD, subcell RD, Neumann,
    FCTProject, IDP and RK4 time stepping, automatic dt): `remhos -gpus 2` against the single-GPU run.
    One halo exchange (face traces + ghost (min,max)) per operator evaluation replaces the
    ExchangeFaceNbrData calls of remhos_lo.cpp:57,131 / remhos_tools.cpp:399 / remhos.cpp:1813."""
    torch = pytest.importorskip('torch')
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    args = ['-no-vis', '-m', mesh(mesh_name)] + flags
    rc1, out1, err1 = run_cli(*args)
    assert rc1 == 0, err1
    rc2, out2, err2 = run_cli(*args, '-gpus', 2)
    assert rc2 == 0, out2[-2000:] + err2[-2000:]
    a, b = parse(out1), parse(out2)
    assert a['n'] == b['n']
    assert abs(a['mass'] - b['mass']) <= 1e-10 * abs(a['mass'])
    assert abs(a['umax'] - b['umax']) <= 1e-10 * max(abs(a['umax']), 1e-300)
    if '-dtc' in flags:
        assert out1.count('Repeat / decrease dt') == out2.count('Repeat / decrease dt')


@pytest.mark.gpu
def test_cli_decomposed_rejections():
    torch = pytest.importorskip('torch')
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    rc, out, err = run_cli('-no-vis', '-m', mesh('periodic-square.mesh'), '-p', 5, '-ho', 3, '-lo', 1, '-fct', 2, '-mono', 1, '-gpus', 2)
    assert rc == 134 and 'decomposed runs' in err
    rc, out, err = run_cli('-no-vis', '-m', mesh('inline-quad.mesh'), '-p', 14, '-ho', 3, '-lo', 1, '-fct', 1, '-gpus', 2)
    assert rc == 134 and 'decomposed remap runs' in err
