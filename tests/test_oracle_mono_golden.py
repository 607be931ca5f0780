"""The oracle against the reference's known answers kept in tests/golden/known_answers.json
(numbers copied from autotest/out_baseline.dat; = README.md runs 12, 13 for the monolithic solver):
final mass and maximum (or mass loss) to the printed digits.  The monolithic-solver rows also pin
NonlinFluxLumping, the SmoothnessIndicator, the inflow boundary state (including the "high order
projection" of remhos.cpp:628-635 for problem 7) and the steady-state stopping rule
(remhos.cpp:1276-1295); the FCTProject rows pin ElementFCTProjection and -dtc 1; the product_remap
rows pin the oracle's restatement of the product-field remap (-ps: ComputeRatio, masked bounds,
CalcCompatibleLOProduct, ScaleProductBounds, CalcFCTProduct, IDP Runge-Kutta on the (u, us) pair)."""
import json
import os

import pytest

from helpers import oracle_run

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'known_answers.json')) as f:
    ROWS = json.load(f)['rows']
if os.environ.get('RMH_SLOW_TESTS', '0') != '1':
    ROWS = [r for r in ROWS if not r.get('slow')]


@pytest.mark.parametrize('row', ROWS, ids=[r['name'] for r in ROWS])
def test_known_answer(row):
    run = oracle_run(row['mesh'], **row['options'])
    run.run()
    if 'mass' in row:
        assert float('%.10g' % run.final_mass) == row['mass']
    if 'mass_us' in row:
        assert float('%.10g' % run.final_mass_us) == row['mass_us']
    if 'mass_loss_us' in row:
        assert float('%.6g' % abs(run.mass0_us - run.final_mass_us)) == row['mass_loss_us']
    if 'max' in row:
        assert float('%.10g' % run.u.max()) == row['max']
    if 'mass_loss' in row:
        assert float('%.6g' % abs(run.mass0 - run.final_mass)) == row['mass_loss']
    if row['options'].get('problem') in (6, 7):
        assert run.residual < 1e-12 and run.t >= 1.0       # stopped by the steady-state criterion
