"""BASELINE configs[2] (mgModeErrorScaling, src/multigrid.c:1734-1851): the multigrid solve of rho = k^2 sin(kx)
(gFillSin, src/grid.c:1563-1603, with the reference's PI = 3.14159265, src/grid.h:584) against the analytic solution sin(kx)
(gFillSinSol :1610).  The 7-point Laplacian is second order: the discrete solution is (k^2 / (2 - 2 cos k)) sin(kx), so
the error is k^2/12 + O(k^4) and falls by 4 per doubling of N.  Also: the device result equals that closed form to the
solver tolerance, and the V-cycle count matches the oracle's solve of the same problem."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu


def test_error_against_analytic_solution_is_second_order(gpu_lib):
    import mg_bench
    recs, orders = mg_bench.error_scaling(gpu_lib, (16, 32, 64))
    for r in recs:
        N = int(r["N"])
        k = 2 * mg_bench.PI / N
        amp = k * k / (2 - 2 * np.cos(k)) - 1          # error amplitude of the discrete solution; rms of a sine = amp/sqrt(2)
        assert r["barRes_last"] <= 1e-10
        assert abs(r["rms_error_vs_analytic"] - amp / np.sqrt(2)) <= 2e-3 * amp, (N, r["rms_error_vs_analytic"], amp / np.sqrt(2))
    assert len(orders) == 2
    for o in orders:
        assert 1.95 <= o <= 2.05, orders
