"""Non-periodic boundaries on the device: gBnd with gDirichlet / gNeumann edges (src/grid.c:921-1023),
mgGS3D and one mgVRecursive V-cycle with them, gSetBndSlices and mgRestrictBnd, through the C-ABI on one rank and on four
thread ranks, against the oracle (which tests/test_oracle_bnd.py pins to the reference's own sources).  Scenario:
tests/bnd_common.py.  phi of every rank, ghost layers included, after every stage: <= 1e-12 relative (bit-identical
arithmetic per node; gNeutralizeGrid's mean is summed in another order)."""
import ctypes as C

import numpy as np
import pytest

import bnd_common as bc
from helpers import small_cfg
from pinc_b200 import abi, sim
from test_oracle_bnd import run_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sub,boundaries", bc.CASES)
def test_device_boundaries_match_oracle(gpu_lib, sub, boundaries):
    L = gpu_lib
    text, cfg = small_cfg("warm", **bc.overrides(sub, boundaries))
    init = bc.fields(cfg)
    O = run_oracle(cfg, init)
    W = sim.World(cfg)
    got = {"A": {}, "B": {}, "C": {}}
    try:
        def setup(r, st):
            g = st.phi.contents
            nmax = int(g.sizeProd[4]) // int(g.size[0])
            b = np.ctypeslib.as_array(g.bndSlice, shape=(8 * nmax,))
            m = st.mpi.contents
            vals = bc.slice_values(cfg, r, nmax)
            for d in range(1, 4):
                for bd in (d, d + 4):
                    edge = m.subdomain[d - 1] == 0 if bd < 4 else m.subdomain[d - 1] == m.nSubdomains[d - 1] - 1
                    if edge and g.bnd[bd] != abi.PERIODIC:
                        assert np.all(b[bd * nmax:(bd + 1) * nmax] == (1.0 if g.bnd[bd] == abi.DIRICHLET else 2.0))     # gSetBndSlices
                        b[bd * nmax:(bd + 1) * nmax] = vals[bd * nmax:(bd + 1) * nmax]
                    else:
                        b[bd * nmax:(bd + 1) * nmax] = 0
            abi.grid_array(st.phi.contents).reshape(-1)[:] = init[r][0]
            abi.grid_array(st.rho.contents).reshape(-1)[:] = init[r][1]
            L.pincSyncGridToDevice(st.phi)               # values and boundary slices
            L.pincSyncGridToDevice(st.rho)
            L.mgRestrictBnd(st.solver.contents.mgPhi)
            for q in range(cfg.mgLevels):                # every level's boundary values as the oracle has them
                gq = st.solver.contents.mgPhi.contents.grids[q].contents
                nq = int(gq.sizeProd[4])
                mine = np.ctypeslib.as_array(gq.bndSlice, shape=(8 * nq,))
                for bd in (1, 2, 3, 5, 6, 7):
                    assert np.array_equal(mine[bd * nq:(bd + 1) * nq], O["bnd"][r][q][bd * nq:(bd + 1) * nq]), (r, q, bd)
        W.run(setup)

        def stage(name, fn):
            W.run(fn)
            for r in range(cfg.nRanks):
                got[name][r] = W.grid(r, "phi").reshape(-1)
        stage("A", lambda r, st: L.gBnd(st.phi, st.mpi))
        stage("B", lambda r, st: L.mgGS3D(st.phi, st.rho, 2, st.mpi))
        sol = lambda st: st.solver.contents
        stage("C", lambda r, st: L.mgVRecursive(0, cfg.mgLevels - 1, 0, sol(st).mgRho, sol(st).mgPhi, sol(st).mgRes, st.mpi))
        for name in "ABC":
            for r in range(cfg.nRanks):
                a, b = got[name][r], O[name][r]
                err = np.abs(a - b).max() / np.abs(b).max()
                assert err <= 1e-12, (name, r, err)
    finally:
        W.close()
