"""The solver's other select() targets on the device: mgVRegular, mgW (src/multigrid.c:1559-1683) and the
"jacobian" smoother, one kernel per reference call, against the oracle (which tests/test_oracle_cycles.py pins to the
reference's own mgVRegular/mgW; the Jacobi smoother is unpinned, see tests/cycles_common.py).  One cycle, one and four thread
ranks; phi (ghost layers included) and the true nodes of res <= 1e-12.  Also: mgSolve with multigrid.cycle = mgW set through
the solver struct runs on the ops path and converges to the tolerance with the V-cycle count of the oracle."""
import ctypes as C

import numpy as np
import pytest

import cycles_common as cc
from helpers import small_cfg
from pinc_b200 import abi, sim
from test_oracle_cycles import run_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sub,cycle,smoothers", cc.CASES + cc.JACOBI_CASES)
def test_device_cycles_match_oracle(gpu_lib, sub, cycle, smoothers):
    L = gpu_lib
    text, cfg = small_cfg("warm", **cc.overrides(sub))
    init = cc.fields(cfg)
    # ghost layers of phi consistent with its true nodes at the start, as they are whenever a solve begins (mgVRecursive's
    # first smoother call reads them before any gHaloOp; the multi-rank smoother reads periodic images instead of ghosts in
    # the dimensions that are not decomposed, which is the same thing only then)
    O = run_oracle(cfg, init, cycle, smoothers, halo_first=True)
    fn = {0: C.cast(L.mgGS3D, C.c_void_p), 1: C.cast(L.mgJacob3D, C.c_void_p)}
    W = sim.World(cfg)
    try:
        def setup(r, st):
            mg = st.solver.contents.mgRho.contents
            mg.preSmooth, mg.postSmooth, mg.coarseSolv = fn[smoothers[0]], fn[smoothers[1]], fn[smoothers[2]]
            abi.grid_array(st.phi.contents).reshape(-1)[:] = init[r][0]
            abi.grid_array(st.rho.contents).reshape(-1)[:] = init[r][1]
            L.pincSyncGridToDevice(st.phi)
            L.pincSyncGridToDevice(st.rho)
        W.run(setup)
        W.run(lambda r, st: L.gHaloOp(W.set_slice, st.phi, st.mpi, abi.TOHALO))
        sol = lambda st: st.solver.contents
        b = cfg.mgLevels - 1
        if cycle == "smoother":
            W.run(lambda r, st: L.mgJacob3D(st.phi, st.rho, 3, st.mpi))
        else:
            f = {"mgVRegular": L.mgVRegular, "mgW": L.mgW, "mgVRecursive": L.mgVRecursive}[cycle]
            W.run(lambda r, st: f(0, b, 0, sol(st).mgRho, sol(st).mgPhi, sol(st).mgRes, st.mpi))
        sz = tuple(t + 2 for t in cfg.trueSize)[::-1]
        for r in range(cfg.nRanks):
            a, ref = W.grid(r, "phi").reshape(-1), O[r][0]
            assert np.abs(a - ref).max() / np.abs(ref).max() <= 1e-12, (r, "phi")
            if cycle != "smoother":
                a, ref = W.grid(r, "res").reshape(sz)[1:-1, 1:-1, 1:-1], O[r][1].reshape(sz)[1:-1, 1:-1, 1:-1]
                assert np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-300) <= 1e-12, (r, "res")
    finally:
        W.close()
