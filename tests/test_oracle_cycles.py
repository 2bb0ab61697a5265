"""The oracle's mgVRegular / mgW / mgJacob3D against the reference's own functions run live from oracle/_ref (one thread per
rank).  Scenario: tests/cycles_common.py.  phi and res of the finest level after one cycle: <= 1e-13 relative (the only
difference is the order of gNeutralizeGrid's sum over ranks)."""
import ctypes as C

import numpy as np
import pytest

import cycles_common as cc
from helpers import small_cfg
from oracle import orc, ref
from pinc_b200 import abi

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libpinc_ref.so not built (needs /root/reference)")


def run_reference(text, cfg, init, cycle, smoothers):
    W = ref.RefWorld(text, cfg.nRanks)
    lib = W.lib
    cyc_args = [C.c_int, C.c_int, C.c_int, C.POINTER(abi.Multigrid), C.POINTER(abi.Multigrid), C.POINTER(abi.Multigrid), C.POINTER(abi.MpiInfo)]
    lib.mgVRegular.argtypes = cyc_args
    lib.mgW.argtypes = cyc_args
    lib.mgJacob3D.argtypes = [C.POINTER(abi.Grid), C.POINTER(abi.Grid), C.c_int, C.POINTER(abi.MpiInfo)]
    fn = {0: C.cast(lib.mgGS3D, C.c_void_p), 1: C.cast(lib.mgJacob3D, C.c_void_p)}

    def setup(r, st):
        mg = st.solver.contents.mgRho.contents
        mg.preSmooth, mg.postSmooth, mg.coarseSolv = fn[smoothers[0]], fn[smoothers[1]], fn[smoothers[2]]
        abi.grid_array(st.phi.contents).reshape(-1)[:] = init[r][0]
        abi.grid_array(st.rho.contents).reshape(-1)[:] = init[r][1]
    W.run_serial(setup)
    sol = lambda st: st.solver.contents
    b = cc.LEVELS - 1
    if cycle == "smoother":
        W.run(lambda r, st: lib.mgJacob3D(st.phi, st.rho, 3, st.mpi))
    else:
        f = {"mgVRegular": lib.mgVRegular, "mgW": lib.mgW, "mgVRecursive": lib.mgVRecursive}[cycle]
        W.run(lambda r, st: f(0, b, 0, sol(st).mgRho, sol(st).mgPhi, sol(st).mgRes, st.mpi))
    out = [(W.grid(r, "phi").reshape(-1).copy(), abi.grid_array(W.ranks[r].res.contents).reshape(-1).copy()) for r in range(cfg.nRanks)]
    W.close()
    return out


def run_oracle(cfg, init, cycle, smoothers, halo_first=False):
    O = orc.OrcWorld(cfg)
    O.lib.orc_mg_set_smoothers(O.mg, *smoothers)
    for r in range(cfg.nRanks):
        O.phi[r][:] = init[r][0]
        O.rho[r][:] = init[r][1]
    k = O._keep
    if halo_first:
        O.lib.orc_halo(C.byref(O.topo), k["phi"], orc.ip(O.size), 1, 0, 0)
    if cycle == "smoother":
        O.lib.orc_jacobi3d(C.byref(O.topo), k["phi"], k["rho"], orc.ip(O.size), 3, None, None)
    else:
        {"mgVRegular": O.lib.orc_mg_vregular, "mgW": O.lib.orc_mg_wcycle, "mgVRecursive": O.lib.orc_mg_vcycle}[cycle](O.mg, k["rho"], k["phi"], k["res"])
    return [(O.phi[r].copy(), O.res[r].copy()) for r in range(cfg.nRanks)]


@needs_ref
@pytest.mark.parametrize("sub,cycle,smoothers", cc.CASES)
def test_oracle_cycles_match_reference(sub, cycle, smoothers):
    text, cfg = small_cfg("warm", **cc.overrides(sub))
    init = cc.fields(cfg)
    R = run_reference(text, cfg, init, cycle, smoothers)
    O = run_oracle(cfg, init, cycle, smoothers)
    sz = tuple(t + 2 for t in cfg.trueSize)[::-1]
    for r in range(cfg.nRanks):
        err = np.abs(O[r][0] - R[r][0]).max() / np.abs(R[r][0]).max()
        assert err <= 1e-13, (r, "phi", err)                    # ghost layers included
        # res holds the prolonged correction of the level below; its ghost nodes are whatever the reference's last
        # interpolation pass left there (the next reader, gAddTo/gSubFrom + gHaloOp, never uses them): true nodes only
        a, b = O[r][1].reshape(sz)[1:-1, 1:-1, 1:-1], R[r][1].reshape(sz)[1:-1, 1:-1, 1:-1]
        err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
        assert err <= 1e-13, (r, "res", err)


def test_oracle_jacobi_closed_form():
    """One Jacobi sweep of the oracle on phi = cos(2 pi j / N), rho = 0: every true node becomes
    (2 cos(2 pi/N) + 4)/6 times its old value (periodic), minus the mean, which is zero."""
    text, cfg = small_cfg("warm", **cc.overrides("1,1,1"))
    O = orc.OrcWorld(cfg)
    sz = tuple(t + 2 for t in cfg.trueSize)[::-1]
    N = cfg.trueSize[0]
    j = np.arange(sz[2]) - 1
    phi = np.broadcast_to(np.cos(2 * np.pi * j / N), sz).copy()
    O.phi[0][:] = phi.reshape(-1)
    O.rho[0][:] = 0
    O.lib.orc_jacobi3d(C.byref(O.topo), O._keep["phi"], O._keep["rho"], orc.ip(O.size), 1, None, None)
    want = (2 * np.cos(2 * np.pi / N) + 4) / 6 * phi
    got = O.phi[0].reshape(sz)
    assert np.abs(got[1:-1, 1:-1, 1:-1] - want[1:-1, 1:-1, 1:-1]).max() <= 1e-15
    assert np.abs(got[:, :, 0] - got[:, :, N]).max() == 0 and np.abs(got[:, :, N + 1] - got[:, :, 1]).max() == 0     # ghost layers refreshed
