"""Non-periodic boundaries: the oracle's restatement (orc_bnd, orc_gs3d_bnd, orc_mg_restrict_bnd) against the reference's
own gBnd / mgGS3D / mgVRecursive / mgRestrictBnd / gSetBndSlices run live from oracle/_ref (one thread per rank), on one
and four sub-domains with mixed DIRICHLET / NEUMANN / PERIODIC edges.  Scenario: tests/bnd_common.py.  Bit-exact where no
global mean is involved, 1e-13 where gNeutralizeGrid's sum is taken in another order."""
import ctypes as C

import numpy as np
import pytest

import bnd_common as bc
from helpers import small_cfg
from oracle import orc, ref
from pinc_b200 import abi

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libpinc_ref.so not built (needs /root/reference)")


def run_reference(text, cfg, init):
    # the reference's own ini path refuses non-periodic boundaries (gGetGlobalSize, src/grid.c:581-585, called by
    # uNormalize): allocate everything PERIODIC and set Grid::bnd of every level of phi afterwards
    from pinc_b200 import config
    ini = config.Ini(text)
    ini.d["grid:boundaries"] = "PERIODIC"
    W = ref.RefWorld(ini.dump(), cfg.nRanks)
    kinds = {"PERIODIC": abi.PERIODIC, "DIRICHLET": abi.DIRICHLET, "NEUMANN": abi.NEUMANN}
    b8 = [abi.NONE] + [kinds[x] for x in cfg.boundaries[:3]] + [abi.NONE] + [kinds[x] for x in cfg.boundaries[3:]]
    lib = W.lib
    lib.gSetBndSlices.argtypes = [C.POINTER(abi.Grid), C.POINTER(abi.MpiInfo)]
    lib.mgRestrictBnd.argtypes = [C.POINTER(abi.Multigrid)]
    out = {"A": [None] * cfg.nRanks, "B": [None] * cfg.nRanks, "C": [None] * cfg.nRanks, "bnd": [None] * cfg.nRanks}

    def setup(r, st):
        for q in range(st.solver.contents.mgPhi.contents.nLevels):
            gq = st.solver.contents.mgPhi.contents.grids[q].contents
            for i in range(8):
                gq.bnd[i] = b8[i]
        g = st.phi.contents
        nmax = max(int(g.sizeProd[4]) // int(g.size[d]) for d in range(4))
        lib.gSetBndSlices(st.phi, st.mpi)
        b = np.ctypeslib.as_array(g.bndSlice, shape=(8 * nmax,))
        const = b.copy()
        vals = bc.slice_values(cfg, r, nmax)
        m = st.mpi.contents
        for d in range(1, 4):
            for bd in (d, d + 4):
                edge = m.subdomain[d - 1] == 0 if bd < 4 else m.subdomain[d - 1] == m.nSubdomains[d - 1] - 1
                if edge and g.bnd[bd] != abi.PERIODIC:
                    assert np.all(const[bd * nmax:(bd + 1) * nmax] == (1.0 if g.bnd[bd] == abi.DIRICHLET else 2.0))
                    b[bd * nmax:(bd + 1) * nmax] = vals[bd * nmax:(bd + 1) * nmax]
                else:
                    b[bd * nmax:(bd + 1) * nmax] = 0
        b[:nmax] = 0
        b[4 * nmax:5 * nmax] = 0
        # the reference leaves the coarse levels' slices uninitialised (src/multigrid.c:177): define them
        for q in range(1, st.solver.contents.mgPhi.contents.nLevels):
            gq = st.solver.contents.mgPhi.contents.grids[q].contents
            nq = max(int(gq.sizeProd[4]) // int(gq.size[d]) for d in range(4))
            np.ctypeslib.as_array(gq.bndSlice, shape=(8 * nq,))[:] = 0
        lib.mgRestrictBnd(st.solver.contents.mgPhi)
        out["bnd"][r] = [np.ctypeslib.as_array(st.solver.contents.mgPhi.contents.grids[q].contents.bndSlice,
                                                shape=(8 * max(int(st.solver.contents.mgPhi.contents.grids[q].contents.sizeProd[4]) // int(st.solver.contents.mgPhi.contents.grids[q].contents.size[d]) for d in range(4)),)).copy()
                         for q in range(st.solver.contents.mgPhi.contents.nLevels)]
        abi.grid_array(st.phi.contents).reshape(-1)[:] = init[r][0]
        abi.grid_array(st.rho.contents).reshape(-1)[:] = init[r][1]
    W.run_serial(setup)

    def stage(name, fn):
        W.run(fn)
        for r in range(cfg.nRanks):
            out[name][r] = W.grid(r, "phi").reshape(-1).copy()
    stage("A", lambda r, st: lib.gBnd(st.phi, st.mpi))
    stage("B", lambda r, st: lib.mgGS3D(st.phi, st.rho, 2, st.mpi))
    sol = lambda st: st.solver.contents
    stage("C", lambda r, st: lib.mgVRecursive(0, bc.LEVELS - 1, 0, sol(st).mgRho, sol(st).mgPhi, sol(st).mgRes, st.mpi))
    W.close()
    return out


def run_oracle(cfg, init):
    O = orc.OrcWorld(cfg)
    sl = O.set_boundaries(cfg.boundaries)
    nmax = O.lib.orc_slice_max(orc.ip(O.size))
    sub = lambda r: [r % cfg.nSubdomains[0], (r // cfg.nSubdomains[0]) % cfg.nSubdomains[1], r // (cfg.nSubdomains[0] * cfg.nSubdomains[1])]
    for r in range(cfg.nRanks):
        vals = bc.slice_values(cfg, r, nmax)
        for d in range(1, 4):
            for bd in (d, d + 4):
                edge = sub(r)[d - 1] == 0 if bd < 4 else sub(r)[d - 1] == cfg.nSubdomains[d - 1] - 1
                if edge and O.bnd[bd] != 1:
                    sl[r][bd * nmax:(bd + 1) * nmax] = vals[bd * nmax:(bd + 1) * nmax]
                else:
                    sl[r][bd * nmax:(bd + 1) * nmax] = 0
        O.phi[r][:] = init[r][0]
        O.rho[r][:] = init[r][1]
    O.lib.orc_mg_restrict_bnd(O.mg)
    out = {"bnd": [[np.ctypeslib.as_array(O.lib.orc_mg_bnd_slice(O.mg, q, r), shape=(8 * O.lib.orc_slice_max(orc.ip(np.array([t // 2 ** q + 2 for t in cfg.trueSize], dtype=np.int32))),)).copy()
                    for q in range(cfg.mgLevels)] for r in range(cfg.nRanks)]}
    k = O._keep
    O.lib.orc_bnd(C.byref(O.topo), k["phi"], orc.ip(O.size), orc.ip(O.bnd), k["bnd0"])
    out["A"] = [p.copy() for p in O.phi]
    O.lib.orc_gs3d_bnd(C.byref(O.topo), k["phi"], k["rho"], orc.ip(O.size), 2, orc.ip(O.bnd), k["bnd0"])
    out["B"] = [p.copy() for p in O.phi]
    O.lib.orc_mg_vcycle(O.mg, k["rho"], k["phi"], k["res"])
    out["C"] = [p.copy() for p in O.phi]
    return out


@needs_ref
@pytest.mark.parametrize("sub,boundaries", bc.CASES)
def test_oracle_boundaries_match_reference(sub, boundaries):
    text, cfg = small_cfg("warm", **bc.overrides(sub, boundaries))
    init = bc.fields(cfg)
    R = run_reference(text, cfg, init)
    O = run_oracle(cfg, init)
    for r in range(cfg.nRanks):
        for q in range(cfg.mgLevels):               # gSetBndSlices + mgRestrictBnd: the boundary values of every level
            for bd in (1, 2, 3, 5, 6, 7):
                n = len(R["bnd"][r][q]) // 8
                assert np.array_equal(R["bnd"][r][q][bd * n:(bd + 1) * n], O["bnd"][r][q][bd * n:(bd + 1) * n]), (r, q, bd)
        for stage in "ABC":
            a, b = O[stage][r], R[stage][r]
            err = np.abs(a - b).max() / np.abs(b).max()
            assert err <= 1e-13, (stage, r, err)


def test_library_boundary_slices_on_host_arrays():
    """gSetBndSlices (src/grid.c:608) and mgRestrictBnd (src/multigrid.c:1314) of libpinc_b200 work on the host arrays of the
    structs (no device needed): same slices as the oracle's on every level, for every rank of a 1,2,2 decomposition."""
    from helpers import ia
    from pinc_b200 import lib as plib
    L = plib.load()
    sub, boundaries = bc.CASES[1]
    text, cfg = small_cfg("warm", **bc.overrides(sub, boundaries))
    O = orc.OrcWorld(cfg)
    O.set_boundaries(cfg.boundaries)
    kinds = {"PERIODIC": abi.PERIODIC, "DIRICHLET": abi.DIRICHLET, "NEUMANN": abi.NEUMANN}
    bnd = ia([kinds[b] for b in cfg.boundaries])
    for r in range(cfg.nRanks):
        ts, gl = ia(cfg.trueSize), ia(cfg.nGhostLayers)
        mpi = L.pincMpiAlloc(3, cfg.nSpecies, ia(cfg.nSubdomains), gl, ts, r, cfg.nRanks)
        rho = L.pincGridAlloc(3, ts, gl, abi.SCALAR, bnd)
        phi = L.pincGridAlloc(3, ts, gl, abi.SCALAR, bnd)
        solver = L.pincMgAllocSolver(rho, phi, cfg.mgLevels, 1, 2, 2, 2)
        L.gSetBndSlices(phi, mpi)
        L.mgRestrictBnd(solver.contents.mgPhi)
        for q in range(cfg.mgLevels):
            g = solver.contents.mgPhi.contents.grids[q].contents
            n = int(g.sizeProd[4])
            mine = np.ctypeslib.as_array(g.bndSlice, shape=(8 * n,))
            want = np.ctypeslib.as_array(O.lib.orc_mg_bnd_slice(O.mg, q, r), shape=(8 * n,))
            for bd in (1, 2, 3, 5, 6, 7):
                assert np.array_equal(mine[bd * n:(bd + 1) * n], want[bd * n:(bd + 1) * n]), (r, q, bd)
        L.mgFreeSolver(solver)
        L.pincGridFree(rho); L.pincGridFree(phi)
