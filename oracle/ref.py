"""Harness around oracle/_ref/libpinc_ref.so — the reference's OWN hot-path sources compiled in
place under shims (oracle/Makefile).  TEST INFRASTRUCTURE ONLY.

One host thread per MPI rank (the shim maps ranks to threads).  Every rank opens the ini file
with the reference's iniOpen, normalises with uAlloc/uNormalize, and allocates its structs with
gAllocMpi / pAlloc / gAlloc / mgAllocSolver / gCreateNeighborhood, exactly as src/main.c:84-99.
Initial conditions are injected (GSL is absent).  Grids are zeroed after allocation (quirk Q5).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import queue
import tempfile
import threading

import numpy as np

from pinc_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libpinc_ref.so")


def available() -> bool:
    return os.path.exists(REF_SO)


def load():
    lib = C.CDLL(REF_SO, mode=os.RTLD_LOCAL)
    P = C.POINTER
    lib.iniOpen.restype = C.c_void_p
    lib.iniOpen.argtypes = [C.c_int, P(C.c_char_p)]
    lib.uAlloc.restype = C.c_void_p
    lib.uAlloc.argtypes = [C.c_void_p]
    lib.uNormalize.argtypes = [C.c_void_p, C.c_void_p]
    lib.gAllocMpi.restype = P(abi.MpiInfo)
    lib.gAllocMpi.argtypes = [C.c_void_p]
    lib.pAlloc.restype = P(abi.Population)
    lib.pAlloc.argtypes = [C.c_void_p]
    lib.gAlloc.restype = P(abi.Grid)
    lib.gAlloc.argtypes = [C.c_void_p, C.c_int]
    lib.mgAllocSolver.restype = P(abi.MultigridSolver)
    lib.mgAllocSolver.argtypes = [C.c_void_p, P(abi.Grid), P(abi.Grid)]
    lib.gCreateNeighborhood.argtypes = [C.c_void_p, P(abi.MpiInfo), P(abi.Grid)]
    lib.puMove.argtypes = [P(abi.Population), C.c_void_p]
    for f in ("puAcc3D1", "puAcc3D1KE"):
        getattr(lib, f).argtypes = [P(abi.Population), P(abi.Grid)]
    for f in ("puBoris3D1", "puBoris3D1KE"):
        getattr(lib, f).argtypes = [P(abi.Population), P(abi.Grid), abi.c_double_p, abi.c_double_p]
    lib.puDistr3D1.argtypes = [P(abi.Population), P(abi.Grid)]
    for f in ("puAccND1", "puAccND1KE", "puAccND0", "puAccND0KE", "puDistrND1", "puDistrND0"):
        getattr(lib, f).argtypes = [P(abi.Population), P(abi.Grid)]
    lib.puExtractEmigrantsND.argtypes = [P(abi.Population), P(abi.MpiInfo)]
    lib.puExtractEmigrants3D.argtypes = [P(abi.Population), P(abi.MpiInfo)]
    lib.puMigrate.argtypes = [P(abi.Population), P(abi.MpiInfo), P(abi.Grid)]
    lib.puNeighborToRank.argtypes = [P(abi.MpiInfo), C.c_int]
    lib.puRankToNeighbor.argtypes = [P(abi.MpiInfo), C.c_int]
    lib.puNeighborToReciprocal.argtypes = [C.c_int, C.c_int]
    lib.gHaloOp.argtypes = [C.c_void_p, P(abi.Grid), P(abi.MpiInfo), C.c_int]
    lib.gHaloOpDim.argtypes = [C.c_void_p, P(abi.Grid), P(abi.MpiInfo), C.c_int, C.c_int]
    lib.gFinDiff1st.argtypes = [P(abi.Grid), P(abi.Grid)]
    lib.gMul.argtypes = [P(abi.Grid), C.c_double]
    lib.gZero.argtypes = [P(abi.Grid)]
    lib.gNeutralizeGrid.argtypes = [P(abi.Grid), P(abi.MpiInfo)]
    lib.gBnd.argtypes = [P(abi.Grid), P(abi.MpiInfo)]
    lib.gPotEnergy.argtypes = [P(abi.Grid), P(abi.Grid), P(abi.Population)]
    lib.gSumTruegrid.restype = C.c_double
    lib.gSumTruegrid.argtypes = [P(abi.Grid)]
    lib.pSumKinEnergy.argtypes = [P(abi.Population)]
    lib.mgSolve.argtypes = [P(abi.MultigridSolver), P(abi.Grid), P(abi.Grid), P(abi.MpiInfo)]
    lib.mgGS3D.argtypes = [P(abi.Grid), P(abi.Grid), C.c_int, P(abi.MpiInfo)]
    lib.mgHalfRestrict3D.argtypes = [P(abi.Grid), P(abi.Grid)]
    lib.mgBilinProl3D.argtypes = [P(abi.Grid), P(abi.Grid), P(abi.MpiInfo)]
    lib.mgResidual.argtypes = [P(abi.Grid), P(abi.Grid), P(abi.Grid), P(abi.MpiInfo)]
    lib.mgVRecursive.argtypes = [C.c_int, C.c_int, C.c_int, P(abi.Multigrid), P(abi.Multigrid), P(abi.Multigrid), P(abi.MpiInfo)]
    lib.mgSumTrueSquared.restype = C.c_double
    lib.mgSumTrueSquared.argtypes = [P(abi.Grid), P(abi.MpiInfo)]
    lib.gTotTruesize.restype = C.c_long
    lib.gTotTruesize.argtypes = [P(abi.Grid), P(abi.MpiInfo)]
    lib.gAddTo.argtypes = [P(abi.Grid), P(abi.Grid)]
    lib.pPosLattice.argtypes = [C.c_void_p, P(abi.Population), P(abi.MpiInfo)]
    lib.pPosPerturb.argtypes = [C.c_void_p, P(abi.Population), P(abi.MpiInfo)]
    lib.pVelZero.argtypes = [P(abi.Population)]
    lib.pincShimInit.argtypes = [C.c_int]
    lib.pincShimSetRank.argtypes = [C.c_int]
    return lib


class RefRank:
    """Per-rank reference structs (all owned by the reference's allocators)."""
    pass


class RefWorld:
    """The reference PINC running `nRanks` sub-domains as threads of this process."""

    def __init__(self, ini_text: str, n_ranks: int):
        self.lib = load()
        self.n = n_ranks
        self.lib.pincShimInit(n_ranks)
        self._tmp = tempfile.NamedTemporaryFile("w", suffix=".ini", delete=False)
        self._tmp.write(ini_text)
        self._tmp.close()
        self.ranks = [RefRank() for _ in range(n_ranks)]
        self._q = [queue.Queue() for _ in range(n_ranks)]
        self._done = queue.Queue()
        self._threads = [threading.Thread(target=self._worker, args=(r,), daemon=True) for r in range(n_ranks)]
        for t in self._threads:
            t.start()
        self.run_serial(self._alloc)

    def _worker(self, r):
        self.lib.pincShimSetRank(r)
        while True:
            fn = self._q[r].get()
            if fn is None:
                return
            try:
                fn(r, self.ranks[r])
                self._done.put((r, None))
            except Exception as e:  # pragma: no cover
                self._done.put((r, e))

    def run(self, fn):
        """Execute fn(rank, state) concurrently on every rank thread (a collective phase)."""
        for r in range(self.n):
            self._q[r].put(fn)
        for _ in range(self.n):
            r, err = self._done.get()
            if err is not None:
                raise err

    def run_serial(self, fn):
        """Execute fn on every rank thread, one rank at a time.  Needed for everything that
        reads the ini dictionary: iniparser 3.1 keeps static scratch buffers (not thread safe).
        Only valid for phases without blocking MPI calls (allocation, initial conditions)."""
        for r in range(self.n):
            self._q[r].put(fn)
            rr, err = self._done.get()
            if err is not None:
                raise err

    def close(self):
        for r in range(self.n):
            self._q[r].put(None)
        os.unlink(self._tmp.name)

    def _alloc(self, r, st):
        lib = self.lib
        argv = (C.c_char_p * 2)(b"pinc", self._tmp.name.encode())
        st.ini = lib.iniOpen(2, argv)
        st.units = lib.uAlloc(st.ini)
        lib.uNormalize(st.ini, st.units)
        st.mpi = lib.gAllocMpi(st.ini)
        st.pop = lib.pAlloc(st.ini)
        st.E = lib.gAlloc(st.ini, abi.VECTOR)
        st.rho = lib.gAlloc(st.ini, abi.SCALAR)
        st.phi = lib.gAlloc(st.ini, abi.SCALAR)
        st.solver = lib.mgAllocSolver(st.ini, st.rho, st.phi)
        lib.gCreateNeighborhood(st.ini, st.mpi, st.rho)
        # quirk Q5: canonical initial state = every grid value zero
        for g in (st.E, st.rho, st.phi):
            abi.grid_array(g.contents)[...] = 0
        sol = st.solver.contents
        for mg in (sol.mgRho, sol.mgPhi, sol.mgRes):
            for q in range(mg.contents.nLevels):
                abi.grid_array(mg.contents.grids[q].contents)[...] = 0
        st.res = sol.res
        pos, vel = abi.pop_arrays(st.pop.contents)
        pos[...] = 0
        vel[...] = 0

    # ---- convenience -------------------------------------------------------------------
    def set_particles(self, per_rank):
        """per_rank[r] = list over species of (pos[n,3] local frame, vel[n,3])."""
        for r, st in enumerate(self.ranks):
            p = st.pop.contents
            pos, vel = abi.pop_arrays(p)
            for s, (ps, vs) in enumerate(per_rank[r]):
                i0 = p.iStart[s]
                n = len(ps)
                assert i0 + n <= p.iStart[s + 1], "species capacity exceeded"
                pos[i0:i0 + n] = ps
                vel[i0:i0 + n] = vs
                p.iStop[s] = i0 + n

    def particles(self, r):
        p = self.ranks[r].pop.contents
        pos, vel = abi.pop_arrays(p)
        return [(pos[p.iStart[s]:p.iStop[s]].copy(), vel[p.iStart[s]:p.iStop[s]].copy()) for s in range(p.nSpecies)]

    def grid(self, r, name):
        return abi.grid_array(getattr(self.ranks[r], name).contents)

    def field_solve(self, history=None):
        """distr .. gMul(E,-1): src/main.c:225-247 with one fold and one solve (Q7)."""
        lib = self.lib
        set_slice = C.cast(lib.setSlice, C.c_void_p)
        add_slice = C.cast(lib.addSlice, C.c_void_p)

        def solve(r, st):
            if history is None:
                lib.mgSolve(st.solver, st.rho, st.phi, st.mpi)
                return
            # the tolerance loop of mgSolveRaw (src/multigrid.c:1696-1705) driven call by call, so that the
            # reference's own barRes per V-cycle (which it does not keep) can be recorded
            sol = st.solver.contents
            rho0, phi0, res0 = (m.contents.grids[0] for m in (sol.mgRho, sol.mgPhi, sol.mgRes))
            bar, hist = 2.0, []
            while bar > 1e-10:
                lib.mgVRecursive(0, sol.mgRho.contents.nLevels - 1, 0, sol.mgRho, sol.mgPhi, sol.mgRes, st.mpi)
                lib.mgResidual(res0, rho0, phi0, st.mpi)
                lib.gHaloOp(set_slice, res0, st.mpi, abi.TOHALO)
                bar = lib.mgSumTrueSquared(res0, st.mpi)
                bar /= lib.gTotTruesize(rho0, st.mpi)
                bar = math.sqrt(bar)
                hist.append(bar)
            if r == 0:
                history[:] = hist

        def phase(r, st):
            lib.puDistr3D1(st.pop, st.rho)
            lib.gHaloOp(add_slice, st.rho, st.mpi, abi.FROMHALO)
            solve(r, st)
            lib.gHaloOp(set_slice, st.phi, st.mpi, abi.TOHALO)
            lib.gFinDiff1st(st.phi, st.E)
            lib.gHaloOp(set_slice, st.E, st.mpi, abi.TOHALO)
            lib.gMul(st.E, -1.0)
        self.run(phase)

    def half_kick(self):
        lib = self.lib

        def phase(r, st):                      # src/main.c:184-186
            lib.gMul(st.E, 0.5)
            lib.puAcc3D1KE(st.pop, st.E)
            lib.gMul(st.E, 2.0)
        self.run(phase)

    def migrate(self):
        lib = self.lib

        def phase(r, st):
            lib.puExtractEmigrants3D(st.pop, st.mpi)
            lib.puMigrate(st.pop, st.mpi, st.rho)
        self.run(phase)

    def step(self, history=None):
        """src/main.c:212-261 minus object calls / HDF5 (canonical driver)."""
        lib = self.lib

        def move(r, st):
            lib.puMove(st.pop, None)
        self.run(move)
        self.migrate()
        self.field_solve(history)

        def acc(r, st):
            lib.puAcc3D1KE(st.pop, st.E)
            lib.pSumKinEnergy(st.pop)
            lib.gPotEnergy(st.rho, st.phi, st.pop)
        self.run(acc)

    def energies(self):
        ns = self.ranks[0].pop.contents.nSpecies
        ke = sum(st.pop.contents.kinEnergy[ns] for st in self.ranks)
        pe = sum(st.pop.contents.potEnergy[ns] for st in self.ranks)
        return ke, pe
