"""Host driver: the reference's `regular()` time loop (/root/reference/src/main.c:50-304) restated on top of the
PINC-named C entry points of libpinc_b200.so, in the canonical order of SURVEY 8c (one rho fold and one
solve per step, no object calls, no HDF5 inside the loop; quirks Q6/Q7).

A `World` owns the sub-domains ("ranks") this process drives:
  * 1 rank                  -> calls run inline on the caller's thread;
  * R ranks, one process    -> one host thread per rank sharing the visible GPU(s); exchanges go through the
                               library's thread transport (tests of the multi-rank path on one GPU);
  * one rank per process    -> `World(cfg, rank=r, world_size=R, nccl_id=...)` under torchrun; NCCL over NVLink.
"""
from __future__ import annotations

import ctypes as C
import queue
import threading

import numpy as np

from . import abi, lib as _lib


class RankState:
    pass


def _ia(vals):
    return (C.c_int * len(vals))(*[int(v) for v in vals])


def _la(vals):
    return (C.c_long * len(vals))(*[int(v) for v in vals])


def _da(vals):
    return (C.c_double * len(vals))(*[float(v) for v in vals])


class World:
    def __init__(self, cfg, rank=None, world_size=None, nccl_id: bytes | None = None, devices=None):
        self.lib = L = _lib.load()
        self.cfg = cfg
        self.R = cfg.nRanks
        if rank is None:
            self.local = list(range(self.R))
        else:
            assert world_size == self.R, "world size must equal the product of grid:nSubdomains"
            self.local = [rank]
        self.nccl_id = nccl_id
        self.devices = devices
        self.ranks = {r: RankState() for r in self.local}
        self.set_slice = _lib.fn_ptr(L, "setSlice")
        self.add_slice = _lib.fn_ptr(L, "addSlice")
        self._threads = {}
        self._moved = False
        if len(self.local) > 1:
            self._q = {r: queue.Queue() for r in self.local}
            self._done = queue.Queue()
            for r in self.local:
                t = threading.Thread(target=self._worker, args=(r,), daemon=True)
                t.start()
                self._threads[r] = t
        self.run(self._create_ctx)
        if len(self.local) > 1:
            arr = (C.c_void_p * self.R)(*[self.ranks[r].ctx for r in range(self.R)])
            L.pincCommInitThreads(arr, self.R)
        elif rank is not None and self.R > 1:
            L.pincCommInitNccl(self.ranks[rank].ctx, nccl_id)
        self.run(self._alloc)

    # ---- execution of a collective phase on every local rank ---------------------------------
    def _worker(self, r):
        while True:
            fn = self._q[r].get()
            if fn is None:
                return
            try:
                fn(r, self.ranks[r])
                self._done.put((r, None))
            except BaseException as e:  # pragma: no cover
                self._done.put((r, e))

    def run(self, fn):
        if len(self.local) == 1:
            r = self.local[0]
            fn(r, self.ranks[r])
            return
        for r in self.local:
            self._q[r].put(fn)
        err = None
        for _ in self.local:
            _, e = self._done.get()
            err = err or e
        if err is not None:
            raise err

    def close(self):
        def fin(r, st):
            L = self.lib
            L.mgFreeSolver(st.solver)
            for g in (st.E, st.rho, st.phi):
                L.pincGridFree(g)
            L.pincPopFree(st.pop)
            L.pincMpiFree(st.mpi)
            L.pincCtxDestroy(st.ctx)
        self.run(fin)
        for r in self._threads:
            self._q[r].put(None)

    # ---- set-up: src/main.c:84-99 with plain arguments -------------------------------------------
    def _create_ctx(self, r, st):
        import os
        if self.devices is not None:
            dev = self.devices[self.local.index(r) % len(self.devices)]
        elif len(self.local) == 1 and self.R > 1:
            dev = int(os.environ.get("LOCAL_RANK", "0"))
        else:
            dev = int(os.environ.get("PINC_B200_DEVICE", "0"))
        st.ctx = self.lib.pincCtxCreate(dev, r, self.R)

    def _alloc(self, r, st):
        L, cfg = self.lib, self.cfg
        nD, nS = cfg.nDims, cfg.nSpecies
        kinds = {"PERIODIC": abi.PERIODIC, "DIRICHLET": abi.DIRICHLET, "NEUMANN": abi.NEUMANN}
        for b in cfg.boundaries:
            if b not in kinds:
                raise ValueError(f"{b} invalid value for grid:boundaries")          # src/grid.c:479
        bnd = _ia([kinds[b] for b in cfg.boundaries])                                 # lower x,y,z then upper x,y,z
        ts, gl = _ia(cfg.trueSize), _ia(cfg.nGhostLayers)
        st.mpi = L.pincMpiAlloc(nD, nS, _ia(cfg.nSubdomains), gl, ts, r, self.R)
        per_rank = [-(-a // self.R) for a in cfg.nAlloc]                    # population.c:58-64
        st.pop = L.pincPopAlloc(nS, nD, _la(per_rank), _da(cfg.charge), _da(cfg.mass))
        st.E = L.pincGridAlloc(nD, ts, gl, abi.VECTOR, bnd)
        st.rho = L.pincGridAlloc(nD, ts, gl, abi.SCALAR, bnd)
        st.phi = L.pincGridAlloc(nD, ts, gl, abi.SCALAR, bnd)
        st.solver = L.pincMgAllocSolver(st.rho, st.phi, cfg.mgLevels, cfg.mgCycles, cfg.nPreSmooth, cfg.nPostSmooth, cfg.nCoarseSolve)
        st.res = st.solver.contents.res
        if any(b != "PERIODIC" for b in cfg.boundaries):
            L.gSetBndSlices(st.phi, st.mpi)                                # src/main.c:102
            L.mgRestrictBnd(st.solver.contents.mgPhi)                      # the coarse levels' boundary values (see mgRestrictBnd)
        ne = cfg.nEmigrantsAlloc
        L.pincCreateNeighborhood(st.mpi, st.rho, _la(ne), len(ne), _da(cfg.thresholds))
        err = C.create_string_buffer(256)
        for name in ("puAcc3D1KE", "puDistr3D1", "puExtractEmigrants3D"):        # the X_set validators (puSanity)
            if L.pincPuSanity(name.encode(), nD, gl, _da(cfg.thresholds), 3, 1, err, 256):
                raise ValueError(err.value.decode())

    # ---- data access --------------------------------------------------------------------------------
    def set_particles(self, per_rank):
        """per_rank[r] = list over species of (pos[n,3] local frame, vel[n,3]); r indexes local ranks by global id."""
        def put(r, st):
            p = st.pop.contents
            pos, vel = abi.pop_arrays(p)
            for s, (ps, vs) in enumerate(per_rank[r]):
                i0, n = p.iStart[s], len(ps)
                assert i0 + n <= p.iStart[s + 1], "species capacity exceeded"
                pos[i0:i0 + n] = ps
                vel[i0:i0 + n] = vs
                p.iStop[s] = i0 + n
            self.lib.pincSyncPopToDevice(st.pop)
        self.run(put)
        self._moved = False

    def init_on_device(self, kind, seed=20261018):
        """Initial conditions generated by the library's kernels (pincPosLattice/Perturb/VelZero or
        pincPosUniform/VelMaxwell) instead of host arrays."""
        L, cfg = self.lib, self.cfg

        def gen(r, st):
            n, ts = _la(cfg.nParticles), _ia(cfg.trueSize)
            if kind == "lattice":
                L.pincPosLattice(st.pop, st.mpi, n, ts)
                L.pincPosPerturb(st.pop, st.mpi, _da(cfg.perturbAmplitude), _da(cfg.perturbMode), ts)
                L.pincVelZero(st.pop)
            else:
                L.pincPosUniform(st.pop, st.mpi, n, ts, seed)
                L.pincVelMaxwell(st.pop, st.mpi, _da(cfg.drift), _da(cfg.thermalVelocity), seed + 1)
        self.run(gen)
        self._moved = False

    def particles(self, r):
        st = self.ranks[r]
        self._on(r, lambda: self.lib.pincSyncPopToHost(st.pop))
        p = st.pop.contents
        pos, vel = abi.pop_arrays(p)
        return [(pos[p.iStart[s]:p.iStop[s]].copy(), vel[p.iStart[s]:p.iStop[s]].copy()) for s in range(p.nSpecies)]

    def grid(self, r, name):
        st = self.ranks[r]
        g = getattr(st, name)
        self._on(r, lambda: self.lib.pincSyncGridToHost(g))
        return abi.grid_array(g.contents).copy()

    def n_particles(self):
        tot = 0
        for st in self.ranks.values():
            p = st.pop.contents
            tot += sum(p.iStop[s] - p.iStart[s] for s in range(p.nSpecies))
        return tot

    def _on(self, r, thunk):
        """Run thunk on rank r's thread (its device context is thread-bound)."""
        if len(self.local) == 1:
            thunk()
            return
        box = {}

        def fn(rr, st):
            if rr == r:
                thunk()
        # every rank runs the phase so the worker protocol stays in lock-step
        self.run(fn)
        return box

    # ---- phases of src/main.c:155-186 and :197-274 ---------------------------------------------------------
    def migrate(self):
        L = self.lib

        def phase(r, st):
            L.puExtractEmigrants3D(st.pop, st.mpi)
            L.puMigrate(st.pop, st.mpi, st.rho)
        self.run(phase)

    def field_solve(self):
        L = self.lib

        def phase(r, st):
            L.puDistr3D1(st.pop, st.rho)
            L.gHaloOp(self.add_slice, st.rho, st.mpi, abi.FROMHALO)
            L.mgSolve(st.solver, st.rho, st.phi, st.mpi)
            L.gHaloOp(self.set_slice, st.phi, st.mpi, abi.TOHALO)
            L.gFinDiff1st(st.phi, st.E)
            L.gHaloOp(self.set_slice, st.E, st.mpi, abi.TOHALO)
            L.gMul(st.E, -1.0)
        self.run(phase)

    def half_kick(self):
        L = self.lib

        def phase(r, st):                              # src/main.c:184-186
            L.gMul(st.E, 0.5)
            L.puAcc3D1KE(st.pop, st.E)
            L.gMul(st.E, 2.0)
        self.run(phase)

    def step(self, fused=False):
        """One time step.  fused=False: the reference's call sequence, entry point by entry point.
        fused=True: puAcc3D1KE, the next step's puMove, the emigrant classification and the deposition of the
        particles that stay run as one pass over the particles (pincAccMoveDistr3D1KE); fused="nodeposit" leaves the
        deposition to puDistr3D1 (pincAccMove3D1KE).  Same arithmetic, bit-identical fields."""
        L = self.lib

        def phase(r, st):
            if not self._moved:
                L.puMove(st.pop, None)
            L.puExtractEmigrants3D(st.pop, st.mpi)
            L.puMigrate(st.pop, st.mpi, st.rho)
            L.puDistr3D1(st.pop, st.rho)
            L.gHaloOp(self.add_slice, st.rho, st.mpi, abi.FROMHALO)
            L.mgSolve(st.solver, st.rho, st.phi, st.mpi)
            L.gHaloOp(self.set_slice, st.phi, st.mpi, abi.TOHALO)
            L.gFinDiff1st(st.phi, st.E)
            L.gHaloOp(self.set_slice, st.E, st.mpi, abi.TOHALO)
            L.gMul(st.E, -1.0)
            if fused == "nodeposit":
                L.pincAccMove3D1KE(st.pop, st.E, st.mpi)
            elif fused:
                L.pincAccMoveDistr3D1KE(st.pop, st.E, st.rho, st.mpi)
            else:
                L.puAcc3D1KE(st.pop, st.E)
            L.pSumKinEnergy(st.pop)
            L.gPotEnergy(st.rho, st.phi, st.pop)
        self.run(phase)
        self._moved = bool(fused)

    def energies(self):
        ns = self.cfg.nSpecies
        ke = sum(st.pop.contents.kinEnergy[ns] for st in self.ranks.values())
        pe = sum(st.pop.contents.potEnergy[ns] for st in self.ranks.values())
        return ke, pe

    def history(self, r=None):
        r = self.local[0] if r is None else r
        out = {}

        def get():
            buf = (C.c_double * 256)()
            n = self.lib.pincMgLastHistory(buf, 256)
            out["h"] = [buf[i] for i in range(min(n, 256))]
        self._on(r, get)
        return out["h"]

    def mg_path(self, r=None):
        """pincMgLastPath of rank r: 0 distributed ops, 1 all-SM kernel, 2 cluster kernel, +4 replicated global solve."""
        r = self.local[0] if r is None else r
        out = {}
        self._on(r, lambda: out.__setitem__("p", self.lib.pincMgLastPath()))
        return out["p"]

    def sync(self):
        self.run(lambda r, st: self.lib.pincDeviceSynchronize())
