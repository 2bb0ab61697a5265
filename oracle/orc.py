"""ctypes wrapper of oracle/libpinc_oracle.so — our CPU restatement of the reference algorithm.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by pinc_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORC_SO = os.path.join(HERE, "libpinc_oracle.so")

c_double_p = C.POINTER(C.c_double)
c_long_p = C.POINTER(C.c_long)
c_int_p = C.POINTER(C.c_int)


class OrcTopo(C.Structure):
    _fields_ = [("nRanks", C.c_int), ("nSub", C.c_int * 3), ("trueSize", C.c_int * 3)]


class OrcSim(C.Structure):
    _fields_ = [
        ("topo", OrcTopo), ("nSpecies", C.c_int),
        ("charge", C.c_double * 8), ("mass", C.c_double * 8),
        ("thresholds", C.c_double * 6), ("size", C.c_int * 3),
        ("pos", C.POINTER(c_double_p)), ("vel", C.POINTER(c_double_p)),
        ("iStart", C.POINTER(c_long_p)), ("iStop", C.POINTER(c_long_p)),
        ("rho", C.POINTER(c_double_p)), ("phi", C.POINTER(c_double_p)),
        ("res", C.POINTER(c_double_p)), ("E", C.POINTER(c_double_p)),
        ("emigrants", C.POINTER(C.POINTER(c_double_p))),
        ("nEmigrants", C.POINTER(c_long_p)), ("nImmigrants", C.POINTER(c_long_p)),
        ("mg", C.c_void_p),
        ("kinEnergy", C.c_double * 9), ("potEnergy", C.c_double),
        ("lastCycles", C.c_int), ("lastBarRes", C.c_double * 256),
    ]


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(ORC_SO):
        build()
    lib = C.CDLL(ORC_SO, mode=os.RTLD_LOCAL)
    P = C.POINTER
    PP = P(c_double_p)
    lib.orc_move.argtypes = [c_double_p, c_double_p, C.c_int, c_long_p, c_long_p]
    lib.orc_acc3d1.argtypes = [c_double_p, c_double_p, C.c_int, c_long_p, c_long_p, c_double_p, c_double_p, c_double_p, c_int_p, c_double_p]
    lib.orc_boris3d1.argtypes = [c_double_p, c_double_p, C.c_int, c_long_p, c_long_p, c_double_p, c_double_p, c_double_p, c_int_p, c_double_p, c_double_p, c_double_p, C.c_int]
    lib.orc_rotation_parameters.argtypes = [C.c_int, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]
    lib.orc_distr3d1.argtypes = [c_double_p, C.c_int, c_long_p, c_long_p, c_double_p, c_double_p, c_int_p]
    lib.orc_acc_nd.argtypes = [c_double_p, c_double_p, C.c_int, c_long_p, c_long_p, c_double_p, c_double_p, c_double_p, c_int_p, C.c_int, c_double_p]
    lib.orc_distr_nd.argtypes = [c_double_p, C.c_int, c_long_p, c_long_p, c_double_p, c_double_p, c_int_p, C.c_int]
    lib.orc_extract3d.argtypes = [c_double_p, c_double_p, C.c_int, c_long_p, c_long_p, c_double_p, PP, c_long_p]
    lib.orc_pnew.argtypes = [c_double_p, c_double_p, c_long_p, c_long_p, C.c_int, c_double_p, c_double_p]
    lib.orc_pnew.restype = C.c_int
    lib.orc_pcut.argtypes = [c_double_p, c_double_p, c_long_p, C.c_int, C.c_long, c_double_p, c_double_p]
    lib.orc_neighbor_to_rank.argtypes = [P(OrcTopo), C.c_int, C.c_int]
    lib.orc_neighbor_to_reciprocal.argtypes = [C.c_int]
    lib.orc_rank_to_neighbor.argtypes = [P(OrcTopo), C.c_int, C.c_int]
    lib.orc_thresholds.argtypes = [c_int_p, c_double_p, c_double_p]
    lib.orc_migrate.argtypes = [P(OrcTopo), PP, PP, C.c_int, P(c_long_p), P(PP), P(c_long_p), P(c_long_p)]
    lib.orc_halo_dim.argtypes = [P(OrcTopo), PP, c_int_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.orc_halo.argtypes = [P(OrcTopo), PP, c_int_p, C.c_int, C.c_int, C.c_int]
    lib.orc_neutralize.argtypes = [P(OrcTopo), PP, c_int_p]
    lib.orc_findiff1st.argtypes = [c_double_p, c_double_p, c_int_p]
    lib.orc_gmul.argtypes = [c_double_p, C.c_long, C.c_double]
    lib.orc_gs3d.argtypes = [P(OrcTopo), PP, PP, c_int_p, C.c_int]
    lib.orc_residual.argtypes = [c_double_p, c_double_p, c_double_p, c_int_p]
    lib.orc_half_restrict3d.argtypes = [c_double_p, c_int_p, c_double_p, c_int_p]
    lib.orc_bilin_prol3d.argtypes = [P(OrcTopo), PP, c_int_p, PP, c_int_p]
    lib.orc_sum_true.restype = C.c_double
    lib.orc_sum_true.argtypes = [c_double_p, c_int_p]
    lib.orc_pot_energy.restype = C.c_double
    lib.orc_pot_energy.argtypes = [c_double_p, c_double_p, c_int_p]
    lib.orc_mg_alloc.restype = C.c_void_p
    lib.orc_mg_alloc.argtypes = [P(OrcTopo), C.c_int, C.c_int, C.c_int, C.c_int]
    lib.orc_mg_free.argtypes = [C.c_void_p]
    lib.orc_mg_solve.restype = C.c_int
    lib.orc_mg_solve.argtypes = [C.c_void_p, PP, PP, PP, C.c_double, C.c_int, c_double_p, C.c_int]
    lib.orc_mg_vcycle.argtypes = [C.c_void_p, PP, PP, PP]
    lib.orc_mg_wcycle.argtypes = [C.c_void_p, PP, PP, PP]
    lib.orc_mg_vregular.argtypes = [C.c_void_p, PP, PP, PP]
    lib.orc_mg_set_smoothers.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.orc_jacobi3d.argtypes = [P(OrcTopo), PP, PP, c_int_p, C.c_int, c_int_p, PP]
    lib.orc_mg_level.restype = c_double_p
    lib.orc_mg_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, c_int_p]
    lib.orc_slice_max.restype = C.c_long
    lib.orc_slice_max.argtypes = [c_int_p]
    lib.orc_set_bnd_slices.argtypes = [P(OrcTopo), C.c_int, c_int_p, c_int_p, c_double_p]
    lib.orc_bnd.argtypes = [P(OrcTopo), PP, c_int_p, c_int_p, PP]
    lib.orc_gs3d_bnd.argtypes = [P(OrcTopo), PP, PP, c_int_p, C.c_int, c_int_p, PP]
    lib.orc_mg_set_bnd.argtypes = [C.c_void_p, c_int_p]
    lib.orc_mg_bnd_slice.restype = c_double_p
    lib.orc_mg_bnd_slice.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.orc_mg_restrict_bnd.argtypes = [C.c_void_p]
    lib.orc_step.argtypes = [P(OrcSim)]
    lib.orc_field_solve.argtypes = [P(OrcSim)]
    lib.orc_accelerate.argtypes = [P(OrcSim), C.c_double]
    _lib = lib
    return lib


def dp(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_double_p)


def lp(a):
    assert a.dtype == np.int64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_long_p)


def ip(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_int_p)


def ptr_array(arrs):
    """double** from a list of float64 arrays."""
    return (c_double_p * len(arrs))(*[dp(a) for a in arrs])


def lptr_array(arrs):
    return (c_long_p * len(arrs))(*[lp(a) for a in arrs])


def make_topo(nSub, trueSize):
    t = OrcTopo()
    t.nRanks = int(np.prod(nSub))
    t.nSub[:] = list(nSub)
    t.trueSize[:] = list(trueSize)
    return t


class OrcWorld:
    """All sub-domains of a PINC run held in one process, stepped in lock-step by the oracle."""

    def __init__(self, cfg, emig_cap=None):
        lib = self.lib = load()
        self.cfg = cfg
        R = self.n = cfg.nRanks
        self.topo = make_topo(cfg.nSubdomains, cfg.trueSize)
        self.size = np.array([t + 2 for t in cfg.trueSize], dtype=np.int32)
        nS = self.nS = cfg.nSpecies
        ngrid = int(np.prod(self.size))
        per_rank_alloc = [-(-a // R) for a in cfg.nAlloc]                     # population.c:58-64
        self.iStart = [np.concatenate([[0], np.cumsum(per_rank_alloc)]).astype(np.int64) for _ in range(R)]
        self.iStop = [self.iStart[r][:nS].copy() for r in range(R)]
        ntot = int(self.iStart[0][nS])
        self.pos = [np.zeros(3 * ntot) for _ in range(R)]
        self.vel = [np.zeros(3 * ntot) for _ in range(R)]
        self.rho = [np.zeros(ngrid) for _ in range(R)]
        self.phi = [np.zeros(ngrid) for _ in range(R)]
        self.res = [np.zeros(ngrid) for _ in range(R)]
        self.E = [np.zeros(3 * ngrid) for _ in range(R)]
        # emigrant buffers (grid.c:1045-1107): corner/edge/face sizing
        alloc = neighbor_alloc(cfg.nEmigrantsAlloc)
        if emig_cap is not None:
            alloc = [min(a, emig_cap) for a in alloc]
        self.emig = [[np.zeros(6 * max(a, 1)) for a in alloc] for _ in range(R)]
        self.nEmig = [np.zeros(27 * nS, dtype=np.int64) for _ in range(R)]
        self.nImm = [np.zeros(27 * nS, dtype=np.int64) for _ in range(R)]
        self.mg = lib.orc_mg_alloc(C.byref(self.topo), cfg.mgLevels, cfg.nPreSmooth, cfg.nPostSmooth, cfg.nCoarseSolve)

        s = self.sim = OrcSim()
        s.topo = self.topo
        s.nSpecies = nS
        for i in range(nS):
            s.charge[i] = cfg.charge[i]
            s.mass[i] = cfg.mass[i]
        thr = np.zeros(6)
        lib.orc_thresholds(ip(self.size), dp(np.array(cfg.thresholds, dtype=np.float64)), dp(thr))
        self.thresholds = thr
        s.thresholds[:] = list(thr)
        s.size[:] = list(self.size)
        self._keep = dict(
            pos=ptr_array(self.pos), vel=ptr_array(self.vel),
            iStart=lptr_array(self.iStart), iStop=lptr_array(self.iStop),
            rho=ptr_array(self.rho), phi=ptr_array(self.phi), res=ptr_array(self.res), E=ptr_array(self.E),
            emig_rows=[ptr_array(e) for e in self.emig],
            nEmig=lptr_array(self.nEmig), nImm=lptr_array(self.nImm))
        k = self._keep
        k["emig"] = (C.POINTER(c_double_p) * R)(*[C.cast(row, C.POINTER(c_double_p)) for row in k["emig_rows"]])
        s.pos, s.vel, s.iStart, s.iStop = k["pos"], k["vel"], k["iStart"], k["iStop"]
        s.rho, s.phi, s.res, s.E = k["rho"], k["phi"], k["res"], k["E"]
        s.emigrants, s.nEmigrants, s.nImmigrants = k["emig"], k["nEmig"], k["nImm"]
        s.mg = self.mg

    def set_boundaries(self, names):
        """grid:boundaries (lower x,y,z then upper x,y,z) for the multigrid levels: gSetBndSlices on level 0, then mgRestrictBnd.
        Returns the per-rank numpy views of the level-0 boundary slices (8 x orc_slice_max doubles)."""
        kinds = {"PERIODIC": 1, "DIRICHLET": 2, "NEUMANN": 3}
        b = [kinds[x] for x in names]
        self.bnd = np.array([0x10] + b[:3] + [0x10] + b[3:], dtype=np.int32)
        self.lib.orc_mg_set_bnd(self.mg, ip(self.bnd))
        nmax = self.lib.orc_slice_max(ip(self.size))
        self.bnd_slices = []
        for r in range(self.n):
            p = self.lib.orc_mg_bnd_slice(self.mg, 0, r)
            self.lib.orc_set_bnd_slices(C.byref(self.topo), r, ip(self.size), ip(self.bnd), p)
            self.bnd_slices.append(np.ctypeslib.as_array(p, shape=(8 * nmax,)))
        self.lib.orc_mg_restrict_bnd(self.mg)
        self._keep["bnd0"] = (c_double_p * self.n)(*[self.lib.orc_mg_bnd_slice(self.mg, 0, r) for r in range(self.n)])
        return self.bnd_slices

    def set_particles(self, per_rank):
        for r in range(self.n):
            for sidx, (ps, vs) in enumerate(per_rank[r]):
                i0 = int(self.iStart[r][sidx])
                n = len(ps)
                assert i0 + n <= self.iStart[r][sidx + 1], "species capacity exceeded"
                self.pos[r][3 * i0:3 * (i0 + n)] = np.asarray(ps, dtype=np.float64).reshape(-1)
                self.vel[r][3 * i0:3 * (i0 + n)] = np.asarray(vs, dtype=np.float64).reshape(-1)
                self.iStop[r][sidx] = i0 + n

    def particles(self, r):
        out = []
        for s in range(self.nS):
            a, b = int(self.iStart[r][s]), int(self.iStop[r][s])
            out.append((self.pos[r][3 * a:3 * b].reshape(-1, 3).copy(), self.vel[r][3 * a:3 * b].reshape(-1, 3).copy()))
        return out

    def grid(self, r, name):
        a = getattr(self, name)[r]
        sz = tuple(int(x) for x in self.size[::-1])
        return a.reshape(sz + ((3,) if name == "E" else (1,)))

    def migrate(self):
        lib, s = self.lib, self.sim
        for r in range(self.n):
            lib.orc_extract3d(dp(self.pos[r]), dp(self.vel[r]), self.nS, lp(self.iStart[r]), lp(self.iStop[r]),
                              dp(self.thresholds), self._keep["emig_rows"][r], lp(self.nEmig[r]))
        lib.orc_migrate(C.byref(self.topo), s.pos, s.vel, self.nS, s.iStop, s.emigrants, s.nEmigrants, s.nImmigrants)

    def field_solve(self):
        self.lib.orc_field_solve(C.byref(self.sim))

    def half_kick(self):
        self.lib.orc_accelerate(C.byref(self.sim), 0.5)

    def step(self):
        self.lib.orc_step(C.byref(self.sim))

    def energies(self):
        return self.sim.kinEnergy[self.nS], self.sim.potEnergy

    def history(self):
        n = self.sim.lastCycles
        return [self.sim.lastBarRes[i] for i in range(min(n, 256))]


def neighbor_alloc(spec):
    """grid.c:1045-1082: 1 value (all), 3 values (corner, edge, face) or 27 values."""
    out = []
    for ne in range(27):
        if ne == 13:
            out.append(0)
            continue
        if len(spec) == 1:
            out.append(spec[0])
        elif len(spec) == 27:
            out.append(spec[ne])
        else:
            digits = [(ne // 3 ** d) % 3 for d in range(3)]
            interface_dims = sum(1 for x in digits if x == 1)
            out.append(spec[interface_dims])
    return out
