"""ctypes restatement of the PINC structs in include/pinc_b200.h (= /root/reference/src/core.h:72-86,
112-138, 261-277 and src/multigrid.h:27-57).

The same layouts serve libpinc_b200.so (the product) and, in tests only, the reference's own
objects compiled under shims (oracle/_ref): both take `Population*`, `Grid*`, `MpiInfo*`.
"""
import ctypes as C

import numpy as np

hid_t = C.c_int64
c_long_p = C.POINTER(C.c_long)
c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)

PERIODIC, DIRICHLET, NEUMANN, NONE = 0x01, 0x02, 0x03, 0x10
TOHALO, FROMHALO = 0, 1
SCALAR, VECTOR = 1, -1


class Population(C.Structure):
    _fields_ = [
        ("pos", c_double_p), ("vel", c_double_p),
        ("iStart", c_long_p), ("iStop", c_long_p),
        ("objVicinity", c_long_p), ("collisions", c_long_p),
        ("charge", c_double_p), ("mass", c_double_p),
        ("kinEnergy", c_double_p), ("potEnergy", c_double_p),
        ("nSpecies", C.c_int), ("nDims", C.c_int),
        ("h5", hid_t),
    ]


class MpiInfo(C.Structure):
    _fields_ = [
        ("mpiRank", C.c_int), ("mpiSize", C.c_int), ("nDims", C.c_int),
        ("subdomain", c_int_p), ("nSubdomains", c_int_p), ("nSubdomainsProd", c_int_p),
        ("offset", c_int_p), ("posToSubdomain", c_double_p),
        ("nSpecies", C.c_int), ("nNeighbors", C.c_int), ("neighborhoodCenter", C.c_int),
        ("migrants", C.POINTER(c_long_p)), ("migrantsDummy", C.POINTER(c_long_p)),
        ("nEmigrants", c_long_p), ("nEmigrantsAlloc", c_long_p), ("nImmigrants", c_long_p),
        ("nImmigrantsAlloc", C.c_long),
        ("emigrants", C.POINTER(c_double_p)), ("emigrantsDummy", C.POINTER(c_double_p)),
        ("immigrants", c_double_p), ("thresholds", c_double_p),
        ("send", C.c_void_p), ("recv", C.c_void_p),
    ]


class Grid(C.Structure):
    _fields_ = [
        ("val", c_double_p), ("rank", C.c_int),
        ("size", c_int_p), ("trueSize", c_int_p), ("sizeProd", c_long_p), ("nGhostLayers", c_int_p),
        ("sendSlice", c_double_p), ("recvSlice", c_double_p), ("bndSlice", c_double_p),
        ("h5", hid_t), ("h5MemSpace", hid_t), ("h5FileSpace", hid_t),
        ("bnd", c_int_p),
    ]


class Multigrid(C.Structure):
    _fields_ = [
        ("grids", C.POINTER(C.POINTER(Grid))),
        ("nLevels", C.c_int), ("nMGCycles", C.c_int),
        ("nPreSmooth", C.c_int), ("nPostSmooth", C.c_int), ("nCoarseSolve", C.c_int),
        ("coarseSolv", C.c_void_p), ("postSmooth", C.c_void_p), ("preSmooth", C.c_void_p),
        ("restrictor", C.c_void_p), ("prolongator", C.c_void_p),
    ]


class MultigridSolver(C.Structure):
    _fields_ = [
        ("res", C.POINTER(Grid)),
        ("mgRho", C.POINTER(Multigrid)), ("mgPhi", C.POINTER(Multigrid)), ("mgRes", C.POINTER(Multigrid)),
        ("mgAlgo", C.c_void_p),
    ]


def np_view(ptr, n, dtype=np.float64):
    """numpy view (no copy) of n elements behind a ctypes pointer."""
    if n == 0:
        return np.empty(0, dtype=dtype)
    ct = {np.float64: C.c_double, np.int64: C.c_long, np.int32: C.c_int}[dtype]
    arr = C.cast(ptr, C.POINTER(ct * n)).contents
    return np.frombuffer(arr, dtype=dtype)


def grid_shape(g):
    """(size tuple incl. ghosts [nValues, nx, ny, nz], total element count) of a Grid."""
    rank = g.rank
    size = [g.size[i] for i in range(rank)]
    return size, int(g.sizeProd[rank])


def grid_array(g):
    """numpy view of grid->val shaped (nz, ny, nx, nValues) (memory order of core.h:261-277)."""
    size, n = grid_shape(g)
    return np_view(g.val, n).reshape(tuple(reversed(size)))


def pop_arrays(p):
    """numpy views (pos, vel) shaped (nAllocTotal, 3) of a Population."""
    ns = p.nSpecies
    ntot = int(p.iStart[ns])
    return np_view(p.pos, 3 * ntot).reshape(ntot, 3), np_view(p.vel, 3 * ntot).reshape(ntot, 3)
