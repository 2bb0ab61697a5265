"""Shared helpers of the parity tests: host structs through the C-ABI constructors, small configs."""
import ctypes as C
import os

import numpy as np

from pinc_b200 import abi, config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ia(v):
    return (C.c_int * len(v))(*[int(x) for x in v])


def la(v):
    return (C.c_long * len(v))(*[int(x) for x in v])


def da(v):
    return (C.c_double * len(v))(*[float(x) for x in v])


def small_ini(name="cold", **over):
    """ini text of configs/<name>.ini with overrides given as section__key=value."""
    ini = config.Ini(open(os.path.join(ROOT, "configs", name + ".ini")).read())
    for k, v in over.items():
        ini.d[k.replace("__", ":").lower()] = str(v)
    return ini.dump()


def small_cfg(name="cold", **over):
    text = small_ini(name, **over)
    return text, config.load_config(config.Ini(text))


class GridH:
    """A Grid allocated by pincGridAlloc with a numpy view of its host values (nz, ny, nx, nv)."""

    def __init__(self, lib, true, nv=1):
        self.lib = lib
        self.ptr = lib.pincGridAlloc(3, ia(true), ia([1] * 6), nv, ia([abi.PERIODIC] * 6))
        self.a = abi.grid_array(self.ptr.contents)
        self.size = np.array([t + 2 for t in true], dtype=np.int32)

    def up(self):
        self.lib.pincSyncGridToDevice(self.ptr)
        return self

    def down(self):
        self.lib.pincSyncGridToHost(self.ptr)
        return self.a

    def flat(self):
        return self.a.reshape(-1)

    def free(self):
        self.lib.pincGridFree(self.ptr)


def single_mpi(lib, true, nS=2, thresholds=(0.1,) * 6):
    m = lib.pincMpiAlloc(3, nS, ia([1, 1, 1]), ia([1] * 6), ia(true), 0, 1)
    return m


def true_view(a):
    return a[1:-1, 1:-1, 1:-1]


def sorted_particles(pos, vel):
    """Canonical order of a particle multiset: lexicographic by (x, y, z, vx, vy, vz)."""
    rec = np.concatenate([pos, vel], axis=1)
    # sort on keys rounded to 1e-7 so that two runs that differ by a few ulp order equal-coordinate particles
    # (lattice starts have thousands of them) the same way
    idx = np.lexsort(np.round(rec, 7).T[::-1])
    return rec[idx]


def multiset_close(pos_a, vel_a, pos_b, vel_b, tol):
    """Two particle sets are equal as multisets up to `tol` per coordinate.  Sized for tens of millions of particles:
    one argsort on x per side, then only the rows that disagree (near-ties in x that the two sides ordered
    differently) are matched by a full lexicographic sort.  Returns the worst coordinate difference."""
    if len(pos_a) != len(pos_b):
        return np.inf
    a = np.concatenate([pos_a, vel_a], axis=1)[np.argsort(pos_a[:, 0], kind="stable")]
    b = np.concatenate([pos_b, vel_b], axis=1)[np.argsort(pos_b[:, 0], kind="stable")]
    d = np.abs(a - b).max(axis=1) if len(a) else np.zeros(0)
    bad = d > tol
    worst = float(d[~bad].max()) if (~bad).any() else 0.0
    if bad.any():
        if bad.sum() > max(1000, len(a) // 1000):
            return float(d.max())
        worst = max(worst, float(np.abs(sorted_particles(a[bad, :3], a[bad, 3:]) - sorted_particles(b[bad, :3], b[bad, 3:])).max()))
    return worst
