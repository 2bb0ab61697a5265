"""host/pinc_h5.c - the format-level HDF5 writer of the C host - against an independent reader (tests/h5mini.py).

Neither libhdf5 nor h5py exist in this image, so the chain of evidence is:
  1. h5mini reads a file written by the REAL HDF5 library (scipy ships a MATLAB v7.3 fixture, which is an HDF5 file behind a
     512-byte user block) and finds the known contents -> the reader's idea of superblock, symbol-table entries, object headers,
     B-tree, symbol-table nodes, local heap, dataspace / datatype / attribute messages is the library's;
  2. the writer's IEEE-double datatype message is byte-identical to the one the real library wrote into that file;
  3. h5mini reads what pinc_h5.c writes: names, nesting, extents, values (bit-exact), attributes, for small and for large
     groups (multi-level B-trees: every name must lie between the keys that bracket it, as libhdf5's lookup requires).
CPU only; no device."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np
import pytest

import h5mini
from helpers import ROOT

LIB = os.path.join(ROOT, "host", "libpinc_h5.so")


def fixture_of_the_real_library():
    try:
        import scipy.io.matlab
    except Exception:
        return None
    hits = glob.glob(os.path.join(os.path.dirname(scipy.io.matlab.__file__), "tests", "data", "testhdf5_7.4_GLNX86.mat"))
    return hits[0] if hits else None


@pytest.fixture(scope="module")
def h5():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(ROOT, "host", "pinc_h5.c")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "host"), LIB])
    L = C.CDLL(LIB)
    L.ph5Create.restype = C.c_void_p
    L.ph5Create.argtypes = [C.c_char_p]
    L.ph5Attr.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.c_int]
    L.ph5Group.argtypes = [C.c_void_p, C.c_char_p]
    L.ph5Dataset.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_ulonglong)]
    L.ph5Write.argtypes = [C.c_void_p, C.c_int, C.c_ulonglong, C.c_ulonglong, C.POINTER(C.c_double)]
    L.ph5XYCreate.argtypes = [C.c_void_p, C.c_char_p]
    L.ph5XYAppend.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
    L.ph5Close.argtypes = [C.c_void_p]
    return L


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def dataset(L, f, path, arr):
    dims = (C.c_ulonglong * arr.ndim)(*arr.shape)
    d = L.ph5Dataset(f, path.encode(), arr.ndim, dims)
    assert d >= 0, path
    flat = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1)
    assert L.ph5Write(f, d, 0, flat.size, dp(flat)) == 0
    return d


def test_reader_on_a_file_written_by_libhdf5():
    path = fixture_of_the_real_library()
    if path is None:
        pytest.skip("scipy's MATLAB v7.3 fixture is not installed")
    f = h5mini.File(path)
    assert (f.sb, f.base, f.leaf_k, f.int_k) == (512, 512, 4, 16) and f.eof + 0 == len(f.buf)
    assert list(f.root.links) == ["testdouble"]
    d = f.root.links["testdouble"]
    assert d.shape == (9, 1) and d.dtype == np.dtype("<f8") and d.layout == "contiguous"
    assert np.array_equal(d.data.reshape(-1), np.arange(9) * (np.pi / 4))          # MATLAB's 0:pi/4:2*pi
    assert d.attrs["MATLAB_class"].tobytes().rstrip(b"\0") == b"double"


def test_datatype_message_equals_the_real_librarys(h5, tmp_path):
    path = fixture_of_the_real_library()
    if path is None:
        pytest.skip("scipy's MATLAB v7.3 fixture is not installed")
    g = h5mini.File(path)
    ref_oh = [oh for name, oh in _entries(g, g_root_oh(g))][0]
    ref_dt = next(d for t, _f, d in g._messages(ref_oh) if t == 0x0003)
    out = str(tmp_path / "one.h5")
    f = h5.ph5Create(out.encode())
    dataset(h5, f, "/x", np.arange(3.0))
    assert h5.ph5Close(f) == 0
    m = h5mini.File(out)
    mine_oh = [oh for name, oh in _entries(m, g_root_oh(m))][0]
    mine = {t: d for t, _f, d in m._messages(mine_oh)}
    assert bytes(mine[0x0003]) == bytes(ref_dt)                       # IEEE binary64 little endian, as H5T_IEEE_F64LE is stored
    ref_fill = next(d for t, _f, d in g._messages(ref_oh) if t == 0x0005)
    assert bytes(mine[0x0005]) == bytes(ref_fill)


def g_root_oh(f):
    import struct
    return struct.unpack_from("<Q", f.buf, f.sb + 24 + 32 + 8)[0]


def _entries(f, oh):
    import struct
    d = next(d for t, _fl, d in f._messages(oh) if t == 0x0011)
    bt, hp = struct.unpack_from("<QQ", d, 0)
    out = []
    f._btree_group(bt, f._heap(hp), out)
    return out


def test_grid_population_and_history_files_round_trip(h5, tmp_path):
    """The three kinds of file PINC's host writes (src/grid.c:1161-1270, src/population.c:497-651, src/io.c:657-733)."""
    rng = np.random.default_rng(1)
    # grid file: one dataset "/n=%.1f" per step, extents reversed (z, y, x, component), two attributes on the file
    out = str(tmp_path / "rho.grid.h5")
    f = h5.ph5Create(out.encode())
    one = np.array([0.005])
    assert h5.ph5Attr(f, b"Axis denormalization factor", dp(one), 1) == 0
    q = np.array([1.6e-19])
    assert h5.ph5Attr(f, b"Quantity denormalization factor", dp(q), 1) == 0
    steps = {n: rng.standard_normal((6, 5, 4, 1)) for n in (0.0, 1.0, 2.0, 10.0)}
    for n, a in steps.items():
        dataset(h5, f, "/n=%.1f" % n, a)
    assert h5.ph5Close(f) == 0
    g = h5mini.File(out)
    assert g.eof == len(g.buf) and g.base == 0
    assert sorted(g.root.links) == sorted("n=%.1f" % n for n in steps)
    for n, a in steps.items():
        d = g.get("/n=%.1f" % n)
        assert d.shape == a.shape and d.dtype == np.dtype("<f8") and np.array_equal(d.data, a)
    assert g.root.attrs["Axis denormalization factor"].tolist() == [0.005]
    assert g.root.attrs["Quantity denormalization factor"].tolist() == [1.6e-19]

    # population file: groups /pos, /vel, /pos/specie %i, datasets (N, 3) per step; written in pieces (one piece per rank in the reference)
    out = str(tmp_path / "pop.pop.h5")
    f = h5.ph5Create(out.encode())
    for s in range(2):
        assert h5.ph5Group(f, b"/pos/specie %d" % s) == 0 and h5.ph5Group(f, b"/vel/specie %d" % s) == 0
    pos = rng.random((1000, 3))
    dims = (C.c_ulonglong * 2)(1000, 3)
    d = h5.ph5Dataset(f, b"/pos/specie 1/n=3.0", 2, dims)
    for a, b in ((0, 400), (400, 1000)):
        piece = np.ascontiguousarray(pos[a:b]).reshape(-1)
        assert h5.ph5Write(f, d, 3 * a, piece.size, dp(piece)) == 0
    assert h5.ph5Write(f, d, 2999, 2, dp(pos.reshape(-1))) != 0                     # beyond the extent: refused
    assert h5.ph5Dataset(f, b"/pos/specie 1/n=3.0", 2, dims) < 0                    # exists already
    assert h5.ph5Close(f) == 0
    g = h5mini.File(out)
    assert sorted(g.root.links) == ["pos", "vel"] and sorted(g.get("/pos").links) == ["specie 0", "specie 1"]
    assert g.get("/vel/specie 0").links == {}
    assert np.array_equal(g.get("/pos/specie 1/n=3.0").data, pos)

    # history file: rows (x, y) appended step by step, one dataset per quantity
    out = str(tmp_path / "history.xy.h5")
    f = h5.ph5Create(out.encode())
    kin = h5.ph5XYCreate(f, b"/energy/kinetic/total")
    pot = h5.ph5XYCreate(f, b"/energy/potential/total")
    empty = h5.ph5XYCreate(f, b"/energy/kinetic/specie 0")
    assert min(kin, pot, empty) >= 0
    rows = rng.standard_normal((45, 2))
    for x, y in rows:
        assert h5.ph5XYAppend(f, kin, x, y) == 0
    assert h5.ph5XYAppend(f, pot, 1.0, 2.5) == 0
    assert h5.ph5Close(f) == 0
    g = h5mini.File(out)
    assert np.array_equal(g.get("/energy/kinetic/total").data, rows)
    assert g.get("/energy/potential/total").data.tolist() == [[1.0, 2.5]]
    assert g.get("/energy/kinetic/specie 0").shape == (0, 2)


@pytest.mark.parametrize("count", [0, 1, 8, 9, 256, 257, 3000])
def test_groups_of_any_size_keep_libhdf5s_lookup_invariants(h5, tmp_path, count):
    """One dataset per time step makes large groups: 9 names need two symbol-table nodes, 257 a second B-tree level, 3000 a
    third.  h5mini walks the tree and raises unless every name lies inside the keys that bracket its node and the names ascend."""
    out = str(tmp_path / ("many%d.h5" % count))
    f = h5.ph5Create(out.encode())
    vals = {}
    order = np.random.default_rng(count).permutation(count)
    for i in order:
        name = "n=%.1f" % (0.5 * i)
        vals[name] = np.array([float(i), -float(i)])
        dataset(h5, f, "/" + name, vals[name])
    assert h5.ph5Close(f) == 0
    g = h5mini.File(out)
    assert g.names_checked == count and len(g.root.links) == count
    assert list(g.root.links) == sorted(vals)                                   # stored in strcmp order
    for name, v in vals.items():
        assert np.array_equal(g.root.links[name].data, v)


def test_c_host_writes_the_reference_files(tmp_path):
    """host/pinc_b200_host with files:output set writes <prefix>_rho.grid.h5 ... as src/main.c does - exercised on the GPU box
    (tests/test_gpu_chost.py); here only: the option exists in the host's source and the writer links into it."""
    src = open(os.path.join(ROOT, "host", "pinc_main.c")).read()
    assert "files:output" in src and "ph5Create" in src
