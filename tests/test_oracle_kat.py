"""The oracle against the reference's own known-answer vectors (tests/golden/kat_*.json, transcribed from
/root/reference/test/pusher.test.c and test/grid.test.c; each entry cites the test it comes from)."""
import ctypes as C
import json
import os

import numpy as np

from oracle import orc

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat_pusher.json")))


def i32(v):
    return np.array(v, dtype=np.int32)


def i64(v):
    return np.array(v, dtype=np.int64)


def f64(v):
    return np.array(v, dtype=np.float64)


def test_interpolation_and_kick():
    k = KAT["acc3d1"]
    O = orc.load()
    size = i32(k["size"])
    E = np.arange(3 * int(np.prod(size)), dtype=np.float64)
    pos, vel = f64(k["pos"]).reshape(-1), f64(k["vel"]).reshape(-1)
    O.orc_acc3d1(orc.dp(pos), orc.dp(vel), 1, orc.lp(i64([0, 10])), orc.lp(i64([2])), orc.dp(f64(k["charge"])),
                 orc.dp(f64(k["mass"])), orc.dp(E), orc.ip(size), None)
    assert np.abs(vel[:3] - f64(k["expect_vel0"])).max() < k["tol"]
    assert abs(vel[3] - k["expect_vel1_x"]) < k["tol"]


def test_constant_field_leapfrog():
    k = KAT["constE"]
    O = orc.load()
    size = i32([256, 256, 256])[:]
    size = i32([8, 8, 8])                     # the field is uniform: the extent does not enter the answer
    n = int(np.prod(size))
    E = np.tile(f64(k["E"]), n)
    nS = 3
    iStart = i64([0, 10, 20, 30]); iStop = i64([1, 11, 21])
    pos = np.zeros(90); vel = np.zeros(90)
    for s in range(nS):
        pos[30 * s:30 * s + 3] = [4.0, 4.0, 4.0]
    q, m = f64(k["charge"]), f64(k["mass"])
    x0 = 4.0
    O.orc_gmul(orc.dp(E), E.size, 0.5)
    O.orc_acc3d1(orc.dp(pos), orc.dp(vel), nS, orc.lp(iStart), orc.lp(iStop), orc.dp(q), orc.dp(m), orc.dp(E), orc.ip(size), None)
    O.orc_gmul(orc.dp(E), E.size, 2.0)
    # 2 steps keep the particle inside the 8^3 box (x = 4 + 0.5 n^2); the reference test runs 5 in a 256^3 box
    for nstep in (1, 2):
        O.orc_move(orc.dp(pos), orc.dp(vel), nS, orc.lp(iStart), orc.lp(iStop))
        O.orc_acc3d1(orc.dp(pos), orc.dp(vel), nS, orc.lp(iStart), orc.lp(iStop), orc.dp(q), orc.dp(m), orc.dp(E), orc.ip(size), None)
        for s in range(nS):
            assert abs(pos[30 * s] - (x0 + k["coef"][s] * nstep ** 2)) < 1e-14


def test_deposition_single_species():
    k = KAT["distr3d1"]
    O = orc.load()
    size = i32(k["size"])
    rho = np.zeros(int(np.prod(size)))
    pos = f64(k["pos"]).reshape(-1)
    O.orc_distr3d1(orc.dp(pos), 1, orc.lp(i64([0, 10])), orc.lp(i64([4])), orc.dp(f64(k["charge"])), orc.dp(rho), orc.ip(size))
    for node, val in k["expect"].items():
        assert abs(rho[int(node)] - val) < k["tol"], node
    assert abs(rho.sum() - 4.0) < 1e-13


def test_deposition_species_renormalisation():
    k = KAT["distr3d1_renorm"]
    O = orc.load()
    size = i32(k["size"])
    rho = np.zeros(int(np.prod(size)))
    pos = np.zeros(90)
    for s in range(3):
        pos[30 * s:30 * s + 3] = k["pos"][s][0]
    O.orc_distr3d1(orc.dp(pos), 3, orc.lp(i64([0, 10, 20, 30])), orc.lp(i64([1, 11, 21])), orc.dp(f64(k["charge"])), orc.dp(rho), orc.ip(size))
    for node, val in k["expect"].items():
        assert abs(rho[int(node)] - val) < k["tol"], node


def extraction_population():
    """The particles of test/pusher.test.c:384-400 in the order pNew placed them."""
    pos = {0: [], 1: []}
    x = 0.0
    while x <= 10:
        for s in (0, 1):
            pos[s].append([x, 5.0, 5.0])
        x += 0.5
    for z in (-1, 0, 1):
        for y in (-1, 0, 1):
            for xx in (-1, 0, 1):
                for s in (0, 1):
                    pos[s].append([5 + xx * 4.5, 5 + y * 4.5, 5 + z * 4.5])
    return pos


def test_emigrant_extraction_exact_order():
    k = KAT["extract3d"]
    O = orc.load()
    nS = 3
    iStart = i64([0, 100, 200, 300]); iStop = i64([0, 100, 200])
    pos = np.zeros(900); vel = np.zeros(900)
    pp = extraction_population()
    for s in (0, 1):
        a = np.array(pp[s])
        n = len(a)
        pos[3 * iStart[s]:3 * (iStart[s] + n)] = a.reshape(-1)
        vel[3 * iStart[s]:3 * (iStart[s] + n)] = np.tile(k["vel"], n)
        iStop[s] = iStart[s] + n
    emig = [np.zeros(6 * 10) for _ in range(27)]
    nEm = np.zeros(27 * nS, dtype=np.int64)
    O.orc_extract3d(orc.dp(pos), orc.dp(vel), nS, orc.lp(iStart), orc.lp(iStop), orc.dp(f64(k["thresholds"])), orc.ptr_array(emig), orc.lp(nEm))
    assert list(nEm) == k["nEmigrants"]
    assert list(emig[12][:36:6]) == k["emigrants12_x"]
    assert list(emig[14][:48:6]) == k["emigrants14_x"]
    assert list(emig[14][1:48:6]) == [5.0] * 8 and list(emig[14][3:48:6]) == [1.0] * 8
    for z in (-1, 0, 1):
        for y in (-1, 0, 1):
            for x in (-1, 0, 1):
                ne = (x + 1) + (y + 1) * 3 + (z + 1) * 9
                if ne < 12 or ne > 14:
                    rec = [5 + x * 4.5, 5 + y * 4.5, 5 + z * 4.5, 1.0, 2.0, 3.0]
                    assert list(emig[ne][:12]) == rec + rec
    assert list(iStop) == k["iStop"]
    for s in (0, 1):
        assert list(pos[3 * iStart[s]:3 * iStop[s]:3]) == k["left_x"]


def test_neighbour_maps():
    k = KAT["rank_neighbor"]
    O = orc.load()
    t = orc.make_topo(k["nSubdomains"], [4, 4, 4])
    for ne, rank in k["neighborToRank"].items():
        assert O.orc_neighbor_to_rank(C.byref(t), k["rank"], int(ne)) == rank
        assert O.orc_rank_to_neighbor(C.byref(t), k["rank"], rank) == int(ne)
    for ne, rec in k["reciprocal"].items():
        assert O.orc_neighbor_to_reciprocal(int(ne)) == rec


def test_gradient_and_laplacian_analytic():
    """Analytic cases of test/grid.test.c:62-143, 203-262: phi = x^2 - z -> (2x, 0, -1); x - 2y^2 + 4z^3 -> -4 + 24z."""
    O = orc.load()
    size = i32([7, 6, 5])
    z, y, x = np.meshgrid(*[np.arange(s, dtype=float) for s in size[::-1]], indexing="ij")
    phi = (x * x - z).reshape(-1)
    E = np.zeros(3 * phi.size)
    O.orc_findiff1st(orc.dp(phi), orc.dp(E), orc.ip(size))
    Ev = E.reshape(size[2], size[1], size[0], 3)[1:-1, 1:-1, 1:-1]
    assert np.array_equal(Ev[..., 0], (2 * x)[1:-1, 1:-1, 1:-1])
    assert not Ev[..., 1].any() and np.all(Ev[..., 2] == -1)
    f = (x - 2 * y * y + 4 * z ** 3).reshape(-1)
    res = np.zeros(f.size)
    O.orc_residual(orc.dp(res), orc.dp(np.zeros(f.size)), orc.dp(f), orc.ip(size))
    rv = res.reshape(size[2], size[1], size[0])[1:-1, 1:-1, 1:-1]
    assert np.abs(rv - (-4 + 24 * z)[1:-1, 1:-1, 1:-1]).max() < 1e-11


def test_pnew_pcut_backfill():
    """test/population.test.c:10-56: pNew appends at iStop[s]; pCut returns the particle and back-fills with the last."""
    k = KAT["pcut"]
    O = orc.load()
    cap = k["nAlloc"]
    iStart = i64([0, cap[0], cap[0] + cap[1]])
    iStop = i64([0, cap[0]])
    pos, vel = np.zeros(3 * (cap[0] + cap[1])), np.zeros(3 * (cap[0] + cap[1]))
    for p3, v3 in k["new"]:
        assert O.orc_pnew(orc.dp(pos), orc.dp(vel), orc.lp(iStart), orc.lp(iStop), 1, orc.dp(f64(p3)), orc.dp(f64(v3))) == 1
    assert iStop[1] == cap[0] + 4
    for which in ("first", "second"):
        p3, v3 = np.zeros(3), np.zeros(3)
        O.orc_pcut(orc.dp(pos), orc.dp(vel), orc.lp(iStop), 1, k["cut_flat_index"], orc.dp(p3), orc.dp(v3))
        assert list(p3) == k[which]["pos"] and list(v3) == k[which]["vel"] and iStop[1] == k[which]["iStop1"]
    # a full species ignores the new particle (population.c:436-438)
    iStop[1] = iStart[2]
    assert O.orc_pnew(orc.dp(pos), orc.dp(vel), orc.lp(iStart), orc.lp(iStop), 1, orc.dp(f64([1, 1, 1])), orc.dp(f64([0, 0, 0]))) == 0
