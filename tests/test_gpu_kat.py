"""The reference's own known-answer vectors (tests/golden/kat_pusher.json, from /root/reference/test/pusher.test.c)
through the CUDA path (C-ABI), plus edge cases: an empty species, a single particle, a population that is all
emigrants.  The KAT grids (5x4x3 nodes, no ghost layers in the old tests) are the same arrays as a 3x2x1 true grid
with one ghost layer per side, which is what the library requires (puSanity)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from helpers import ROOT, GridH, da, ia, la, single_mpi
from pinc_b200 import abi

pytestmark = pytest.mark.gpu
KAT = json.load(open(os.path.join(ROOT, "tests", "golden", "kat_pusher.json")))
TRUE = (3, 2, 1)          # + ghosts = 5 x 4 x 3 nodes


def make_pop(L, species_particles, charge, mass, cap=100):
    nS = len(species_particles)
    p = L.pincPopAlloc(nS, 3, la([cap] * nS), da(charge), da(mass))
    pc = p.contents
    pos, vel = abi.pop_arrays(pc)
    for s, (ps, vs) in enumerate(species_particles):
        a = pc.iStart[s]
        n = len(ps)
        if n:
            pos[a:a + n] = ps
            vel[a:a + n] = vs
        pc.iStop[s] = a + n
    L.pincSyncPopToDevice(p)
    return p


def test_kat_interpolation_and_kick(gpu_lib):
    k, L = KAT["acc3d1"], gpu_lib
    E = GridH(L, TRUE, 3)
    E.a.reshape(-1)[:] = np.arange(E.a.size, dtype=float)
    E.up()
    p = make_pop(L, [(np.array(k["pos"]), np.array(k["vel"], dtype=float))], k["charge"], k["mass"])
    L.puAcc3D1(p, E.ptr)
    L.pincSyncPopToHost(p)
    _, vel = abi.pop_arrays(p.contents)
    assert np.abs(vel[0] - np.array(k["expect_vel0"])).max() < k["tol"]
    assert abs(vel[1, 0] - k["expect_vel1_x"]) < k["tol"]
    L.pincPopFree(p)


def test_kat_constant_field_leapfrog(gpu_lib):
    k, L = KAT["constE"], gpu_lib
    true = (6, 6, 6)
    E = GridH(L, true, 3)
    E.a[...] = np.array(k["E"], dtype=float)
    E.up()
    x0 = 3.0
    parts = [(np.array([[x0, 3.0, 3.0]]), np.zeros((1, 3))) for _ in range(3)]
    p = make_pop(L, parts, k["charge"], k["mass"])
    L.gMul(E.ptr, 0.5); L.puAcc3D1(p, E.ptr); L.gMul(E.ptr, 2.0)
    for n in (1, 2):
        L.puMove(p, None)
        L.puAcc3D1(p, E.ptr)
        L.pincSyncPopToHost(p)
        pos, _ = abi.pop_arrays(p.contents)
        for s in range(3):
            assert abs(pos[p.contents.iStart[s], 0] - (x0 + k["coef"][s] * n * n)) < 1e-14
    L.pincPopFree(p)


def test_kat_deposition(gpu_lib):
    L = gpu_lib
    for name in ("distr3d1", "distr3d1_renorm"):
        k = KAT[name]
        rho = GridH(L, TRUE, 1).up()
        if name == "distr3d1":
            parts = [(np.array(k["pos"]), np.zeros((len(k["pos"]), 3)))]
            mass = [1.0]
        else:
            parts = [(np.array(ps), np.zeros((len(ps), 3))) for ps in k["pos"]]
            mass = [1.0] * 3
        p = make_pop(L, parts, k["charge"], mass)
        L.puDistr3D1(p, rho.ptr)
        val = rho.down().reshape(-1)
        for node, v in k["expect"].items():
            assert abs(val[int(node)] - v) < k["tol"], (name, node)
        L.pincPopFree(p)
        rho.free()


def test_kat_extraction_counts_and_buffers(gpu_lib):
    """test/pusher.test.c:360-545: exact emigrant table and population sizes; the emigrant records and the remaining
    particles as multisets (order is an artefact of the reference's serial back-fill)."""
    k, L = KAT["extract3d"], gpu_lib
    true = (8, 8, 8)
    from test_oracle_kat import extraction_population
    pp = extraction_population()
    vel = np.array(k["vel"], dtype=float)
    parts = [(np.array(pp[0]), np.tile(vel, (len(pp[0]), 1))), (np.array(pp[1]), np.tile(vel, (len(pp[1]), 1))), (np.zeros((0, 3)), np.zeros((0, 3)))]
    p = make_pop(L, parts, [-1.0, 1.0, 2.0], [10.0, 1.0, 10.0])
    rho = GridH(L, true, 1)
    m = single_mpi(L, true, nS=3)
    L.pincCreateNeighborhood(m, rho.ptr, la([10]), 1, da([1.0] * 6))          # lower 1, upper size-1-1 = 8 ... the KAT uses 9
    for d in range(3):
        m.contents.thresholds[3 + d] = 9.0
    L.puExtractEmigrants3D(p, m)
    assert [m.contents.nEmigrants[i] for i in range(81)] == k["nEmigrants"]
    assert [p.contents.iStop[s] for s in range(3)] == k["iStop"]
    L.pincSyncPopToHost(p)
    pos, _ = abi.pop_arrays(p.contents)
    for s in (0, 1):
        a, b = p.contents.iStart[s], p.contents.iStop[s]
        assert sorted(pos[a:b, 0]) == sorted(k["left_x"])
        assert np.all(pos[a:b, 1:] == 5.0)
    L.puMigrate(p, m, rho.ptr)                   # single rank: everybody comes back through the periodic wrap
    assert [p.contents.iStop[s] - p.contents.iStart[s] for s in range(3)] == [len(pp[0]), len(pp[1]), 0]
    L.pincPopFree(p)


def test_empty_species_and_single_particle(gpu_lib):
    L = gpu_lib
    true = (4, 4, 4)
    E, rho = GridH(L, true, 3).up(), GridH(L, true, 1).up()
    m = single_mpi(L, true, nS=2)
    L.pincCreateNeighborhood(m, rho.ptr, la([4]), 1, da([0.1] * 6))
    p = make_pop(L, [(np.zeros((0, 3)), np.zeros((0, 3))), (np.array([[2.25, 2.5, 2.75]]), np.array([[0.1, 0.0, -0.1]]))], [-1.0, 2.0], [1.0, 4.0])
    L.puAcc3D1KE(p, E.ptr)
    assert p.contents.kinEnergy[0] == 0.0 and abs(p.contents.kinEnergy[1] - 0.5 * 4.0 * 0.02) < 1e-15
    L.puMove(p, None)
    L.puExtractEmigrants3D(p, m)
    L.puMigrate(p, m, rho.ptr)
    L.puDistr3D1(p, rho.ptr)
    val = rho.down()
    assert abs(val.sum() - 2.0) < 1e-13 and (val != 0).sum() == 8
    L.pincPopFree(p)


def test_everybody_emigrates(gpu_lib):
    """All particles sit in the emigration bands: the live range empties and refills through the periodic wrap."""
    L = gpu_lib
    true = (4, 4, 4)
    rho = GridH(L, true, 1).up()
    m = single_mpi(L, true, nS=1)
    L.pincCreateNeighborhood(m, rho.ptr, la([64]), 1, da([0.1] * 6))
    rng = np.random.default_rng(2)
    n = 50
    pos = np.column_stack([0.05 * rng.random(n), 1 + 3 * rng.random(n), 4.95 + 0.04 * rng.random(n)])
    p = make_pop(L, [(pos, np.zeros((n, 3)))], [1.0], [1.0])
    L.puExtractEmigrants3D(p, m)
    assert p.contents.iStop[0] == p.contents.iStart[0]
    assert sum(m.contents.nEmigrants[i] for i in range(27)) == n
    L.puMigrate(p, m, rho.ptr)
    assert p.contents.iStop[0] - p.contents.iStart[0] == n
    L.pincSyncPopToHost(p)
    got, _ = abi.pop_arrays(p.contents)
    exp = pos + np.array([4.0, 0.0, -4.0])
    assert np.array_equal(np.sort(got[:n], axis=0), np.sort(exp, axis=0))
    L.pincPopFree(p)


def test_debug_scans_pass_and_fail_like_the_reference(gpu_lib):
    """pPosAssertInLocalFrame / pVelAssertMax (src/population.c:316-365): silent when satisfied, msg(ERROR) + exit
    otherwise (checked in a child process, because the error convention is exit(EXIT_FAILURE))."""
    import subprocess, sys, textwrap
    L = gpu_lib
    true = (4, 4, 4)
    rho = GridH(L, true, 1)
    p = make_pop(L, [(np.array([[1.5, 2.5, 3.5], [4.9, 0.1, 2.0]]), np.array([[0.5, -2.0, 0.1], [0.9, 0.0, -0.9]]))], [1.0], [1.0])
    L.pPosAssertInLocalFrame(p, rho.ptr)
    L.pVelAssertMax(p, 1.0)
    L.pincPopFree(p)
    code = textwrap.dedent("""
        import sys, numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        from pinc_b200 import lib as plib
        from helpers import GridH
        from test_gpu_kat import make_pop
        L = plib.load()
        rho = GridH(L, (4, 4, 4), 1)
        p = make_pop(L, [(np.array([[1.5, 2.5, 3.5], [2.0, %s, 2.0]]), np.array([[0.5, 0.0, 0.1], [0.0, %s, 0.0]]))], [1.0], [1.0])
        L.pPosAssertInLocalFrame(p, rho.ptr)
        L.pVelAssertMax(p, 1.0)
        print("survived")
    """)
    for posy, vely, msg in (("5.5", "0.0", "is out of bounds in dimension 1"), ("2.0", "1.5", "travels too fast in dimension 1")):
        r = subprocess.run([sys.executable, "-c", code % (ROOT, os.path.join(ROOT, "tests"), posy, vely)], capture_output=True, text=True, timeout=300)
        assert r.returncode != 0 and "survived" not in r.stdout
        assert "Particle i=1 (of specie 0) " + msg in r.stderr, r.stderr[-500:]


def test_kat_pnew_pcut(gpu_lib):
    """test/population.test.c:10-56 through the device population: pNew x 4, pCut(pop,1,33) twice."""
    L = gpu_lib
    k = KAT["pcut"]
    p = L.pincPopAlloc(2, 3, la(k["nAlloc"]), da([-1.0, 1.0]), da([1.0, 100.0]))
    L.pincSyncPopToDevice(p)
    for p3, v3 in k["new"]:
        L.pNew(p, 1, da(p3), da(v3))
    assert p.contents.iStop[1] == k["nAlloc"][0] + 4
    for which in ("first", "second"):
        p3, v3 = (C.c_double * 3)(), (C.c_double * 3)()
        L.pCut(p, 1, k["cut_flat_index"], p3, v3)
        assert list(p3) == k[which]["pos"] and list(v3) == k[which]["vel"] and p.contents.iStop[1] == k[which]["iStop1"]
    # what is left is particles 0 and 2 of the four, in the reference's back-filled order
    L.pincSyncPopToHost(p)
    pos, vel = abi.pop_arrays(p.contents)
    a = p.contents.iStart[1]
    assert pos[a:a + 2].tolist() == [[0, 1, 2], [6, 7, 8]] and vel[a:a + 2].tolist() == [[0, 10, 20], [60, 70, 80]]
    L.pincPopFree(p)
