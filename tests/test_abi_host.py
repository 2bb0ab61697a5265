"""CPU-side checks of the drop-in boundary: libpinc_b200.so loads without a GPU and exports every symbol that
include/pinc_b200.h declares; the ctypes mirrors have the C struct layouts; the host-struct constructors and the
neighbour maps reproduce the reference's known answers (tests/golden/kat_*.json).  No compute entry point is
called here (they need a device and have no CPU fallback)."""
import ctypes as C
import json
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from helpers import ROOT, da, ia, la
from pinc_b200 import abi, lib as plib

HEADER = os.path.join(ROOT, "include", "pinc_b200.h")
KATG = json.load(open(os.path.join(ROOT, "tests", "golden", "kat_grid.json")))
KATP = json.load(open(os.path.join(ROOT, "tests", "golden", "kat_pusher.json")))


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:[A-Za-z_][\w \*]*?)[\s\*]+(\w+)\s*\([^;{]*\)\s*;", src, flags=re.M)
    return sorted(set(n for n in names if n not in ("funPtr", "void", "int", "double")))


def test_library_exports_every_declared_symbol():
    L = plib.load()
    names = declared_functions()
    assert len(names) > 60
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/pinc_b200.h but not exported"
        assert n in plib.SIGNATURES, f"{n} has no ctypes signature in pinc_b200/lib.py"
    assert L.pincVersion().startswith(b"pinc-b200")


def test_ctypes_struct_layout_matches_header():
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "pinc_b200.h"
int main(){
 printf("%zu %zu %zu %zu %zu\n", sizeof(Population), sizeof(MpiInfo), sizeof(Grid), sizeof(Multigrid), sizeof(MultigridSolver));
 printf("%zu %zu %zu %zu\n", offsetof(Population,nSpecies), offsetof(MpiInfo,thresholds), offsetof(Grid,bnd), offsetof(Multigrid,prolongator));
 return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    got = [C.sizeof(abi.Population), C.sizeof(abi.MpiInfo), C.sizeof(abi.Grid), C.sizeof(abi.Multigrid), C.sizeof(abi.MultigridSolver),
           abi.Population.nSpecies.offset, abi.MpiInfo.thresholds.offset, abi.Grid.bnd.offset, abi.Multigrid.prolongator.offset]
    assert got == [int(x) for x in out]


def test_grid_alloc_known_answer():
    k = KATG["galloc"]
    L = plib.load()
    g = L.pincGridAlloc(3, ia(k["trueSize"]), ia(k["nGhost"]), k["nValues"], None)
    gc = g.contents
    assert gc.rank == 4
    assert [gc.size[i] for i in range(4)] == k["size"]
    assert [gc.sizeProd[i] for i in range(5)] == k["sizeProd"]
    assert [gc.nGhostLayers[i] for i in range(8)] == [0] * 8
    assert not abi.grid_array(gc).any()                      # zero-initialised (quirk Q5)
    L.pincGridFree(g)


def test_neighbourhood_known_answer():
    k = KATG["neighborhood"]
    L = plib.load()
    g = L.pincGridAlloc(3, ia(k["trueSize"]), ia(k["nGhost"]), 1, None)
    for spec, expect in ((k["alloc_full"], k["alloc_full"]), (k["alloc_smart_in"], k["alloc_smart"]),
                         (k["alloc_equal_in"], [4] * 13 + [0] + [4] * 13)):
        m = L.pincMpiAlloc(3, 2, ia([1, 1, 1]), ia(k["nGhost"]), ia(k["trueSize"]), 0, 1)
        spec_in = list(spec)
        L.pincCreateNeighborhood(m, g, la(spec_in), len(spec_in), da(k["thresholds_current_in"]))
        mc = m.contents
        assert mc.nNeighbors == 27 and mc.neighborhoodCenter == 13
        assert [mc.thresholds[i] for i in range(6)] == k["thresholds_out"]
        assert [mc.nEmigrantsAlloc[i] for i in range(27)] == list(expect)
        L.pincMpiFree(m)
    L.pincGridFree(g)


def test_rank_neighbour_maps_known_answer():
    k = KATP["rank_neighbor"]
    L = plib.load()
    m = L.pincMpiAlloc(3, 2, ia(k["nSubdomains"]), ia([1] * 6), ia([4, 4, 4]), k["rank"], 60)
    assert [m.contents.subdomain[i] for i in range(3)] == k["subdomain"]
    for ne, rank in k["neighborToRank"].items():
        assert L.puNeighborToRank(m, int(ne)) == rank
        assert L.puRankToNeighbor(m, rank) == int(ne)
    for ne, rec in k["reciprocal"].items():
        assert L.puNeighborToReciprocal(int(ne), 3) == rec
    L.pincMpiFree(m)


def test_pu_sanity_mirrors_puSanity():
    L = plib.load()
    err = C.create_string_buffer(200)
    ok = L.pincPuSanity(b"puAcc3D1KE", 3, ia([1] * 6), da([0.1] * 6), 3, 1, err, 200)
    assert ok == 0 and err.value == b""
    assert L.pincPuSanity(b"puAcc3D1KE", 2, ia([1] * 4), da([0.1] * 4), 3, 1, err, 200) == 1
    assert b"only supports grid:nDims=3" in err.value
    assert L.pincPuSanity(b"puDistr3D1", 3, ia([0] * 6), da([0.1] * 6), 3, 1, err, 200) == 2
    assert L.pincPuSanity(b"puDistr3D1", 3, ia([1] * 6), da([-0.1] * 6), 3, 1, err, 200) == 3
    assert L.pincPuSanity(b"puDistr3D1", 3, ia([1] * 6), da([0.7] * 6), 3, 1, err, 200) == 4


def test_rotation_parameters_and_solver_structs_on_host():
    L = plib.load()
    T, S = np.zeros(6), np.zeros(6)
    B, q, m = np.array([0.0, 0.0, 2.0]), np.array([-1.0, 1.0]), np.array([1.0, 4.0])
    dp = lambda a: a.ctypes.data_as(abi.c_double_p)
    L.pincGet3DRotationParameters(2, dp(B), dp(q), dp(m), dp(T), dp(S))
    assert np.allclose(T, [0, 0, -1.0, 0, 0, 0.25]) and np.allclose(S, [0, 0, -1.0, 0, 0, 0.5 / 1.0625])
    rho = L.pincGridAlloc(3, ia([16, 8, 8]), ia([1] * 6), 1, ia([abi.PERIODIC] * 6))
    phi = L.pincGridAlloc(3, ia([16, 8, 8]), ia([1] * 6), 1, ia([abi.PERIODIC] * 6))
    s = L.pincMgAllocSolver(rho, phi, 3, 1, 10, 10, 10)
    sc = s.contents
    assert sc.mgRho.contents.nLevels == 3
    # level 0 aliases the caller's grids (src/multigrid.c:142), coarser levels halve (src/multigrid.c:148)
    assert C.addressof(sc.mgRho.contents.grids[0].contents) == C.addressof(rho.contents)
    assert C.addressof(sc.mgPhi.contents.grids[0].contents) == C.addressof(phi.contents)
    g2 = sc.mgPhi.contents.grids[2].contents
    assert [g2.trueSize[i] for i in range(4)] == [1, 4, 2, 2] and [g2.size[i] for i in range(4)] == [1, 6, 4, 4]
    L.mgFreeSolver(s)
    L.pincGridFree(rho); L.pincGridFree(phi)


def test_ini_entry_points_from_a_pinc_style_host(tmp_path):
    """The entry points that take `dictionary *ini` (X_set selectors of io.h:105 select(), the mgSolver triple,
    mgAllocSolver, puGet3DRotationParameters) driven from a C host that brings its own iniGet* layer, as PINC's
    main.c:55-99 does; wrong configurations end like msg(ERROR)."""
    exe = str(tmp_path / "ini_host")
    so_dir = os.path.join(ROOT, "pinc_b200")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "ini_host.c"),
                           "-o", exe, "-L", so_dir, "-lpinc_b200", "-Wl,-rpath," + so_dir])
    r = subprocess.run([exe, "ok"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ini-host-ok" in r.stdout, (r.returncode, r.stderr)
    r = subprocess.run([exe, "badcycle"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "PINC-B200 ERROR" in r.stderr and "multigrid:cycle = mgFMG" in r.stderr
    r = subprocess.run([exe, "baddims"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "only supports grid:nDims=3" in r.stderr


def test_ini_entry_points_need_the_hosts_ini_layer():
    """Without iniGet* in the process (this Python host) the ini-taking entry points fail loudly instead of guessing."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from pinc_b200 import lib\n"
            "L = lib.load()\nL.puAcc3D1_set(None)\nprint('survived')\n") % ROOT
    r = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "survived" not in r.stdout and "ini layer" in r.stderr


def test_missing_device_fails_loudly():
    """No CPU fallback: without a GPU a compute entry point terminates with the reference's ERROR convention."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from pinc_b200 import lib, abi\nimport ctypes as C\n"
            "L = lib.load()\n"
            "g = L.pincGridAlloc(3, (C.c_int*3)(4,4,4), (C.c_int*6)(1,1,1,1,1,1), 1, None)\n"
            "L.gZero(g)\nprint('survived')\n") % ROOT
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode != 0 and "survived" not in r.stdout
    assert "PINC-B200 ERROR" in r.stderr and "no CPU path" in r.stderr
