"""Drop-in proof: oracle/_ref/dropin_host is a PINC host compiled against the reference's own headers and
linked with the reference's own io.o, units.o, aux.o, population.o, grid.o and iniparser (tests/c/dropin_host.c, built by
`make -C oracle dropin` where /root/reference exists; the binaries travel in oracle/_ref).  It reads the ini with the
reference's iniOpen, picks its methods with the reference's select() -> X_set(ini), allocates Units, MpiInfo, Population,
Grids with the reference's uAlloc / gAllocMpi / pAlloc / gAlloc / gCreateNeighborhood, fills the particles with the
reference's pPosLattice / pPosPerturb / pVelZero - and every compute entry point of the time step (puMove,
puExtractEmigrants3D, puMigrate, puDistr3D1, gHaloOp, mgSolve, gFinDiff1st, gMul, puAcc3D1KE, pSumKinEnergy, gPotEnergy)
and mgAllocSolver(ini, rho, phi) resolve to libpinc_b200.so, which reads the dictionary through the reference's iniGet*.

Its CPU twin dropin_host_ref is the same source linked with the reference's pusher.o and multigrid.o.  Same ini, same
steps: particle counts equal, kinetic and potential energy of every step equal to 1e-10 relative, final phase-space
checksum equal to 1e-12 relative."""
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu
GPU = os.path.join(ROOT, "oracle", "_ref", "dropin_host")
CPU = os.path.join(ROOT, "oracle", "_ref", "dropin_host_ref")


def run(exe, ini, over):
    r = subprocess.run([exe, os.path.join(ROOT, "configs", ini)] + over, capture_output=True, text=True, timeout=900, cwd="/tmp")
    assert r.returncode == 0, (exe, r.stdout[-1500:], r.stderr[-1500:])
    rows = re.findall(r"n=(\d+) kinetic=(\S+) potential=(\S+) particles=(\d+)", r.stdout)
    cs = float(re.search(r"checksum=(\S+)", r.stdout).group(1))
    return np.array([[float(k), float(p), float(n)] for _, k, p, n in rows]), cs, r.stdout


@pytest.mark.parametrize("ini,over", [
    ("cold.ini", ["grid:nSubdomains=1,1,1", "grid:trueSize=32,16,16", "multigrid:mgLevels=4", "population:nParticles=8 pc",
                  "population:nAlloc=16 pc", "population:perturbAmplitude=2e-3,0,0,0,0,0", "grid:nEmigrantsAlloc=4 pc", "time:nTimeSteps=10"]),
    ("cold.ini", ["grid:nSubdomains=1,1,1", "grid:trueSize=16,8,8", "multigrid:mgLevels=3", "population:nParticles=27 pc",
                  "population:nAlloc=32 pc", "population:perturbAmplitude=5e-3,0,0,0,0,0", "grid:nEmigrantsAlloc=4 pc", "time:nTimeSteps=12"]),
    # the N-dimensional select() targets (src/main.c:58-70) through the reference's select()
    ("cold.ini", ["grid:nSubdomains=1,1,1", "grid:trueSize=16,8,8", "multigrid:mgLevels=3", "population:nParticles=27 pc",
                  "population:nAlloc=32 pc", "population:perturbAmplitude=5e-3,0,0,0,0,0", "grid:nEmigrantsAlloc=4 pc", "time:nTimeSteps=10",
                  "methods:acc=puAccND1KE", "methods:distr=puDistrND1", "methods:migrate=puExtractEmigrantsND"]),
])
def test_reference_host_runs_on_the_library(ini, over):
    if not (os.path.exists(GPU) and os.path.exists(CPU)):
        pytest.skip("oracle/_ref/dropin_host not built (needs /root/reference at build time: make -C oracle dropin)")
    got, cs_g, out = run(GPU, ini, over)
    ref, cs_r, _ = run(CPU, ini, over)
    assert "library=pinc-b200" in out and re.search(r"mg_path=[12]", out), out[-300:]
    assert got.shape == ref.shape and len(ref) >= 10
    assert np.array_equal(got[:, 2], ref[:, 2])
    for col, name in ((0, "kinetic"), (1, "potential")):
        err = np.abs(got[:, col] - ref[:, col]).max() / np.abs(ref[:, col]).max()
        assert err <= 1e-10, (name, err)
    assert abs(cs_g - cs_r) <= 1e-12 * abs(cs_r), (cs_g, cs_r)
