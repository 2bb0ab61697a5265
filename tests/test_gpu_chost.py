"""The C host (host/pinc_main.c: PINC's regular() over the C-ABI, own ini reader and unit normalisation in C)
against the Python driver on the same ini: identical lattice + perturbation start, so the energy history must agree
to rounding (the two hosts call the same entry points in the same order)."""
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import ROOT, small_cfg
from pinc_b200 import initial, sim

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "host", "pinc_b200_host")
OVER = ["grid:nSubdomains=1,1,1", "grid:trueSize=16,8,8", "multigrid:mgLevels=3", "population:nParticles=8 pc",
        "population:nAlloc=16 pc", "population:perturbAmplitude=2e-3,0,0,0,0,0", "grid:nEmigrantsAlloc=4 pc", "time:nTimeSteps=6"]


@pytest.mark.parametrize("fused", [0, 1])
def test_c_host_matches_python_driver(fused):
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "host")])
    r = subprocess.run([EXE, os.path.join(ROOT, "configs", "cold.ini")] + OVER + [f"methods:fused={fused}"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows = re.findall(r"n=(\d+) kinetic=(\S+) potential=(\S+) particles=(\d+)", r.stdout)
    assert len(rows) == 6
    got = np.array([[float(k), float(p)] for _, k, p, _ in rows])
    over = {k.replace(":", "__").lower(): v for k, v in (o.split("=", 1) for o in OVER)}
    text, cfg = small_cfg("cold", **over)
    W = sim.World(cfg)
    W.set_particles(initial.perturb(cfg, initial.lattice(cfg)))
    W.migrate(); W.field_solve(); W.half_kick()
    ref = []
    for _ in range(6):
        W.step(fused=bool(fused))
        ref.append(W.energies())
    n_py = W.n_particles()
    W.close()
    ref = np.array(ref)
    assert int(rows[-1][3]) * 2 == n_py
    assert np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max()
    assert '"transport": "self"' in r.stdout
