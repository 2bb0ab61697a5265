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


def test_c_host_writes_the_reference_hdf5_files(tmp_path):
    """files:output: <prefix>_rho/phi/E.grid.h5, <prefix>_pop.pop.h5, <prefix>_history.xy.h5 as src/main.c:120-131, 262-266 writes them
    (format-level writer host/pinc_h5.c; read back with tests/h5mini.py, which is pinned to a file of the real library in
    tests/test_h5_writer.py): names, extents (z, y, x, component), attributes, and the contents against the Python driver."""
    import h5mini
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < os.path.getmtime(os.path.join(ROOT, "host", "pinc_main.c")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "host")])
    prefix = str(tmp_path / "run")
    r = subprocess.run([EXE, os.path.join(ROOT, "configs", "cold.ini")] + OVER + ["files:output=" + prefix, "files:particles=1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows = re.findall(r"n=(\d+) kinetic=(\S+) potential=(\S+) particles=(\d+)", r.stdout)
    over = {k.replace(":", "__").lower(): v for k, v in (o.split("=", 1) for o in OVER)}
    text, cfg = small_cfg("cold", **over)
    W = sim.World(cfg)
    W.set_particles(initial.perturb(cfg, initial.lattice(cfg)))
    W.migrate(); W.field_solve(); W.half_kick()
    names = ["n=%d.0" % n for n in range(1, 7)]
    files = {q: h5mini.File("%s_%s.grid.h5" % (prefix, q)) for q in ("rho", "phi", "E")}
    for q, f in files.items():
        assert sorted(f.root.links) == sorted(names)
        assert f.root.attrs["Quantity denormalization factor"].tolist() == [1.0]
        assert f.root.attrs["Axis denormalization factor"].shape == (1,) and f.root.attrs["Axis denormalization factor"][0] > 0
    pop = h5mini.File(prefix + "_pop.pop.h5")
    assert sorted(pop.root.links) == ["pos", "vel"] and sorted(pop.get("/pos").links) == ["specie 0", "specie 1"]
    for n in range(1, 7):
        W.step(fused=False)
        for q in ("rho", "phi", "E"):
            d = files[q].get("/n=%d.0" % n)
            ref = W.grid(0, q)                                    # (z, y, x, component) with the ghost layers
            true = ref[1:-1, 1:-1, 1:-1, :]
            assert d.shape == true.shape and d.shape[:3] == (8, 8, 16)
            assert np.abs(d.data - true).max() <= 1e-9 * max(np.abs(true).max(), 1e-300), (q, n)
        got = W.particles(0)
        for s in range(2):
            p = pop.get("/pos/specie %d/n=%d.0" % (s, n)).data
            v = pop.get("/vel/specie %d/n=%d.5" % (s, n)).data
            assert p.shape == got[s][0].shape and v.shape == got[s][1].shape
            # global frame = local frame + offset (src/population.c:727-763); offset = -nGhostLayers on the only rank
            assert np.abs(np.sort(p[:, 0]) - np.sort(got[s][0][:, 0] - 1.0)).max() <= 1e-9
            assert np.abs(np.sort(v[:, 0]) - np.sort(got[s][1][:, 0])).max() <= 1e-12
    W.close()
    hist = h5mini.File(prefix + "_history.xy.h5")
    kin = hist.get("/energy/kinetic/total").data
    pot = hist.get("/energy/potential/total").data
    assert kin.shape == (6, 2) and kin[:, 0].tolist() == [1, 2, 3, 4, 5, 6]
    assert np.array_equal(kin[:, 1], np.array([float(k) for _, k, _p, _n in rows]))          # %.17g on stdout is the same double
    assert np.array_equal(pot[:, 1], np.array([float(p) for _, _k, p, _n in rows]))
    assert sorted(hist.get("/energy/kinetic").links) == ["specie 0", "specie 1", "total"]
