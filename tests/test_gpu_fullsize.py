"""Size-independent properties at the BASELINE configs[1] size (64^3 cells, 36.7 M particles), where the oracle is
too slow to step: exact charge conservation of the fixed-point deposition, cell order after the sort, conservation
of the particle count through migration, idempotence of the ghost fill, run-to-run bit reproducibility of one step,
a residual history that ends below the reference's tolerance."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import ROOT
from pinc_b200 import abi, config, initial, sim

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world():
    ini = config.Ini(open(os.path.join(ROOT, "configs", "warm.ini")).read())
    cfg = config.load_config(ini)
    W = sim.World(cfg)
    W.set_particles({0: initial.maxwellian(cfg, seed=5, ranks=[0])[0]})
    W.migrate(); W.field_solve(); W.half_kick()
    yield W, cfg
    W.close()


def test_full_size_step_properties(world):
    W, cfg = world
    L, st = W.lib, W.ranks[0]
    n0 = W.n_particles()
    assert n0 == sum(cfg.nParticles) == 36700160
    for _ in range(2):
        W.step(fused="nodeposit")
    assert W.n_particles() == n0                                   # migration neither loses nor duplicates particles
    hist = W.history()
    assert 1 <= len(hist) < 200 and hist[-1] <= 1e-10 and all(h > 1e-10 for h in hist[:-1])
    # deposition: sum over ALL nodes (ghosts included, before the fold) of rho = sum_s q_s N_s up to the final
    # per-node fp64 conversion: the integer accumulation itself is exact
    L.puDistr3D1(st.pop, st.rho)
    rho = W.grid(0, "rho")
    q = np.array(cfg.charge)
    p = st.pop.contents
    ns = np.array([p.iStop[s] - p.iStart[s] for s in range(2)], dtype=float)
    assert abs(rho.sum() - (q * ns).sum()) <= 1e-9 * np.abs(q * ns).sum()
    again = W.grid(0, "rho")
    L.puDistr3D1(st.pop, st.rho)
    assert np.array_equal(W.grid(0, "rho"), again)                 # bit-reproducible
    # ghost fill is idempotent
    L.gHaloOp(W.set_slice, st.rho, st.mpi, abi.TOHALO)
    a = W.grid(0, "rho")
    L.gHaloOp(W.set_slice, st.rho, st.mpi, abi.TOHALO)
    assert np.array_equal(W.grid(0, "rho"), a)


def test_full_size_cell_order_after_extraction(world):
    W, cfg = world
    L, st = W.lib, W.ranks[0]
    L.puMove(st.pop, None)
    L.puExtractEmigrants3D(st.pop, st.mpi)
    p = st.pop.contents
    stay = [p.iStop[s] - p.iStart[s] for s in range(2)]
    L.puMigrate(st.pop, st.mpi, st.rho)
    L.pincSyncPopToHost(st.pop)
    pos, _ = abi.pop_arrays(p)
    size = np.array(cfg.trueSize) + 2
    for s in range(2):
        a = p.iStart[s]
        x = pos[a:a + stay[s]]
        cell = x[:, 0].astype(np.int64) + (size[0] - 1) * (x[:, 1].astype(np.int64) + (size[1] - 1) * x[:, 2].astype(np.int64))
        assert np.all(np.diff(cell) >= 0)                          # stayers are in cell order
        imm = pos[a + stay[s]:p.iStop[s]]
        assert np.all((imm >= 0.1) & (imm < size - 1.1))           # immigrants landed inside the thresholds
    # put the population back into a steppable state for other tests
    L.puDistr3D1(st.pop, st.rho)
