"""The multi-rank solve at the grid geometry `bench.py --gpus 2/8` times: sub-domains of 64^3 cells decomposed 1,1,2
and 2,2,2 (global 64x64x128 and 128^3), ranks as host threads on one GPU (thread transport), one whole field solve
(puDistr3D1 -> gHaloOp(add) -> mgSolve -> gHaloOp(set) -> gFinDiff1st) against the oracle's DISTRIBUTED solve of
the same problem (the reference's own sources, oracle/_ref, one host thread per rank, where that library exists;
else the oracle restatement, single-threaded):

  * V-cycle count: exact (121 on these grids); residual norm per V-cycle: 1e-6 relative above the rounding floor;
  * rho, phi, E of every rank, ghost layers included: <= 1e-10 relative;
  * the solve ran replicated (pincMgLastPath >= 4), i.e. through the kernel the N > 1 bench lines time.

Few particles per cell (the particle path is covered elsewhere); the oracle's solve takes ~10-30 s per case."""
import numpy as np
import pytest

from helpers import small_cfg
from oracle import orc, ref
from pinc_b200 import config
from pinc_b200 import initial, sim

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("sub", ["1,1,2", "2,2,2"])
def test_replicated_solve_at_bench_geometry(gpu_lib, sub):
    text, cfg = small_cfg("warm", grid__nsubdomains=sub, population__nparticles="0.5 pc", population__nalloc="2 pc",
                          population__thermalvelocitycells="0.02,0.00046", grid__nemigrantsalloc="0.25 pc, 0.5 pc, 2 pc")
    assert cfg.trueSize == [64, 64, 64] and cfg.mgLevels == 5
    per_rank = initial.maxwellian(cfg, seed=3)
    use_ref = ref.available()
    W = sim.World(cfg)
    O = ref.RefWorld(config.Ini(text).dump(), cfg.nRanks) if use_ref else orc.OrcWorld(cfg)
    try:
        ho = []
        for X in (W, O):
            X.set_particles(per_rank)
            X.migrate()
            if X is O and use_ref:
                X.field_solve(history=ho)
            else:
                X.field_solve()
        if not use_ref:
            ho = O.history()
        hw, ho = np.array(W.history()), np.array(ho)
        assert len(hw) == len(ho) and len(ho) > 60, (len(hw), len(ho))
        assert np.all(np.abs(hw - ho) <= 1e-6 * ho + 1e-13), np.abs(hw / ho - 1).max()
        assert W.mg_path(0) >= 4
        for r in range(cfg.nRanks):
            for name in ("rho", "phi", "E"):
                err = rel(W.grid(r, name), O.grid(r, name))
                assert err <= 1e-10, (r, name, err)
    finally:
        W.close()
        if use_ref:
            O.close()
