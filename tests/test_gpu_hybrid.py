"""The hybrid multi-rank multigrid solve (multigrid.cu: hybridSolve, k_mg_solve<false,true>): every rank smooths its own
sub-domain of the finest level block-resident, block faces across sub-domain boundaries travel through tagged slots in the
neighbour rank's peer-mapped arena, the coarser levels are replicated.  Ranks are host threads sharing one GPU (thread
transport; their persistent kernels run side by side on disjoint SMs), so the cross-rank slot protocol - the code the
NCCL ranks of `bench.py --gpus N` run over NVLink - is exercised on a one-GPU box.

One whole field solve (puDistr3D1 -> gHaloOp(add) -> mgSolve -> gHaloOp(set) -> gFinDiff1st) against the oracle's
DISTRIBUTED solve of the same problem: V-cycle count exact, residual norm per V-cycle 1e-6 relative above the rounding
floor, rho/phi/E of every rank (ghost layers included) <= 1e-10 relative; a second solve checks the persisted tags."""
import numpy as np
import pytest

from helpers import small_cfg
from oracle import orc
from pinc_b200 import initial, sim

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("sub,true,levels", [("1,1,2", "32,32,32", 4), ("2,2,2", "32,32,32", 4), ("1,2,1", "24,32,40", 3),
                                              ("2,1,2", "16,32,16", 3)])
def test_hybrid_solve_matches_oracle(gpu_lib, sub, true, levels):
    text, cfg = small_cfg("warm", grid__nsubdomains=sub, grid__truesize=true, multigrid__mglevels=levels,
                          population__nparticles="2 pc", population__nalloc="8 pc",
                          population__thermalvelocitycells="0.02,0.00046", grid__nemigrantsalloc="1 pc, 2 pc, 4 pc")
    per_rank = initial.maxwellian(cfg, seed=5)
    gpu_lib.pincMgSetHybrid(1)
    W = sim.World(cfg)
    O = orc.OrcWorld(cfg)
    try:
        for X in (W, O):
            X.set_particles(per_rank)
            X.migrate()
        for solve in range(2):
            for X in (W, O):
                X.field_solve()
            hw, ho = np.array(W.history()), np.array(O.history())
            assert len(hw) == len(ho) and (solve > 0 or len(ho) > 5), (len(hw), len(ho))
            assert np.all(np.abs(hw - ho) <= 1e-6 * ho + 1e-13), np.abs(hw / ho - 1).max()
            assert W.mg_path(0) == 9, W.mg_path(0)          # 1 (all-SM kernel) + 8 (hybrid)
            for r in range(cfg.nRanks):
                for name in ("rho", "phi", "E"):
                    err = rel(W.grid(r, name), O.grid(r, name))
                    assert err <= 1e-10, (solve, r, name, err)
    finally:
        W.close()
