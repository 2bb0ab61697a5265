"""Cell-slotted particle storage (particles.cu "slotted mode"): the steady state of
pincAccMove3D1KE -> puExtractEmigrants3D -> puMigrate -> puDistr3D1, in which only the particles that change cell move.

Against the oracle (reference order, CPU), on one sub-domain and on four thread ranks with cross-rank migration:
  * the population really is slotted while the loop runs (pincPopLayout == 1);
  * rho, phi, E <= 1e-10 relative, energies <= 1e-10, migrant tables and population sizes exact, every step;
  * phase space as a multiset after leaving the mode (any other entry point restores the contiguous planes);
  * fields bit-identical to the counting-sort path (the deposition is integer arithmetic, the push per particle);
  * a cell that outgrows its slots (head room forced to zero) falls back to the sort for a step and nothing is lost."""
import numpy as np
import pytest

from helpers import small_cfg, sorted_particles
from oracle import orc
from pinc_b200 import initial, sim

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def warm(sub, true="16,16,16", levels=3, ppc=24, vth="0.08,0.004"):
    return small_cfg("warm", grid__nsubdomains=sub, grid__truesize=true, multigrid__mglevels=levels, population__nparticles=f"{ppc} pc",
                     population__nalloc=f"{3 * ppc} pc", population__thermalvelocitycells=vth, grid__nemigrantsalloc=f"{ppc} pc")


def layouts(W):
    out = {}
    for r, st in W.ranks.items():
        W._on(r, lambda st=st, r=r: out.__setitem__(r, W.lib.pincPopLayout(st.pop)))
    return out


def run_against_oracle(L, cfg, steps, expect_slotted=True, per_rank=None, fused="nodeposit"):
    per_rank = initial.maxwellian(cfg, seed=17) if per_rank is None else per_rank
    W, O = sim.World(cfg), orc.OrcWorld(cfg)
    try:
        for X in (W, O):
            X.set_particles(per_rank)
            X.migrate(); X.field_solve(); X.half_kick()
        seen = 0
        for it in range(steps):
            if not fused:
                # the debug scans of the reference's main loop (src/main.c:207, 221) have slotted forms: they must pass and
                # must not cost the layout
                W.run(lambda r, st: (L.pVelAssertMax(st.pop, 1.0), L.pPosAssertInLocalFrame(st.pop, st.rho)))
            W.step(fused=fused); O.step()
            seen += sum(layouts(W).values())
            for r in range(cfg.nRanks):
                for name in ("rho", "phi", "E"):
                    assert rel(W.grid(r, name), O.grid(r, name)) <= 1e-10, (it, r, name)
                st = W.ranks[r]
                nS = cfg.nSpecies
                assert [st.mpi.contents.nEmigrants[i] for i in range(27 * nS)] == list(O.nEmig[r]), (it, r)
                assert [st.mpi.contents.nImmigrants[i] for i in range(27 * nS)] == list(O.nImm[r]), (it, r)
            kw, pw = W.energies(); ko, po = O.energies()
            assert abs(kw - ko) <= 1e-10 * abs(ko) and abs(pw - po) <= 1e-10 * abs(po), it
        if expect_slotted:
            assert seen >= cfg.nRanks * (steps - 2), seen       # slotted from the second step on (the first one has to sort once)
        # the fused pass leaves the positions one puMove ahead of the oracle: bring the oracle there, then compare phase space
        for r in range(cfg.nRanks if fused else 0):
            O.lib.orc_move(orc.dp(O.pos[r]), orc.dp(O.vel[r]), cfg.nSpecies, orc.lp(O.iStart[r]), orc.lp(O.iStop[r]))
        for r in range(cfg.nRanks):
            got, ref = W.particles(r), O.particles(r)           # (reading the particles leaves slotted mode)
            for s in range(cfg.nSpecies):
                assert len(got[s][0]) == len(ref[s][0]), (r, s)
                a, b = sorted_particles(*got[s]), sorted_particles(*ref[s])
                assert np.abs(a - b).max() <= 1e-10 * max(1.0, np.abs(b).max()), (r, s)
        assert sum(layouts(W).values()) == 0
        return {r: {n: W.grid(r, n) for n in ("rho", "phi", "E")} for r in range(cfg.nRanks)}
    finally:
        W.close()


@pytest.mark.parametrize("sub,fused", [("1,1,1", "nodeposit"), ("1,2,2", "nodeposit"), ("1,1,1", False), ("2,1,2", False),
                                       ("1,1,1", True), ("1,2,2", True), ("2,1,2", True)])
def test_slotted_steps_match_oracle(gpu_lib, sub, fused):
    """fused="nodeposit": pincAccMove3D1KE (kick + move + re-binning in one pass); False: the reference's call order, puAcc3D1KE
    and puMove as separate slotted passes; True: pincAccMoveDistr3D1KE (the push also deposits the particles that keep their
    cell, movers and immigrants are deposited where they arrive, puDistr3D1 only converts the accumulators)."""
    L = gpu_lib
    L.pincSetSlotted(1, 25, 16)
    text, cfg = warm(sub)
    before = L.pincSlottedOverflows()
    a = run_against_oracle(L, cfg, 6, fused=fused)
    assert L.pincSlottedOverflows() == before
    L.pincSetSlotted(0, -1, -1)                                 # the counting sort every step: bit-identical fields
    try:
        b = run_against_oracle(L, cfg, 6, expect_slotted=False, fused=fused)
    finally:
        L.pincSetSlotted(1, 25, 16)
    for r in a:
        for n in a[r]:
            assert np.array_equal(a[r][n], b[r][n]), (r, n)


@pytest.mark.parametrize("fused", ["nodeposit", True])
def test_slot_overflow_falls_back_to_the_sort(gpu_lib, fused):
    L = gpu_lib
    L.pincSetSlotted(1, 0, 0)                                   # no head room beyond the fullest cell's count
    try:
        # every particle drifts towards the plane x = L/2 at 0.3 cells per step: the cells there fill up beyond any head room
        text, cfg = warm("1,1,1", ppc=12, vth="0.02,0.001")
        per_rank = initial.maxwellian(cfg, seed=23)
        for r in range(cfg.nRanks):
            for s, (pos, vel) in enumerate(per_rank[r]):
                vel[:, 0] = -0.3 * np.sign(pos[:, 0] - (1 + cfg.trueSize[0] / 2))
        before = L.pincSlottedOverflows()
        run_against_oracle(L, cfg, 6, expect_slotted=False, per_rank=per_rank, fused=fused)
        assert L.pincSlottedOverflows() > before
    finally:
        L.pincSetSlotted(1, 25, 16)
