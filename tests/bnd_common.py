"""Shared pieces of the non-periodic boundary tests (gDirichlet/gNeumann in gBnd, src/grid.c:921-1023;
mgRestrictBnd, src/multigrid.c:1314-1379): the scenario and how each of the three implementations is driven through it.

Scenario (all values seeded): boundaries per case; phi, rho random on every rank; the boundary slices of the finest level
are first set by gSetBndSlices (constants 1/2) and then overwritten on the global-domain edges with smooth non-constant
values (so that the element order of the slices and mgRestrictBnd's every-second-value rule matter); mgRestrictBnd fills
the coarser levels.  Then, recording phi (ghost layers included) after each stage:
  A  gBnd(phi)                       B  mgGS3D(phi, rho, 2 cycles)               C  one mgVRecursive V-cycle.
Whole mgSolve runs are not part of it: with a Dirichlet lower edge the reference overwrites TRUE nodes (slice 1,
src/grid.c:940), the residual there never vanishes and its tolerance loop (src/multigrid.c:1697) does not terminate."""
import numpy as np

CASES = [("1,1,1", "DIRICHLET,PERIODIC,NEUMANN,DIRICHLET,PERIODIC,NEUMANN"),
         ("1,2,2", "NEUMANN,DIRICHLET,PERIODIC,DIRICHLET,NEUMANN,PERIODIC"),
         ("2,1,2", "DIRICHLET,NEUMANN,DIRICHLET,NEUMANN,DIRICHLET,NEUMANN")]
TRUE = "8,8,16"
LEVELS = 3


def overrides(sub, boundaries):
    return dict(grid__nsubdomains=sub, grid__truesize=TRUE, multigrid__mglevels=LEVELS, grid__boundaries=boundaries,
                population__nparticles="1 pc", population__nalloc="2 pc", grid__nemigrantsalloc="1 pc")


def fields(cfg, seed=31):
    """(phi, rho) per rank, flat ghost-inclusive arrays."""
    rng = np.random.default_rng(seed)
    n = int(np.prod([t + 2 for t in cfg.trueSize]))
    return [(rng.standard_normal(n), rng.standard_normal(n)) for _ in range(cfg.nRanks)]


def slice_values(cfg, r, nmax):
    """Non-constant boundary values for the 8 slices of rank r (slices 0 and 4 are unused)."""
    s = np.arange(nmax, dtype=np.float64)
    return np.concatenate([(0.3 + 0.1 * b) * np.cos(0.37 * s + b + 0.5 * r) + 0.05 * b for b in range(8)])
