"""Shared scenario of the tests of the solver's other select() targets: the cycles mgVRegular and mgW
(src/multigrid.c:1559-1683) and the Jacobi smoother mgJacob3D (:500-551, ini name "jacobian").  Periodic grids, random rho
and phi on every level's finest arrays; ONE cycle per case (whether the reference's tolerance loop terminates with these
variants is not something its own tests establish: mgVRegular subtracts the coarse correction, undamped Jacobi does not
smooth the checkerboard mode), recording phi and the residual grid of the finest level afterwards."""
import numpy as np

# (nSubdomains, cycle, (pre, post, coarse) smoothers: 0 gaussSeidelRB, 1 jacobian)
CASES = [("1,1,1", "mgVRegular", (0, 0, 0)), ("1,2,2", "mgVRegular", (0, 0, 0)), ("1,1,1", "mgW", (0, 0, 0)), ("2,1,2", "mgW", (0, 0, 0))]
# The Jacobi cases have no reference to compare with: mgJacob3D (src/multigrid.c:500-551) never advances its write index
# (`tempVal[g]` with g fixed), starts its neighbour indices at +-sizeProd[d] instead of g +- sizeProd[d] (it reads phiVal[-1])
# and then copies the uninitialised scratch array over phi - undefined behaviour.  What the library and the oracle provide
# under that name is the iteration its comments describe (every node from the old values of its six neighbours, then
# gHaloOp and gBnd); PARITY UNPINNED for this one function, device and oracle are only compared with each other and with a
# closed form.
JACOBI_CASES = [("1,1,1", "mgVRecursive", (1, 1, 1)), ("1,2,2", "mgVRecursive", (1, 0, 1)), ("1,1,1", "smoother", (1, 1, 1))]
TRUE = "16,8,8"
LEVELS = 3


def overrides(sub):
    return dict(grid__nsubdomains=sub, grid__truesize=TRUE, multigrid__mglevels=LEVELS, multigrid__npresmooth=3,
                multigrid__npostsmooth=2, multigrid__ncoarsesolve=4,
                population__nparticles="1 pc", population__nalloc="2 pc", grid__nemigrantsalloc="1 pc")


def fields(cfg, seed=41):
    rng = np.random.default_rng(seed)
    n = int(np.prod([t + 2 for t in cfg.trueSize]))
    return [(rng.standard_normal(n), rng.standard_normal(n)) for _ in range(cfg.nRanks)]
