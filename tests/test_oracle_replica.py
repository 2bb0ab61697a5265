"""What the replicated multi-rank solve relies on, checked on the CPU oracle: the distributed
multigrid solve of PINC over R sub-domains (exchanging ghost layers, all-reducing gBnd's means and the residual norm) is
the same function as ONE solve of the global periodic grid; results differ only by the order of those sums.  The GPU
tests check the CUDA path against the distributed oracle; this pins the equivalence itself."""
import ctypes as C

import numpy as np
import pytest

from oracle import orc


def _solve(O, nsub, local, levels, rho_global, cycles_cap=100):
    """Run the oracle's solve over nsub sub-domains of `local` true nodes; returns (V-cycles, history, global phi)."""
    t = orc.make_topo(nsub, local)
    R = int(np.prod(nsub))
    size = [n + 2 for n in local]
    npts = int(np.prod(size))
    rhos, phis, ress = [], [], []
    for r in range(R):
        sub = [r % nsub[0], (r // nsub[0]) % nsub[1], r // (nsub[0] * nsub[1])]
        a = np.zeros(size[::-1])                                    # [z][y][x], ghost-inclusive
        z0, y0, x0 = sub[2] * local[2], sub[1] * local[1], sub[0] * local[0]
        a[1:-1, 1:-1, 1:-1] = rho_global[z0:z0 + local[2], y0:y0 + local[1], x0:x0 + local[0]]
        rhos.append(a.reshape(-1).copy()); phis.append(np.zeros(npts)); ress.append(np.zeros(npts))
    mg = O.orc_mg_alloc(C.byref(t), levels, 10, 10, 10)
    hist = np.zeros(128)
    n = O.orc_mg_solve(mg, orc.ptr_array(rhos), orc.ptr_array(phis), orc.ptr_array(ress), 1e-10, cycles_cap, orc.dp(hist), 128)
    O.orc_mg_free(mg)
    G = np.zeros([local[2] * nsub[2], local[1] * nsub[1], local[0] * nsub[0]])
    for r in range(R):
        sub = [r % nsub[0], (r // nsub[0]) % nsub[1], r // (nsub[0] * nsub[1])]
        z0, y0, x0 = sub[2] * local[2], sub[1] * local[1], sub[0] * local[0]
        G[z0:z0 + local[2], y0:y0 + local[1], x0:x0 + local[0]] = phis[r].reshape(size[::-1])[1:-1, 1:-1, 1:-1]
    return n, hist[:n].copy(), G


@pytest.mark.parametrize("nsub,local,levels", [((1, 1, 2), (8, 8, 8), 2), ((1, 2, 2), (16, 8, 8), 3), ((2, 2, 2), (8, 8, 8), 2)])
def test_distributed_solve_equals_global_solve(nsub, local, levels):
    O = orc.load()
    glob = [local[d] * nsub[d] for d in range(3)]
    rho = np.random.default_rng(5).standard_normal(glob[::-1])
    n_d, h_d, phi_d = _solve(O, nsub, local, levels, rho)
    n_g, h_g, phi_g = _solve(O, (1, 1, 1), glob, levels, rho)
    assert n_d == n_g and n_d > 1
    floor = 1e-13 * max(1.0, np.abs(phi_g).max())
    assert np.all(np.abs(h_d - h_g) <= 1e-6 * h_g + floor)
    assert np.abs(phi_d - phi_g).max() <= 1e-12 * np.abs(phi_g).max()
