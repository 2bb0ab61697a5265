#!/usr/bin/env python
"""Stand-alone multigrid benchmark (BASELINE config 3, input/mgErrorScaling.ini repaired as in SURVEY 8d-C3):
rho = k^2 sin(k x) with the reference's PI (src/grid.h:584, quirk Q11) or white noise, N^3 cells, mgLevels =
log2(N) - 1, V(10,10) with 10 coarse sweeps.  Prints V-cycles, barRes history end, device time per solve and
per V-cycle for every execution mode of the solver, and the error against the analytic solution."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import GridH, single_mpi  # noqa: E402
from pinc_b200 import lib as plib  # noqa: E402

PI = 3.14159265
MODES = {0: "ops", 1: "fused-exact", 2: "auto", 3: "auto-exact", 4: "cluster-always", 5: "allsm"}


PROF_NAMES = ["neut_rho", "gs_big", "gs_small", "restrict", "prolong", "neut_phi", "norm", "x_halo", "small_section", "b_load", "b_wait", "r_sync", "b_tail", "r_recv", "r_comp", "r_send",
              "gs16", "res16", "pro16", "-", "gs8", "res8", "pro8", "-", "gs4", "res4", "pro4", "-", "f_neut", "f_resid", "f_restr", "f_prol"]


def run_case(L, spec, kind="sin", mode=2, reps=3):
    """One grid, one solver mode: `reps` cold-started solves (phi = 0 on every level); returns the record main() prints.
    kind "sin": rho = k^2 sin(k x) as gFillSin (src/grid.c:1563-1603, norm = 0) with the reference's PI, analytic solution
    sin(k x) (gFillSinSol :1610); "noise": white noise."""
    true = tuple(int(v) for v in spec.split("x")) if "x" in str(spec) else (int(spec),) * 3
    N = true[0]
    levels = int(np.log2(min(true))) - 1
    L.pincMgSetMode(mode)
    rho, phi = GridH(L, true, 1), GridH(L, true, 1)
    z, y, x = np.meshgrid(*[np.arange(s, dtype=float) for s in rho.size[::-1]], indexing="ij")
    k = 2 * PI / N
    if kind == "sin":
        rho.a[..., 0] = k * k * np.sin(k * (x - 1))
        sol = np.sin(k * (x - 1))
    else:
        rho.a[..., 0] = 8 * np.random.default_rng(0).standard_normal(rho.a.shape[:3])
        sol = None
    solver = L.pincMgAllocSolver(rho.ptr, phi.ptr, levels, 1, 10, 10, 10)
    m = single_mpi(L, true)
    times = []
    ncyc = None
    for rep in range(reps):
        phi.a[...] = 0
        phi.up(); rho.up()
        for q in range(1, levels):           # cold start of the coarse levels as well
            for mg in (solver.contents.mgPhi,):
                g = mg.contents.grids[q]
                np.ctypeslib.as_array(g.contents.val, shape=(int(g.contents.sizeProd[4]),))[:] = 0
                L.pincSyncGridToDevice(g)
        L.pincDeviceSynchronize()
        L.pincTimerStart()
        L.mgSolve(solver, rho.ptr, phi.ptr, m)
        times.append(L.pincTimerStopMs())
        buf = (C.c_double * 256)()
        ncyc = L.pincMgLastHistory(buf, 256)
        last = buf[min(ncyc, 250) - 1]
    err = emax = None
    if sol is not None:
        p = phi.down()[1:-1, 1:-1, 1:-1, 0]
        err = float(np.sqrt(np.mean((p - sol[1:-1, 1:-1, 1:-1]) ** 2)))
        emax = float(np.abs(p - sol[1:-1, 1:-1, 1:-1]).max())
    rec = {"N": str(spec), "levels": levels, "mode": MODES[mode], "rho": kind, "vcycles": ncyc, "barRes_last": last,
           "ms_per_solve": min(times), "us_per_vcycle": 1e3 * min(times) / max(ncyc, 1), "rms_error_vs_analytic": err,
           "max_error_vs_analytic": emax, "path": int(L.pincMgLastPath())}
    if os.environ.get("PINC_B200_MGPROF"):
        buf = (C.c_longlong * 64)()
        L.pincMgProfRead.argtypes = [C.POINTER(C.c_longlong)]
        if L.pincMgProfRead(buf):
            rec["prof_us_per_vcycle"] = {n: round(buf[2 * i] / 1965.0 / (reps * max(ncyc, 1)), 2) for i, n in enumerate(PROF_NAMES) if n != "-"}
            rec["prof_calls_per_vcycle"] = {n: round(buf[2 * i + 1] / (reps * max(ncyc, 1)), 1) for i, n in enumerate(PROF_NAMES) if n != "-"}
    L.mgFreeSolver(solver)
    rho.free(); phi.free()
    L.pincMgSetMode(2)
    return rec


def error_scaling(L, sizes=(16, 32, 64, 128), mode=2):
    """BASELINE configs[2] (mgModeErrorScaling, src/multigrid.c:1734-1851): the sine problem on N^3 cells for a ladder of N.
    Returns the records plus the observed order of the error against the analytic solution, log2(e(N)/e(2N)) - 2 for the
    second-order 7-point Laplacian (discrete eigenvalue (2 - 2 cos k) against k^2: e = k^2/12 + O(k^4))."""
    recs = [run_case(L, n, "sin", mode) for n in sizes]
    orders = []
    for a, b in zip(recs, recs[1:]):
        if int(b["N"]) == 2 * int(a["N"]):
            orders.append(float(np.log2(a["rms_error_vs_analytic"] / b["rms_error_vs_analytic"])))
    return recs, orders


def main():
    L = plib.load()
    # sizes: N (cube) or AxBxC
    sizes = [x for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "16,32,64".split(","))]
    kind = sys.argv[2] if len(sys.argv) > 2 else "sin"
    modes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 2, 3]
    for spec in sizes:
        true = tuple(int(v) for v in spec.split("x")) if "x" in spec else (int(spec),) * 3
        for mode in modes:
            if mode == 0 and max(true) > 64:
                continue
            print(json.dumps(run_case(L, spec, kind, mode)), flush=True)


if __name__ == "__main__":
    main()
