#!/usr/bin/env python
"""bench.py — particle-steps/s of the PIC time step (push + deposit + multigrid solve + migration) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload warm|warm_big|<ini>]

Workload at N=1: BASELINE.json configs[1] — warm Maxwellian electron-ion plasma on ONE sub-domain of 64^3 cells,
70 particles per cell and species (36.7 M particles), configs/warm.ini.  For N>1 every GPU keeps that same
sub-domain (weak scaling): nSubdomains = 1,1,2 / 1,2,2 / 2,2,2, one process per GPU (torchrun), halos and
migrants over NCCL.  A step = one pass of src/main.c:197-274 (canonical order) over all particles.

value   device-resident throughput: fused particle pass (pincAccMoveDistr3D1KE: kick, move, re-binning and the deposition of
        the particles that keep their cell in one pass over the cell slots) + persistent multigrid kernel, CUDA-event timed
        on the library's stream, max over ranks.
e2e     the same K steps as a JOB that starts and ends in HOST buffers, through the PINC-named entry points in
        the reference's call order (puMove, puExtractEmigrants3D, puMigrate, puDistr3D1, gHaloOp, mgSolve,
        gFinDiff1st, puAcc3D1KE, ...): the population is copied host->device from page-locked memory at the
        start and device->host at the end INSIDE the timed region, and every step reads its results
        (energies; rho, phi, E as the reference writes them per step) back to the host.
--impl reference   the reference's own C sources (oracle/_ref, built from /root/reference under an MPI shim)
        on the host cores, one thread-rank per sub-domain, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "particle-steps/s (push+deposit+MG solve)"
UNIT = "particle-steps/s"
SUBDOMAINS = {1: "1,1,1", 2: "1,1,2", 4: "1,2,2", 8: "2,2,2"}


def load_cfg(workload, n_gpus, particles_scale=1.0):
    from pinc_b200 import config
    path = workload if os.path.exists(workload) else os.path.join(ROOT, "configs", workload + ".ini")
    ini = config.Ini(open(path).read())
    if workload == "warm" or n_gpus > 1:
        ini.d["grid:nsubdomains"] = SUBDOMAINS[n_gpus]
    if particles_scale != 1.0:
        for k in ("population:nparticles", "population:nalloc"):
            v = config.atof(ini.d[k])
            ini.d[k] = f"{v * particles_scale:g} pc"
    text = ini.dump()
    return text, config.load_config(config.Ini(text))


class ClockSampler:
    """nvidia-smi polled every 100 ms in the background.  It is started ahead of the warm-up steps (its own start-up takes longer
    than a short timed region) and every line is time-stamped as it arrives; the summary uses the samples that fall between
    mark_start() and mark_stop(), i.e. inside the timed region, and says so when it had to widen to the warm-up steps."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.p = None
        self.rows = []
        self.t0 = self.t1 = None
        self.thread = None

    def _pump(self):
        for line in self.p.stdout:
            self.rows.append((time.perf_counter(), line))

    def start(self):
        import threading
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.p = None

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_stop(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.t1 is None:
            self.mark_stop()
        if self.t0 is None:
            self.t0 = float("-inf")
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        if self.thread is not None:
            self.thread.join(timeout=5)

        def digest(rows):
            sm, mx, reasons = [], [], set()
            for _, line in rows:
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons
        inside = [r for r in self.rows if self.t0 <= r[0] <= self.t1 + 0.1]
        sm, mx, reasons = digest(inside)
        window = "timed region"
        if not sm:
            # a timed region shorter than one polling period: the samples of the warm-up steps right before it (same load)
            sm, mx, reasons = digest([r for r in self.rows if r[0] <= self.t1 + 0.1][-5:])
            window = "warm-up steps + timed region (the timed region is shorter than one 100 ms polling period)"
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------------------
def run_reference(args, n_gpus):
    """The reference's CPU implementation of the same step on the host cores (bounded sample)."""
    from pinc_b200 import initial
    from oracle import orc, ref
    cores = host_cores()
    n_ranks = 8 if cores >= 8 else (4 if cores >= 4 else (2 if cores >= 2 else 1))
    # same global problem as the N=1 GPU workload (64^3 cells), decomposed over the host cores; the sample is
    # bounded by the particle count (ppc scaled down), not by changing the grid
    scale = args.cpu_sample
    from pinc_b200 import config
    ini = config.Ini(open(os.path.join(ROOT, "configs", "warm.ini")).read())
    sub = SUBDOMAINS[n_ranks]
    ini.d["grid:nsubdomains"] = sub
    ini.d["grid:truesize"] = ",".join(str(64 // int(s)) for s in sub.split(","))
    ini.d["population:nparticles"] = f"{70 * scale:g} pc"
    ini.d["population:nalloc"] = f"{128 * scale:g} pc"
    text = ini.dump()
    cfg = config.load_config(config.Ini(text))
    kind = "reference" if ref.available() else "port"
    per_rank = initial.maxwellian(cfg, seed=20261018)
    W = ref.RefWorld(text, cfg.nRanks) if kind == "reference" else orc.OrcWorld(cfg)
    W.set_particles(per_rank)
    W.migrate(); W.field_solve(); W.half_kick()
    n_part = sum(len(p) for r in per_rank for (p, _) in r)
    for _ in range(args.warmup):
        W.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        W.step()
    dt = time.perf_counter() - t0
    value = n_part * args.steps / dt
    used = cfg.nRanks if kind == "reference" else 1
    sample = (f"global 64^3 cells as {sub} sub-domains of {ini.d['grid:truesize']}, {70 * scale:g} particles/cell/species "
              f"({n_part} particles), {args.steps} steps after {args.warmup} warm-up; throughput per particle-step is "
              f"what is compared (the GPU workload has 70/cell/species)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "warm (BASELINE configs[1]): 64^3 cells, Maxwellian electrons+ions, CPU sample", "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    return line


# ------------------------------------------------------------------------------------------------------------
def run_mg_error_scaling(args, n_gpus):
    """BASELINE configs[2] (input/mgErrorScaling.ini, src/multigrid.c:1734-1851): the stand-alone multigrid solve of the sine
    problem on a ladder of grids; V-cycles, time per V-cycle and the error against the analytic solution.  One GPU (for
    N > 1 every rank would run a replica: the path does not shard, so only rank 0 reports)."""
    import torch
    from pinc_b200 import lib as plib
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import mg_bench
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    L = plib.load()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    for _ in range(max(1, args.warmup // 2)):
        mg_bench.run_case(L, 64, "sin", 2, reps=1)
    sampler.mark_start()
    launches0 = L.pincLaunchCount()
    recs, orders = mg_bench.error_scaling(L, (16, 32, 64, 128))
    launches = L.pincLaunchCount() - launches0
    sampler.mark_stop()
    clocks = sampler.stop()
    big = recs[-2]                       # 64^3: the grid of BASELINE configs[1]
    n_fine = 64 ** 3
    value = n_fine * big["vcycles"] / (big["ms_per_solve"] * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ach = (606 + 32) * value / 1e9
    return {"metric": "fine-grid nodes x V-cycles per second (mgSolve, sine problem, 64^3)", "value": value, "unit": "node-V-cycles/s",
            "n_gpus": n_gpus, "steps": 3, "warmup": args.warmup, "ms_per_step": big["ms_per_solve"], "higher_is_better": True,
            "scaling": "replicas only", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "mgErrorScaling (BASELINE configs[2]): rho = k^2 sin(kx) with the reference's PI (src/grid.h:584), N^3 cells for N = 16..128, "
                                   "mgLevels = log2(N) - 1, V(10,10), 10 coarse sweeps, tolerance 1e-10; a step = one cold-started mgSolve (best of 3)",
                       "l2": "all levels are L2/shared-memory resident by design (2.3 MB per array at 64^3); no flush"},
            "cases": [{k: r[k] for k in ("N", "levels", "vcycles", "ms_per_solve", "us_per_vcycle", "rms_error_vs_analytic", "max_error_vs_analytic", "barRes_last", "path")} for r in recs],
            "error_order_observed": orders, "error_order_expected": 2.0,
            "clocks": clocks, "gpu_launches": int(launches),
            "roofline": {"kernel": "k_mg_solve", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                         "note": "latency-bound (DESIGN.md section 4): 638 algorithmic B per fine node and V-cycle; us_per_vcycle is the figure of merit"},
            "e2e": None}


def run_ours(args, n_gpus, rank, world_size):
    import torch
    from pinc_b200 import abi, initial, lib as plib, sim
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    nccl_id = None
    parity = None
    if world_size > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        L = plib.load()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import multi_check
        if not args.no_parity:
            # the NCCL path against the oracle BEFORE anything is timed (small seeded configs, every rank checks its own
            # sub-domain against the oracle's run of the whole world): warm + cold, replicated + distributed solve
            parity = {"tolerance": 1e-10, "cases": []}
            # (kind, replicated?, hybrid?, cells per rank): the hybrid solve needs a finest level of > 4096 nodes per rank
            for kind, replica, hybrid, true in (("warm", 1, 1, "32,16,16"), ("cold", 1, 1, "32,16,16"), ("warm", 1, 0, "16,8,8"), ("warm", 0, 0, "16,8,8")):
                if True:
                    r = multi_check.check(rank, world_size, kind, steps=4, replica=replica, hybrid=hybrid, true=true)
                    t = torch.tensor([r["worst_field_err"], r["worst_particle_err"], 0.0 if r["tables_exact"] else 1.0,
                                      0.0 if r["cycles_equal"] else 1.0, 0.0 if r["sizes_exact"] else 1.0],
                                     dtype=torch.float64, device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    v = t.tolist()
                    parity["cases"].append({"config": f"{kind}, {true.replace(',', 'x')} cells per rank, {multi_check.SUB[world_size]}, 4 steps",
                                            "solve": ("hybrid (finest level distributed over peer-memory mailboxes, coarser levels replicated)" if hybrid else "replicated") if replica else "distributed (one kernel per reference call, peer-memory smoother)",
                                            "mg_path": r["mg_path"], "transport": r["transport"], "vcycles_last": r["vcycles_last"],
                                            "worst_field_err": v[0], "worst_particle_err": v[1], "tables_exact": v[2] == 0.0,
                                            "cycles_equal": v[3] == 0.0, "sizes_exact": v[4] == 0.0})
            parity["ok"] = all(c["worst_field_err"] <= 1e-10 and c["worst_particle_err"] <= 1e-10 and c["tables_exact"]
                               and c["cycles_equal"] and c["sizes_exact"] for c in parity["cases"])
            parity["checker"] = "oracle/pinc_oracle.c (every rank runs the whole world on the CPU and compares its sub-domain)"
        nccl_id = multi_check.fresh_nccl_id(L, rank)
    FUSED = True if args.particle_pass == "full" else "nodeposit"
    text, cfg = load_cfg(args.workload, n_gpus, args.particles_scale)
    assert cfg.nRanks == world_size, (cfg.nSubdomains, world_size)
    os.environ["PINC_B200_DEVICE"] = str(local_rank)
    W = sim.World(cfg) if world_size == 1 else sim.World(cfg, rank=rank, world_size=world_size, nccl_id=nccl_id)
    L = W.lib
    st = W.ranks[rank]
    t_ic = time.perf_counter()
    mine = initial.maxwellian(cfg, seed=20261018, ranks=[rank])[0]
    per_rank = {rank: mine}
    t_ic = time.perf_counter() - t_ic
    # page-lock the host arrays the job copies from/to
    p = st.pop.contents
    nbytes_pop = 3 * 8 * int(p.iStart[p.nSpecies])
    L.pincHostRegister(C.cast(p.pos, C.c_void_p), nbytes_pop)
    L.pincHostRegister(C.cast(p.vel, C.c_void_p), nbytes_pop)
    for g in (st.rho, st.phi, st.E):
        L.pincHostRegister(C.cast(g.contents.val, C.c_void_p), 8 * int(g.contents.sizeProd[4]))

    def barrier():
        L.pincDeviceSynchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---------------- e2e job: host buffers -> K reference-order steps -> host buffers -----------------------
    W.set_particles(per_rank)            # fills the host arrays (and uploads once so the set-up below can run)
    W.migrate(); W.field_solve(); W.half_kick()
    for _ in range(max(1, args.warmup // 2)):
        W.step(fused=False)
    L.pincSyncPopToHost(st.pop)          # the job's input state now lives in the host arrays
    n_live = sum(p.iStop[s] - p.iStart[s] for s in range(p.nSpecies))
    grid_bytes = sum(8 * int(g.contents.sizeProd[4]) for g in (st.rho, st.phi, st.E))
    e2e_steps = max(2, args.steps if args.e2e_steps <= 0 else min(args.steps, args.e2e_steps))
    barrier()
    t0 = time.perf_counter()
    L.pincSyncPopToDevice(st.pop)                               # H2D: 48 B per live particle
    t_h2d = time.perf_counter() - t0
    for _ in range(e2e_steps):
        W.step(fused=False)
        for g in (st.rho, st.phi, st.E):                        # D2H of the step's field results
            L.pincSyncGridToHost(g)
        W.energies()
    t1 = time.perf_counter()
    L.pincSyncPopToHost(st.pop)                                 # D2H: 48 B per live particle
    t_d2h = time.perf_counter() - t1
    barrier()
    t_e2e = allmax(time.perf_counter() - t0)
    n_global_e2e = allsum(float(n_live))
    e2e = {"value": n_global_e2e * e2e_steps / t_e2e, "unit": UNIT, "steps": e2e_steps,
           "h2d_bytes_per_step": int(48 * n_live / e2e_steps),
           "d2h_bytes_per_step": int(48 * n_live / e2e_steps + grid_bytes + 8 * (p.nSpecies + 2)),
           "ms_per_step": 1e3 * t_e2e / e2e_steps,
           # the two population copies happen once per job, not once per step: their cost per call (rank 0) and what remains per step
           "population_h2d_ms": 1e3 * t_h2d, "population_d2h_ms": 1e3 * t_d2h,
           "ms_per_step_without_population_copies": 1e3 * (t_e2e - t_h2d - t_d2h) / e2e_steps,
           "path": "PINC entry points in reference order; population H2D at start and D2H at end inside the timed region (page-locked), fields+energies D2H every step"}

    # ---------------- device-resident throughput (fused particle pass) ------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()                      # ahead of the warm-up: nvidia-smi's own start-up is longer than a short timed region
    for _ in range(args.warmup):
        W.step(fused=FUSED)
    barrier()
    launches0 = L.pincLaunchCount()
    sampler.mark_start()
    L.pincTimerStart()
    cycles = []
    for _ in range(args.steps):
        W.step(fused=FUSED)
    ms = L.pincTimerStopMs()
    barrier()
    sampler.mark_stop()
    clocks = sampler.stop()
    launches = L.pincLaunchCount() - launches0
    ms = allmax(ms)
    n_live = sum(p.iStop[s] - p.iStart[s] for s in range(p.nSpecies))
    n_global = allsum(float(n_live))
    value = n_global * args.steps / (ms * 1e-3)
    hist = W.history()

    # ---------------- per-kernel-class device time (CUDA events around every launch), 3 extra steps ------------
    L.pincProfReset(); L.pincProfEnable(1)
    prof_steps = 3
    for _ in range(prof_steps):
        W.step(fused=FUSED)
    prof = plib.profile(L)
    L.pincProfEnable(0)
    hist = W.history() or hist                      # V-cycles of the last profiled solve
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    kernels = {}
    tot_ms = sum(v[0] for v in prof.values()) or 1.0
    if "mgfused" in prof and len(hist):
        # the library books the algorithmic bytes of ONE V-cycle per launch (it cannot know how many the tolerance loop
        # will take); one launch runs len(hist) of them (606 B x N_fine per V-cycle + 32 B/pt norm check)
        kms, cnt, by = prof["mgfused"]
        n_fine = cfg.trueSize[0] * cfg.trueSize[1] * cfg.trueSize[2]
        prof["mgfused"] = (kms, cnt, (by + 32.0 * n_fine * cnt) * len(hist))
    for k, (kms, cnt, by) in prof.items():
        kernels[k] = {"ms_per_step": kms / prof_steps, "launches_per_step": cnt / prof_steps,
                      "alg_GBps": (by / (kms * 1e-3) / 1e9) if kms > 0 else None, "share": kms / tot_ms}
    dom = max(prof.items(), key=lambda kv: kv[1][0])[0] if prof else None
    # the HBM-bound kernels of the path (what the north star's ">= 60 % of HBM peak on push and deposit" is about):
    # algorithmic GB/s of each class against the measured copy bandwidth
    hbm_kernels = {k: {"achieved": v["alg_GBps"], "frac": v["alg_GBps"] / peak, "ms_per_step": v["ms_per_step"]}
                   for k, v in kernels.items() if k in ("push", "deposit", "move", "sort", "findiff") and v["alg_GBps"]}
    # the whole particle phase against the HBM roofline: SURVEY 8d's 120 algorithmic B per particle-step (push 96 + deposit 24)
    # over everything the phase launches (push, move, deposit, re-binning/sort, emigrant extraction, import)
    pp_ms = sum(kernels[k]["ms_per_step"] for k in ("push", "move", "deposit", "sort", "extract", "import") if k in kernels)
    particle_phase = None
    if pp_ms > 0:
        ach_pp = 120.0 * n_live / (pp_ms * 1e-3) / 1e9
        particle_phase = {"ms_per_step": pp_ms, "alg_bytes_per_particle": 120, "achieved": ach_pp, "peak": peak, "unit": "GB/s", "frac": ach_pp / peak,
                          "layout": "cell-slotted (only particles that change cell move)" if L.pincPopLayout(st.pop) else "contiguous, counting sort every step"}
    roofline = None
    if dom:
        kms, cnt, by = prof[dom]
        ach = by / (kms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
        except Exception:
            pass
        roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "traffic_source": "static: dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture (profiles/traffic.json), not measured in this run",
                    "peak_source": peak_src,
                    "alg_bytes_per_launch": by / cnt, "avg_launch_ms": kms / cnt, "hbm_bound_kernels": hbm_kernels,
                    "particle_phase": particle_phase}
        if dom == "mgfused":
            # the multigrid kernel is bound by the latency of its dependent half-sweeps, not by bytes: say so
            phases = sum(2 * (cfg.nCoarseSolve if q == cfg.mgLevels - 1 else cfg.nPreSmooth + cfg.nPostSmooth)
                         for q in range(cfg.mgLevels))
            roofline["note"] = ("latency-bound: one launch = the whole tolerance loop of the reference's V(10,10) cycle "
                                "(alg_bytes_per_launch = V-cycles of the launch x (606 + 32) B x N_fine); the grids are L2/shared-memory "
                                "resident, so frac against HBM is not the figure of merit, us per dependent half-sweep is")
            roofline["vcycles_per_launch"] = len(hist)
            roofline["dependent_half_sweeps_per_vcycle"] = phases
            roofline["us_per_half_sweep"] = (1e3 * kms / cnt) / max(1, len(hist) * phases)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: per GPU {cfg.trueSize} cells, {cfg.nSpecies} species, "
                                   f"{n_live} particles (Maxwellian, sigma_v={cfg.thermalVelocity} cells/step), nSubdomains={cfg.nSubdomains}",
                       "global_particles": int(n_global), "mgLevels": cfg.mgLevels, "parallelism": f"domain-decomposition x{world_size}",
                       "l2": "particle arrays (48 B x particles per GPU) exceed the 126 MB L2; no explicit flush",
                       "vcycles_last_solve": len(hist), "ic_seconds": t_ic,
                       "particle_pass": "pincAccMoveDistr3D1KE (kick + move + re-binning + deposition of the stayers)" if FUSED is True else "pincAccMove3D1KE + puDistr3D1",
                       "mg_path": {0: "distributed, one kernel per reference call", 1: "all-SM persistent kernel", 2: "cluster kernel",
                                   5: "replicated: global problem on every rank, all-SM persistent kernel",
                                   6: "replicated: global problem on every rank, cluster kernel",
                                   9: "hybrid: finest level distributed (block faces across sub-domains through peer-memory mailboxes), coarser levels replicated, one persistent kernel per rank"}.get(W.mg_path(), "?")},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "kernels": kernels,
            "kernels_source": f"{prof_steps} extra steps AFTER the timed region with CUDA events around every launch (the timed steps carry no per-launch events); "
                              "avg_launch_ms and ms_per_step therefore come from different steps of the same run"}
    if parity is not None:
        line["parity"] = parity
    W.close()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="warm")
    ap.add_argument("--particles-scale", type=float, default=1.0)
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer job (0: the same K as --steps)")
    ap.add_argument("--cpu-sample", type=float, default=1.0, help="fraction of the 70 particles/cell used by the CPU arms")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the NCCL path that N > 1 runs ahead of the timed region")
    ap.add_argument("--particle-pass", default="full", choices=["full", "nodeposit"],
                    help="full (default, measured faster: 1.35 against 1.47 ms per step at 36.7 M particles): kick + move + re-binning + the "
                         "deposition of the particles that keep their cell in one pass (pincAccMoveDistr3D1KE); nodeposit: the deposition is "
                         "puDistr3D1's own pass over the slots (pincAccMove3D1KE)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(run_reference(args, args.gpus)), flush=True)
        return
    assert world == args.gpus, f"--gpus {args.gpus} needs {args.gpus} ranks (torchrun); WORLD_SIZE={world}"
    if args.workload == "mgErrorScaling":
        if rank == 0:
            print(json.dumps(run_mg_error_scaling(args, args.gpus)), flush=True)
        return
    if args.warmup < 3:
        print(f"bench.py: --warmup {args.warmup} raised to 3 (timing rules)", file=sys.stderr)
        args.warmup = 3
    line = run_ours(args, args.gpus, rank, world)
    if rank == 0:
        if args.gpus == 1 and not args.no_cpu_baseline:
            a = argparse.Namespace(**vars(args))
            a.steps, a.warmup = 2, 1
            ref_line = run_reference(a, 1)
            line["cpu_baseline"] = ref_line["cpu_baseline"]
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
