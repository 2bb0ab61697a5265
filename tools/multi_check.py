#!/usr/bin/env python
"""Multi-GPU parity check: one process per GPU (torchrun), NCCL transport, against the oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_check.py [cold|warm] [steps] [replica 1|0] [hybrid 0|1]

Every rank steps its own sub-domain on its GPU; every rank also steps the WHOLE world in the oracle on the CPU
(small config) and compares its own sub-domain: population sizes and migrant tables exact, V-cycle counts equal,
fields and particle phase space to 1e-10 relative.  `check()` is also what `bench.py --gpus N` runs ahead of its
timed region to put a `parity` object into the JSON line."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SUB = {1: "1,1,1", 2: "1,1,2", 4: "1,2,2", 8: "2,2,2"}


def fresh_nccl_id(L, rank):
    """A new ncclUniqueId made by rank 0 and broadcast over torch.distributed (one id per communicator)."""
    import torch
    import torch.distributed as dist
    buf = C.create_string_buffer(128)
    if rank == 0:
        L.pincNcclUniqueId(buf)
    t = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    return bytes(t.cpu().tolist())


def check(rank, world, kind="warm", steps=4, replica=1, tol=1e-10, hybrid=0, true="16,8,8"):
    """Returns {'worst_field_err', 'worst_particle_err', 'tables_exact', 'cycles_equal', 'sizes_exact', 'mg_path', ...}
    for THIS rank; raises nothing (the caller decides).  Collective: every rank must call it."""
    from helpers import small_cfg, sorted_particles
    from oracle import orc
    from pinc_b200 import initial, lib as plib, sim
    L = plib.load()
    L.pincMgSetReplica(int(replica))
    L.pincMgSetHybrid(int(hybrid))
    over = dict(grid__nsubdomains=SUB[world], grid__truesize=true, multigrid__mglevels=3, population__nparticles="8 pc",
                population__nalloc="24 pc", grid__nemigrantsalloc="4 pc")
    if kind == "cold":
        over["population__perturbamplitude"] = "2e-3,0,0,0,0,0"
        text, cfg = small_cfg("cold", **over)
        per_rank = initial.perturb(cfg, initial.lattice(cfg))
    else:
        over["population__thermalvelocitycells"] = "0.08,0.004"
        text, cfg = small_cfg("warm_big", **over)
        per_rank = initial.maxwellian(cfg, seed=7)
    W = sim.World(cfg, rank=rank, world_size=world, nccl_id=fresh_nccl_id(L, rank))
    O = orc.OrcWorld(cfg)
    W.set_particles({rank: per_rank[rank]})
    O.set_particles(per_rank)
    for X in (W, O):
        X.migrate(); X.field_solve(); X.half_kick()
    out = dict(kind=kind, replica=int(replica), steps=steps, worst_field_err=0.0, worst_particle_err=0.0, tables_exact=True,
               cycles_equal=True, sizes_exact=True, transport=L.pincTransportName().decode())
    for it in range(steps):
        W.step(fused=(it % 2 == 1)); O.step()
        out["cycles_equal"] &= len(W.history()) == len(O.history())
        if it % 2 == 1:
            continue                      # after a fused step the positions are one puMove ahead
        for name in ("rho", "phi", "E"):
            a, b = W.grid(rank, name), O.grid(rank, name)
            out["worst_field_err"] = max(out["worst_field_err"], float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)))
        got, ref = W.particles(rank), O.particles(rank)
        for s in range(cfg.nSpecies):
            if len(got[s][0]) != len(ref[s][0]):
                out["sizes_exact"] = False
                continue
            a, b = sorted_particles(*got[s]), sorted_particles(*ref[s])
            if len(a):
                out["worst_particle_err"] = max(out["worst_particle_err"], float(np.abs(a - b).max() / max(1.0, np.abs(b).max())))
        st = W.ranks[rank]
        nS = cfg.nSpecies
        out["tables_exact"] &= [st.mpi.contents.nEmigrants[i] for i in range(27 * nS)] == list(O.nEmig[rank])
        out["tables_exact"] &= [st.mpi.contents.nImmigrants[i] for i in range(27 * nS)] == list(O.nImm[rank])
    out["mg_path"] = int(W.mg_path())
    out["vcycles_last"] = len(O.history())
    out["emigrants_last_step"] = int(sum(O.nEmig[rank]))
    out["ok"] = bool(out["worst_field_err"] <= tol and out["worst_particle_err"] <= tol and out["tables_exact"]
                     and out["cycles_equal"] and out["sizes_exact"])
    W.close()
    L.pincMgSetReplica(1)
    L.pincMgSetHybrid(1)
    return out


def main():
    import torch
    import torch.distributed as dist
    kind = sys.argv[1] if len(sys.argv) > 1 else "warm"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    replica = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    hybrid = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    r = check(rank, world, kind, steps, replica, hybrid=hybrid, true="32,16,16" if hybrid else "16,8,8")
    assert r["ok"], r
    print(f"rank {rank}/{world} [{kind}] ok: {steps} steps, transport={r['transport']}, mg_path={r['mg_path']}, "
          f"worst field error {r['worst_field_err']:.2e}, emigrants last step {r['emigrants_last_step']}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
