#!/usr/bin/env python
"""Multi-GPU parity check: one process per GPU (torchrun), NCCL transport, against the oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_check.py [cold|warm] [steps]

Every rank steps its own sub-domain on its GPU; every rank also steps the WHOLE world in the oracle on the CPU
(small config) and compares its own sub-domain: population sizes and migrant tables exact, fields and particle
phase space to 1e-10 relative."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import small_cfg, sorted_particles  # noqa: E402
from oracle import orc  # noqa: E402
from pinc_b200 import initial, lib as plib, sim  # noqa: E402

SUB = {1: "1,1,1", 2: "1,1,2", 4: "1,2,2", 8: "2,2,2"}


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "warm"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = plib.load()
    buf = C.create_string_buffer(128)
    if rank == 0:
        L.pincNcclUniqueId(buf)
    t = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    nccl_id = bytes(t.cpu().tolist())
    over = dict(grid__nsubdomains=SUB[world], grid__truesize="16,8,8", multigrid__mglevels=3, population__nparticles="8 pc",
                population__nalloc="24 pc", grid__nemigrantsalloc="4 pc")
    if kind == "cold":
        over["population__perturbamplitude"] = "2e-3,0,0,0,0,0"
        text, cfg = small_cfg("cold", **over)
        per_rank = initial.perturb(cfg, initial.lattice(cfg))
    else:
        over["population__thermalvelocitycells"] = "0.08,0.004"
        text, cfg = small_cfg("warm_big", **over)
        per_rank = initial.maxwellian(cfg, seed=7)
    W = sim.World(cfg, rank=rank, world_size=world, nccl_id=nccl_id)
    O = orc.OrcWorld(cfg)
    W.set_particles({rank: per_rank[rank]})
    O.set_particles(per_rank)
    for X in (W, O):
        X.migrate(); X.field_solve(); X.half_kick()
    worst = 0.0
    for it in range(steps):
        W.step(fused=(it % 2 == 1)); O.step()
        if it % 2 == 1:
            continue                      # after a fused step the positions are one puMove ahead
        for name in ("rho", "phi", "E"):
            a, b = W.grid(rank, name), O.grid(rank, name)
            err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
            worst = max(worst, err)
            assert err <= 1e-10, (rank, it, name, err)
        got, ref = W.particles(rank), O.particles(rank)
        for s in range(cfg.nSpecies):
            assert len(got[s][0]) == len(ref[s][0]), (rank, it, s, len(got[s][0]), len(ref[s][0]))
            a, b = sorted_particles(*got[s]), sorted_particles(*ref[s])
            assert np.abs(a - b).max() <= 1e-10 * max(1.0, np.abs(b).max())
        st = W.ranks[rank]
        nS = cfg.nSpecies
        assert [st.mpi.contents.nEmigrants[i] for i in range(27 * nS)] == list(O.nEmig[rank])
        assert [st.mpi.contents.nImmigrants[i] for i in range(27 * nS)] == list(O.nImm[rank])
        assert len(W.history()) == len(O.history())
    moved = int(sum(O.nEmig[rank]))
    print(f"rank {rank}/{world} [{kind}] ok: {steps} steps, transport={L.pincTransportName().decode()}, "
          f"worst field error {worst:.2e}, emigrants last step {moved}", flush=True)
    dist.barrier()
    W.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
