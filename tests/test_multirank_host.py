"""Host-side logic of the N>1 path on the CPU: two processes, torch.distributed with the gloo backend.

What runs between ranks in the product is (a) the bootstrap of bench.py / tools/multi_check.py (a 128-byte id
broadcast from rank 0), (b) the sub-domain <-> rank maps every rank derives for itself (pincMpiAlloc,
puNeighborToRank, puNeighborToReciprocal: rank a's neighbour ne must see a as neighbour reciprocal(ne)), and (c)
the rank-local initial conditions, which must tile the global domain.  No device entry point is called."""
import os
import subprocess
import sys
import textwrap

from helpers import ROOT

WORKER = textwrap.dedent('''
    import ctypes as C, os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
    from helpers import small_cfg, ia
    from pinc_b200 import initial, lib as plib
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # (a) bootstrap: 128 bytes from rank 0 reach every rank unchanged
    payload = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        payload = (torch.arange(128, dtype=torch.int64) * 3 %% 251).to(torch.uint8)
    dist.broadcast(payload, 0)
    assert payload.tolist() == [(i * 3) %% 251 for i in range(128)]
    # (b) topology
    text, cfg = small_cfg("warm_big", grid__nsubdomains="1,1,2", grid__truesize="16,8,8", multigrid__mglevels=3,
                          population__nparticles="8 pc", population__nalloc="16 pc", grid__nemigrantsalloc="4 pc")
    assert cfg.nRanks == world
    L = plib.load()
    m = L.pincMpiAlloc(3, 2, ia(cfg.nSubdomains), ia(cfg.nGhostLayers), ia(cfg.trueSize), rank, world)
    sub = [m.contents.subdomain[d] for d in range(3)]
    off = [m.contents.offset[d] for d in range(3)]
    assert off == initial.rank_offset(rank, cfg) and sub == initial.rank_subdomain(rank, cfg.nSubdomains)
    table = torch.tensor([L.puNeighborToRank(m, ne) for ne in range(27)], dtype=torch.int64)
    gathered = [torch.zeros(27, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, table)
    for ne in range(27):
        peer = int(table[ne])
        rec = L.puNeighborToReciprocal(ne, 3)
        assert int(gathered[peer][rec]) == rank, (rank, ne, peer, rec)
    # (c) rank-local initial conditions tile the global domain
    mine = initial.maxwellian(cfg, seed=3, ranks=[rank])[0]
    n_mine = torch.tensor([len(mine[s][0]) for s in range(2)], dtype=torch.int64)
    dist.all_reduce(n_mine)
    assert n_mine.tolist() == [n // world * world for n in cfg.nParticles]
    for s in range(2):
        g = mine[s][0] + np.array(off, dtype=float)            # global frame
        lo = np.array(sub) * np.array(cfg.trueSize)
        assert (g >= lo).all() and (g < lo + np.array(cfg.trueSize)).all()
    all_ranks = initial.maxwellian(cfg, seed=3)
    assert np.array_equal(all_ranks[rank][0][0], mine[0][0])    # a rank's draw does not depend on who else draws
    dist.barrier()
    print(f"rank{rank}-ok", flush=True)
''') % (ROOT, ROOT)


def test_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", str(script)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("-ok") == 2
