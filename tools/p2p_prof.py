#!/usr/bin/env python
"""Cycle accounting of the multi-rank peer-memory smoother (run under torchrun with PINC_B200_MGPROF=1)."""
import ctypes as C, os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from pinc_b200 import config, initial, lib as plib, sim
import bench
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L = plib.load()
buf = C.create_string_buffer(128)
if rank == 0:
    L.pincNcclUniqueId(buf)
t = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda"); dist.broadcast(t, 0)
text, cfg = bench.load_cfg("warm", world, 0.1)
W = sim.World(cfg, rank=rank, world_size=world, nccl_id=bytes(t.cpu().tolist()))
W.set_particles({rank: initial.maxwellian(cfg, seed=1, ranks=[rank])[0]})
W.migrate(); W.field_solve(); W.half_kick()
W.step()
out = (C.c_longlong * 32)()
L.pincMgProfRead.argtypes = [C.POINTER(C.c_longlong)]
L.pincMgProfRead(out)
W.step()
L.pincMgProfRead(out)
n = max(out[21], 1)
if rank == 0:
    print({"half_sweeps": out[21], "us_sweep_tail": out[20] / 1965 / n, "us_sys_fence": out[22] / 1965 / n,
           "us_ticket_and_wait": out[24] / 1965 / n, "us_sweep_thread0": out[26] / 1965 / n, "vcycles": len(W.history())})
dist.barrier(); W.close()
