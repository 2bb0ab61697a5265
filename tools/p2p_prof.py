#!/usr/bin/env python
"""In-kernel cycle accounting of a multi-rank multigrid solve (run under torchrun with PINC_B200_MGPROF=1):
    PINC_B200_MGPROF=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/p2p_prof.py [hybrid 1|0]
Prints us per V-cycle of every accounted phase of rank 0's persistent kernel (thread 0 of CTA 0 / of the first block CTA)."""
import ctypes as C, json, os, sys
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
L.pincMgSetHybrid(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
buf = C.create_string_buffer(128)
if rank == 0:
    L.pincNcclUniqueId(buf)
t = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda"); dist.broadcast(t, 0)
text, cfg = bench.load_cfg("warm", world, 0.1)
W = sim.World(cfg, rank=rank, world_size=world, nccl_id=bytes(t.cpu().tolist()))
W.set_particles({rank: initial.maxwellian(cfg, seed=1, ranks=[rank])[0]})
W.migrate(); W.field_solve(); W.half_kick()
W.step()
out = (C.c_longlong * 64)()
L.pincMgProfRead.argtypes = [C.POINTER(C.c_longlong)]
L.pincMgProfRead(out)
solves = 3
L.pincTimerStart()
for _ in range(solves):
    W.step()
ms = L.pincTimerStopMs()
L.pincMgProfRead(out)
names = ["neut_rho", "gs_big", "gs_small", "restrict", "prolong", "neut_phi", "norm", "x_halo", "small_section", "b_load", "b_wait", "r_sync", "b_tail", "xh_send", "xh_recv", "xh_sync",
         "gs16", "res16", "pro16", "-", "gs8", "res8", "pro8", "-", "gs4", "res4", "pro4", "-", "f_neut", "f_resid", "f_restr", "f_prol"]
ncyc = len(W.history())
if rank == 0:
    print(json.dumps({"ranks": world, "mg_path": W.mg_path(), "vcycles": ncyc, "ms_per_step": ms / solves,
                      "us_per_vcycle": {n: round(out[2 * i] / 1965.0 / (solves * ncyc), 2) for i, n in enumerate(names) if n != "-"},
                      "calls_per_vcycle": {n: round(out[2 * i + 1] / (solves * ncyc), 1) for i, n in enumerate(names) if n != "-"}}))
dist.barrier(); W.close()
