"""torchrun on 2 GPUs: wall time of each phase of a time-sharded step."""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ghost_b200 import sharding

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
n_local, halo = 324000000, 236759
core = torch.randn((1, n_local), dtype=torch.float32, device=dev)

def t(label, fn, reps=3):
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    if rank == 0: print("%-28s %.2f ms" % (label, dt * 1e3), flush=True)
    return r

t("all_reduce(2 doubles)", lambda: sharding.global_means(torch.ones(1, dtype=torch.float64, device=dev), n_local))
t("all_gather lens + item", lambda: [int(v.item()) for v in (lambda L: (dist.all_gather(L, torch.tensor([n_local], device=dev)), L)[1])([torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)])])
t("torch.empty padded", lambda: torch.empty((1, n_local + halo), dtype=torch.float32, device=dev))
t("exchange_halos (all)", lambda: sharding.exchange_halos(core, halo, rank, world))
def p2p():
    ops = []
    other = 1 - rank
    s = core[:, :halo].contiguous(); r = torch.empty((1, halo), dtype=torch.float32, device=dev)
    ops.append(dist.P2POp(dist.isend, s, other)); ops.append(dist.P2POp(dist.irecv, r, other))
    for q in dist.batch_isend_irecv(ops): q.wait()
t("batch_isend_irecv 0.95 MB", p2p)
dist.barrier(); dist.destroy_process_group()
