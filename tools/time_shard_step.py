"""torchrun on 2 GPUs: per-rank timing of the cfg4 bench step, phase by phase."""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ghost_b200 import sharding, Morse
from ghost_b200.engine import CwtPlan, scale_tables
import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
wl = dict(bench.WORKLOADS["cfg4"])
fs, n, tile = wl["fs"], wl["n"], wl["tile"]
freqs = bench.plan_frequencies(wl)
m = Morse(fs=fs); om = freqs / (fs / 2.0) * np.pi; L = m.compute_lengths(om)
k0, nt, terms = scale_tables(m, om, L)
plan = CwtPlan(L, k0, nt, terms, dtype=np.float32, output="power", device=local)
core = torch.randn((1, n), dtype=torch.float32, device=dev)
out = plan.alloc_out(1, tile)
halo = sharding.required_halo(plan)
for step in range(4):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    T = [time.perf_counter()]
    sums = plan.channel_means(core) * float(n); means = sharding.global_means(sums, n)
    torch.cuda.synchronize(); T.append(time.perf_counter())
    padded, hl, hr = sharding.exchange_halos(core, halo, rank, world)
    torch.cuda.synchronize(); T.append(time.perf_counter())
    per_tile = []
    for a in range(0, n, tile):
        b = min(n, a + tile)
        t0 = time.perf_counter()
        plan.execute(padded, out, means=means, start=hl + a, stop=hl + b, halo_left=min(halo, hl + a),
                     halo_right=min(halo, n - b + hr), out_start=0)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        per_tile.append(((t1 - t0) * 1e3, (time.perf_counter() - t0) * 1e3))
    T.append(time.perf_counter())
    print("rank %d step %d: means %.1f ms, halos %.1f ms, tiles %.1f ms; per tile (launch ms, total ms): %s" % (
        rank, step, (T[1] - T[0]) * 1e3, (T[2] - T[1]) * 1e3, (T[3] - T[2]) * 1e3,
        ["%.1f/%.1f" % p for p in per_tile]), flush=True)
dist.barrier(); dist.destroy_process_group()
