"""Host logic of the multi-GPU paths on CPU tensors, world_size 2 and 3, gloo backend.

The compute kernel is replaced by a numpy FIR with the same halo semantics as
gcwt_execute (zero outside the readable range), so what is checked is the plumbing:
partitioning, global mean all-reduce and halo exchange reproduce the unsharded result.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ghost_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class FakePlan:
    """Stand-in for CwtPlan on the CPU: one 'scale', a fixed complex FIR, 'same' alignment."""

    def __init__(self, taps):
        self.taps = taps
        self.max_length = len(taps)
        self.n_scales = 1

    def channel_means(self, x):
        return x.double().mean(dim=1)

    def alloc_out(self, n_ch, n):
        return torch.zeros((n_ch, 1, n), dtype=torch.complex128)

    def execute(self, x, out, *, means, start, stop, halo_left, halo_right, out_start=None):
        L = len(self.taps)
        adv = (L - 1) // 2
        out_start = start if out_start is None else out_start
        for c in range(x.shape[0]):
            seg = x[c, start - halo_left:stop + halo_right].double().numpy() - float(means[c])
            full = np.convolve(seg, self.taps)                 # zero outside the readable range
            first = halo_left + adv
            out[c, 0, out_start:out_start + stop - start] = torch.from_numpy(full[first:first + stop - start])
        return out


def _worker(rank, world, port, n, n_ch, ntaps, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        x = rng.standard_normal((n_ch, n)) + 3.0
        taps = rng.standard_normal(ntaps) + 1j * rng.standard_normal(ntaps)
        plan = FakePlan(taps)
        # ---- time sharding
        lo, hi = sharding.time_block(n, rank, world, align=16)
        core = torch.from_numpy(x[:, lo:hi].copy())
        got = sharding.run_time_shard(plan, core, rank, world)
        whole = FakePlan(taps).execute(torch.from_numpy(x), plan.alloc_out(n_ch, n),
                                       means=torch.from_numpy(x.mean(axis=1)), start=0, stop=n,
                                       halo_left=0, halo_right=0)
        err_t = float((got - whole[:, :, lo:hi]).abs().max())
        # ---- channel sharding
        clo, chi = sharding.channel_block(n_ch, rank, world)
        mine = torch.from_numpy(x[clo:chi].copy())
        gotc = plan.execute(mine, plan.alloc_out(chi - clo, n), means=plan.channel_means(mine), start=0,
                            stop=n, halo_left=0, halo_right=0)
        err_c = float((gotc - whole[clo:chi]).abs().max()) if chi > clo else 0.0
        # ---- mean
        m = sharding.global_means(core.double().sum(dim=1), hi - lo)
        err_m = float((m - torch.from_numpy(x.mean(axis=1))).abs().max())
        ret[rank] = (err_t, err_c, err_m, lo, hi, clo, chi)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,n_ch,ntaps", [(2, 4000, 3, 257), (3, 5000, 2, 600), (2, 1000, 1, 31)])
def test_time_and_channel_sharding_gloo(world, n, n_ch, ntaps):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n, n_ch, ntaps, ret), nprocs=world, join=True)
    assert len(ret) == world
    covered = 0
    chans = 0
    for r in range(world):
        err_t, err_c, err_m, lo, hi, clo, chi = ret[r]
        assert err_t < 1e-9, (r, err_t)
        assert err_c < 1e-9 and err_m < 1e-12
        assert lo == covered
        covered = hi
        assert clo == chans
        chans = chi
    assert covered == n and chans == n_ch


def test_partition_helpers():
    for n, w in [(64, 8), (10, 3), (5, 8), (256, 8)]:
        blocks = [sharding.channel_block(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks[:-1], blocks[1:]))
        sizes = [b[1] - b[0] for b in blocks]
        assert max(sizes) - min(sizes) <= 1
    blocks = [sharding.time_block(2592000000, r, 8, align=16384) for r in range(8)]
    assert blocks[0][0] == 0 and blocks[-1][1] == 2592000000
    assert all(a[1] == b[0] and a[1] % 16384 == 0 for a, b in zip(blocks[:-1], blocks[1:]))
