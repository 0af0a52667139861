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


def _worker_tiled(rank, world, port, n, ntaps, tile, ret, in_place=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)
        x = rng.standard_normal((2, n)) - 1.0
        taps = rng.standard_normal(ntaps) + 1j * rng.standard_normal(ntaps)
        plan = FakePlan(taps)
        lo, hi = sharding.time_block(n, rank, world, align=8)
        core = torch.from_numpy(x[:, lo:hi].copy())
        if in_place:                                   # the shard keeps room for its halos: they are received in place
            shard = sharding.TimeShard(2, hi - lo, sharding.required_halo(plan), dtype=torch.float64)
            shard.core.copy_(core)
            core = shard
        got = torch.zeros((2, 1, hi - lo), dtype=torch.complex128)
        order = []

        def consumer(out, a, b):
            got[:, :, a:b] = out[:, :, :b - a]
            order.append((a, b))

        done = sharding.run_time_shard_tiled(plan, core, rank, world, tile, consumer=consumer)
        whole = FakePlan(taps).execute(torch.from_numpy(x), plan.alloc_out(2, n), means=torch.from_numpy(x.mean(axis=1)),
                                       start=0, stop=n, halo_left=0, halo_right=0)
        ret[rank] = (float((got - whole[:, :, lo:hi]).abs().max()), done, hi - lo, order)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("in_place", [False, True])
@pytest.mark.parametrize("world,n,ntaps,tile", [(2, 6000, 301, 500), (3, 9000, 257, 1000), (2, 4000, 129, 4000)])
def test_time_shard_tiles_overlap_the_exchange(world, n, ntaps, tile, in_place):
    """Tiled time shards: interior tiles are transformed while the halos travel, the tiles next to a seam
    afterwards from small edge buffers; the union equals the unsharded transform."""
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker_tiled, args=(world, port, n, ntaps, tile, ret, in_place), nprocs=world, join=True)
    for r in range(world):
        err, done, n_local, order = ret[r]
        assert err < 1e-9, (r, err)
        assert done == 2 * n_local
        assert sorted(order) == [(a, min(n_local, a + tile)) for a in range(0, n_local, tile)]
        if tile < n_local and world > 1:
            seam_tiles = [t for t in order if (r > 0 and t[0] < ntaps - 1) or (r < world - 1 and n_local - t[1] < ntaps - 1)]
            assert order[-len(seam_tiles):] == seam_tiles            # the tiles that need neighbour data come last


def _worker_bad(rank, world, port, n, ntaps, align, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = np.random.default_rng(1).standard_normal((1, n))
        plan = FakePlan(np.ones(ntaps, dtype=complex))
        lo, hi = sharding.time_block(n, rank, world, align=align)
        core = torch.from_numpy(x[:, lo:hi].copy())
        try:
            sharding.run_time_shard(plan, core, rank, world)      # also on a rank whose shard is empty
            ret[rank] = "no error"
        except ValueError as e:
            ret[rank] = str(e)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,ntaps,align,needle", [(1000, 801, 1, "must supply a halo"), (1000, 31, 512, "is empty")])
def test_time_shards_too_short_or_empty_raise_on_every_rank(n, ntaps, align, needle):
    """a shard shorter than the halo (or empty) must not be transformed with silent zero
    padding (or hang the collectives): every rank raises the same ValueError."""
    world = 3 if needle == "is empty" else 2
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker_bad, args=(world, port, n, ntaps, align, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert needle in ret[r], (r, ret[r])
