"""Multi-GPU execution: one process per GPU, torch.distributed for the plumbing.

Two ways the path shards:

* **channels** -- fully independent (the reference handles one channel per call and the
  mean is per channel).  Each rank transforms a contiguous block of channels; there is
  no data-path collective.
* **time** -- one very long recording is cut into contiguous time shards.  The reference
  filter is an exact L-tap FIR, so a shard only needs ``halo`` real samples from each
  neighbour (``halo >= plan.max_length - 1`` also covers the reach of the decimation
  pyramid) plus the global mean: one all-reduce of (sum, count) and one send/recv pair
  with each neighbour.  Results stay sharded.

Everything here works on CPU tensors with the gloo backend too (that is how the host
logic is tested without GPUs); only ``run_*`` touch the device.
"""
from __future__ import annotations

import numpy as np

__all__ = ["channel_block", "time_block", "global_means", "exchange_halos", "run_channel_shard",
           "run_time_shard", "run_time_shard_tiled", "required_halo"]


def _dist():
    import torch.distributed as dist
    return dist


def channel_block(n_channels, rank, world):
    """Contiguous channel range [lo, hi) of ``rank``; sizes differ by at most one."""
    base, extra = divmod(int(n_channels), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def time_block(n_samples, rank, world, align=1):
    """Contiguous sample range [lo, hi) of ``rank``; interior cuts are multiples of ``align``."""
    per = -(-int(n_samples) // int(world))
    per = -(-per // align) * align
    lo = min(n_samples, rank * per)
    return lo, min(n_samples, lo + per)


def required_halo(plan):
    """Samples a time shard needs from each neighbour: the kernel's reach on either side
    is (L_max - 1) / 2, the half-band pyramid adds less than 0.2 L_max at the level the
    longest kernel runs on; L_max - 1 covers both."""
    return int(plan.max_length) - 1


def global_means(local_sum, local_count, group=None):
    """Per-channel mean over all ranks' samples from per-rank float64 sums (the reference
    subtracts the mean of the whole recording, transforms.py:143)."""
    import torch
    dist = _dist()
    buf = torch.cat([local_sum.to(torch.float64).reshape(-1),
                     torch.tensor([float(local_count)], dtype=torch.float64, device=local_sum.device)])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf[:-1] / buf[-1]


def exchange_halos(core, halo, rank, world, group=None):
    """Return ``(padded, halo_left, halo_right)``: ``core`` (channels, n_local) extended by
    up to ``halo`` samples from rank-1 on the left and rank+1 on the right.

    A neighbour shorter than ``halo`` contributes what it has (the caller should size
    shards so that they are longer than the halo).  Rank 0 / world-1 get no halo on their
    outer side: the true signal edges are zero-padded by the kernels, like the reference.
    """
    import torch
    dist = _dist()
    n_ch, n_local = core.shape
    lens = [torch.zeros(1, dtype=torch.int64, device=core.device) for _ in range(world)]
    mine = torch.tensor([n_local], dtype=torch.int64, device=core.device)
    if world > 1:
        dist.all_gather(lens, mine, group=group)
    else:
        lens = [mine]
    lens = [int(v.item()) for v in lens]
    hl = min(halo, lens[rank - 1]) if rank > 0 else 0
    hr = min(halo, lens[rank + 1]) if rank < world - 1 else 0
    padded = torch.empty((n_ch, hl + n_local + hr), dtype=core.dtype, device=core.device)
    padded[:, hl:hl + n_local] = core
    ops = []
    send_r = send_l = recv_l = recv_r = None
    if rank < world - 1:                                  # my tail is the right neighbour's left halo
        k = min(halo, n_local)
        send_r = core[:, n_local - k:].contiguous()
        ops.append(dist.P2POp(dist.isend, send_r, rank + 1, group))
        recv_r = torch.empty((n_ch, hr), dtype=core.dtype, device=core.device)
        ops.append(dist.P2POp(dist.irecv, recv_r, rank + 1, group))
    if rank > 0:
        k = min(halo, n_local)
        send_l = core[:, :k].contiguous()
        ops.append(dist.P2POp(dist.isend, send_l, rank - 1, group))
        recv_l = torch.empty((n_ch, hl), dtype=core.dtype, device=core.device)
        ops.append(dist.P2POp(dist.irecv, recv_l, rank - 1, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if recv_l is not None:
        padded[:, :hl] = recv_l
    if recv_r is not None:
        padded[:, hl + n_local:] = recv_r
    return padded, hl, hr


def run_channel_shard(plan, x_local, out=None):
    """Transform this rank's channel block; no communication."""
    return plan.execute(x_local, out)


def run_time_shard(plan, core, rank, world, out=None, group=None):
    """Time-sharded transform of this rank's ``core`` (channels, n_local) CUDA tensor.

    Returns the (channels, scales, n_local) coefficients of the core samples.  Collectives:
    one all-reduce (global mean), one all-gather of shard lengths, one batched send/recv
    of halos with the two neighbours.
    """
    import torch
    n_local = core.shape[1]
    sums = plan.channel_means(core) * float(n_local)
    means = global_means(sums, n_local, group)
    padded, hl, hr = exchange_halos(core, required_halo(plan), rank, world, group)
    if out is None:
        out = plan.alloc_out(core.shape[0], n_local)
    plan.execute(padded, out, means=means, start=hl, stop=hl + n_local, halo_left=hl, halo_right=hr,
                 out_start=0)
    return out


def run_time_shard_tiled(plan, core, rank, world, tile, out=None, consumer=None, group=None):
    """As :func:`run_time_shard` for shards whose coefficients do not fit in device memory: after
    the mean all-reduce and the halo exchange the shard is transformed in time tiles into a
    reused (channels, scales, tile) buffer (see ``CwtPlan.execute_tiled``).  Returns the number
    of coefficients produced on this rank."""
    n_local = core.shape[1]
    sums = plan.channel_means(core) * float(n_local)
    means = global_means(sums, n_local, group)
    padded, hl, hr = exchange_halos(core, required_halo(plan), rank, world, group)
    tile = int(min(tile, n_local))
    if out is None:
        out = plan.alloc_out(core.shape[0], tile)
    halo = required_halo(plan)
    done = 0
    for a in range(0, n_local, tile):
        b = min(n_local, a + tile)
        plan.execute(padded, out, means=means, start=hl + a, stop=hl + b,
                     halo_left=min(halo, hl + a), halo_right=min(halo, n_local - b + hr), out_start=0)
        if consumer is not None:
            consumer(out, a, b)
        done += (b - a) * core.shape[0] * plan.n_scales
    return done
