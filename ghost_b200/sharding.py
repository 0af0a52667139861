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
           "run_time_shard", "run_time_shard_tiled", "required_halo", "HaloExchange", "check_time_shards",
           "gather_shard_lengths", "TimeShard"]


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


def gather_shard_lengths(n_local, rank, world, device, group=None):
    """Every rank's shard length (one all-gather of an int64)."""
    import torch
    dist = _dist()
    mine = torch.tensor([int(n_local)], dtype=torch.int64, device=device)
    if world > 1:
        lens = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
        dist.all_gather(lens, mine, group=group)
    else:
        lens = [mine]
    return [int(v.item()) for v in lens]


def check_time_shards(lens, halo):
    """Raise -- on every rank alike, since all ranks hold the same ``lens`` -- when the partition cannot
    be transformed exactly: an empty shard would leave a rank without work (and the others waiting in
    the collectives), and a shard shorter than ``halo`` next to an interior seam cannot supply its
    neighbour's halo (the samples of rank +-2 would silently read as zero padding)."""
    for r, n in enumerate(lens):
        if n <= 0:
            raise ValueError("time shard of rank {} is empty ({} ranks for these samples): use fewer ranks"
                             " or a smaller alignment".format(r, len(lens)))
    for r in range(len(lens) - 1):
        if lens[r] < halo or lens[r + 1] < halo:
            raise ValueError("time shards of ranks {} and {} hold {} and {} samples but each must supply a halo"
                             " of {} samples (the widest wavelet's support): use fewer ranks".format(
                                 r, r + 1, lens[r], lens[r + 1], halo))


class TimeShard:
    """A rank's time shard stored with room for its halos: ``buf`` is (channels, halo + n_local + halo) and
    ``core`` the view of the n_local samples this rank owns.  Fill ``core`` (e.g. ``shard.core.copy_(...)``);
    the transforms below then receive the neighbours' halos straight into the margins, so the tiles next to a
    seam are transformed in place -- no edge buffers, no copies of the shard."""

    def __init__(self, n_channels, n_local, halo, dtype=None, device=None):
        import torch
        self.halo, self.n_local = int(halo), int(n_local)
        self.buf = torch.zeros((int(n_channels), self.n_local + 2 * self.halo),
                               dtype=torch.float32 if dtype is None else dtype, device=device)
        self.core = self.buf[:, self.halo:self.halo + self.n_local]


class HaloExchange:
    """The halo send / receive pairs with both neighbours, posted at construction and completed by
    ``wait()`` -- so that the tiles which need no neighbour data can be transformed in between.

    ``left`` / ``right`` (after ``wait``): (channels, halo) tensors holding the last ``halo`` samples of
    rank - 1 and the first ``halo`` samples of rank + 1, or None at the true ends of the recording
    (which the kernels zero-pad, like the reference)."""

    def __init__(self, core, halo, rank, world, group=None, into=None):
        import torch
        dist = _dist()
        n_ch, n_local = core.shape
        self.left = self.right = None
        self._reqs = []
        self._keep = []
        self._into = into                    # a TimeShard: wait() moves the received halos into its margins
        ops = []
        if halo > 0 and rank < world - 1:
            tail = core[:, n_local - halo:].contiguous()
            self.right = torch.empty((n_ch, halo), dtype=core.dtype, device=core.device)
            ops += [dist.P2POp(dist.isend, tail, rank + 1, group), dist.P2POp(dist.irecv, self.right, rank + 1, group)]
            self._keep.append(tail)
        if halo > 0 and rank > 0:
            head = core[:, :halo].contiguous()
            self.left = torch.empty((n_ch, halo), dtype=core.dtype, device=core.device)
            ops += [dist.P2POp(dist.isend, head, rank - 1, group), dist.P2POp(dist.irecv, self.left, rank - 1, group)]
            self._keep.append(head)
        if ops:
            self._reqs = dist.batch_isend_irecv(ops)

    def wait(self):
        for req in self._reqs:
            req.wait()
        self._reqs = []
        if self._into is not None:           # halo-sized copies (<= 1 MB per channel) into the shard's own margins
            sh = self._into
            if self.left is not None:
                sh.buf[:, :sh.halo].copy_(self.left)
            if self.right is not None:
                sh.buf[:, sh.halo + sh.n_local:].copy_(self.right)
            self._into = None
        return self.left, self.right


def exchange_halos(core, halo, rank, world, group=None):
    """Return ``(padded, halo_left, halo_right)``: ``core`` (channels, n_local) extended by ``halo``
    samples from rank-1 on the left and rank+1 on the right (blocking form; the transform paths below use
    :class:`HaloExchange` and never copy the whole shard).  Raises on every rank when a shard is empty or
    shorter than ``halo`` (:func:`check_time_shards`)."""
    import torch
    lens = gather_shard_lengths(core.shape[1], rank, world, core.device, group)
    check_time_shards(lens, halo)
    left, right = HaloExchange(core, halo, rank, world, group).wait()
    parts = ([left] if left is not None else []) + [core] + ([right] if right is not None else [])
    return torch.cat(parts, dim=1), (0 if left is None else halo), (0 if right is None else halo)


def run_channel_shard(plan, x_local, out=None):
    """Transform this rank's channel block; no communication."""
    return plan.execute(x_local, out)


def _time_shard_tiles(plan, core, rank, world, tile, emit, group=None, means=None, lens=None):
    """Shared driver of the time-sharded transforms.  ``emit(a, b)`` returns ``(out, out_start)`` for the
    tile of core samples [a, b) and is told when the tile is complete through ``emit.done(out, a, b)``.

    Collectives: one all-reduce (global mean, reference transforms.py:143), one all-gather of shard
    lengths, one batched send/recv of ``L_max - 1`` samples with each neighbour.  The exchange is posted
    first; the interior tiles -- every tile further than the halo from both ends of the shard -- are
    transformed while it is in flight; the tiles next to a seam follow from small edge buffers
    (halo + tile samples), so the shard itself is never copied."""
    import torch
    shard = core if isinstance(core, TimeShard) else None
    if shard is not None:
        core = shard.core
    n_ch, n_local = core.shape
    halo = required_halo(plan)
    if shard is not None and shard.halo < halo:
        raise ValueError("TimeShard was made with a halo of {} samples but this plan needs {}".format(shard.halo, halo))
    # lengths first: an unusable partition must raise on every rank before any rank enters another collective
    # (``lens`` from an earlier call on the same partition skips the all-gather and its host synchronisation)
    if lens is None:
        lens = gather_shard_lengths(n_local, rank, world, core.device, group)
    check_time_shards(lens, halo)
    if means is None:
        sums = plan.channel_means(core) * float(n_local)
        means = global_means(sums, n_local, group)
    xch = HaloExchange(core, halo, rank, world, group, into=shard)
    has_l, has_r = rank > 0 and halo > 0, rank < world - 1 and halo > 0
    tile = int(min(tile, n_local))
    tiles = [(a, min(n_local, a + tile)) for a in range(0, n_local, tile)]
    edge = []
    done = 0
    for a, b in tiles:
        if (has_l and a < halo) or (has_r and n_local - b < halo):
            edge.append((a, b))
            continue
        out, o0 = emit(a, b)
        plan.execute(core, out, means=means, start=a, stop=b, halo_left=min(halo, a), halo_right=min(halo, n_local - b),
                     out_start=o0)
        emit.done(out, a, b)
        done += (b - a) * n_ch * plan.n_scales
    left, right = xch.wait()
    for a, b in edge:
        if shard is not None:
            # the halos sit in the shard's own margins: transform the tile where it is
            o = shard.halo                                        # index of core sample 0 in shard.buf
            hl = min(halo, a + (halo if has_l else 0))
            hr = min(halo, n_local - b + (halo if has_r else 0))
            out, o0 = emit(a, b)
            plan.execute(shard.buf, out, means=means, start=o + a, stop=o + b, halo_left=hl, halo_right=hr, out_start=o0)
            emit.done(out, a, b)
            done += (b - a) * n_ch * plan.n_scales
            continue
        # [left halo | core[lo:hi] | right halo]: only as much of the core as this tile can reach
        lo, hi = max(0, a - halo), min(n_local, b + halo)
        parts, off = [], 0
        if has_l and a < halo:
            parts.append(left[:, a:])                      # the last (halo - a) samples of rank - 1
            off = halo - a
        parts.append(core[:, lo:hi])
        hr = 0
        if has_r and n_local - b < halo:
            hr = halo - (n_local - b)
            parts.append(right[:, :hr])
        buf = torch.cat(parts, dim=1)
        start = off + (a - lo)
        out, o0 = emit(a, b)
        plan.execute(buf, out, means=means, start=start, stop=start + (b - a), halo_left=min(halo, start),
                     halo_right=min(halo, buf.shape[1] - (start + (b - a))), out_start=o0)
        emit.done(out, a, b)
        done += (b - a) * n_ch * plan.n_scales
    return done


class _Emit:
    def __init__(self, get, done=None):
        self._get, self._done = get, done

    def __call__(self, a, b):
        return self._get(a, b)

    def done(self, out, a, b):
        if self._done is not None:
            self._done(out, a, b)


def run_time_shard(plan, core, rank, world, out=None, group=None, tile=None):
    """Time-sharded transform of this rank's ``core`` (channels, n_local) CUDA tensor, or of a
    :class:`TimeShard` (halos received in place, no edge buffers).

    Returns the (channels, scales, n_local) coefficients of the core samples (they stay sharded)."""
    inner = core.core if isinstance(core, TimeShard) else core
    n_local = inner.shape[1]
    if out is None:
        out = plan.alloc_out(inner.shape[0], n_local)
    _time_shard_tiles(plan, core, rank, world, n_local if tile is None else tile, _Emit(lambda a, b: (out, a)), group)
    return out


def run_time_shard_tiled(plan, core, rank, world, tile, out=None, consumer=None, group=None, means=None, lens=None):
    """As :func:`run_time_shard` for shards whose coefficients do not fit in device memory: the shard is
    transformed in time tiles into a reused (channels, scales, tile) buffer and every finished tile is
    handed to ``consumer(out, a, b)`` (core sample range [a, b); interior tiles come first, the tiles
    next to a seam last).  ``lens``: the shard lengths of all ranks when they are already known (e.g. from
    :func:`gather_shard_lengths` once per partition).  Returns the number of coefficients produced on this rank."""
    inner = core.core if isinstance(core, TimeShard) else core
    tile = int(min(tile, inner.shape[1]))
    if out is None:
        out = plan.alloc_out(inner.shape[0], tile)
    return _time_shard_tiles(plan, core, rank, world, tile, _Emit(lambda a, b: (out, 0), consumer), group, means, lens)
