"""Run under torchrun on >= 2 GPUs: time-sharded transform (NCCL halo exchange + mean
all-reduce) must equal the unsharded transform of the same recording."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ghost_b200 import Morse, synth, sharding            # noqa: E402
from ghost_b200.engine import CwtPlan, scale_tables      # noqa: E402
from ghost_b200 import ContinuousWaveletTransform        # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fs, n = 30000.0, 6000000
    cwt = ContinuousWaveletTransform(dtype=np.float32)
    cwt.fs = fs
    cwt.wavelet.fs = fs
    f = np.asarray(cwt.plan_frequencies(n, freq_limits=[8.0, 15000.0], voices_per_octave=10))
    m = Morse(fs=fs)
    om = f / (fs / 2.0) * np.pi
    L = m.compute_lengths(om)
    k0, nt, terms = scale_tables(m, om, L)
    x = synth.chirp_pink(n, fs, 0, np.float32)[None, :] + 0.37
    ok = True
    for dtype, bar in ((np.float32, 2e-6), (np.float64, 1e-11)):
        if dtype == np.float64:
            sel = slice(0, len(f), 12)
            plan = CwtPlan(L[sel], k0[sel], nt[sel], np.concatenate([terms[sum(nt[:i]):sum(nt[:i + 1])] for i in range(len(f))[sel]]),
                           dtype=dtype, device=local)
        else:
            plan = CwtPlan(L, k0, nt, terms, dtype=dtype, device=local)
        lo, hi = sharding.time_block(n, rank, world, align=16384)
        core = torch.from_numpy(x[:, lo:hi].copy()).cuda()
        got = sharding.run_time_shard(plan, core, rank, world)
        whole = plan.execute(torch.from_numpy(x).cuda())[:, :, lo:hi]
        err = (torch.linalg.vector_norm((got - whole).double(), dim=2) / torch.linalg.vector_norm(whole.double(), dim=2)).max().item()
        print("rank %d %s shard [%d,%d) halo %d scales %d: max rel-L2 vs unsharded %.3e" % (
            rank, np.dtype(dtype).name, lo, hi, sharding.required_halo(plan), plan.n_scales, err), flush=True)
        ok = ok and err <= bar
        del got, whole, plan
        torch.cuda.empty_cache()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
