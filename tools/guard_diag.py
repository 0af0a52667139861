"""GPU box: which (tile, scale) pairs the accuracy guard re-computes on a benchmark workload, and why
(run with GCWT_GUARD_DUMP=1 to get the per-term breakdown on stderr).
    python tools/guard_diag.py cfg3 [channels]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from ghost_b200 import Morse, synth
from ghost_b200.engine import CwtPlan, scale_tables

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
nch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = dict(bench.WORKLOADS[name])
fs, n = wl["fs"], wl["n"]
freqs = bench.plan_frequencies(wl)
m = Morse(fs=fs)
om = freqs / (fs / 2.0) * np.pi
L = m.compute_lengths(om)
k0, nt, terms = scale_tables(m, om, L)
plan = CwtPlan(L, k0, nt, terms, dtype=np.float32, output=wl["output"])
x = torch.from_numpy(np.stack([synth.chirp_pink(n, fs, c, np.float32) for c in range(nch)])).cuda()
means = plan.channel_means(x)
tile = int(wl.get("tile", n))
out = plan.alloc_out(nch, tile)
halo = plan.max_length - 1
for a in range(0, n, tile):
    b = min(n, a + tile)
    plan.execute(x, out, means=means, start=a, stop=b, halo_left=min(halo, a), halo_right=min(halo, n - b), out_start=0)
    st = plan.guard_stats()
    print("tile [%d, %d): re-computed pairs %d, scales %s" % (a, b, st["last"], np.flatnonzero(st["scales"]).tolist()), flush=True)
