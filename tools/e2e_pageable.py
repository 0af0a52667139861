"""GPU box: the host path with a pageable destination (what transform() allocates itself) against a pinned one."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from ghost_b200 import ContinuousWaveletTransform, synth

wl = bench.WORKLOADS["cfg2"]
nch, n, fs = 8, wl["n"], wl["fs"]
X = np.stack([synth.chirp_pink(n, fs, c, np.float32) for c in range(nch)])
kw = dict(fs=fs, freq_limits=wl["freq_limits"], voices_per_octave=wl["vpo"], multichannel=True)
cwt = ContinuousWaveletTransform(dtype=np.float32)
cwt.transform(X[:1], **kw)
S = cwt.frequencies.size
gb = nch * S * n * 4 / 1e9
for label, make in (("pageable, fresh np.empty per call (transform() default)", lambda: None),
                    ("pageable, reused array", "reuse"),
                    ("pinned, reused array", "pinned")):
    out = None
    if make == "reuse":
        out = np.empty((nch, S, n), dtype=np.float32); out[:] = 0
    elif make == "pinned":
        out = torch.empty((nch, S, n), dtype=torch.float32).pin_memory().numpy()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        cwt.transform(X, out=out, **kw)
        a = cwt.amplitude
        ts.append(time.perf_counter() - t0)
    print("%-58s %.3f s best of 3 = %.1f GB/s  %s" % (label, min(ts), gb / min(ts), cwt.last_plan.host_stats()), flush=True)
