"""GPU box: what the accuracy guard costs when it fires.  Config-2 shape with 8 channels of white noise high-passed
at 0.2 Nyquist (every scale below the cut is re-computed in fp64) next to the benchmark's chirp + pink input."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from ghost_b200 import synth

wl = dict(bench.WORKLOADS["cfg2"])
nch, n, fs = 8, wl["n"], wl["fs"]
rng = np.random.default_rng(1)
spec = np.fft.rfft(rng.standard_normal(n))
f = np.fft.rfftfreq(n, 1 / fs)
hp = np.fft.irfft(spec * (f >= 0.2 * fs / 2), n=n).astype(np.float32)
inputs = {"chirp_pink": np.stack([synth.chirp_pink(n, fs, c, np.float32) for c in range(nch)]),
          "white_hp02": np.stack([np.roll(hp, 1009 * c) for c in range(nch)])}
for guard in (True, False):
    plan, freqs = bench.build_plan(wl, 0, guard=guard)
    for name, X in inputs.items():
        x = torch.from_numpy(X).cuda()
        out = plan.alloc_out(nch, n)
        plan.execute(x, out)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            plan.execute(x, out)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 3 * 1e3
        st = plan.guard_stats()
        print("guard %-5s %-11s %8.2f ms/step  %.3e coeff/s  re-computed pairs %d of %d" % (
            guard, name, ms, nch * n * len(freqs) / (ms * 1e-3), st["last"], nch * len(freqs)), flush=True)
        del x, out
    plan.close()
