"""GPU box: effect of the planner's band tolerance on class assignment, accuracy and speed (config 2, 8 channels)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from ghost_b200 import Morse, synth, ContinuousWaveletTransform
from ghost_b200.engine import CwtPlan, scale_tables
from oracle import cwt_oracle as orc

fs, n = 1250.0, 2250000
cw = ContinuousWaveletTransform(dtype=np.float32); cw.fs = fs; cw.wavelet.fs = fs
f = np.asarray(cw.plan_frequencies(n, freq_limits=[0.40, 300.0]))
m = Morse(fs=fs); om = f / (fs / 2) * np.pi; L = m.compute_lengths(om)
k0, nt, terms = scale_tables(m, om, L)
# accuracy on a short recording against the oracle
ns = 300000
xs = synth.chirp_pink(ns, fs, 0, np.float32)
sel = np.arange(0, len(f), 3)
W, _, _ = orc.cwt_complex(xs, fs, frequencies=f[sel], parallel=True)
amp = np.abs(W)
X = torch.from_numpy(synth.recording(8, n, fs, np.float32)).cuda()
for tol in (3e-7, 1e-6, 3e-6):
    plan = CwtPlan(L, k0, nt, terms, dtype=np.float32, band_tol=tol)
    lev = plan.levels()
    got = plan.execute(torch.from_numpy(xs[None, :]).cuda())[0].cpu().numpy().astype(np.float64)[sel]
    err = np.linalg.norm(got - amp, axis=1) / np.linalg.norm(amp, axis=1)
    out = plan.alloc_out(8, n)
    for _ in range(3): plan.execute(X, out)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): plan.execute(X, out)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print("tol %.0e: levels %s  max relL2 %.2e  %.3f ms/step (8 ch)" % (
        tol, {int(k): int((lev == k).sum()) for k in sorted(set(lev.tolist()))}, err.max(), dt * 1e3), flush=True)
    del plan, out
