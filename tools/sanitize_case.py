"""Small end-to-end case for compute-sanitizer: every kernel family, both precisions."""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
from ghost_b200 import ContinuousWaveletTransform, Morse, synth
from ghost_b200.engine import CwtPlan, scale_tables

fs, n = 1000.0, 70001
x = synth.chirp_pink(n, fs, 0, np.float32)
ts = np.arange(n) / fs
ts[30000:] += 1.0
for dtype in (np.float32, np.float64):
    for output in ("amplitude", "power", "complex"):
        cwt = ContinuousWaveletTransform(dtype=dtype, output=output)
        cwt.transform(x, fs=fs, timestamps=ts, voices_per_octave=4 if dtype == np.float64 else 10)
        print(np.dtype(dtype).name, output, cwt.amplitude.shape, float(np.abs(cwt.amplitude).max()),
              sorted(set(cwt.last_plan.levels().tolist())))
# aligned rows (n multiple of 4) so that the vector-store interpolation path runs, plus halos and tiles
n = 80000
X = synth.recording(3, n, fs, np.float32)
m = Morse(fs=fs)
cw = ContinuousWaveletTransform(dtype=np.float32); cw.fs = fs; cw.wavelet.fs = fs
f = np.asarray(cw.plan_frequencies(n))
om = f / (fs / 2) * np.pi
L = m.compute_lengths(om)
k0, nt, terms = scale_tables(m, om, L)
plan = CwtPlan(L, k0, nt, terms, dtype=np.float32, output="power")
xd = torch.from_numpy(X).cuda()
out = plan.execute(xd)
cnt = plan.execute_tiled(xd, 30000)
torch.cuda.synchronize()
print("tiled ok", cnt, float(out.max()))
