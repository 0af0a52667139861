"""API-level latency of the drop-in on config 1 (1 ch x 60 s x 1 kHz), first and repeated calls."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from ghost_b200 import ContinuousWaveletTransform, synth
from oracle import cwt_oracle as orc

x = synth.chirp_pink(60000, 1000.0, 0, np.float32)
torch.zeros(1).cuda()
for dtype in (np.float32, np.float64):
    cwt = ContinuousWaveletTransform(dtype=dtype)
    ts = []
    for i in range(4):
        t0 = time.perf_counter(); cwt.transform(x, fs=1000.0); a = cwt.amplitude; ts.append(time.perf_counter() - t0)
    print(np.dtype(dtype).name, "transform()+amplitude seconds: first %.3f then %s" % (ts[0], ["%.4f" % t for t in ts[1:]]))
t0 = time.perf_counter(); orc.cwt_amplitude(x, 1000.0, parallel=True); print("oracle port parallel: %.3f s" % (time.perf_counter() - t0))
t0 = time.perf_counter(); orc.cwt_amplitude(x, 1000.0, parallel=False); print("oracle port serial: %.3f s" % (time.perf_counter() - t0))
