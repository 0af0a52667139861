"""GPU box: randomized check of the fp64 path (four-step FFT) against the CPU oracle, hostile spectra included.
Bar: per scale max|a - b| / max|b| <= 1e-10.      python tools/fuzz_fp64_vs_oracle.py [n_cases] [seed]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from ghost_b200 import ContinuousWaveletTransform, Morse
from oracle import cwt_oracle as orc

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
worst, fails = 0.0, []
for case in range(n_cases):
    gamma = float(rng.choice([1, 2, 3, 3, 4, 6, 9]))
    beta = float(rng.choice([1, 3, 5, 10, 20, 40, 80]))
    fs = float(rng.choice([200.0, 1000.0, 30000.0]))
    n = int(rng.choice([700, 1024, 4097, 20000, 65536, 150001, 262144]))
    vpo = int(rng.choice([4, 8, 10, 16]))
    kind = str(rng.choice(["walk", "tilt", "highpass", "tones", "white"]))
    spec = np.fft.rfft(rng.standard_normal(n))
    fr = np.fft.rfftfreq(n, 1.0 / fs)
    g = np.ones_like(fr)
    if kind == "tilt":
        g = (np.maximum(fr, fr[1]) / fr[-1]) ** float(rng.uniform(-1.5, 4.0))
    elif kind == "highpass":
        g = (fr >= float(rng.uniform(0.01, 0.4)) * fs / 2).astype(float)
    x = np.fft.irfft(spec * g, n=n)
    x = x / x.std()
    if kind == "walk":
        x = np.cumsum(x) * 0.05 + rng.standard_normal(n)
    if kind == "tones":
        t = np.arange(n) / fs
        x = 1e-3 * x + np.sin(2 * np.pi * 0.31 * fs * t) + 1e-4 * np.sin(2 * np.pi * 0.004 * fs * t)
    x = x + float(rng.uniform(-3, 3))
    ts = None
    if n >= 4097 and rng.random() < 0.4:
        ts = np.arange(n) / fs
        ts[int(n * rng.uniform(0.3, 0.7)):] += 10.0 / fs
    W, f, _ = orc.cwt_complex(x, fs, gamma=gamma, beta=beta, voices_per_octave=vpo, timestamps=ts, parallel=True)
    cwt = ContinuousWaveletTransform(wavelet=Morse(gamma=gamma, beta=beta), output="complex")
    kw = dict(fs=fs, voices_per_octave=vpo)
    if ts is not None:
        kw["timestamps"] = ts
    cwt.transform(x, **kw)
    if len(f) == 0:
        print("case %d: no scales" % case); continue
    assert cwt.frequencies.tolist() == f.tolist()
    err = np.max(np.abs(cwt.coefficients - W), axis=1) / np.max(np.abs(W), axis=1)
    worst = max(worst, float(err.max()))
    tag = "ok " if err.max() <= 1e-10 else "BAD"
    if err.max() > 1e-10:
        fails.append((case, gamma, beta, fs, n, vpo, kind, float(err.max())))
    print("case %2d %s %-8s g=%g b=%g fs=%g n=%d vpo=%d epochs=%d S=%d err=%.2e" % (
        case, tag, kind, gamma, beta, fs, n, vpo, 2 if ts is not None else 1, len(f), err.max()), flush=True)
print("worst %.2e, failures %d" % (worst, len(fails)))
for f_ in fails:
    print("FAIL", f_)
sys.exit(1 if fails else 0)
