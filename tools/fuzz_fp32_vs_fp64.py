"""GPU box: randomized differential test, fp32 fused paths vs the fp64 generic path (both on
the device; the fp64 path is pinned to the oracle by the parity suite).

    python tools/fuzz_fp32_vs_fp64.py [n_cases] [seed]
"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from ghost_b200 import ContinuousWaveletTransform, Morse

import os
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
only = os.environ.get("FUZZ_ONLY")                              # comma-separated case numbers to run (RNG still advances)
only = set(int(v) for v in only.split(",")) if only else None
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
worst = 0.0
fails = []
for case in range(n_cases):
    gamma = float(rng.choice([1, 2, 3, 3, 3, 4, 6, 9]))
    beta = float(rng.choice([1, 3, 5, 10, 20, 20, 40, 80]))
    fs = float(rng.choice([200.0, 1000.0, 1250.0, 30000.0]))
    n = int(rng.choice([257, 1000, 4097, 20000, 65536, 250001, 1000003]))
    nch = int(rng.choice([1, 1, 2, 5]))
    vpo = int(rng.choice([4, 8, 10, 16]))
    output = str(rng.choice(["amplitude", "power", "complex"]))
    kind = str(rng.choice(["walk", "walk", "tilt", "tilt", "highpass", "tones"])) if os.environ.get("FUZZ_HOSTILE") else "walk"
    if kind == "walk":
        x = rng.standard_normal((nch, n)).astype(np.float32).cumsum(axis=1) * 0.05 + rng.standard_normal((nch, n)).astype(np.float32) \
            + float(rng.uniform(-3, 3))
    else:
        spec = np.fft.rfft(rng.standard_normal((nch, n)), axis=1)
        fr = np.fft.rfftfreq(n, 1.0 / fs)
        if kind == "tilt":
            g = (np.maximum(fr, fr[1] if n > 2 else 1.0) / fr[-1]) ** float(rng.uniform(-1.5, 4.0))
        elif kind == "highpass":
            g = (fr >= float(rng.uniform(0.01, 0.4)) * fs / 2).astype(float)
        else:
            g = np.full_like(fr, float(10 ** rng.uniform(-4, -1)))
        x = np.fft.irfft(spec * g, n=n, axis=1)
        if kind == "tones":
            t = np.arange(n) / fs
            x = x + np.sin(2 * np.pi * float(rng.uniform(0.05, 0.45)) * fs * t) \
                + float(10 ** rng.uniform(-4, -1)) * np.sin(2 * np.pi * float(rng.uniform(0.001, 0.02)) * fs * t)
        x = (x / max(x.std(), 1e-30) + float(rng.uniform(-3, 3))).astype(np.float32)
    ts = None
    if n >= 4097 and rng.random() < 0.4:                       # a gap -> two epochs
        ts = np.arange(n) / fs
        ts[int(n * rng.uniform(0.3, 0.7)):] += 10.0 / fs
    kw = dict(fs=fs, voices_per_octave=vpo, multichannel=True)
    if ts is not None:
        kw["timestamps"] = ts
    if rng.random() < 0.5:
        kw["freq_limits"] = [fs / 2000.0, fs / 2.5]
    if only is not None and case not in only:
        continue
    res = {}
    lev = np.zeros(0, dtype=np.int32)
    try:
        for dt in (np.float64, np.float32):
            cwt = ContinuousWaveletTransform(wavelet=Morse(gamma=gamma, beta=beta), dtype=dt, output=output)
            cwt.transform(x, **kw)
            res[dt] = cwt.coefficients if output == "complex" else (cwt.amplitude if output == "amplitude" else cwt.power)
            if cwt.last_plan is not None:
                lev = cwt.last_plan.levels()
    except Exception as e:                                      # noqa: BLE001
        fails.append((case, gamma, beta, fs, n, nch, vpo, output, repr(e)))
        print("case %d EXC %r" % (case, e), flush=True)
        continue
    a, b = res[np.float32].astype(res[np.float64].dtype), res[np.float64]
    if a.shape[1] == 0:
        print("case %d: no scales (n=%d)" % (case, n)); continue
    num = np.linalg.norm((a - b).reshape(a.shape[0], a.shape[1], -1), axis=2)
    den = np.linalg.norm(b.reshape(b.shape[0], b.shape[1], -1), axis=2)
    rel = num / den
    err = float(rel.max())
    ci, si = np.unravel_index(int(np.argmax(rel)), rel.shape)
    bar = 2e-5 if output == "power" else 1e-5
    worst = max(worst, err)
    tag = "ok " if err <= bar else "BAD"
    if err > bar:
        fails.append((case, gamma, beta, fs, n, nch, vpo, output, err))
    rer = cwt.last_plan.guard_stats()["last"] if cwt.last_plan is not None else -1
    print("case %2d %s %-8s g=%g b=%g fs=%g n=%d ch=%d vpo=%d %s epochs=%d S=%d levels=%s err=%.2e recomputed=%d" % (
        case, tag, kind, gamma, beta, fs, n, nch, vpo, output, 2 if ts is not None else 1, a.shape[1],
        sorted(set(lev.tolist())), err, rer), flush=True)
    if err > 0.5 * bar:
        order = np.argsort(rel.max(axis=0))[::-1][:6]
        print("      worst scales (index, level, L, err): " + ", ".join(
            "(%d, %d, %.2e)" % (int(i), int(lev[i]) if len(lev) else 99, float(rel[:, i].max())) for i in order), flush=True)
print("worst %.2e, failures %d" % (worst, len(fails)))
for f in fails:
    print("FAIL", f)
sys.exit(1 if fails else 0)
