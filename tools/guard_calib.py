"""GPU box: per-scale error of the fp32 fused paths against the fp64 device path on inputs whose
spectra are hostile to a band-limited method (blue / high-passed / strong out-of-band tones), next
to the benign ones.  Writes gpurun_out/calib/<config>_<signal>.npz with everything needed to
calibrate the execute-time accuracy guard offline (per-scale error, output energy, levels).

    python tools/guard_calib.py [extra ContinuousWaveletTransform kwargs as k=v ...]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from ghost_b200 import Morse                                            # noqa: E402
from ghost_b200.engine import CwtPlan, scale_tables                      # noqa: E402
from ghost_b200 import ContinuousWaveletTransform, synth                 # noqa: E402


def shaped_noise(rng, n, fs, power, f_hp=None):
    """White noise with amplitude spectrum ~ f**power, optionally brick-wall high-passed at f_hp."""
    spec = np.fft.rfft(rng.standard_normal(n))
    f = np.fft.rfftfreq(n, 1.0 / fs)
    g = np.ones_like(f)
    if power:
        g = (f / f[-1]) ** power
    if f_hp:
        g = g * (f >= f_hp)
    x = np.fft.irfft(spec * g, n=n)
    return x / x.std()


def signals(n, fs, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    out = {
        "white": rng.standard_normal(n),
        "chirp_pink": synth.chirp_pink(n, fs, seed, np.float64),
        "randwalk": np.cumsum(rng.standard_normal(n)) * 0.05 + rng.standard_normal(n) + 2.5,
        "violet": shaped_noise(rng, n, fs, 2.0),
        "f4": shaped_noise(rng, n, fs, 4.0),
        "white_hp02": shaped_noise(rng, n, fs, 0.0, f_hp=0.2 * fs / 2),
        "white_hp001": shaped_noise(rng, n, fs, 0.0, f_hp=0.01 * fs / 2),
        "tones": 1e-4 * np.sin(2 * np.pi * (0.003 * fs) * t) + np.sin(2 * np.pi * (0.1234 * fs) * t)
                 + 1e-4 * rng.standard_normal(n),
        "tones_mild": 1e-2 * np.sin(2 * np.pi * (0.003 * fs) * t) + np.sin(2 * np.pi * (0.1234 * fs) * t)
                      + 1e-2 * rng.standard_normal(n),
    }
    return out


CONFIGS = {
    "a1k": dict(fs=1000.0, n=120000, freq_limits=None),
    "b30k": dict(fs=30000.0, n=1200000, freq_limits=[1.7, 15000.0]),
}


def main():
    os.makedirs("gpurun_out/calib", exist_ok=True)
    extra = dict(kv.split("=") for kv in sys.argv[1:])
    for cname, cfg in CONFIGS.items():
        fs, n = cfg["fs"], cfg["n"]
        cwt = ContinuousWaveletTransform(dtype=np.float32)
        cwt.fs = fs
        cwt.wavelet.fs = fs
        f = np.asarray(cwt.plan_frequencies(n, freq_limits=cfg["freq_limits"], voices_per_octave=10))
        m = Morse(fs=fs)
        om = f / (fs / 2.0) * np.pi
        L = m.compute_lengths(om)
        k0, nt, terms = scale_tables(m, om, L)
        kw = {}
        if "band_tol" in extra:
            kw["band_tol"] = float(extra["band_tol"])
        if "guard" in extra:
            kw["guard"] = bool(int(extra["guard"]))
        if "guard_tol" in extra:
            kw["guard_tol"] = float(extra["guard_tol"])
        p32 = CwtPlan(L, k0, nt, terms, dtype=np.float32, output="amplitude", **kw)
        p32n = CwtPlan(L, k0, nt, terms, dtype=np.float32, output="amplitude", no_interp=True, **kw)
        p64 = CwtPlan(L, k0, nt, terms, dtype=np.float64, output="amplitude")
        lev = p32.levels()
        for sname, x in signals(n, fs, 11).items():
            xd64 = torch.from_numpy(np.ascontiguousarray(x[None, :], dtype=np.float64)).cuda()
            xd32 = xd64.float()
            a64 = p64.execute(xd32)[0]
            res = {}
            for tag, plan in (("interp", p32), ("direct", p32n)):
                a32 = plan.execute(xd32)[0]
                res[tag + "_rerouted"] = plan.guard_stats()["last"]
                num = torch.linalg.vector_norm(a32.double() - a64, dim=1)
                den = torch.linalg.vector_norm(a64, dim=1)
                res[tag] = (num / den).cpu().numpy()
                del a32
            p_out = (a64 * a64).sum(dim=1).cpu().numpy()
            xm = xd32[0].double() - xd32[0].double().mean()
            e_in = float((xm * xm).sum())
            np.savez("gpurun_out/calib/%s_%s.npz" % (cname, sname), err_interp=res["interp"], err_direct=res["direct"],
                     p_out=p_out, e_in=e_in, levels=lev, freqs=f, L=L, fs=fs, n=n)
            bad = int((res["interp"] > 1e-5).sum())
            print("%s %-12s rerouted %3d/%d worst interp %.2e direct %.2e  scales>1e-5: %d  (worst scale %d level %d, out/in rms %.1e)" % (
                cname, sname, res["interp_rerouted"], len(f), res["interp"].max(), res["direct"].max(), bad, int(np.argmax(res["interp"])),
                int(lev[int(np.argmax(res["interp"]))]),
                float(np.sqrt(p_out[int(np.argmax(res["interp"]))] / e_in))), flush=True)
            del a64, xd64, xd32
        del p32, p32n, p64
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
