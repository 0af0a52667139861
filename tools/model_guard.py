"""Numpy prototype of the execute-time accuracy guard (design validation, NOT a product path).

For every scale of a band-limited class the guard compares a bound on the error of the fp32 fused
path with the scale's measured output energy.  This script evaluates the same bound offline from the
calibration runs of tools/guard_calib.py (per-scale errors measured on the B200) so that its constants
can be checked against reality:  python tools/model_guard.py [config] [signal ...]
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from tools.model_fastpath import halfband, multiplier, morse_terms, Pyramid    # noqa: E402
from tools.guard_calib import signals, CONFIGS                                   # noqa: E402

EPS = 2.0 ** -24
KBINS, NCD = 256, 1024


def hb_taps32():
    """The device's taps: fp32, rounded greedily so that the odd taps sum to 0.25 (fast_path.cu::design_halfband)."""
    h = halfband()
    T = (len(h) - 1) // 2
    odd = [float(h[T + t]) for t in range(1, T + 1, 2)]
    done = 0.0
    for k in range(len(odd)):
        f = float(np.float32(np.longdouble(0.25) - np.longdouble(done) - np.longdouble(sum(odd[k + 1:]))))
        odd[k] = f
        done += f
    out = np.zeros_like(h)
    out[T] = 0.5
    for i, t in enumerate(range(1, T + 1, 2)):
        out[T + t] = out[T - t] = odd[i]
    return out


def hb_gain(h, theta):
    T = (len(h) - 1) // 2
    t = np.arange(-T, T + 1)
    return (h[None, :] * np.cos(np.multiply.outer(theta, t).reshape(-1, len(t)))).sum(axis=1).reshape(np.shape(theta))


def alias_table(h, level):
    """B[b][m] = largest pyramid gain with which a component of octave b (|w| in (pi/2^(b+1), pi/2^b])
    reaches bin m of the level's 1024-point grid.  g(k, m) = prod_i |hb(2 pi (k mod 2^i)/2^i + theta_m/2^i)|."""
    D = 1 << level
    theta = 2 * np.pi * np.arange(KBINS) / NCD
    g = np.ones((1, KBINS))
    for i in range(1, level + 1):
        r = np.arange(1 << i)
        f = np.abs(hb_gain(h, 2 * np.pi * r[:, None] / (1 << i) + theta[None, :] / (1 << i)))
        g = g[r % (1 << (i - 1))] * f
    k = np.arange(D)
    kk = np.minimum(k, D - k).astype(float)
    B = np.zeros((level + 1, KBINS))
    for b in range(level):
        sel = (kk * 2 / D > 2.0 ** -(b + 1)) & (kk * 2 / D <= 2.0 ** -b) & (k > 0)
        if sel.any():
            B[b] = g[sel].max(axis=0)
    return B


def drop_envelope(L, k0, X, w):
    """Envelope of |H(w)| of the L-tap kernel away from its band: |sum_k (-1)^k X_k / sin((w - w_k)/2)| / L."""
    k = k0 + np.arange(len(X))
    wk = 2 * np.pi * k / L
    sg = np.where(k % 2 == 0, 1.0, -1.0)
    d = w[:, None] - wk[None, :]
    near = np.abs(d) < 2 * np.pi / L                     # within one grid spacing: the term is at most |X_k|
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(near, 0.0, sg[None, :] * X[None, :] / np.sin(d / 2))
    return np.abs(t.sum(axis=1)) / L + (near * X[None, :]).sum(axis=1)


def guard_terms(x, fs, freqs, Ls, levels, gamma=3.0, beta=20.0):
    """Per scale: (band bound^2, rounding bound^2, q) in units of sum_t |W|^2."""
    h = hb_taps32()
    x = np.asarray(x, dtype=np.float64)
    x = x - x.mean()
    J = max(int(levels.max()), 0)
    pyr = Pyramid(x, J + 2, h)                                # two levels deeper than the plan needs: energies only
    e = [float((pyr.data[j] ** 2).sum()) * 2 ** j for j in range(J + 3)]
    out = []
    tabs = {}
    for s, (f, L, lev) in enumerate(zip(freqs, Ls, levels)):
        om = f / (fs / 2) * np.pi
        k0, X = morse_terms(gamma, beta, om, int(L))
        if lev < 0:
            _, H = multiplier(int(L), k0, X, 4096, np.arange(4096))
            q = float((np.abs(H) ** 2).sum()) / 4096
            # chunk-mean-removed energy of the raw signal (4096-sample chunks)
            nb = len(x) // 4096
            xc = x[:nb * 4096].reshape(nb, 4096)
            ech = float(((xc - xc.mean(axis=1, keepdims=True)) ** 2).sum()) * len(x) / (nb * 4096)
            out.append((0.0, q * ech, q))
            continue
        D = 1 << lev
        if lev not in tabs:
            tabs[lev] = alias_table(h, lev)
        B = tabs[lev]
        G, _ = multiplier(int(L), k0, X, NCD * D, np.arange(KBINS))
        wm = 2 * np.pi * np.arange(KBINS) / (NCD * D)
        Hd = np.ones(KBINS)
        for st in range(lev):
            Hd *= hb_gain(h, wm * (1 << st))
        T = np.abs(G / Hd)
        q = float((T ** 2).sum()) / NCD
        band = 0.0
        for b in range(lev + 1):
            Eb = max(e[b] - e[b + 2], 0.0)
            a_alias = float((B[b] * T).max()) ** 2 if b < lev else 0.0
            lo, hi = np.pi / 2 ** (b + 1), np.pi / 2 ** b
            if b == lev:
                lo = np.pi / (2 * D)
            w = np.geomspace(lo, hi, 24)
            if b == lev:      # just above the kept band: the exact response on a 4x finer grid (as the device does)
                Gx, _ = multiplier(int(L), k0, X, 4 * NCD * D, np.arange(4 * KBINS, 8 * KBINS))
                a_drop = max(1.1 * float(np.abs(Gx).max()), 1.15 * float(drop_envelope(int(L), k0, X, -w).max())) ** 2
            else:
                env = np.concatenate([drop_envelope(int(L), k0, X, w), drop_envelope(int(L), k0, X, -w)])
                a_drop = (1.15 * float(env.max())) ** 2
            band += (a_alias + a_drop) * Eb
        # rounding: chunk FFT on the level's signal (chunk mean removed) + storage rounding of the pyramid stages
        xl = pyr.data[lev]
        nb = max(1, len(xl) // NCD)
        xc = xl[:nb * NCD].reshape(nb, -1) if len(xl) >= NCD else xl[None, :]
        ech = float(((xc - xc.mean(axis=1, keepdims=True)) ** 2).sum()) * D * len(xl) / xc.size
        stage = sum(e[j] * 2.0 ** (j - lev) for j in range(1, lev + 1)) / 3.0
        out.append((band, q * ech, q * stage, q))
    return out, e


def main():
    cname = sys.argv[1] if len(sys.argv) > 1 else "a1k"
    cfg = CONFIGS[cname]
    sigs = signals(cfg["n"], cfg["fs"], 11)
    names = sys.argv[2:] or list(sigs)
    for name in names:
        z = np.load("gpurun_out/calib/%s_%s.npz" % (cname, name))
        x = sigs[name].astype(np.float32).astype(np.float64)
        terms, e = guard_terms(x, cfg["fs"], z["freqs"], z["L"], z["levels"])
        print("== %s %s   pyramid energies / E: %s" % (cname, name, " ".join("%.1e" % (v / e[0]) for v in e)))
        for s, t in enumerate(terms):
            lev = int(z["levels"][s])
            P = float(z["p_out"][s])
            err = float(z["err_direct"][s])
            if lev < 0:
                rb = np.sqrt(t[1] / P) * EPS
                print("  s%3d lev %2d err %.2e | round/eps %.2e -> kappa %.1f" % (s, lev, err, rb / EPS, err / rb))
            else:
                bb = np.sqrt(t[0] / P)
                rb = np.sqrt(t[1] / P) * EPS
                sb = np.sqrt(t[2] / P) * EPS
                print("  s%3d lev %2d err %.2e | band %.2e  round %.2e  stage %.2e | err/band %.2f err/round %.1f" % (
                    s, lev, err, bb, rb, sb, err / bb if bb else np.inf, err / rb))


def summary(cname, kr=6.0, ks=2.0, tau=5e-6):
    cfg = CONFIGS[cname]
    sigs = signals(cfg["n"], cfg["fs"], 11)
    for name in sigs:
        z = np.load("gpurun_out/calib/%s_%s.npz" % (cname, name))
        x = sigs[name].astype(np.float32).astype(np.float64)
        terms, e = guard_terms(x, cfg["fs"], z["freqs"], z["L"], z["levels"])
        err = z["err_direct"]
        tot = np.zeros(len(err)); bnd = np.zeros(len(err)); rnd = np.zeros(len(err))
        for s, t in enumerate(terms):
            P = float(z["p_out"][s])
            if z["levels"][s] < 0:
                rnd[s] = kr * EPS * np.sqrt(t[1] / P)
            else:
                bnd[s] = np.sqrt(t[0] / P)
                rnd[s] = EPS * np.sqrt((kr ** 2 * t[1] + ks ** 2 * t[2]) / P)
            tot[s] = np.hypot(bnd[s], rnd[s])
        flag = tot > tau
        bad = err > 1e-5
        print("%s %-12s flagged %3d/%d  truly>1e-5: %3d  missed: %d  max err unflagged %.2e  max err/bound %.2f (band-dominated %.2f, round-dominated %.2f)" % (
            cname, name, flag.sum(), len(err), bad.sum(), int((bad & ~flag).sum()),
            err[~flag].max() if (~flag).any() else 0.0, (err / tot).max(),
            (err / tot)[bnd > rnd].max() if (bnd > rnd).any() else 0, (err / tot)[bnd <= rnd].max() if (bnd <= rnd).any() else 0))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "summary":
        for c in sys.argv[2:] or ["a1k", "b30k"]:
            summary(c)
    else:
        main()
