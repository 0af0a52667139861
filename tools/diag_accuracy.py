"""Per-scale fp32 error of the device path against the oracle (GPU box diagnostic)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from ghost_b200 import Morse, synth
from ghost_b200.engine import CwtPlan, scale_tables
from oracle import cwt_oracle as orc


def run(fs, n, gamma=3, beta=20, **kw):
    x = synth.chirp_pink(n, fs, 0, np.float32)
    W, f, L = orc.cwt_complex(x, fs, gamma=gamma, beta=beta, parallel=True)
    amp = np.abs(W)
    m = Morse(gamma=gamma, beta=beta, fs=fs)
    om = f / (fs / 2.0) * np.pi
    k0, nt, terms = scale_tables(m, om, L)
    xd = torch.from_numpy(x[None, :]).cuda()
    for name, opts in (("interp", {}), ("no_interp", {"no_interp": True})):
        plan = CwtPlan(L, k0, nt, terms, dtype=np.float32, output="amplitude", **opts)
        got = plan.execute(xd)[0].cpu().numpy().astype(np.float64)
        l2 = np.linalg.norm(got - amp, axis=1) / np.linalg.norm(amp, axis=1)
        mx = np.max(np.abs(got - amp), axis=1) / np.max(amp, axis=1)
        lev = plan.levels()
        print(name, "gamma", gamma, "beta", beta, "n", n)
        for lv in sorted(set(lev.tolist())):
            sel = lev == lv
            print("  level %2d: %2d scales  relL2 max %.2e  maxnorm max %.2e" % (lv, sel.sum(), l2[sel].max(), mx[sel].max()))


if __name__ == "__main__":
    run(1000.0, 120000)
    run(2000.0, 60000, gamma=1, beta=1)
    run(2000.0, 60000, gamma=9, beta=3)
