"""profiles/<tag>_guard_calibration.md from two runs of tools/guard_calib.py on the B200 (guard=0 and default):
    python tools/make_guard_table.py <log without guard> <log with guard> <tag>"""
import re
import sys

pat = re.compile(r"^(\w+) (\w+)\s+rerouted\s+(\d+)/(\d+) worst interp (\S+) direct (\S+)\s+scales>1e-5: (\d+)\s+\(worst scale (\d+) level (-?\d+), out/in rms (\S+)\)")


def rows(path):
    out = {}
    for line in open(path):
        m = pat.match(line.strip())
        if m:
            out[(m.group(1), m.group(2))] = m.groups()
    return out


off, on, tag = rows(sys.argv[1]), rows(sys.argv[2]), sys.argv[3]
lines = ["# Round %s - accuracy guard on hostile and benign spectra (B200)" % tag, "",
         "`tools/guard_calib.py`: fp32 amplitude of the fused paths against the fp64 device path (itself pinned to the oracle at 1e-10),",
         "per scale relative L2, worst scale reported.  a1k: fs 1 kHz, 120 000 samples, default grid (94 scales); b30k: fs 30 kHz,",
         "1.2 M samples, the 128-scale grid of configs 3 / 4.  Signals: white; chirp + pink (the benchmark's); random walk + white +",
         "offset; violet (amplitude ~ f^2); f^4; white noise high-passed at 0.2 and 0.01 Nyquist; a weak tone (1e-4 / 1e-2) beside a",
         "strong out-of-band one.  `re-computed` = (channel, scale) pairs the guard sent to the fp64 path.", "",
         "| config | signal | guard off: worst rel. L2 | scales > 1e-5 | weakest output / input rms | guard on: worst rel. L2 | scales > 1e-5 | re-computed |",
         "|---|---|---|---|---|---|---|---|"]
for key in off:
    o, g = off[key], on.get(key)
    if g is None:
        continue
    lines.append("| %s | %s | %s | %s | %s | %s | %s | %s / %s |" % (key[0], key[1], o[4], o[6], o[9], g[4], g[6], g[2], g[3]))
open("profiles/%s_guard_calibration.md" % tag, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
