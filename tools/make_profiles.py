"""Turn the scratch captures of one gpurun call into the committed round summary under profiles/.

    python tools/make_profiles.py r1f
expects gpurun_out/{launches_TAG.csv, prof_fused_TAG.ncu-rep, bench_TAG.json[, bench_TAG_cfg3.json]}.
"""
import collections
import csv
import io
import json
import shutil
import subprocess
import sys

tag = sys.argv[1]
G, P = "gpurun_out/", "profiles/"

# ---- launch list shares -------------------------------------------------------------------------
rows = [r for r in csv.reader(open(G + "launches_%s.csv" % tag)) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki].split("(")[0][:48]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
tot = sum(a[1] for a in agg.values())
d = json.load(open(G + "bench_%s.json" % tag))
fam, ms = d["roofline"]["families"], d["ms_per_step"]
out = ["# Round %s - launch list and live shares, config 2" % tag, "",
       "`%s_launches_cfg2_8ch.csv`: `ncu --metrics gpu__time_duration.sum --clock-control none` over 3 passes with 8",
       "channels (cold-cache, serialised under the profiler: compare SHARES, not absolutes).", "",
       "| kernel | launches | total ms | share |", "|---|---|---|---|"]
out[2] = out[2] % tag
for k, (n, t) in agg.items():
    out.append("| `%s` | %d | %.3f | %.1f %% |" % (k, n, t / 1e6, t / tot * 100))
out += ["", "Live shares in the bench run itself (`%s_bench_cfg2.json`: CUDA events around every launch group, 64 channels):" % tag, "",
        "| family | scales | ms/step | share of step | achieved GB/s | of measured HBM peak |", "|---|---|---|---|---|---|"]
for k, v in fam.items():
    out.append("| %s | %d | %.3f | %.1f %% | %.0f | %.3f |" % (k, v["scales"], v["ms_per_step"], v["ms_per_step"] / ms * 100,
                                                               v["achieved_gbs"], v["frac"]))
mp = d["roofline"]["mean_pyramid_ms_per_step"]
out.append("| mean + pyramid | - | %.3f | %.1f %% | - | - |" % (mp, mp / ms * 100))
out += ["", "Step %.3f ms = %.3e coeff/s; whole step %.0f GB/s = %.3f of the measured %.1f GB/s; e2e %.3e coeff/s; CPU port %.3e coeff/s on %d cores; clocks %s." % (
    ms, d["value"], d["roofline"]["whole_step"]["achieved"], d["roofline"]["whole_step"]["frac"], d["roofline"]["peak"],
    d["e2e"]["value"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], json.dumps(d["clocks"]))]
open(P + "%s_launch_shares.md" % tag, "w").write("\n".join(out) + "\n")
shutil.copy(G + "launches_%s.csv" % tag, P + "%s_launches_cfg2_8ch.csv" % tag)
shutil.copy(G + "bench_%s.json" % tag, P + "%s_bench_cfg2.json" % tag)
try:
    shutil.copy(G + "bench_%s_cfg3.json" % tag, P + "%s_bench_cfg3_shard.json" % tag)
except FileNotFoundError:
    pass

# ---- ncu summary + measured traffic -----------------------------------------------------------------
subprocess.run([sys.executable, "tools/ncu_summary.py", G + "prof_fused_%s.ncu-rep" % tag, P + "%s_fused_kernels_ncu.md" % tag,
                "Round %s - fused kernels (one launch per scale class), config 2 with 16 channels" % tag], check=True)
raw = subprocess.run(["ncu", "-i", G + "prof_fused_%s.ncu-rep" % tag, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
famb = {}
for r in data:
    k = "fused_full" if "fused_full" in r[ki] else ("fused_banded" if "fused_banded" in r[ki] else "fused_interp")
    famb.setdefault(k, []).append(float(r[ri].replace(",", "")) * mult[units[ri]] + float(r[wi].replace(",", "")) * mult[units[wi]])
ch, n = 16, 2250000
tj = {"source": "profiles/%s_fused_kernels_ncu.md (prof_fused_%s.ncu-rep, config 2 with 16 channels, one launch per class)" % (tag, tag),
      "channels_in_capture": ch}
for k, v in famb.items():
    alg = ch * n * (fam[k]["scales"] * 4 + 4)
    tj[k + "_dram_bytes_per_step_16ch"] = sum(v)
    tj[k + "_algorithmic_bytes_per_step_16ch"] = alg
    tj[k + "_traffic_over_algorithmic"] = sum(v) / alg
    tj[k + "_bytes_per_launch"] = sum(v) / len(v)
json.dump(tj, open(P + "traffic.json", "w"), indent=1)
print("\n".join(out[-12:]))
print({k: round(v, 3) for k, v in tj.items() if k.endswith("over_algorithmic")})
