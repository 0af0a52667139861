"""Turn the scratch captures of one gpurun call (tools/gpu_final_r2.sh) into the committed round summary
under profiles/.

    python tools/make_profiles.py r2a
expects gpurun_out/TAG/{launches_cfg2_8ch.csv, prof_fused.ncu-rep, bench_cfg2.json[, bench_cfg3_shard.json,
bench_f64.json, prof_f64.ncu-rep, bench_reference.json, bench_cfg2_noguard.json]}.
"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

tag = sys.argv[1]
G, P = "gpurun_out/%s/" % tag, "profiles/"

# ---- launch list shares -------------------------------------------------------------------------
rows = [r for r in csv.reader(open(G + "launches_cfg2_8ch.csv")) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki].split("(")[0][:48]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[ui], 1.0)
tot = sum(a[1] for a in agg.values())
d = json.load(open(G + "bench_cfg2.json"))
fam, ms = d["roofline"]["families"], d["roofline"].get("ms_per_step_with_family_spans", d["ms_per_step"])
out = ["# Round %s - launch list and live shares, config 2" % tag, "",
       "`%s_launches_cfg2_8ch.csv`: `ncu --metrics gpu__time_duration.sum --clock-control none` over 3 passes with 8" % tag,
       "channels (cold-cache, serialised under the profiler: compare SHARES, not absolutes).", "",
       "| kernel | launches | total ms | share |", "|---|---|---|---|"]
for k, (n, t) in agg.items():
    out.append("| `%s` | %d | %.3f | %.1f %% |" % (k, n, t / 1e6, t / tot * 100))
out += ["", "Live shares in the bench run itself (`%s_bench_cfg2.json`: CUDA events around every launch group, 64 channels;" % tag,
        "the families are timed in a second pass of the same K steps with the library's event spans on, %.3f ms/step;" % ms,
        "the headline %.3f ms/step is the pass without them):" % d["ms_per_step"], "",
        "| family | scales | ms/step | share of step | achieved GB/s | of measured HBM peak |", "|---|---|---|---|---|---|"]
for k, v in fam.items():
    out.append("| %s | %d | %.3f | %.1f %% | %.0f | %.3f |" % (k, v["scales"], v["ms_per_step"], v["ms_per_step"] / ms * 100,
                                                               v["achieved_gbs"], v["frac"]))
mp = d["roofline"]["mean_pyramid_ms_per_step"]
out.append("| mean + pyramid | - | %.3f | %.1f %% | - | - |" % (mp, mp / ms * 100))
e2e, cpu = d["e2e"], d["cpu_baseline"]
out += ["", "Step %.3f ms = %.3e coeff/s; whole step %.0f GB/s = %.3f of the measured %.1f GB/s (%.3f of the nominal 8 TB/s)." % (
    d["ms_per_step"], d["value"], d["roofline"]["whole_step"]["achieved"], d["roofline"]["whole_step"]["frac"], d["roofline"]["peak"],
    d["roofline"]["whole_step"]["frac_of_nominal"]),
    "End to end through `%s`: %.3e coeff/s (%.1f ms/step, %.1f GB/s over the link = %.3f of the %.1f GB/s this box's plain pinned D2H copies reach)." % (
    e2e["api"].split(" ->")[0], e2e["value"], e2e["ms_per_step"], e2e["link_gbs"], e2e["frac_of_d2h_ceiling"], e2e["d2h_ceiling_gbs"]),
    "CPU arm: %s, %.3e coeff/s on %d cores (%s)." % (cpu["kind"], cpu["value"], cpu["cores"], cpu["sample"]),
    "Guard: %s.  Clocks %s." % (json.dumps(d["guard"]), json.dumps(d["clocks"]))]
if os.path.exists(G + "bench_cfg2_noguard.json"):
    ng = json.load(open(G + "bench_cfg2_noguard.json"))
    out.append("Same run without the accuracy guard (`--no-guard`): %.3f ms/step, %.3e coeff/s: the guard costs %.1f %%." % (
        ng["ms_per_step"], ng["value"], (d["ms_per_step"] / ng["ms_per_step"] - 1) * 100))
    shutil.copy(G + "bench_cfg2_noguard.json", P + "%s_bench_cfg2_noguard.json" % tag)
open(P + "%s_launch_shares.md" % tag, "w").write("\n".join(out) + "\n")
shutil.copy(G + "launches_cfg2_8ch.csv", P + "%s_launches_cfg2_8ch.csv" % tag)
shutil.copy(G + "bench_cfg2.json", P + "%s_bench_cfg2.json" % tag)
for name in ("bench_cfg3_shard.json", "bench_f64.json", "bench_reference.json"):
    if os.path.exists(G + name):
        shutil.copy(G + name, P + "%s_%s" % (tag, name))

# ---- ncu summaries + measured traffic -----------------------------------------------------------------
subprocess.run([sys.executable, "tools/ncu_summary.py", G + "prof_fused.ncu-rep", P + "%s_fused_kernels_ncu.md" % tag,
                "Round %s - mean, pyramid, fused kernels and guard verdict (one launch per scale class), config 2 with 16 channels" % tag],
               check=True)
if os.path.exists(G + "prof_f64.ncu-rep"):
    subprocess.run([sys.executable, "tools/ncu_summary.py", G + "prof_f64.ncu-rep", P + "%s_f64_kernels_ncu.md" % tag,
                    "Round %s - fp64 four-step path (config 5: 8 channels x 262144 samples, complex128 out)" % tag], check=True)
raw = subprocess.run(["ncu", "-i", G + "prof_fused.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
famb = {}
for r in data:
    if "fused_" not in r[ki]:
        continue
    k = "fused_full" if "fused_full" in r[ki] else ("fused_banded" if "fused_banded" in r[ki] else "fused_interp")
    famb.setdefault(k, []).append(float(r[ri].replace(",", "")) * mult[units[ri]] + float(r[wi].replace(",", "")) * mult[units[wi]])
ch, n = 16, 2250000
tj = {"source": "profiles/%s_fused_kernels_ncu.md (prof_fused.ncu-rep of gpu_final_r2.sh, config 2 with 16 channels, one launch per class; "
                "dram__bytes_read.sum + dram__bytes_write.sum)" % tag,
      "channels_in_capture": ch}
for k, v in famb.items():
    alg = ch * n * (fam[k]["scales"] * 4 + 4)
    tj[k + "_dram_bytes_per_step_16ch"] = sum(v)
    tj[k + "_algorithmic_bytes_per_step_16ch"] = alg
    tj[k + "_traffic_over_algorithmic"] = sum(v) / alg
    tj[k + "_bytes_per_launch"] = sum(v) / len(v)
json.dump(tj, open(P + "traffic.json", "w"), indent=1)
print("\n".join(out[-14:]))
print({k: round(v, 3) for k, v in tj.items() if k.endswith("over_algorithmic")})
