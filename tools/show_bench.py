"""Print the per-family summary of bench JSON lines: python tools/show_bench.py gpurun_out/<tag>/bench_*.json"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f)); r = d["roofline"]
    except Exception as e:  # noqa: BLE001
        print(f, "unreadable:", e); continue
    fam = {k: (round(v["ms_per_step"], 3), round(v["frac"], 3)) for k, v in r["families"].items()}
    print("%-44s %.4g coeff/s %.2f ms %s pyr %.2f whole %.3f" % (f.split("/")[-1], d["value"], d["ms_per_step"], fam,
          r.get("mean_pyramid_ms_per_step", 0), r["whole_step"]["frac"]))
