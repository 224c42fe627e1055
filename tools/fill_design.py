#!/usr/bin/env python
"""Fill the @PLACEHOLDER@ numbers of DESIGN.md / README.md from a bench.py JSON line.

    python tools/fill_design.py profiles/r01_bench_n1_final.json [eager_steps_per_s] [cpu_steps_per_s]
"""
import json
import sys

d = json.load(open(sys.argv[1]))
eager = float(sys.argv[2]) if len(sys.argv) > 2 else 26.7
cpu = float(sys.argv[3]) if len(sys.argv) > 3 else (d.get("cpu_baseline") or {}).get("value", 0.81)
bd = d["breakdown"]
conv = bd["conv2d_tf32"]["ms_per_step"]
att = sum(v["ms_per_step"] for k, v in bd.items() if k.startswith("attention"))
fa = d["roofline_filtered_act"]
vals = {
    "VALUE": f"{d['value']:.1f}", "MS": f"{d['ms_per_step']:.2f}",
    "E2E": f"{d['e2e']['value']:.1f}", "E2EMS": f"{1000.0 / d['e2e']['value']:.2f}",
    "XEAGER": f"{d['value'] / eager:.1f}", "XCPU": f"{d['value'] / cpu:.0f}",
    "LAUNCHES": str(d["launches_per_step"]),
    "VAE": f"{(d.get('vae_decode') or {}).get('images_per_s', float('nan')):.0f}",
    "CONV": f"{conv:.2f}", "CONVTF": f"{d['roofline']['achieved']:.0f}",
    "FACT": f"{fa['ms_per_step']:.2f}", "FACTGB": f"{fa['achieved']:.0f}", "FACTFRAC": f"{100 * fa['frac']:.0f}",
    "ATT": f"{att:.2f}",
}
for path in ("DESIGN.md", "README.md"):
    s = open(path).read()
    for k, v in vals.items():
        s = s.replace(f"@{k}@", v)
    open(path, "w").write(s)
print(vals)
