#!/usr/bin/env python
"""Summarise an ncu launch list (gpu__time_duration.sum + dram bytes) into profiles/: one full denoising step.

    python tools/launch_summary.py gpurun_out/launches.csv profiles/r01_launches_step_final.csv profiles/r01_traffic.json
"""
import collections
import csv
import json
import re
import sys


def main():
    src, out_csv, out_json = sys.argv[1:4]
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    byid = collections.OrderedDict()
    for r in rows:
        d = byid.setdefault(r["ID"], {"name": r["Kernel Name"]})
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            d["us"] = v / 1e3 if u.startswith("n") else (v if u.startswith("u") else v * 1e3)
        else:
            d[r["Metric Name"]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    L = list(byid.values())
    starts = [i for i, d in enumerate(L) if "timestep_embedding" in d["name"]]
    if len(starts) < 2:
        raise SystemExit(f"need two step starts in the capture window, found {starts} in {len(L)} launches")
    step = L[starts[-2]:starts[-1]]          # the last complete step of the capture: a CUDA-graph replay, not the
                                             # eager recording pass (which also packs weights once)

    def short(n):
        n = n.replace("void ", "").replace("afldm::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        return re.sub(r"\(.*", "", n)[:60]

    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for d in step:
        a = agg[short(d["name"])]
        a[0] += 1
        a[1] += d["us"]
        a[2] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    tot = sum(v[1] for v in agg.values())
    out = ["kernel,launches,total_us,share,dram_bytes_total,dram_bytes_per_launch"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:50s} {v[0]:4d} {v[1]:9.1f} us {v[1] / tot * 100:5.1f}%  dram {v[2] / 1e6:9.1f} MB")
        out.append(f"{k},{v[0]},{v[1]:.1f},{v[1] / tot:.4f},{v[2]:.0f},{v[2] / v[0]:.0f}")
    out.append(f"# one denoising step (B=16, TF32 class) = {len(step)} launches, {tot:.1f} us summed; ncu --metrics "
               "gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
               "(serialised, cold cache: compare shares, not absolutes)")
    open(out_csv, "w").write("\n".join(out) + "\n")
    conv = [v for k, v in agg.items() if k.startswith("conv_tc_kernel") or k.startswith("splitk_reduce")]
    n_main = sum(v[0] for k, v in agg.items() if k.startswith("conv_tc_kernel"))
    fact = [v for k, v in agg.items() if ("resample" in k and k.rstrip(">").endswith(", 0, 1")) or k.startswith("fact_mma_kernel")]
    json.dump({
        "source": f"{out_csv} (ncu dram__bytes_read.sum + dram__bytes_write.sum, one step)",
        "conv2d_tf32": {"launches": n_main, "dram_bytes_per_step": sum(v[2] for v in conv),
                        "dram_bytes_per_launch": sum(v[2] for v in conv) / max(n_main, 1)},
        "filtered_act": {"launches": sum(v[0] for v in fact), "dram_bytes_per_step": sum(v[2] for v in fact),
                         "dram_bytes_per_launch": sum(v[2] for v in fact) / max(sum(v[0] for v in fact), 1)},
    }, open(out_json, "w"), indent=1)
    print(f"{len(step)} launches, {tot:.1f} us")


if __name__ == "__main__":
    main()
