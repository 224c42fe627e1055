#!/bin/bash
# Run bench.py under several conv_tc tuning knobs (static getenv: one process each) and keep the per-layer tables.
# usage: tools/bench_knobs.sh "<NAME=VAL NAME=VAL>" "<...>" ...   (empty string = defaults)
i=0
for cfg in "$@"; do
  tag=$(echo "$cfg" | tr ' =' '__'); [ -z "$tag" ] && tag=default
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-vae --dump-breakdown gpurun_out/knob_$tag.csv > gpurun_out/knob_$tag.json 2> gpurun_out/knob_$tag.err
  python - "$tag" <<'P'
import json,sys
d=json.load(open(f"gpurun_out/knob_{sys.argv[1]}.json"))
b=d["breakdown"]
print(sys.argv[1], round(d["value"],1), "conv tf32", b.get("conv2d_tf32",{}).get("ms_per_step"), "f16", b.get("conv2d_f16",{}).get("ms_per_step"))
P
done
