#!/bin/bash
# Run bench.py under several conv_tc tuning knobs (static getenv: one process each) and keep the per-layer tables.
# usage: tools/bench_knobs.sh "<NAME=VAL NAME=VAL>" "<...>" ...   (empty string = defaults)
i=0
for cfg in "$@"; do
  tag=$(echo "$cfg" | tr ' =' '__'); [ -z "$tag" ] && tag=default
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --dump-breakdown gpurun_out/knob_$tag.csv > gpurun_out/knob_$tag.json 2> gpurun_out/knob_$tag.err
  python - "$tag" <<'P'
import json,sys
d=json.load(open(f"gpurun_out/knob_{sys.argv[1]}.json"))
print(sys.argv[1], round(d["value"],1), "conv", d["breakdown"]["conv2d_tf32"]["ms_per_step"])
P
done
