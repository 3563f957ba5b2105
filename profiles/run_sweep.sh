#!/bin/bash
# Runs on the GPU box: bench.py (value leg) under a list of environment settings.  Usage: run_sweep.sh <tag> "VAR=a VAR=b ..."
TAG=${1:-sweep}; shift
mkdir -p gpurun_out
for setting in "$@"; do
  name=$(echo "$setting" | tr ' =/' '___')
  env $setting timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  env $setting timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --first 300000000 > gpurun_out/${TAG}_${name}_first300m.json 2>> gpurun_out/${TAG}_${name}.err
  python - <<PY
import json
for suf in ("", "_first300m"):
    try:
        d=json.load(open("gpurun_out/${TAG}_${name}%s.json" % suf))
        print("${setting}%s" % suf, "kernel ms %.4f"%d["roofline"]["avg_launch_ms"], "frac %.3f"%d["roofline"]["frac"], "ms/step %.3f"%d["ms_per_step"])
    except Exception as e:
        print("${setting}%s" % suf, "failed", e)
PY
done
