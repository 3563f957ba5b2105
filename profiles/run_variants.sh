#!/bin/bash
# Runs on the GPU box: bench.py (value leg only) against the default library and every build under _variants/.
TAG=${1:-var}
mkdir -p gpurun_out
for lib in default flow-pipeline_b200/_variants/*.so; do
  name=$(basename $lib .so)
  if [ "$lib" = "default" ]; then unset FLOWAGG_LIB; else export FLOWAGG_LIB=$PWD/$lib; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_${name}.json"))
    print("${name}", "kernel ms %.4f"%d["roofline"]["avg_launch_ms"], "frac %.3f"%d["roofline"]["frac"], "ms/step %.3f"%d["ms_per_step"])
except Exception as e:
    print("${name}", "failed", e)
PY
done
