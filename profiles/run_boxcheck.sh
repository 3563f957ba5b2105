#!/bin/bash
# Runs on the GPU box: bench.py with the box-wide check forced at N=1 (value + e2e + box), then the full GPU test log.
TAG=${1:-box}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --box-check > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value %.2f G/s ms/step %.3f kernel %.4f frac %.3f e2e %.3f" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"]/1e9))
print({k:v for k,v in d.items() if k.startswith("box") or k.startswith("sketch")})
PY
