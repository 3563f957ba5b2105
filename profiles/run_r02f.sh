#!/bin/bash
# GPU call: deferred offsets-free host submits + the e2e leg without shipped offsets.
TAG=${1:-r02s2e}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value %.4g ms/step %.3f kernel %.4f ms frac %.3f"%(d["value"],d["ms_per_step"],d["roofline"]["avg_launch_ms"],d["roofline"]["frac"]))
e=d["e2e"]; print("e2e %.4g (%s) ms/step %.2f h2d %d"%(e["value"],e.get("input"),e["ms_per_step"],e["h2d_bytes_per_step"]))
o=e.get("other_input")
if o: print("other e2e %.4g (%s) ms/step %.2f"%(o["value"],o.get("input"),o["ms_per_step"]))
print("cpu", d["cpu_baseline"]["value"], "framed_stream", d.get("framed_stream",{}).get("ms_per_slab"))
PY
