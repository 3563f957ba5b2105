#!/bin/bash
# GPU call: framing kernels (speculation from shared memory, L2 prefetch along the walks).
TAG=${1:-r02s2f}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "null_offsets or one_behind or host_inserter or framing or readme or mocker_10k" > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -2 gpurun_out/${TAG}_pytest_gpu.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
e=d["e2e"]; o=e.get("other_input",{})
print("value %.4g e2e %.4g (%s; %.2f ms) other %.4g framed_stream %.3f ms/slab vs %.3f"%(d["value"],e["value"],e.get("input"),e["ms_per_step"],o.get("value",0),d["framed_stream"]["ms_per_slab"],d["framed_stream"]["with_offsets_ms_per_slab"]))
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_framed_launches.csv python profiles/frame_probe.py > gpurun_out/${TAG}_framed_launches.log 2>&1
grep -E "k_frame|k_tile" gpurun_out/${TAG}_framed_launches.csv | awk -F'","' '{print $5, $NF}' | tail -12
