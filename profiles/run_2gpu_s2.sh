#!/bin/bash
# Second session, 2-GPU box: the whole GPU suite (nothing skipped), bench.py at N=2 (both e2e inputs, verified box-wide queries),
# and a memcheck pass over the tests of the new table layout and of the deferred offsets-free submit.
TAG=${1:-r02s2_2gpu}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv > gpurun_out/${TAG}_gpus.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -rs > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 \
    > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${TAG}_bench.json") if l.startswith("{")][-1])
e=d["e2e"]
print("N=%d value %.2f G/s ms/step %.3f e2e %.3f G/s (%s) other %.3f" % (d["n_gpus"], d["value"]/1e9, d["ms_per_step"], e["value"]/1e9, e.get("input"), e.get("other_input",{}).get("value",0)/1e9))
print({k:v for k,v in d.items() if k.startswith("box") or k.startswith("sketch")})
PY
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -x -q -k "wide_key or table_full or one_behind" > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/${TAG}_memcheck.log
