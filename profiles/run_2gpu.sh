#!/bin/bash
# Runs on a 2-GPU box: the whole GPU suite (nothing skipped) and bench.py at N=2 with the verified box-wide queries.
TAG=${1:-r02_2gpu}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv > gpurun_out/${TAG}_gpus.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -rs > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -6 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 \
    > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${TAG}_bench.json") if l.startswith("{")][-1])
print("N=%d value %.2f G/s ms/step %.3f e2e %.3f G/s" % (d["n_gpus"], d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9))
print({k:v for k,v in d.items() if k.startswith("box") or k.startswith("sketch")})
PY
