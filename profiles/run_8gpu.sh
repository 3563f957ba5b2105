#!/bin/bash
# Runs on an 8-GPU box: bench.py at N=8 (value, e2e, verified box-wide queries) and BASELINE configs[3] at full size.
TAG=${1:-r02_8gpu}; N=${2:-8}; TOTAL=${3:-10000000000}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
BENCH_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu \
    > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
grep "bench debug" gpurun_out/${TAG}_bench.err | cut -c1-260 | head -40
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${TAG}_bench.json") if l.startswith("{")][-1])
print("N=%d value %.2f G/s ms/step %.3f e2e %.3f G/s (%.1f ms/step)" % (d["n_gpus"], d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"]))
print({k:v for k,v in d.items() if k.startswith("box") or k.startswith("sketch")})
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 profiles/run_config3.py $TOTAL \
    > gpurun_out/${TAG}_config3.json 2> gpurun_out/${TAG}_config3.err
tail -2 gpurun_out/${TAG}_config3.err | cut -c1-300
cat gpurun_out/${TAG}_config3.json | cut -c1-1500
