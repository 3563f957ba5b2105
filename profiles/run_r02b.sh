#!/bin/bash
# Round-2 second session, GPU call 1: wide-key table redesign (configs[4]), sample-and-hold admission (configs[2]), slot-hash variant.
TAG=${1:-r02s2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 400 python profiles/run_configs.py "configs[2]" "configs[4]" > gpurun_out/${TAG}_configs.json 2> gpurun_out/${TAG}_configs.err
tail -3 gpurun_out/${TAG}_configs.err | cut -c1-900
FLOWAGG_LIB=$PWD/flow-pipeline_b200/_variants/admit_estimate.so timeout 300 python profiles/run_configs.py "bounded" > gpurun_out/${TAG}_configs_admit_estimate.json 2> gpurun_out/${TAG}_configs_admit_estimate.err
tail -1 gpurun_out/${TAG}_configs_admit_estimate.err | cut -c1-600
for v in default cheap_hash; do
  if [ "$v" = "default" ]; then unset FLOWAGG_LIB; else export FLOWAGG_LIB=$PWD/flow-pipeline_b200/_variants/$v.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_bench_${v}.json 2> gpurun_out/${TAG}_bench_${v}.err
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --first 300000000 > gpurun_out/${TAG}_bench300m_${v}.json 2> gpurun_out/${TAG}_bench300m_${v}.err
  python - <<PY
import json
for f in ("bench","bench300m"):
    try:
        d=json.load(open("gpurun_out/${TAG}_%s_${v}.json"%f))
        print("${v}", f, "kernel ms %.4f"%d["roofline"]["avg_launch_ms"], "frac %.3f"%d["roofline"]["frac"], "ms/step %.3f"%d["ms_per_step"], "value %.3g"%d["value"])
    except Exception as e:
        print("${v}", f, "failed", e)
PY
done
unset FLOWAGG_LIB
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum \
    --clock-control none -k regex:k_tile -c 4 --csv --log-file gpurun_out/${TAG}_5tuple_traffic.csv python profiles/prof_configs2.py 5tuple > gpurun_out/${TAG}_5tuple_traffic.log 2>&1
tail -6 gpurun_out/${TAG}_5tuple_traffic.csv | cut -c1-400
ls -la gpurun_out/ | grep ${TAG} | head -30
