#!/bin/bash
# GPU call 4: warp-level key combining + register budget of the weighted address-key kernel (configs[2]).
TAG=${1:-r02s2d}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "topk or srcaddr or cms or all_key_modes or box" > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -2 gpurun_out/${TAG}_pytest_gpu.log
for v in default w4_8 w4_5 nocombine; do
  if [ "$v" = "default" ]; then unset FLOWAGG_LIB; else export FLOWAGG_LIB=$PWD/flow-pipeline_b200/_variants/$v.so; fi
  FA_CONFIG2_FLOWS=402653184 timeout 200 python profiles/run_configs.py "bounded" > gpurun_out/${TAG}_cfg2_${v}.json 2> gpurun_out/${TAG}_cfg2_${v}.err
  echo "$v: $(grep -o '"flows_per_s": [0-9.]*' gpurun_out/${TAG}_cfg2_${v}.err) $(grep -o '"per_launch_ms": \[[^]]*\]' gpurun_out/${TAG}_cfg2_${v}.err)"
done
