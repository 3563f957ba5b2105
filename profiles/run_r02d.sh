#!/bin/bash
# GPU call 3: Counters split by cache line + overlapped replica/candidate probes (configs[2]); wide-key table with verified sector stores.
TAG=${1:-r02s2c}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 400 python profiles/run_configs.py "configs[2]" "configs[4]" > gpurun_out/${TAG}_configs.json 2> gpurun_out/${TAG}_configs.err
cut -c1-420 gpurun_out/${TAG}_configs.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print('bench kernel ms %.4f frac %.3f ms/step %.3f value %.3g'%(d['roofline']['avg_launch_ms'],d['roofline']['frac'],d['ms_per_step'],d['value']))"
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum \
    --clock-control none -k regex:k_tile -c 4 --csv --log-file gpurun_out/${TAG}_5tuple_traffic.csv python profiles/prof_configs2.py 5tuple > gpurun_out/${TAG}_5tuple_traffic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile -s 2 -c 1 -f -o gpurun_out/${TAG}_bounded python profiles/prof_configs2.py bounded > gpurun_out/${TAG}_bounded_prof.log 2>&1
tail -1 gpurun_out/${TAG}_bounded_prof.log
