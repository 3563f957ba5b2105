#!/bin/bash
# GPU call 2: 256-bit key-record stores (configs[4]) + full ncu capture of the bounded-candidate kernel (configs[2]).
TAG=${1:-r02s2b}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "5tuple or wide_key or table_full or all_key_modes" > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python profiles/run_configs.py "configs[4]" > gpurun_out/${TAG}_configs.json 2> gpurun_out/${TAG}_configs.err
tail -1 gpurun_out/${TAG}_configs.err | cut -c1-700
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum \
    --clock-control none -k regex:k_tile -c 4 --csv --log-file gpurun_out/${TAG}_5tuple_traffic.csv python profiles/prof_configs2.py 5tuple > gpurun_out/${TAG}_5tuple_traffic.log 2>&1
grep -E "dram__bytes|gpu__time" gpurun_out/${TAG}_5tuple_traffic.csv | tail -3 | cut -c150-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile -s 2 -c 1 -f -o gpurun_out/${TAG}_bounded python profiles/prof_configs2.py bounded > gpurun_out/${TAG}_bounded_prof.log 2>&1
tail -2 gpurun_out/${TAG}_bounded_prof.log
ls -la gpurun_out/ | grep ${TAG}
