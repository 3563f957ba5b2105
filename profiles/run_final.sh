#!/bin/bash
# Runs on the GPU box: the round's final evidence -- default bench lines (both arms), launch list, full ncu capture of the
# fused kernel and of the 5-tuple instantiation, the other BASELINE configs, the framed-stream launch list.
TAG=${1:-r02_final}
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -2 gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_default_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_default_bench_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile -s 2 -c 2 -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 1 --warmup 1 --flows 33554432 --no-e2e --no-cpu --no-frame-leg > gpurun_out/${TAG}_prof_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_framed_launches.csv \
    python profiles/frame_probe.py > gpurun_out/${TAG}_framed_launches.log 2>&1
timeout 900 python profiles/run_configs.py > gpurun_out/${TAG}_other_configs.json 2> gpurun_out/${TAG}_other_configs.err
timeout 600 ncu --set full --clock-control none -k regex:k_tile -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_5tuple \
    python profiles/prof_configs2.py 5tuple > gpurun_out/${TAG}_prof_5tuple.log 2>&1
ls -la gpurun_out | grep ${TAG}
