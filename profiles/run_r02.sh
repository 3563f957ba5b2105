#!/bin/bash
# Runs on the GPU box (under gpurun).  Usage: profiles/run_r02.sh <tag> [tests|notests] [prof|noprof]
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
if [ "${2:-tests}" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
  tail -5 gpurun_out/${TAG}_pytest_gpu.log
fi
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json | cut -c1-1500
FA_SHAPE=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_bench_noshape.json 2> gpurun_out/${TAG}_bench_noshape.err
cat gpurun_out/${TAG}_bench_noshape.json | cut -c1-600
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --first 300000000 > gpurun_out/${TAG}_bench_first300m.json 2> gpurun_out/${TAG}_bench_first300m.err
cat gpurun_out/${TAG}_bench_first300m.json | cut -c1-600
if [ "${3:-prof}" = "prof" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 1 --flows 33554432 --no-e2e --no-cpu > gpurun_out/${TAG}_launches_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile -s 2 -c 2 -f -o gpurun_out/${TAG}_prof \
      python bench.py --steps 1 --warmup 1 --flows 33554432 --no-e2e --no-cpu > gpurun_out/${TAG}_prof_bench.log 2>&1
fi
ls -la gpurun_out/ | grep ${TAG}
