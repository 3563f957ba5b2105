#!/bin/bash
# Runs on the GPU box (under gpurun): launch list + one full ncu capture of the fused kernel.
# Usage: profiles/run_profile.sh <tag>     -> gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_prof.ncu-rep
TAG=${1:-r01}
mkdir -p gpurun_out
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --flows 33554432 --no-e2e --no-cpu > gpurun_out/${TAG}_launches_bench.log 2>&1
# the top kernel, full set, 2 launches after warm-up
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile -s 2 -c 2 -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 1 --warmup 1 --flows 33554432 --no-e2e --no-cpu > gpurun_out/${TAG}_prof_bench.log 2>&1
ls -la gpurun_out/
