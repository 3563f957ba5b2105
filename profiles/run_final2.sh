#!/bin/bash
# Round 2, second session: final evidence on one B200 -- full GPU suite, both bench arms, the group-table size sweep, the other
# BASELINE configs, launch list of the default bench, full ncu captures of the wide-key and bounded-candidate kernels.
TAG=${1:-r02s2_final}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -2 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
if [ ! -s gpurun_out/${TAG}_bench.json ]; then
  tail -5 gpurun_out/${TAG}_bench.err
  timeout 900 python bench.py --steps 10 --warmup 3 --no-frame-leg > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_noframe.err
fi
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value %.4g ms/step %.3f kernel %.4f ms frac %.3f"%(d["value"],d["ms_per_step"],d["roofline"]["avg_launch_ms"],d["roofline"]["frac"]))
e=d["e2e"]; print("e2e %.4g (%s) ms/step %.2f"%(e["value"],e.get("input"),e["ms_per_step"]))
o=e.get("other_input"); 
if o: print("other e2e %.4g (%s) ms/step %.2f"%(o["value"],o.get("input"),o["ms_per_step"]))
print("cpu", d["cpu_baseline"]["value"], "framed_stream", d.get("framed_stream",{}).get("ms_per_slab"))
PY
for cap in 1048576 2097152; do
  FA_BENCH_TABLE_CAP=$cap timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-frame-leg > gpurun_out/${TAG}_bench_cap${cap}.json 2> gpurun_out/${TAG}_bench_cap${cap}.err
  python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_cap${cap}.json')); print('cap ${cap}: kernel ms %.4f frac %.3f ms/step %.3f value %.4g'%(d['roofline']['avg_launch_ms'],d['roofline']['frac'],d['ms_per_step'],d['value']))"
done
timeout 600 python profiles/run_configs.py > gpurun_out/${TAG}_other_configs.json 2> gpurun_out/${TAG}_other_configs.err
cut -c1-330 gpurun_out/${TAG}_other_configs.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_default_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_default_bench_launches.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:k_tile -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_5tuple \
    python profiles/prof_configs2.py 5tuple > gpurun_out/${TAG}_prof_5tuple.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_tile -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_bounded \
    python profiles/prof_configs2.py bounded > gpurun_out/${TAG}_prof_bounded.log 2>&1
ls -la gpurun_out | grep ${TAG} | awk '{print $5, $9}'
