import os, sys, time
sys.path.insert(0, '.')
print('cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
for f in ('/sys/fs/cgroup/cpu.max', '/sys/fs/cgroup/cpu/cpu.cfs_quota_us', '/sys/fs/cgroup/cpu/cpu.cfs_period_us'):
    try: print(f, open(f).read().strip())
    except Exception as e: print(f, 'n/a')
import flow_pipeline_b200 as fp
from oracle import oracle as o
import bench
cfg = bench.mocker_cfg(fp)
slabs = bench.host_slabs(fp, cfg, 0, 1 << 24, 1 << 20)
for th in (1, 2, 4, 8, 16, 32, 64, 128):
    rows, res = o.run_slabs(slabs, key_mode='aspair', threads=th)
    print(th, 'threads', round((1 << 24) / res['seconds'] / 1e6, 1), 'M flows/s')
