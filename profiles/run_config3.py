#!/usr/bin/env python
"""BASELINE.json configs[3] at full size: N Kafka partitions, one per GPU (torchrun, one process per GPU), 10 B flows box-wide
(1.25 B per GPU at N = 8), per-GPU count-min sketch d=4 w=2^20, ONE NCCL all-reduce of the sketches (32 MiB, sum, 64-bit),
box-wide top-1000 SrcAddr.  Input is generated on each GPU slab by slab (Zipf-addressed mocker flows, own seed per partition).

Two contexts per rank see the same flows: the sketch workload proper (FA_CFG_TOPK_ONLY: bounded candidate table) and an exact
one (every key kept) whose hash-partitioned exchange (parallel.exchange_rows) yields the EXACT box-wide weights the answer is
checked against: no estimate below the exact weight, recall of the exact top-1000.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 profiles/run_config3.py
"""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_pipeline_b200 as fp  # noqa: E402

par = importlib.import_module("flow-pipeline_b200.parallel")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000_000
per_rank = total // max(world, 1) if len(sys.argv) > 1 else 1_250_000_000
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.current_stream().cuda_stream
SLAB, K = 1 << 24, 1000
d_buf = torch.empty(SLAB * 92, dtype=torch.uint8, device=dev)
d_off = torch.empty(SLAB + 1, dtype=torch.int32, device=dev)
cfg = fp.FaMockerConfig.make(seed=1 + rank, flows_per_second=2_500_000, addr_mode=fp.FA_ADDR_ZIPF24, framed=True)
sk = fp.FlowAgg("srcaddr", device=local, stream=stream, topk_only=True, topk_k=K, cms_depth=4, cms_width_log2=20)
ex = fp.FlowAgg("srcaddr", device=local, stream=stream, cms=False, table_capacity=1 << 25)


def dmax(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def dsum(x):
    t = torch.tensor([int(x)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


ev = []
done = 0
t_wall = time.time()
while done < per_rank:
    n = min(SLAB, per_rank - done)
    nb = sk.mocker_device(cfg, done, n, d_buf, d_buf.numel(), d_off)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sk.submit_device(d_buf, d_off, n, nb)
    e1.record()
    ex.submit_device(d_buf, d_off, n, nb)
    sk.sync()
    ex.sync()
    ev.append((e0, e1))
    done += n
torch.cuda.synchronize()
sketch_ms = sum(a.elapsed_time(b) for a, b in ev)
local_t, glob_t = par.sketch_tensor(sk, fp.FA_CMS_LOCAL), par.sketch_tensor(sk, fp.FA_CMS_GLOBAL)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
par.allreduce_sketch(local_t, glob_t)  # warm-up (NCCL channel set-up)
torch.cuda.synchronize()
a0.record()
par.allreduce_sketch(local_t, glob_t)
a1.record()
torch.cuda.synchronize()
allreduce_ms = dmax(a0.elapsed_time(a1))
tot = torch.tensor([int(local_t.sum().item())], dtype=torch.int64, device=dev)
if world > 1:
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
linear = int(tot.item()) == int(glob_t.sum().item())
q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
q0.record()
top = par.box_topk(sk, K)
q1.record()
torch.cuda.synchronize()
topk_ms = dmax(q0.elapsed_time(q1))
share = par.exchange_rows(ex, device=dev)  # exact box-wide sum(Bytes) of the keys this rank owns
mine = share[np.argsort(share["bytes"], kind="stable")[::-1][:K]]
exact = par.merge_rows(mine, 4, device=dev)
exact = exact[np.argsort(exact["bytes"], kind="stable")[::-1][:K]]
exd = {tuple(r["key"][:4]): int(r["bytes"]) for r in exact}
hits = sum(1 for r in top if tuple(r["key"][:4]) in exd)
under = sum(1 for r in top if tuple(r["key"][:4]) in exd and int(r["estimate"]) < exd[tuple(r["key"][:4])])
n_flows_box = dsum(per_rank)
distinct_box = dsum(len(share))
slowest_sketch_ms = dmax(sketch_ms)
if rank == 0:
    print(json.dumps({
        "config": "configs[3]: %d Kafka partitions sharded per GPU, NCCL sketch all-reduce, %.3g flows box-wide top-%d" % (world, n_flows_box, K),
        "n_gpus": world, "flows_box_wide": n_flows_box, "flows_per_gpu": per_rank,
        "sketch_kernel_ms_per_gpu_max": slowest_sketch_ms, "flows_per_s_box_wide_sketching": n_flows_box / (slowest_sketch_ms * 1e-3),
        "sketch_allreduce_ms": allreduce_ms, "sketch_allreduce_bytes": int(local_t.numel() * 8),
        "allreduce_busbw_GBs": 2 * (world - 1) / max(world, 1) * local_t.numel() * 8 / (allreduce_ms * 1e-3) / 1e9,
        "box_topk_ms": topk_ms, "sketch_linear": bool(linear), "top%d_recall_vs_exact_box_wide" % K: hits / max(len(exd), 1),
        "estimates_below_exact": under, "candidates_in_table_rank0": sk.stats()["n_groups"], "distinct_keys_box_wide": distinct_box,
        "top1": {"estimate": int(top["estimate"][0]), "exact": int(exact["bytes"][0])},
        "wall_s": time.time() - t_wall,
        "check": "ok" if (linear and under == 0 and len(top) == K and hits >= 0.99 * len(exd)) else "MISMATCH"}))
sk.close()
ex.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
