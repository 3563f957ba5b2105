#!/usr/bin/env python
"""Throughput of the other BASELINE.json configs (not the bench line; parity for them is in tests/).
Usage (GPU box): python profiles/run_configs.py > gpurun_out/configs.json"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_pipeline_b200 as fp  # noqa: E402

ONLY = sys.argv[1:]  # substrings of the config names to run (default: all)


def want(name):
    return not ONLY or any(o in name for o in ONLY)


SLAB = 1 << 24
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
d_buf = torch.empty(SLAB * 92, dtype=torch.uint8, device=dev)
d_off = torch.empty(SLAB + 1, dtype=torch.int32, device=dev)
out = {}


def run(name, agg, cfg, n_total, after=None):
    if not want(name):
        return
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range((n_total + SLAB - 1) // SLAB)]
    done = 0
    nbytes = 0
    for i, (e0, e1) in enumerate(ev):
        n = min(SLAB, n_total - done)
        nb = agg.mocker_device(cfg, done, n, d_buf, d_buf.numel(), d_off)
        e0.record()
        agg.submit_device(d_buf, d_off, n, nb)
        e1.record()
        done += n
        nbytes += nb + 4 * (n + 1)
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    st = agg.stats()
    res = {"flows": n_total, "kernel_ms": ms, "per_launch_ms": [round(a.elapsed_time(b), 3) for a, b in ev][:8], "flows_per_s": n_total / ms * 1e3, "GBps": nbytes / ms / 1e6, "stats": st}
    if after:
        res.update(after(agg))
    out[name] = res
    print(name, json.dumps(res), file=sys.stderr, flush=True)


# kernel 1 alone: decode to the 20 columns (16.7M records)
cfg = fp.FaMockerConfig.make(seed=1, flows_per_second=250_000, n_src_as=256, n_dst_as=256, framed=True)
with fp.FlowAgg("flows5m", stream=stream, columns=True, aggregate=False, max_batch_records=SLAB) as a:
    run("kernel 1 alone: decode 2^24 records to 20 columns", a, cfg, SLAB)
    run("kernel 1 alone: decode 2^24 records to 20 columns (2nd pass)", a, cfg, SLAB)
with fp.FlowAgg("aspair", stream=stream, columns=True, max_batch_records=SLAB, table_capacity=1 << 19) as a:
    run("unfused K1 -> K2 (columns then aggregate), 2^24 records", a, cfg, SLAB)
    run("unfused K1 -> K2 (columns then aggregate), 2^24 records (2nd pass)", a, cfg, SLAB)
# mocker-native distribution: 9 AS pairs x 2 slots -> the hot-key worst case for the table atomics
cfg = fp.FaMockerConfig.make(seed=1, flows_per_second=250_000, framed=True)
with fp.FlowAgg("flows5m", stream=stream) as a:
    run("mocker_native_flows5m_100M (18 groups)", a, cfg, 100_000_000)
# configs[1] with the full flows_5m key
cfg = fp.FaMockerConfig.make(seed=1, flows_per_second=250_000, n_src_as=256, n_dst_as=256, framed=True)
with fp.FlowAgg("flows5m", stream=stream, table_capacity=1 << 20) as a:
    run("configs[1] with full flows_5m key 100M (131072 groups)", a, cfg, 100_000_000)
# configs[2]: CMS d=4 w=2^20, top-1000 SrcAddr over 1B flows, Zipf-like addresses
cfg = fp.FaMockerConfig.make(seed=1, flows_per_second=2_500_000, addr_mode=1, framed=True)


def topk(a):
    t0 = time.time()
    top = a.topk_local(1000)
    t1 = time.time()
    rows = a.flush(keep=True, sort=False)
    exact = rows[np.argsort(-rows["bytes"].astype(np.int64), kind="stable")][:1000]
    hit = {bytes(k.tobytes()) for k in top["key"]}
    recall = sum(bytes(k.tobytes()) in hit for k in exact["key"]) / 1000.0
    return {"topk_ms": (t1 - t0) * 1e3, "distinct_keys": int(len(rows)), "top1000_recall_vs_exact": recall,
            "top1_estimate": int(top["estimate"][0]), "top1_exact": int(exact["bytes"][0])}


EXACT_TOP = {}


def topk_exact(a):
    r = topk(a)
    rows = a.flush(keep=True, sort=False)
    exact = rows[np.argsort(-rows["bytes"].astype(np.int64), kind="stable")][:1000]
    EXACT_TOP["keys"] = {bytes(k.tobytes()) for k in exact["key"]}
    return r


def topk_bounded(a):
    t0 = time.time()
    top = a.topk_local(1000)
    t1 = time.time()
    res = {"topk_ms": (t1 - t0) * 1e3, "candidates_in_table": int(a.stats()["n_groups"]), "top1_estimate": int(top["estimate"][0])}
    if "keys" in EXACT_TOP:
        res["top1000_recall_vs_exact"] = sum(bytes(k.tobytes()) in EXACT_TOP["keys"] for k in top["key"]) / 1000.0
    return res


N2 = int(os.environ.get("FA_CONFIG2_FLOWS", 1_000_000_000))
if want("configs[2] CMS d=4 w=2^20 top-1000 SrcAddr over 1B flows, EXACT"):
    with fp.FlowAgg("srcaddr", stream=stream, cms=True, cms_depth=4, cms_width_log2=20, table_capacity=1 << 25) as a:
        run("configs[2] CMS d=4 w=2^20 top-1000 SrcAddr over 1B flows, EXACT candidate table (1.6 GB, every key)", a, cfg, N2, topk_exact)
if want("configs[2] CMS d=4 w=2^20 top-1000 SrcAddr over 1B flows, bounded"):
    # the sketch workload proper: bounded candidate set (FA_CFG_TOPK_ONLY), memory independent of the number of keys
    with fp.FlowAgg("srcaddr", stream=stream, topk_only=True, topk_k=1000, cms_depth=4, cms_width_log2=20) as a:
        run("configs[2] CMS d=4 w=2^20 top-1000 SrcAddr over 1B flows, bounded candidates (FA_CFG_TOPK_ONLY, 12 MiB table)", a, cfg, N2, topk_bounded)
# configs[4]: 100M unique 5-tuples, HBM open-address table at load 0.37
cfg = fp.FaMockerConfig.make(seed=1, flows_per_second=250_000, addr_mode=2, framed=True)
if want("configs[4]"):
    with fp.FlowAgg("5tuple", stream=stream, table_capacity=1 << 28) as a:
        run("configs[4] 100M unique 5-tuples (2^28-slot table in HBM)", a, cfg, 100_000_000)
print(json.dumps(out, indent=1))
