#!/usr/bin/env python
"""One short pass of BASELINE configs[2] (bounded candidates) for ncu: 4 slabs of 2^24 Zipf-addressed flows."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_pipeline_b200 as fp
SLAB = 1 << 24
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
d_buf = torch.empty(SLAB * 92, dtype=torch.uint8, device=dev)
d_off = torch.empty(SLAB + 1, dtype=torch.int32, device=dev)
mode = sys.argv[1] if len(sys.argv) > 1 else "bounded"
cfg = fp.FaMockerConfig.make(seed=1, flows_per_second=2_500_000, addr_mode=1 if mode != "5tuple" else 2, framed=True)
if mode == "bounded":
    a = fp.FlowAgg("srcaddr", stream=stream, topk_only=True, topk_k=1000, cms_depth=4, cms_width_log2=20)
elif mode == "5tuple":
    a = fp.FlowAgg("5tuple", stream=stream, table_capacity=1 << 28)
else:
    a = fp.FlowAgg("srcaddr", stream=stream, cms=True, cms_depth=4, cms_width_log2=20, table_capacity=1 << 25)
for i in range(4):
    nb = a.mocker_device(cfg, i * SLAB, SLAB, d_buf, d_buf.numel(), d_off)
    a.submit_device(d_buf, d_off, SLAB, nb)
a.sync()
print(a.stats())
a.close()
