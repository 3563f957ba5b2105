import sys, time
sys.path.insert(0, '.')
import torch, numpy as np
import flow_pipeline_b200 as fp
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
for n in (1 << 18, 1 << 21, 1 << 24):
    cfg = fp.FaMockerConfig.make(seed=1, flows_per_second=250000, n_src_as=256, n_dst_as=256, framed=True)
    a = fp.FlowAgg("aspair", device=0, stream=stream, table_capacity=1 << 19)
    d_buf = torch.empty(n * 88 + 4096, dtype=torch.uint8, device=dev)
    d_off = torch.empty(n + 1, dtype=torch.int32, device=dev)
    nb = a.mocker_device(cfg, 0, n, d_buf, d_buf.numel(), d_off)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        a.submit_device(d_buf, None, 0, nb)
        a.sync(); t1 = time.time()
        print(n, "records: framed submit", round((t1 - t0) * 1e3, 3), "ms", flush=True)
    a.close()
