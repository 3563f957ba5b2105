"""CPU, world_size 2, gloo: the N>1 plumbing (partition ownership, sketch all-reduce, top-K and
row merges).  The per-rank sketches/rows come from the oracle here; on the GPU box the same
functions are fed from fa_cms_device / fa_flush."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib

    import flow_pipeline_b200 as fp
    from oracle import oracle as o

    par = importlib.import_module("flow-pipeline_b200.parallel")
    n_part, per_part = 4, 3000
    cfg = fp.FaMockerConfig.make(seed=11, flows_per_second=50, n_src_as=8, n_dst_as=8, addr_mode=1, framed=True)
    mine = par.my_partitions(n_part, world, rank)
    assert mine == [p for p in range(n_part) if p % world == rank]
    d, wl = 4, 10
    cms = np.zeros(d << wl, dtype=np.uint64)
    rows_all, cand_all = [], []
    for p in mine:  # partition p = records [p*per_part, (p+1)*per_part)
        buf, offs = fp.mocker_host(cfg, p * per_part, per_part)
        cand, c, _ = o.run_batch(buf, offs, key_mode="srcaddr", cms=(d, wl))
        rows, _, _ = o.run_batch(buf, offs, key_mode="flows5m")
        cms += c
        rows_all.append(rows)
        cand_all.append(cand)
    local = torch.from_numpy(cms.view(np.int64))
    glob = torch.zeros_like(local)
    par.allreduce_sketch(local, glob)
    g = glob.numpy().view(np.uint64)
    cands = par.sum_rows_by_key(np.concatenate(cand_all), 4)
    k = 20
    top_local = o.topk(g, d, wl, 4, cands, k).view(fp.HH_DTYPE)
    top = par.merge_topk(top_local, k, 4)
    rows = par.merge_rows(np.concatenate(rows_all), 4)
    # hash-partitioned exchange (the high-cardinality merge): every rank ends up with the partials of the keys it owns
    part = par.sum_rows_by_key(np.concatenate(cand_all), 4)
    got = par.rows_of_tensor(par.exchange_partial_rows(part, "srcaddr"))
    assert (fp.row_owner("srcaddr", got, world) == rank).all()
    share = par.sum_rows_by_key(got, 4)
    q.put(("share", rank, share.copy()))
    if rank == 0:
        q.put(("main", g.copy(), top.copy(), rows.copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_matches_single_process():
    sys.path.insert(0, ROOT)
    import flow_pipeline_b200 as fp
    from oracle import oracle as o

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    shares, main = {}, None
    while len(shares) < 2 or main is None:
        item = q.get(timeout=120)
        if item[0] == "share":
            shares[item[1]] = item[2]
        else:
            main = item[1:]
    g, top, rows = main
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process answer over all four partitions
    cfg = fp.FaMockerConfig.make(seed=11, flows_per_second=50, n_src_as=8, n_dst_as=8, addr_mode=1, framed=True)
    buf, offs = fp.mocker_host(cfg, 0, 4 * 3000)
    cand, cms, _ = o.run_batch(buf, offs, key_mode="srcaddr", cms=(4, 10))
    want_rows, _, _ = o.run_batch(buf, offs, key_mode="flows5m")
    assert np.array_equal(g, cms)                     # sketch is linear
    want_top = o.topk(cms, 4, 10, 4, cand, 20)
    assert np.array_equal(top["key"], want_top["key"]) and np.array_equal(top["estimate"], want_top["estimate"])
    assert np.array_equal(rows, want_rows)
    # the two shares are disjoint, both non-trivial, and together they are the exact group-by of everything
    import importlib

    par = importlib.import_module("flow-pipeline_b200.parallel")
    assert len(shares[0]) and len(shares[1]) and len(shares[0]) + len(shares[1]) == len(cand)
    assert np.array_equal(par.sum_rows_by_key(np.concatenate([shares[0], shares[1]]), 4), cand)
