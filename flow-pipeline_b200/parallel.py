"""Multi-GPU plumbing: one process per GPU, torch.distributed for the collectives.

The path shards with no data-path collective: Kafka partition p belongs to rank
p mod world (sarama runs one ConsumeClaim per claimed partition,
inserter/inserter.go:176), every rank decodes and aggregates its own partitions.
Two exchange steps exist, both at query time only:
  * box-wide top-K: all-reduce(sum, 64-bit) of the fixed-size count-min sketch
    (the sketch is linear, so the reduced sketch is the sketch of the union),
    then each rank ranks its own candidate keys by the GLOBAL estimate and the
    per-rank lists are merged;
  * exact roll-up: the per-rank flows_5m rows (a few MiB at most for the
    AS-level keys) are gathered and summed by key on rank 0.
Backend: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist

from .flowagg import FA_CMS_GLOBAL, FA_CMS_LOCAL, HH_DTYPE, ROW_DTYPE, FlowAgg


def partition_owner(partition: int, world: int) -> int:
    return partition % world


def my_partitions(n_partitions: int, world: int, rank: int):
    return [p for p in range(n_partitions) if partition_owner(p, world) == rank]


class _CudaView:
    """Zero-copy view of library-owned device memory as a torch tensor (int64: NCCL's
    64-bit sum is the same bits as the uint64 wrap-around sum)."""

    def __init__(self, ptr, n_words):
        self.__cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i8", "data": (ptr, False), "version": 3}


def sketch_tensor(agg: FlowAgg, which=FA_CMS_LOCAL, device=None):
    ptr, n = agg.cms_device(which)
    return torch.as_tensor(_CudaView(ptr, n), device=device if device is not None else f"cuda:{agg.cfg.device}")


def allreduce_sketch(local: torch.Tensor, out: torch.Tensor, group=None):
    """out <- sum over ranks of local (out of place, so repeated queries do not double count)."""
    out.copy_(local)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def box_topk(agg: FlowAgg, k: int, group=None):
    """Box-wide heavy hitters: sketch all-reduce + per-rank candidates + merge (every rank returns the list)."""
    local = sketch_tensor(agg, FA_CMS_LOCAL)
    glob = sketch_tensor(agg, FA_CMS_GLOBAL)
    agg.sync()
    torch.cuda.current_stream().synchronize()
    allreduce_sketch(local, glob, group)
    torch.cuda.synchronize()
    mine = agg.topk_local(k, FA_CMS_GLOBAL)
    return merge_topk(mine, k, agg.kw, group, device=glob.device)


def merge_topk(mine: np.ndarray, k: int, key_words: int, group=None, device="cpu"):
    """Gather every rank's (<= k) candidates and keep the global top-k (estimate desc, key asc)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return FlowAgg.topk_merge([mine], key_words, k)
    words = HH_DTYPE.itemsize // 8
    send = np.zeros(k, dtype=HH_DTYPE)
    send[: len(mine)] = mine
    cnt = torch.tensor([len(mine)], dtype=torch.int64, device=device)
    t = torch.from_numpy(send.view(np.int64).reshape(k, words).copy()).to(device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    parts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    dist.all_gather(parts, t, group=group)
    lists = []
    for c, p in zip(cnts, parts):
        a = np.ascontiguousarray(p.cpu().numpy()).view(HH_DTYPE).reshape(-1)[: int(c.item())]
        lists.append(a)
    return FlowAgg.topk_merge(lists, key_words, k)


def merge_rows(rows: np.ndarray, key_words: int, group=None, device="cpu"):
    """Sum the per-rank roll-up rows by key (SummingMergeTree's merge, create.sh:88-90) on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        words = ROW_DTYPE.itemsize // 8
        cnt = torch.tensor([len(rows)], dtype=torch.int64, device=device)
        cnts = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(cnts, cnt, group=group)
        m = int(max(c.item() for c in cnts))
        send = np.zeros(m, dtype=ROW_DTYPE)
        send[: len(rows)] = rows
        t = torch.from_numpy(send.view(np.int64).reshape(m, words).copy()).to(device)
        parts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(parts, t, group=group)
        rows = np.concatenate([np.ascontiguousarray(p.cpu().numpy()).view(ROW_DTYPE).reshape(-1)[: int(c.item())]
                               for c, p in zip(cnts, parts)])
    return sum_rows_by_key(rows, key_words)


def partition_rows(rows: np.ndarray, key_mode, world: int):
    """Rows grouped by owner rank (fa_row_owner's hash partition) and the per-owner counts."""
    from .flowagg import row_owner

    owner = row_owner(key_mode, rows, world)
    order = np.argsort(owner, kind="stable")
    return rows[order], np.bincount(owner, minlength=world).astype(np.int64)


def exchange_partial_rows(rows: np.ndarray, key_mode, group=None, device="cpu"):
    """All-to-all of partial roll-up rows by key owner.  Returns the rows this rank owns, one partial per sender,
    as an int64 tensor [n, words] on `device` (view it as ROW_DTYPE on the host)."""
    world = dist.get_world_size(group)
    rows, counts = partition_rows(rows, key_mode, world)
    words = ROW_DTYPE.itemsize // 8
    send_counts = torch.from_numpy(counts).to(device)
    recv_counts = torch.zeros_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    rc = [int(x) for x in recv_counts.cpu().tolist()]
    sc = [int(x) for x in counts.tolist()]
    send = torch.from_numpy(np.ascontiguousarray(rows).view(np.int64).reshape(len(rows), words).copy()).to(device)
    recv = torch.empty((sum(rc), words), dtype=torch.int64, device=device)
    dist.all_to_all_single(recv, send, output_split_sizes=rc, input_split_sizes=sc, group=group)
    return recv


def rows_of_tensor(t: torch.Tensor) -> np.ndarray:
    return np.ascontiguousarray(t.cpu().numpy()).view(ROW_DTYPE).reshape(-1)


def exchange_rows(agg: FlowAgg, group=None, device="cpu"):
    """The one exchange step of an exact box-wide roll-up when every rank is its own process:
    each rank flushes its partial rows, sends every row to the rank that owns its key (all-to-all; NCCL
    send/recv over NVLink when `device` is a CUDA device), folds what it receives into its own table on the GPU
    (fa_merge_rows) and flushes again.  Returns this rank's share: the exact, fully merged rows of the keys it
    owns, in ORDER BY order.  The shares of different ranks are disjoint."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return agg.flush()
    recv = exchange_partial_rows(agg.flush(sort=False), agg.key_mode, group, device)
    if recv.is_cuda:
        torch.cuda.current_stream(recv.device).synchronize()
        agg.merge_rows(recv, n=recv.shape[0])  # straight from device memory
        agg.sync()
    else:
        agg.merge_rows(rows_of_tensor(recv))
    return agg.flush()


def sum_rows_by_key(rows: np.ndarray, key_words: int):
    if len(rows) == 0:
        return rows
    keys = rows["key"][:, :key_words]
    order = np.lexsort(tuple(keys[:, i] for i in reversed(range(key_words))))
    rows = rows[order]
    keys = rows["key"][:, :key_words]
    new = np.ones(len(rows), dtype=bool)
    new[1:] = np.any(keys[1:] != keys[:-1], axis=1)
    idx = np.flatnonzero(new)
    out = rows[idx].copy()
    for f in ("bytes", "packets", "count"):
        out[f] = np.add.reduceat(rows[f], idx)  # uint64 wrap-around
    return out
