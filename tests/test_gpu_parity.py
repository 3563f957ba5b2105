"""GPU parity: the sm_100a kernels, driven through the C ABI (include/flowagg.h), against the
CPU oracle and the golden vectors.  Bit-exact: everything on this path is integer/byte work."""
import os

import numpy as np
import pytest

from conftest import GOLD2ORACLE, concat_records, frame

pytestmark = pytest.mark.gpu

COLS = ["time_received", "sampling_rate", "time_flow_start", "bytes", "packets", "type", "sequence_num", "src_as", "dst_as",
        "etype", "proto", "src_port", "dst_port", "src_addr", "dst_addr", "sampler_addr"]


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), "these tests need the B200"
    return torch


def gpu_columns(fp, blob, offs, framed, **kw):
    with fp.FlowAgg(columns=True, aggregate=False, max_batch_records=max(len(offs) - 1, 1024), **kw) as a:
        a.submit(blob, offs, framed=framed)
        cols, st = a.columns()
    n = len(offs) - 1
    out = {}
    for k, v in cols.items():
        out[k] = v[: n * 16].reshape(n, 16) if k in ("src_addr", "dst_addr", "sampler_addr") else v[:n]
    return out, st


def assert_columns_equal(gpu, ora, n):
    valid = ora["valid"].astype(bool)
    assert np.array_equal(gpu["valid"].astype(bool), valid)
    for k in COLS:
        assert np.array_equal(gpu[k][valid], ora[k][valid].astype(gpu[k].dtype)), k
        if k.endswith("_addr"):
            assert np.array_equal(gpu[k + "_len"][valid], np.minimum(ora[k + "_len"][valid], 255)), k
    # skipped records are zero rows (the inserter appends nothing for them, inserter.go:125-126)
    for k in COLS:
        assert not gpu[k][~valid].any(), k


# ---------------------------------------------------------------- kernel 1: decode


def test_edge_cases_decode(fp, oracle, edge_cases, torch_cuda):
    msgs = [bytes.fromhex(c["hex"]) for c in edge_cases["cases"]]
    blob, offs = concat_records(msgs)
    gpu, st = gpu_columns(fp, blob, offs, framed=False)
    want_ok = np.array([c["go_ok"] for c in edge_cases["cases"]])
    bad = [c["name"] for c, g in zip(edge_cases["cases"], gpu["valid"]) if bool(g) != c["go_ok"]]
    assert not bad, bad
    assert st["n_bad"] == int((~want_ok).sum()) and st["n_records"] == len(msgs)
    ora = oracle.decode_columns(blob, offs, framed=False)
    assert_columns_equal(gpu, ora, len(msgs))
    # and framed
    fblob, foffs = concat_records(frame(msgs))
    gpu2, _ = gpu_columns(fp, fblob, foffs, framed=True)
    for k in gpu:
        assert np.array_equal(gpu[k], gpu2[k]), k


def test_fuzz_2k_decode_vs_golden_and_oracle(fp, oracle, fuzz_2k, torch_cuda):
    g = fuzz_2k
    gpu, st = gpu_columns(fp, g["blob"], g["offsets"], framed=False)
    valid = g["valid"].astype(bool)
    assert np.array_equal(gpu["valid"].astype(bool), valid)
    for gk, ok in GOLD2ORACLE.items():
        want = g[gk][valid]
        assert np.array_equal(gpu[ok][valid], want.astype(gpu[ok].dtype)), gk
    assert_columns_equal(gpu, oracle.decode_columns(g["blob"], g["offsets"], framed=False), len(valid))


def test_mocker_10k_config0(fp, oracle, mocker_10k, torch_cuda):
    """BASELINE.json configs[0]: 10k mocker messages, 1 partition, bare values (Postgres path)."""
    g = mocker_10k
    gpu, st = gpu_columns(fp, g["blob"], g["offsets"], framed=False)
    assert st["n_bad"] == 0 and gpu["valid"].all()
    for gk, ok in GOLD2ORACLE.items():
        assert np.array_equal(gpu[ok], g[gk].astype(gpu[ok].dtype)), gk
    with fp.FlowAgg("flows5m") as a:
        a.submit(g["blob"], g["offsets"], framed=False)
        rows = a.flush()
        assert a.stats()["n_groups"] == 0  # flush resets the table
    assert np.array_equal(rows["key"][:, :4], g["rollup_key"])           # pandas groupby
    assert np.array_equal(np.stack([rows["bytes"], rows["packets"], rows["count"]], 1), g["rollup_val"])
    want, _, _ = oracle.run_batch(g["blob"], g["offsets"], framed=False, key_mode="flows5m")
    assert np.array_equal(rows, want)                                     # oracle, byte for byte
    # the inserter's 14-column row for a few records (inserter.go:142-157)
    for i in (0, 17, 9999):
        assert oracle.ip_string(bytes(gpu["src_addr"][i][: gpu["src_addr_len"][i]])).startswith("2001:db8:0:1::")


# ---------------------------------------------------------------- fused kernel 1 -> 2


def mixed_batch(fp, fuzz_2k, n_mocker=60000, addr_mode=0, n_as=3, seed=5):
    """mocker records with the fuzz set (bad records, SamplingRate != 1, odd addresses) spliced in"""
    cfg = fp.FaMockerConfig.make(seed=seed, flows_per_second=100, n_src_as=n_as, n_dst_as=n_as, addr_mode=addr_mode, framed=True)
    buf, offs = fp.mocker_host(cfg, 1000, n_mocker)
    g = fuzz_2k
    fmsgs = frame([bytes(g["blob"][g["offsets"][i]:g["offsets"][i + 1]]) for i in range(len(g["offsets"]) - 1)])
    fb, fo = concat_records(fmsgs)
    blob = np.concatenate([buf[: offs[n_mocker // 2]], fb, buf[offs[n_mocker // 2]:]])
    o = np.concatenate([offs[: n_mocker // 2], fo[:-1] + offs[n_mocker // 2], offs[n_mocker // 2:] + fo[-1]]).astype(np.uint32)
    return blob, o


@pytest.mark.parametrize("mode", ["flows5m", "aspair", "srcaddr", "dstaddr", "5tuple", "srcport", "dstport"])
@pytest.mark.parametrize("scale", [False, True])
def test_fused_rollup_all_key_modes(fp, oracle, fuzz_2k, torch_cuda, mode, scale):
    blob, offs = mixed_batch(fp, fuzz_2k, addr_mode=1 if mode == "srcaddr" else 0)
    want, _, res = oracle.run_batch(blob, offs, framed=True, key_mode=mode, scale=scale)
    with fp.FlowAgg(mode, scale_sampling=scale, table_capacity=1 << 18) as a:
        a.submit(blob, offs, framed=True)
        st = a.stats()
        rows = a.flush()
    assert st["n_bad"] == res["n_bad"] and st["n_nokey"] == res["n_nokey"] and st["n_groups"] == len(want)
    assert np.array_equal(rows, want)
    assert rows["count"].sum() == len(offs) - 1 - res["n_bad"] - res["n_nokey"]


def test_unfused_k1_k2_equals_fused(fp, oracle, fuzz_2k, torch_cuda):
    blob, offs = mixed_batch(fp, fuzz_2k)
    for mode in ("flows5m", "5tuple"):
        with fp.FlowAgg(mode, table_capacity=1 << 18) as a, fp.FlowAgg(mode, columns=True, table_capacity=1 << 18,
                                                                         max_batch_records=len(offs)) as b:
            a.submit(blob, offs)
            b.submit(blob, offs)
            assert np.array_equal(a.flush(), b.flush())


def test_sketch_and_topk(fp, oracle, torch_cuda):
    cfg = fp.FaMockerConfig.make(seed=9, flows_per_second=1000, addr_mode=1, framed=True)
    buf, offs = fp.mocker_host(cfg, 0, 300000)
    d, wl = 4, 14
    cand, cms, _ = oracle.run_batch(buf, offs, key_mode="srcaddr", cms=(d, wl))
    with fp.FlowAgg("srcaddr", cms=True, cms_depth=d, cms_width_log2=wl, table_capacity=1 << 19) as a:
        a.submit(buf, offs)
        got = a.cms_read()
        assert np.array_equal(got, cms)                       # counters bit-equal
        top = a.topk_local(1000)
        rows = a.flush(keep=True)
        # linearity on the device: a second pass doubles every counter
        a.submit(buf, offs)
        assert np.array_equal(a.cms_read(), cms * 2)
    assert np.array_equal(rows, cand)
    want = oracle.topk(cms, d, wl, 4, cand, 1000)
    assert np.array_equal(top["key"], want["key"]) and np.array_equal(top["estimate"], want["estimate"])
    # recall of the sketch top-100 against the exact heavy hitters (viz-ch.json:233 semantic)
    exact = cand[np.lexsort((cand["key"][:, 3], -cand["bytes"].astype(np.int64)))][:100]
    hit = {bytes(k.tobytes()) for k in top["key"][:200]}
    assert sum(bytes(k.tobytes()) in hit for k in exact["key"]) >= 90


# ---------------------------------------------------------------- ingest paths and edge shapes


@pytest.mark.parametrize("n", [0, 1, 31, 255, 256, 257, 1023, 5000])
def test_ragged_sizes(fp, oracle, torch_cuda, n):
    cfg = fp.FaMockerConfig.make(seed=2, flows_per_second=7, n_src_as=5, n_dst_as=4, framed=True)
    buf, offs = fp.mocker_host(cfg, 0, n)
    want, _, _ = oracle.run_batch(buf, offs, key_mode="flows5m")
    with fp.FlowAgg("flows5m") as a:
        a.submit(buf, offs)
        assert np.array_equal(a.flush(), want)


def test_host_batching_and_slabs(fp, oracle, torch_cuda):
    cfg = fp.FaMockerConfig.make(seed=4, flows_per_second=50, n_src_as=16, n_dst_as=16, framed=True)
    buf, offs = fp.mocker_host(cfg, 0, 100000)
    want, _, _ = oracle.run_batch(buf, offs, key_mode="flows5m")
    # 64 KiB / 500-record staging forces ~170 pipelined batches
    with fp.FlowAgg("flows5m", max_batch_bytes=64 << 10, max_batch_records=500) as a:
        a.submit(buf, offs)
        st = a.stats()
        assert st["n_submits"] > 100 and st["n_records"] == 100000
        assert np.array_equal(a.flush(), want)
    # the pinned slabs a Go host would memcpy into, alternating, slab-relative offsets
    with fp.FlowAgg("flows5m", max_batch_bytes=1 << 20, max_batch_records=8192) as a:
        r = 0
        slot = 0
        while r < 100000:
            hb, ho = a.host_buffer(slot)
            r1 = min(r + 8000, 100000)
            nb = int(offs[r1] - offs[r])
            hb[:nb] = buf[offs[r]:offs[r1]]
            ho[: r1 - r + 1] = offs[r:r1 + 1] - offs[r]
            a.submit(hb, ho[: r1 - r + 1], nbytes=nb)
            r = r1
            slot ^= 1
        assert np.array_equal(a.flush(), want)


def test_device_submit_and_device_mocker(fp, oracle, torch_cuda):
    torch = torch_cuda
    n = 200000
    cfg = fp.FaMockerConfig.make(seed=6, flows_per_second=500, n_src_as=256, n_dst_as=256, framed=True)
    hbuf, hoffs = fp.mocker_host(cfg, 77, n)
    want, _, _ = oracle.run_batch(hbuf, hoffs, key_mode="aspair")
    with fp.FlowAgg("aspair", stream=torch.cuda.current_stream().cuda_stream) as a:
        d_buf = torch.empty(n * 96, dtype=torch.uint8, device="cuda")
        d_off = torch.empty(n + 1, dtype=torch.int32, device="cuda")
        nb = a.mocker_device(cfg, 77, n, d_buf, d_buf.numel(), d_off)
        assert nb == len(hbuf)
        assert np.array_equal(d_buf[:nb].cpu().numpy(), hbuf)             # same bytes on CPU and GPU
        assert np.array_equal(d_off.cpu().numpy().view(np.uint32), hoffs)
        a.submit_device(d_buf, d_off, n, nb)
        assert np.array_equal(a.flush(), want)


def test_tile_too_big_for_shared_memory_falls_back_to_global(fp, oracle, torch_cuda):
    cfg = fp.FaMockerConfig.make(seed=8, flows_per_second=10, framed=False)
    buf, offs = fp.mocker_host(cfg, 0, 3000)
    msgs = [bytes(buf[offs[i]:offs[i + 1]]) for i in range(3000)]
    big = b"\xc2\x3e" + b"\xe0\xd4\x03" + b"z" * 60000  # unknown field 1000, 60 000 bytes
    for i in (5, 300, 301, 1500, 2999):
        msgs[i] = big + msgs[i] + big
    for framed in (False, True):
        blob, o = concat_records(frame(msgs) if framed else msgs)
        want, _, res = oracle.run_batch(blob, o, framed=framed, key_mode="5tuple")
        assert res["n_bad"] == 0
        with fp.FlowAgg("5tuple", table_capacity=1 << 14) as a:
            a.submit(blob, o, framed=framed)
            assert np.array_equal(a.flush(), want)


def test_corrupt_offsets_are_bad_records_not_crashes(fp, oracle, torch_cuda):
    cfg = fp.FaMockerConfig.make(seed=10, flows_per_second=10, framed=True)
    buf, offs = fp.mocker_host(cfg, 0, 2000)
    bad = offs.copy()
    bad[700] = bad[703]          # record 699 swallows three, 700..702 become empty/negative spans
    bad[1200] = 0xFFFFFFF0       # far outside the buffer
    bad[1500] = bad[1499] - 5    # decreasing
    with fp.FlowAgg("flows5m") as a:
        a.submit(buf, bad)
        st = a.stats()
        rows = a.flush()
    assert 4 <= st["n_bad"] <= 12 and rows["count"].sum() == 2000 - st["n_bad"]


def test_null_offsets_framing_on_gpu(fp, oracle, fuzz_2k, torch_cuda):
    blob, offs = mixed_batch(fp, fuzz_2k, n_mocker=200000)
    want, _, res = oracle.run_batch(blob, offs, framed=True, key_mode="flows5m")
    with fp.FlowAgg("flows5m") as a:
        a.submit(blob, None, framed=True)
        st = a.stats()
        assert st["n_records"] == len(offs) - 1 and st["n_bad"] == res["n_bad"]
        assert np.array_equal(a.flush(), want)
    # adversarial payloads: long zero runs and embedded bytes that look like record headers
    rng = np.random.default_rng(3)
    msgs = []
    for i in range(20000):
        k = int(rng.integers(0, 4))
        pay = [bytes(int(rng.integers(0, 300))), b"\x50" * int(rng.integers(0, 200)), bytes(rng.integers(0, 256, int(rng.integers(0, 400)), dtype=np.uint8)),
               b"\x02\x48\x05" * int(rng.integers(0, 90))][k]
        n = len(pay)
        ln = b""
        while True:
            b = n & 0x7F
            n >>= 7
            if n:
                ln += bytes([b | 0x80])
            else:
                ln += bytes([b])
                break
        msgs.append(b"\x62" + ln + pay + b"\x48" + bytes([i % 100 + 1]) + b"\x70\x01\x78\x02")  # NextHop blob, Bytes, SrcAS, DstAS
    fblob, foffs = concat_records(frame(msgs))
    want, _, res = oracle.run_batch(fblob, foffs, framed=True, key_mode="aspair")
    with fp.FlowAgg("aspair") as a:
        a.submit(fblob, None, framed=True)
        assert a.stats()["n_records"] == 20000
        assert np.array_equal(a.flush(), want)
        # a stream cut inside the last record: the tail is one bad record
        a.submit(fblob[:-3], None, framed=True)
        st = a.stats()
        assert st["n_records"] == 40000 and st["n_bad"] == 1


def test_offsets_free_host_submits_run_one_behind(fp, oracle, torch_cuda):
    """fa_submit(offsets = NULL) stages its batch and finishes the previous one (include/flowagg.h): a run of such submits, a
    with-offsets submit in between, and every reader of the context (stats, flush_begin/_end, flush) must see exactly the
    records submitted so far -- rows == oracle over the whole stream, whatever call finished which batch."""
    cfg = fp.FaMockerConfig.make(seed=41, flows_per_second=1000, n_src_as=64, n_dst_as=64, framed=True)
    parts = [fp.mocker_host(cfg, i * 30000, 30000) for i in range(5)]
    blob = np.concatenate([b for b, _ in parts])
    offs = np.concatenate([[0]] + [o[1:].astype(np.int64) + sum(len(bb) for bb, _ in parts[:i]) for i, (_, o) in enumerate(parts)]).astype(np.uint32)
    want, _, res = oracle.run_batch(blob, offs, framed=True, key_mode="aspair")
    with fp.FlowAgg("aspair", max_batch_bytes=4 << 20) as a:
        a.submit(parts[0][0], None, framed=True)               # staged only
        a.submit(parts[1][0], None, framed=True)               # finishes 0, stages 1
        assert a.stats()["n_records"] == 60000                 # a reader finishes 1
        a.submit(parts[2][0], None, framed=True)               # staged
        a.submit(*parts[3], framed=True)                       # a with-offsets submit finishes 2 first, then runs
        a.submit(parts[4][0], None, framed=True)               # staged
        a.flush_begin()                                        # finishes 4 before the tables swap
        rows = a.flush_end()
        assert np.array_equal(rows, want)
        assert a.stats()["n_records"] == 150000 and a.stats()["n_bad"] == 0
        # a window of its own after the swap, closed by the synchronous flush
        a.submit(parts[0][0], None, framed=True)
        w0, _, _ = oracle.run_batch(*parts[0], framed=True, key_mode="aspair")
        assert np.array_equal(a.flush(), w0)
        # a piece larger than max_batch_bytes is refused, not truncated
        with pytest.raises(fp.FlowAggError):
            a.submit(blob, None, framed=True)


def test_table_full_is_reported_not_silent(fp, torch_cuda):
    cfg = fp.FaMockerConfig.make(seed=12, flows_per_second=10, addr_mode=2, framed=True)
    buf, offs = fp.mocker_host(cfg, 0, 5000)
    with fp.FlowAgg("5tuple", table_capacity=1024) as a:
        a.submit(buf, offs)
        st = a.stats()
        assert st["n_groups"] == 1024 and st["n_dropped"] == 5000 - 1024
        rows = a.flush(allow_full=True)
        assert len(rows) == 1024 and rows["count"].sum() == 1024


def test_wide_key_table_at_high_load_with_repeats(fp, oracle, torch_cuda):
    """The 5-tuple table (heads + key records, fingerprint-filtered probes): 3 500 distinct keys in 4 096 slots -- long probe
    chains, every probe passing occupied heads of other keys -- each key seen three times over three submits, so the repeat
    path (fingerprint match -> acquire -> key record compare) carries two thirds of the updates.  Rows == oracle, bytewise."""
    cfg = fp.FaMockerConfig.make(seed=77, flows_per_second=10, addr_mode=2, framed=True)
    buf, offs = fp.mocker_host(cfg, 0, 3500)
    blob = np.concatenate([buf, buf, buf])
    o = np.concatenate([offs, offs[1:] + offs[-1], offs[1:] + 2 * offs[-1]]).astype(np.uint32)
    want, _, res = oracle.run_batch(blob, o, framed=True, key_mode="5tuple")
    assert len(want) == 3500 and res["n_bad"] == 0
    with fp.FlowAgg("5tuple", table_capacity=4096) as a:
        for _ in range(3):
            a.submit(buf, offs, framed=True)
        st = a.stats()
        rows = a.flush()
    assert st["n_groups"] == 3500 and st["n_dropped"] == 0
    assert np.array_equal(rows, want)
    assert (rows["count"] == 3).all()


def test_box_topk_single_process_multi_gpu(fp, oracle, torch_cuda):
    if torch_cuda.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    cfg = fp.FaMockerConfig.make(seed=13, flows_per_second=1000, addr_mode=1, framed=True)
    parts = [fp.mocker_host(cfg, p * 50000, 50000) for p in range(2)]
    ctxs = [fp.FlowAgg("srcaddr", device=p, cms=True, cms_width_log2=14, table_capacity=1 << 18) for p in range(2)]
    for c, (b, o) in zip(ctxs, parts):
        c.submit(b, o)
    top = fp.FlowAgg.topk(ctxs, 100)
    buf, offs = fp.mocker_host(cfg, 0, 100000)
    cand, cms, _ = oracle.run_batch(buf, offs, key_mode="srcaddr", cms=(4, 14))
    want = oracle.topk(cms, 4, 14, 4, cand, 100)
    assert np.array_equal(top["key"], want["key"]) and np.array_equal(top["estimate"], want["estimate"])
    top2 = fp.FlowAgg.topk(ctxs, 100)  # repeated query must not double count
    assert np.array_equal(top2["estimate"], want["estimate"])
    for c in ctxs:
        c.close()


# ---------------------------------------------------------------- BASELINE.json configs[1] at full size


def test_config1_full_size_properties(fp, oracle, torch_cuda):
    """100M mocker FlowMessages, (SrcAS,DstAS) group-by, 64k pairs: size-independent properties."""
    torch = torch_cuda
    n_total, slab = 100_000_000, 1 << 24
    cfg = fp.FaMockerConfig.make(seed=1, flows_per_second=250_000, n_src_as=256, n_dst_as=256, framed=True)
    d_buf = torch.empty(slab * 90, dtype=torch.uint8, device="cuda")
    d_off = torch.empty(slab + 1, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    with fp.FlowAgg("aspair", stream=s) as a, fp.FlowAgg("aspair", stream=s) as first:
        done = 0
        first_rows = None
        while done < n_total:
            n = min(slab, n_total - done)
            nb = a.mocker_device(cfg, done, n, d_buf, d_buf.numel(), d_off)
            a.submit_device(d_buf, d_off, n, nb)
            if done == 0:
                # slab 0 again, in a second context fed two halves in reverse order (order/split invariance)
                half = n // 2
                offs = d_off.cpu().numpy().view(np.uint32)
                first.submit_device(d_buf, d_off[half:], n - half, nb)
                first.submit_device(d_buf, d_off, half, int(offs[half]))
                first_rows = first.flush()
                # oracle on the first 2M records of the slab
                m = 2_000_000
                hb = d_buf[: int(offs[m])].cpu().numpy()
                want, _, _ = oracle.run_batch(hb, offs[: m + 1], key_mode="aspair", threads=8)
                first.submit_device(d_buf, d_off, m, int(offs[m]))
                assert np.array_equal(first.flush(), want)
            a.sync()
            done += n
        st = a.stats()
        slab0 = None
        rows = a.flush()
    assert st["n_records"] == n_total and st["n_bad"] == 0 and st["n_groups"] == 65536
    assert rows["count"].sum() == n_total and len(rows) == 65536
    assert rows["count"].min() > 1000 and rows["count"].max() < 2100       # uniform pairs: ~1526 each
    assert first_rows["count"].sum() == slab and (rows["count"] >= first_rows["count"]).all()
    # Bytes ~ U[0,1500), Packets ~ U[0,100): mocker.go:59-60
    assert abs(rows["bytes"].sum() / n_total - 749.5) < 0.5 and abs(rows["packets"].sum() / n_total - 49.5) < 0.05


def test_host_inserter_mirror_rows_equal_oracle(fp, oracle, torch_cuda, tmp_path):
    """flowagg-inserter (C++ mirror of inserter.go over the C ABI): two partitions -> flows_5m TSV."""
    import os
    import subprocess

    from conftest import ROOT

    exe = os.path.join(ROOT, "flow-pipeline_b200", "host", "flowagg-inserter")
    cfg = fp.FaMockerConfig.make(seed=21, flows_per_second=40, n_src_as=5, n_dst_as=5, framed=True)
    files, rows = [], []
    for p in range(2):
        buf, offs = fp.mocker_host(cfg, p * 20000, 20000)
        f = tmp_path / f"claim{p}.bin"
        f.write_bytes(buf.tobytes())
        files.append(str(f))
        want, _, _ = oracle.run_batch(buf, offs, key_mode="flows5m")
        rows.append(want)
    out = tmp_path / "rows.tsv"
    r = subprocess.run([exe, "-claim.file", ",".join(files), "-flush.count", "0", "-flush.dur", "1h", "-out", str(out)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = sorted(tuple(l.split("\t")[2:5] + l.split("\t")[8:11] + [l.split("\t")[1]]) for l in out.read_text().splitlines())
    import time

    want = []
    for part in rows:
        for w in part:
            ts = time.strftime("%Y-%m-%d %H:%M:%S", time.gmtime(int(w["key"][0])))
            want.append((str(w["key"][1]), str(w["key"][2]), f"[{w['key'][3]}]", str(w["bytes"]), str(w["packets"]), str(w["count"]), ts))
    assert got == sorted(want)
    # -flush.box: the closing flush is one exact roll-up over both partitions, already in ORDER BY order
    buf, offs = fp.mocker_host(cfg, 0, 40000)
    merged, _, _ = oracle.run_batch(buf, offs, key_mode="flows5m")
    out2 = tmp_path / "rows_box.tsv"
    r = subprocess.run([exe, "-claim.file", ",".join(files), "-flush.count", "0", "-flush.dur", "1h", "-flush.box", "-out", str(out2)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got2 = [tuple(l.split("\t")[1:5] + l.split("\t")[8:11]) for l in out2.read_text().splitlines()]
    want2 = [(time.strftime("%Y-%m-%d %H:%M:%S", time.gmtime(int(w["key"][0]))), str(w["key"][1]), str(w["key"][2]), f"[{w['key'][3]}]",
              str(w["bytes"]), str(w["packets"]), str(w["count"])) for w in merged]
    assert got2 == want2
    # -offsets.gpu: the slabs go to the library without their offsets (boundaries found on the GPU, one slab behind): same rows
    out2b = tmp_path / "rows_box_nooff.tsv"
    r = subprocess.run([exe, "-claim.file", ",".join(files), "-flush.count", "0", "-flush.dur", "1h", "-flush.box", "-offsets.gpu", "-out", str(out2b)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert out2b.read_text() == out2.read_text()
    # -format rowbinary: what `INSERT INTO flows_5m FORMAT RowBinary` takes (create.sh:70-87), 70 bytes per row
    out3 = tmp_path / "rows_box.bin"
    r = subprocess.run([exe, "-claim.file", ",".join(files), "-flush.count", "0", "-flush.dur", "1h", "-flush.box", "-format", "rowbinary",
                        "-out", str(out3)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rb = np.dtype([("Date", "<u2"), ("Timeslot", "<u4"), ("SrcAS", "<u4"), ("DstAS", "<u4"), ("n0", "u1"), ("EType", "<u4"), ("n1", "u1"),
                   ("MapBytes", "<u8"), ("n2", "u1"), ("MapPackets", "<u8"), ("n3", "u1"), ("MapCount", "<u8"), ("Bytes", "<u8"),
                   ("Packets", "<u8"), ("Count", "<u8")])
    assert rb.itemsize == 70
    got3 = np.frombuffer(out3.read_bytes(), dtype=rb)
    assert len(got3) == len(merged)
    assert np.array_equal(got3["Timeslot"], merged["key"][:, 0]) and np.array_equal(got3["Date"], merged["key"][:, 0] // 86400)
    assert np.array_equal(got3["SrcAS"], merged["key"][:, 1]) and np.array_equal(got3["DstAS"], merged["key"][:, 2])
    assert np.array_equal(got3["EType"], merged["key"][:, 3]) and (got3["n0"] == 1).all() and (got3["n3"] == 1).all()
    for a, b in (("Bytes", "bytes"), ("Packets", "packets"), ("Count", "count"), ("MapBytes", "bytes"), ("MapPackets", "packets"), ("MapCount", "count")):
        assert np.array_equal(got3[a], merged[b]), a


@pytest.mark.parametrize("mode,addr_mode,cms", [("flows5m", 0, False), ("aspair", 0, False), ("srcaddr", 1, True), ("dstport", 0, False)])
def test_hot_keys_take_the_replica_path(fp, oracle, torch_cuda, mode, addr_mode, cms):
    """Skewed keys (the mocker's own 9 AS pairs; Zipf addresses): from the second submit on a context
    sends its updates through the hot-key replicas, folded into the main table before it is read.
    Same rows, same sketch."""
    cfg = fp.FaMockerConfig.make(seed=31, flows_per_second=300, addr_mode=addr_mode, framed=True)
    n = 120_000
    buf, offs = fp.mocker_host(cfg, 0, n)
    want, wcms, res = oracle.run_batch(buf, offs, key_mode=mode, cms=(4, 12) if cms else None)
    with fp.FlowAgg(mode, cms=cms, cms_depth=4, cms_width_log2=12, table_capacity=1 << 18) as a:
        third = n // 3
        for lo, hi in ((0, third), (third, 2 * third), (2 * third, n)):   # three submits: direct, then combined
            a.submit(buf[offs[lo]:offs[hi]], (offs[lo:hi + 1] - offs[lo]).astype(np.uint32))
        st = a.stats()
        assert st["n_records"] == n and st["n_bad"] == 0
        if cms:
            assert np.array_equal(a.cms_read(), wcms)
        assert np.array_equal(a.flush(), want)


def test_topk_ties_are_broken_by_key(fp, oracle, torch_cuda):
    """20 000 distinct addresses with identical weights: thousands of estimates tie at the cut, so the
    device-side selection has to fall back to ordering every candidate (estimate desc, key asc)."""
    msgs = []
    rng = np.random.default_rng(5)
    addrs = rng.permutation(20000)
    for a in addrs:
        addr = bytes.fromhex("20010db800000001") + int(a).to_bytes(8, "big")
        msgs.append(b"\x18\x01" + b"\x32\x10" + addr + b"\x48\x64")  # SamplingRate=1, SrcAddr, Bytes=100
    blob, offs = concat_records(frame(msgs))
    cand, cms, _ = oracle.run_batch(blob, offs, key_mode="srcaddr", cms=(4, 16))
    want = oracle.topk(cms, 4, 16, 4, cand, 50)
    with fp.FlowAgg("srcaddr", cms=True, cms_depth=4, cms_width_log2=16, table_capacity=1 << 16) as a:
        a.submit(blob, offs)
        top = a.topk_local(50)
    assert np.array_equal(top["key"], want["key"]) and np.array_equal(top["estimate"], want["estimate"])
    assert (np.diff(top["estimate"].astype(np.int64)) <= 0).all()


def _nccl_worker(rank, world, port, q):
    import importlib
    import os
    import sys

    import torch
    import torch.distributed as dist

    from conftest import ROOT

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import flow_pipeline_b200 as fp

    par = importlib.import_module("flow-pipeline_b200.parallel")
    cfg = fp.FaMockerConfig.make(seed=13, flows_per_second=1000, addr_mode=1, framed=True)
    n_part, per = 4, 40000
    with fp.FlowAgg("srcaddr", device=rank, cms=True, cms_width_log2=14, table_capacity=1 << 18,
                    stream=torch.cuda.current_stream().cuda_stream) as a, fp.FlowAgg("flows5m", device=rank) as b:
        for p in par.my_partitions(n_part, world, rank):   # Kafka partition p -> rank p mod world
            buf, offs = fp.mocker_host(cfg, p * per, per)
            a.submit(buf, offs)
            b.submit(buf, offs)
        top = par.box_topk(a, 100)                        # NCCL all-reduce of the sketches + merge
        top2 = par.box_topk(a, 100)                       # a second query must not double count
        rows = par.merge_rows(b.flush(keep=True), 4, device=torch.device("cuda", rank))
        # the same roll-up through the hash-partitioned exchange: all-to-all over NCCL, merged on the GPU
        share = par.exchange_rows(b, device=torch.device("cuda", rank))
        assert (fp.row_owner("flows5m", share, world) == rank).all()
    q.put(("share", rank, share.copy()))
    if rank == 0:
        q.put(("main", top.copy(), top2.copy(), rows.copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_one_process_per_gpu_box_topk_and_row_merge_over_nccl(fp, oracle, torch_cuda):
    if torch_cuda.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    import os

    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    shares, main = {}, None
    while len(shares) < 2 or main is None:
        item = q.get(timeout=300)
        if item[0] == "share":
            shares[item[1]] = item[2]
        else:
            main = item[1:]
    top, top2, rows = main
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    cfg = fp.FaMockerConfig.make(seed=13, flows_per_second=1000, addr_mode=1, framed=True)
    buf, offs = fp.mocker_host(cfg, 0, 4 * 40000)
    cand, cms, _ = oracle.run_batch(buf, offs, key_mode="srcaddr", cms=(4, 14))
    want = oracle.topk(cms, 4, 14, 4, cand, 100)
    for t in (top, top2):
        assert np.array_equal(t["key"], want["key"]) and np.array_equal(t["estimate"], want["estimate"])
    want_rows, _, _ = oracle.run_batch(buf, offs, key_mode="flows5m")
    assert np.array_equal(rows, want_rows)
    both = np.concatenate([shares[0], shares[1]])
    assert len(both) == len(want_rows)
    order = np.lexsort(tuple(both["key"][:, i] for i in reversed(range(4))))
    assert np.array_equal(both[order], want_rows)


def _random_wire_messages(seed, n):
    """Wire-format fuzz that needs no schema: random tags / wire types / lengths, plus pure noise."""
    rng = np.random.default_rng(seed)

    def varint(v):
        out = bytearray()
        while True:
            b = v & 0x7F
            v >>= 7
            if v:
                out.append(b | 0x80)
            else:
                out.append(b)
                return bytes(out)

    known = [1, 2, 3, 4, 6, 7, 9, 10, 11, 14, 15, 20, 21, 22, 30, 38, 100, 101]
    msgs = []
    for i in range(n):
        kind = rng.integers(0, 10)
        if kind == 0:  # pure noise
            msgs.append(bytes(rng.integers(0, 256, int(rng.integers(0, 60)), dtype=np.uint8)))
            continue
        parts = []
        for _ in range(int(rng.integers(0, 14))):
            num = int(rng.choice(known)) if rng.random() < 0.7 else int(rng.choice([5, 8, 16, 127, 128, 2047, 2048, 1 << 20, (1 << 29) - 1, 5, 8, 16, 0, 1 << 29]))
            wt = int(rng.choice([0] * 12 + [2] * 8 + [1, 1, 5, 5, 3, 3, 4, 6, 7]))
            tag = varint((num << 3) | wt)
            if rng.random() < 0.03:
                tag = tag[:-1] + bytes([tag[-1] | 0x80]) + b"\x00"  # over-long tag encoding
            if wt == 0:
                bits = int(rng.choice([1, 7, 8, 14, 21, 28, 29, 35, 36, 49, 63, 64]))
                body = varint(int(rng.integers(0, 1 << 62)) >> (62 - min(bits, 62)) if bits < 64 else (1 << 64) - 1)
                if rng.random() < 0.03:
                    body = b"\xff" * 9 + bytes([int(rng.choice([0, 1, 2, 0x7f]))])
            elif wt == 2:
                ln = int(rng.choice([0, 1, 4, 15, 16, 17, 40, 130]))
                payload = bytes(rng.integers(0, 256, ln, dtype=np.uint8)) if rng.random() < 0.7 else bytes(ln)
                if num in (100, 101) and rng.random() < 0.6:
                    payload = "héllo→日本"[: ln or 1].encode()[:ln]
                    ln = len(payload)
                body = varint(ln) + payload
            elif wt == 1:
                body = bytes(8)
            elif wt == 5:
                body = bytes(4)
            elif wt == 3:
                inner = varint((int(rng.integers(1, 50)) << 3) | 0) + varint(int(rng.integers(0, 1000)))
                body = inner + varint((num << 3) | 4) if rng.random() < 0.8 else inner
            else:
                body = b""
            parts.append(tag + body)
        m = b"".join(parts)
        if rng.random() < 0.1 and len(m) > 1:
            m = m[: int(rng.integers(1, len(m)))]
        msgs.append(m)
    return msgs


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_schemaless_wire_fuzz_gpu_equals_oracle(fp, oracle, torch_cuda, seed):
    msgs = _random_wire_messages(seed, 30000)
    for framed in (False, True):
        blob, offs = concat_records(frame(msgs) if framed else msgs)
        ora = oracle.decode_columns(blob, offs, framed=framed)
        gpu, st = gpu_columns(fp, blob, offs, framed=framed)
        assert 0.2 < ora["valid"].mean() < 0.95          # the fuzz produces both kinds
        assert_columns_equal(gpu, ora, len(msgs))
        assert st["n_bad"] == int((ora["valid"] == 0).sum())
    # and through the fused kernels, every key mode
    blob, offs = concat_records(frame(msgs))
    for mode in ("flows5m", "srcaddr", "5tuple"):
        want, _, res = oracle.run_batch(blob, offs, key_mode=mode, scale=True)
        with fp.FlowAgg(mode, scale_sampling=True, table_capacity=1 << 17) as a:
            a.submit(blob, offs)
            st = a.stats()
            assert (st["n_bad"], st["n_nokey"]) == (res["n_bad"], res["n_nokey"])
            assert np.array_equal(a.flush(), want)


def _addr_text(oracle, cols, name, i):
    """net.IP.String() of the column value (inserter.go:131-140); longer than 16 bytes: "?" + hex of the 16 kept bytes + ".."."""
    ln = int(cols[name + "_len"][i])
    if ln > 16:
        return "?" + bytes(cols[name][i][:16]).hex() + ".."
    return oracle.ip_string(bytes(cols[name][i][:ln]))


def test_host_inserter_copy_sink_and_metrics_scrape(fp, oracle, torch_cuda, fuzz_2k, tmp_path):
    """-sink copy: the inserter's 14-column row (inserter.go:51-66,142-157) as a Postgres COPY script, checked field by field
    against the oracle; -metrics: insert_count (inserter.go:44-49) and the bad-record counter scraped from the endpoint
    (inserter.go:69-73) equal the oracle's counts, and the GPU-busy counter is alive."""
    import os
    import socket
    import subprocess
    import time
    import urllib.request

    from conftest import ROOT

    exe = os.path.join(ROOT, "flow-pipeline_b200", "host", "flowagg-inserter")
    g = fuzz_2k
    msgs = [bytes(g["blob"][g["offsets"][i]:g["offsets"][i + 1]]) for i in range(len(g["offsets"]) - 1)]
    msgs += [b"\x32\x04" + bytes([192, 168, 1, 1]) + b"\x3a\x10" + bytes(10) + b"\xff\xff" + bytes([10, 0, 0, 7]) + b"\x48\x05" + b"\xb0\x02\x8e\xb0\xdf\xf3\x05",
             b"\x32\x11" + bytes(range(17)) + b"\x48\x07", b"\xff\xff\xff", b"\x48"]       # a 17-byte address; two undecodable messages
    blob, offs = concat_records(frame(msgs))
    f = tmp_path / "claim.bin"
    f.write_bytes(blob.tobytes())
    out = tmp_path / "rows.copy"
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    proc = subprocess.Popen([exe, "-claim.file", str(f), "-sink", "copy", "-flush.count", "0", "-out", str(out), "-metrics", "-metrics.addr",
                             f":{port}", "-linger", "60s"], stderr=subprocess.PIPE, text=True)
    cols = oracle.decode_columns(blob, offs, framed=True)
    n_good, n_bad = int(cols["valid"].sum()), int((~cols["valid"].astype(bool)).sum())
    assert n_bad >= 2
    scraped = {}
    try:
        deadline = time.time() + 50
        while time.time() < deadline:
            try:
                body = urllib.request.urlopen(f"http://127.0.0.1:{port}/metrics", timeout=2).read().decode()
            except OSError:
                time.sleep(0.2)
                continue
            scraped = {l.split()[0]: float(l.split()[1]) for l in body.splitlines() if l and not l.startswith("#")}
            if scraped.get("flowagg_records_total", 0) == len(msgs):
                break
            time.sleep(0.2)
    finally:
        proc.terminate()
        err = proc.communicate(timeout=30)[1]
    assert scraped.get("flowagg_records_total") == len(msgs), (scraped, err[-500:])
    assert scraped["insert_count"] == n_good and scraped["flowagg_bad_records_total"] == n_bad
    assert scraped["flowagg_gpu_busy_seconds_total"] > 0
    lines = out.read_text().splitlines()
    assert lines[0].startswith("COPY flows (date_inserted, time_flow, type, sampling_rate, src_ip, dst_ip, bytes, packets, src_port, dst_port, "
                               "etype, proto, src_as, dst_as) FROM stdin;")
    body = lines[1:]
    if body and body[-1] == "\\.":            # the end-of-data marker is written on a clean shutdown
        body = body[:-1]
    want = []
    for i in range(len(msgs)):
        if not cols["valid"][i]:
            continue
        ty = int(cols["type"][i])
        ty = ty - (1 << 32) if ty >= (1 << 31) else ty
        tf = time.strftime("%Y-%m-%d %H:%M:%S+00", time.gmtime(int(cols["time_flow_start"][i]) & ((1 << 63) - 1))) if int(cols["time_flow_start"][i]) < (1 << 40) else None
        want.append((tf, [str(ty), str(int(cols["sampling_rate"][i])), _addr_text(oracle, cols, "src_addr", i), _addr_text(oracle, cols, "dst_addr", i),
                          str(int(cols["bytes"][i])), str(int(cols["packets"][i])), str(int(cols["src_port"][i])), str(int(cols["dst_port"][i])),
                          str(int(cols["etype"][i])), str(int(cols["proto"][i])), str(int(cols["src_as"][i])), str(int(cols["dst_as"][i]))]))
    assert len(body) == len(want) == n_good
    for line, (tf, rest) in zip(body, want):
        parts = line.split("\t")
        assert len(parts) == 14 and parts[2:] == rest
        assert len(parts[0]) == 22 and parts[0].endswith("+00")          # date_inserted: NOW() at the time of the copy
        if tf is not None:
            assert parts[1] == tf
    assert any("\t192.168.1.1\t10.0.0.7\t" in l for l in body) and any("\t?000102030405060708090a0b0c0d0e0f..\t" in l for l in body)


def test_host_inserter_mirror_row_sink_matches_the_inserters_14_columns(fp, oracle, torch_cuda, fuzz_2k, tmp_path):
    """-sink rows: one row per decoded flow, columns and IP formatting of inserter.go:131-157
    (net.IP.String incl. v4, v4-mapped, "?hex" and the "<nil>" -> 0.0.0.0 patch)."""
    import os
    import subprocess

    from conftest import ROOT

    exe = os.path.join(ROOT, "flow-pipeline_b200", "host", "flowagg-inserter")
    g = fuzz_2k
    msgs = [bytes(g["blob"][g["offsets"][i]:g["offsets"][i + 1]]) for i in range(len(g["offsets"]) - 1)]
    msgs += [b"\x32\x04" + bytes([192, 168, 1, 1]) + b"\x3a\x10" + bytes(10) + b"\xff\xff" + bytes([10, 0, 0, 7]) + b"\x48\x05",
             b"\x32\x10" + bytes.fromhex("20010db8000000010000000000000080") + b"\x3a\x03abc" + b"\xb0\x02\x8e\xb0\xdf\xf3\x05"]
    blob, offs = concat_records(frame(msgs))
    f = tmp_path / "claim.bin"
    f.write_bytes(blob.tobytes())
    out = tmp_path / "rows.tsv"
    r = subprocess.run([exe, "-claim.file", str(f), "-sink", "rows", "-flush.count", "0", "-out", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = out.read_text().splitlines()
    cols = oracle.decode_columns(blob, offs, framed=True)
    want = []
    for i in range(len(msgs)):
        if not cols["valid"][i]:
            continue
        ty = int(cols["type"][i])
        ty = ty - (1 << 32) if ty >= (1 << 31) else ty
        sip, dip = _addr_text(oracle, cols, "src_addr", i), _addr_text(oracle, cols, "dst_addr", i)
        want.append("\t".join(["NOW()", str(int(cols["time_flow_start"][i])), str(ty), str(int(cols["sampling_rate"][i])), sip, dip,
                               str(int(cols["bytes"][i])), str(int(cols["packets"][i])), str(int(cols["src_port"][i])),
                               str(int(cols["dst_port"][i])), str(int(cols["etype"][i])), str(int(cols["proto"][i])),
                               str(int(cols["src_as"][i])), str(int(cols["dst_as"][i]))]))
    assert got == want
    assert any("\t192.168.1.1\t10.0.0.7\t" in l for l in got) and any("\t2001:db8:0:1::80\t?616263\t" in l for l in got)


def test_repeated_flushes_speculative_row_count(fp, oracle, torch_cuda):
    """From the second flush on the row count is guessed from the previous flush (one synchronisation);
    a roll-up that grows, shrinks or stays must still come back exact and ordered."""
    sizes = [(3, 2000), (3, 2000), (64, 50000), (64, 50000), (2, 500), (200, 120000), (200, 120000)]
    with fp.FlowAgg("flows5m", table_capacity=1 << 17) as a:
        first = 0
        for n_as, n in sizes:
            cfg = fp.FaMockerConfig.make(seed=17, flows_per_second=37, n_src_as=n_as, n_dst_as=n_as, framed=True)
            buf, offs = fp.mocker_host(cfg, first, n)
            first += n
            want, _, _ = oracle.run_batch(buf, offs, key_mode="flows5m")
            a.submit(buf, offs)
            got = a.flush()
            assert np.array_equal(got, want), (n_as, n, len(got), len(want))
            assert a.stats()["n_groups"] == 0


_STRIDE_SNIPPET = r"""
import sys, numpy as np
sys.path.insert(0, {root!r})
import flow_pipeline_b200 as fp
from oracle import oracle
cfg = fp.FaMockerConfig.make(seed=9, flows_per_second=250_000, n_src_as=64, n_dst_as=64, framed=True)
buf, offs = fp.mocker_host(cfg, {first}, 70_001)          # 273 full tiles of 256 records and a ragged tail
for mode in ("aspair", "flows5m"):
    want, _, _ = oracle.run_batch(buf, offs, key_mode=mode)
    with fp.FlowAgg(mode, table_capacity=1 << 16) as a:
        a.submit(buf, offs)
        assert a.stats()["n_bad"] == 0
        assert np.array_equal(a.flush(), want), mode
print("ok")
"""


@pytest.mark.parametrize("stride,first", [("", 0), ("", 300_000_000), ("2", 0), ("4", 300_000_000), ("8", 0), ("7", 300_000_000), ("37", 0)])
def test_record_to_lane_stride_is_result_neutral(torch_cuda, stride, first):
    """The host picks how many records lie between neighbouring lanes of a warp (shared-memory bank conflicts of
    near-constant record sizes: SequenceNum >= 2^28 makes every mocker record 85-86 bytes, 11 lanes per bank).
    Whatever it picks -- or FA_LANE_STRIDE forces -- every record is decoded exactly once: same rows."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env.pop("FA_LANE_STRIDE", None)
    if stride:
        env["FA_LANE_STRIDE"] = stride
    r = subprocess.run([sys.executable, "-c", _STRIDE_SNIPPET.format(root=root, first=first)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def _all_ones_port_records(n):
    """SrcPort = 0xFFFFFFFF (a legal uint32): the one key that equals the table's empty marker."""
    return [b"\x48" + bytes([1 + i % 100]) + b"\x50\x02" + b"\xa8\x01\xff\xff\xff\xff\x0f" for i in range(n)]


@pytest.mark.parametrize("mode", ["aspair", "srcaddr", "5tuple", "srcport", "flows5m"])
def test_merge_rows_is_the_summing_merge(fp, oracle, torch_cuda, mode):
    """fa_merge_rows: rows that are already aggregates are summed into the table by key (SummingMergeTree,
    create.sh:88-90) -- from host memory, from device memory, and filtered by key owner."""
    cfg = fp.FaMockerConfig.make(seed=17, flows_per_second=40, n_src_as=32, n_dst_as=32, addr_mode=1, framed=True)
    n = 60_000
    buf, offs = fp.mocker_host(cfg, 0, n)
    extra = frame(_all_ones_port_records(7)) if mode == "srcport" else []
    if extra:
        eb, eo = concat_records(extra)
        buf = np.concatenate([buf[: offs[n]], eb])
        offs = np.concatenate([offs, (eo[1:] + offs[n]).astype(np.uint32)])
    total = len(offs) - 1
    half = total // 2
    want, _, _ = oracle.run_batch(buf, offs, key_mode=mode)
    lo = (buf[: offs[half]], offs[: half + 1])
    hi = (buf[offs[half]: offs[total]], (offs[half:] - offs[half]).astype(np.uint32))
    with fp.FlowAgg(mode, table_capacity=1 << 18) as a, fp.FlowAgg(mode, table_capacity=1 << 18) as b:
        a.submit(*lo)
        b.submit(*hi)
        part = a.flush(sort=False)
        b.merge_rows(part)                                   # host rows
        assert np.array_equal(b.flush(keep=True), want)
        # again from device memory, split over three "owners": every row lands exactly once
        b.reset()
        b.submit(*hi)
        d_rows = torch_cuda.from_numpy(part.view(np.int64).reshape(len(part), -1).copy()).cuda()
        for owner in range(3):
            b.merge_rows(d_rows, n=len(part), owner=owner, n_owners=3)
        b.sync()
        assert np.array_equal(b.flush(), want)
        owners = fp.row_owner(mode, part, 3)
        assert set(np.unique(owners)) <= {0, 1, 2} and (len(part) < 30 or len(np.unique(owners)) == 3)


@pytest.mark.parametrize("mode", ["flows5m", "5tuple", "srcport"])
def test_flush_box_is_the_exact_group_by_of_every_context(fp, oracle, torch_cuda, mode):
    """fa_flush_box: contexts that saw different partitions hold partial sums of the same keys; the hash-partitioned
    exchange (peer loads across GPUs when the box has several) leaves the exact roll-up of everything, in
    ORDER BY order, and empty contexts."""
    n_dev = torch_cuda.cuda.device_count()
    n_ctx = max(3, min(n_dev, 4))
    cfg = fp.FaMockerConfig.make(seed=23, flows_per_second=30, n_src_as=16, n_dst_as=16, addr_mode=1, framed=True)
    per = 30_000
    buf, offs = fp.mocker_host(cfg, 0, n_ctx * per)
    if mode == "srcport":
        eb, eo = concat_records(frame(_all_ones_port_records(5)))
        buf = np.concatenate([buf[: offs[-1]], eb])
        offs = np.concatenate([offs, (eo[1:] + offs[-1]).astype(np.uint32)])
    want, _, _ = oracle.run_batch(buf, offs, key_mode=mode)
    ctxs = [fp.FlowAgg(mode, device=i % n_dev, table_capacity=1 << 18) for i in range(n_ctx)]
    try:
        total = len(offs) - 1
        cuts = [min(i * per, total) for i in range(n_ctx)] + [total]
        for i, c in enumerate(ctxs):
            a, b = cuts[i], cuts[i + 1]
            c.submit(buf[offs[a]: offs[b]], (offs[a: b + 1] - offs[a]).astype(np.uint32))
        rows = fp.FlowAgg.flush_box(ctxs)
        assert np.array_equal(rows, want)
        assert all(len(c.flush()) == 0 for c in ctxs)         # every context was reset
        ctxs[0].submit(buf[: offs[100]], offs[:101])           # and is usable again
        w2, _, _ = oracle.run_batch(buf[: offs[100]], offs[:101], key_mode=mode)
        assert np.array_equal(fp.FlowAgg.flush_box(ctxs), w2)
    finally:
        for c in ctxs:
            c.close()


def test_readme_sample_rows_on_the_gpu(fp, oracle, torch_cuda):
    """The README's published flows_5m rows (README.md:180-183), same hand-made flows as the oracle's own pin test."""
    from test_oracle_golden import _flow

    t0 = 1584912398
    msgs = [_flow(t0 + i % 2, 65001, d, b, p) for d, flows in ((65000, [(1000, 50), (1000, 50), (900, 50), (30, 2)]),
                                                              (65001, [(1000, 90), (900, 90), (35, 10)]),
                                                              (65002, [(1000, 48)] * 4 + [(400, 48), (420, 48)]))
            for i, (b, p) in enumerate(flows)]
    blob, offs = concat_records(msgs)
    want, _, _ = oracle.run_batch(blob, offs, key_mode="flows5m", framed=False)
    with fp.FlowAgg("flows5m") as a:
        a.submit(blob, offs, framed=False)
        rows = a.flush()
    assert np.array_equal(rows, want)
    assert [(int(r["key"][0]), int(r["key"][2]), int(r["bytes"]), int(r["packets"]), int(r["count"])) for r in rows] == \
        [(1584912300, 65000, 2930, 152, 4), (1584912300, 65001, 1935, 190, 3), (1584912300, 65002, 4820, 288, 6)]


@pytest.mark.gpu
def test_rollup_boundary_vectors_on_the_gpu(fp, oracle, torch_cuda, rollup_boundaries):
    """tests/golden/rollup_boundaries.json through the C ABI: slot edges, day rollover, DateTime wrap, UInt64 wrap of the sums,
    the all-ones key (the table's reserved side slot) -- rows bytewise equal to the independently computed expectation."""
    from conftest import rollup_case_rows

    for case in rollup_boundaries["cases"]:
        msgs = [bytes.fromhex(h) for h in case["messages_hex"]]
        for framed in (False, True):
            blob, offs = concat_records(frame(msgs) if framed else msgs)
            with fp.FlowAgg("flows5m", device=0) as a:
                a.submit(blob, offs, framed=framed)
                rows = a.flush()
                got = [(tuple(int(x) for x in r["key"][:4]), int(r["bytes"]), int(r["packets"]), int(r["count"])) for r in rows]
                assert got == rollup_case_rows(case), (case["name"], framed)
                # the same flows again, twice, into the emptied table: every sum doubles modulo 2^64
                a.submit(blob, offs, framed=framed)
                a.submit(blob, offs, framed=framed)
                rows = a.flush()
                got = [(tuple(int(x) for x in r["key"][:4]), int(r["bytes"]), int(r["packets"]), int(r["count"])) for r in rows]
                want = [(k, (2 * b) & (2 ** 64 - 1), (2 * p) & (2 ** 64 - 1), 2 * c) for k, b, p, c in rollup_case_rows(case)]
                assert got == want, (case["name"], framed)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["aspair", "flows5m", "srcaddr", "5tuple", "dstport"])
def test_asynchronous_flush_equals_the_synchronous_one(fp, oracle, torch_cuda, mode):
    """fa_flush_begin / fa_flush_end: the filled table is swapped for a spare and drained on a side stream while the next
    submits already aggregate into the spare.  Window by window the rows equal the oracle's (and fa_flush's), whatever is
    submitted between the two halves; row-count guesses that are too small, caller arrays that are too small, peeks and
    unsorted dumps included."""
    cfg = fp.FaMockerConfig.make(seed=5, flows_per_second=50, n_src_as=40, n_dst_as=40, framed=True, addr_mode=fp.FA_ADDR_ZIPF24 if mode == "srcaddr" else 0)
    windows = [fp.mocker_host(cfg, i * 30000, n) for i, n in enumerate((30000, 500, 30000, 30000, 7))]
    want = [oracle.run_batch(b, o, key_mode=mode)[0] for b, o in windows]
    with fp.FlowAgg(mode, device=0, table_capacity=1 << 17) as a:
        a.submit(*windows[0], framed=True)
        a.flush_begin()
        a.submit(*windows[1], framed=True)              # lands in the spare table while window 0 drains
        st = a.stats()                                  # allowed between the halves: counts the CURRENT table
        assert st["n_groups"] == len(want[1])
        assert np.array_equal(a.flush_end(), want[0])
        a.flush_begin()                                 # window 1 is tiny: the next guess (its row count + 1/8) is far too small ...
        a.submit(*windows[2], framed=True)
        assert np.array_equal(a.flush_end(), want[1])
        a.flush_begin()                                 # ... for window 2: the exact path re-sorts from scratch
        a.submit(*windows[3], framed=True)
        small = np.empty(3, dtype=fp.ROW_DTYPE)         # a caller array that is too small: rows are kept, the call repeats
        assert np.array_equal(a.flush_end(out=small), want[2])
        assert np.array_equal(a.flush(), want[3])       # the synchronous flush on the swapped-in table
        a.submit(*windows[4], framed=True)
        a.flush_begin(keep=True)                        # a peek: drained in place by flush_end
        assert np.array_equal(a.flush_end(), want[4])
        a.flush_begin(sort=False)
        rows = a.flush_end()
        assert np.array_equal(rows[np.lexsort([rows["key"][:, k] for k in range(11, -1, -1)])], want[4])
        assert len(a.flush()) == 0
    # one flush in flight: a second begin, a plain flush or a box query in between is refused, not queued
    with fp.FlowAgg(mode, device=0, table_capacity=1 << 17) as a:
        a.submit(*windows[0], framed=True)
        a.flush_begin()
        with pytest.raises(fp.FlowAggError):
            a.flush_begin()
        with pytest.raises(fp.FlowAggError):
            a.flush()
        assert np.array_equal(a.flush_end(), want[0])


@pytest.mark.gpu
def test_topk_only_keeps_a_bounded_candidate_set(fp, oracle, torch_cuda):
    """FA_CFG_TOPK_ONLY (BASELINE configs[2] as a real sketch workload): the sketch is bit-equal to the oracle's; the group
    table only ever holds keys whose estimate reached 1/(64 K) of the weight seen -- far fewer than the distinct keys of the
    stream, whatever their number -- and the top-K it yields is the oracle's (exact candidates) up to >= 99 % recall, with
    identical estimates for the keys both report."""
    cfg = fp.FaMockerConfig.make(seed=31, flows_per_second=5000, addr_mode=fp.FA_ADDR_ZIPF24, framed=True)
    d, wl, K, n_sub, per = 4, 18, 200, 6, 500_000
    slabs = [fp.mocker_host(cfg, i * per, per) for i in range(n_sub)]
    buf = np.concatenate([b for b, _ in slabs])
    offs = np.concatenate([[0]] + [o[1:].astype(np.int64) + sum(len(bb) for bb, _ in slabs[:i]) for i, (_, o) in enumerate(slabs)]).astype(np.uint32)
    cand, cms, _ = oracle.run_batch(buf, offs, key_mode="srcaddr", cms=(d, wl))
    want = oracle.topk(cms, d, wl, 4, cand, K)
    assert len(cand) > 200_000                                                   # distinct keys of the stream
    with fp.FlowAgg("srcaddr", topk_only=True, topk_k=K, cms_depth=d, cms_width_log2=wl) as a:
        for b, o in slabs:                                                       # several submits: the table is pruned after each
            a.submit(b, o)
        assert np.array_equal(a.cms_read(), cms)                                 # counters bit-equal, whatever the candidates
        st = a.stats()
        assert 0 < st["n_groups"] <= 4 * 64 * 256 and st["n_groups"] < len(cand) // 8   # bounded: 64 K' slots' worth, K' = 2^ceil(log2 K)
        top = a.topk_local(K)
    assert len(top) == K
    got = {bytes(r["key"][:4].tobytes()): int(r["estimate"]) for r in top}
    exp = {bytes(r["key"][:4].tobytes()): int(r["estimate"]) for r in want}
    common = set(got) & set(exp)
    assert len(common) >= 0.99 * K
    assert all(got[k] == exp[k] for k in common)                                 # same sketch, same estimate
    # a context may not combine the flag with an exact roll-up's options, and needs an address key
    for bad in (dict(key_mode="aspair", topk_only=True), dict(key_mode="srcaddr", topk_only=True, scale_sampling=True),
                dict(key_mode="srcaddr", topk_only=True, table_capacity=1024)):
        with pytest.raises(fp.FlowAggError):
            fp.FlowAgg(**bad)
