"""CPU: host-side logic and the C-ABI surface (no compute calls -- there is no GPU here)."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_every_symbol_of_the_header(fp):
    hdr = open(os.path.join(ROOT, "include", "flowagg.h")).read()
    declared = set(re.findall(r"^(?:int|void|const char \*)\s*\*?(fa_[a-z_0-9]+)\s*\(", hdr, re.M))
    assert len(declared) >= 20
    import ctypes as C

    L = C.CDLL(fp.lib_path())
    for name in sorted(declared):
        assert hasattr(L, name), f"libflowagg.so does not export {name}"
    assert declared == set(__import__("flow_pipeline_b200").flowagg.exported_symbols()), "binding and header diverge"
    L.fa_build_info.restype = C.c_char_p
    assert b"sm_100a" in L.fa_build_info()


def test_library_carries_sm100a_code_only(fp):
    import subprocess

    out = subprocess.run(["cuobjdump", "-lelf", fp.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_key_record_stores_are_whole_sectors(fp):
    """The wide-key (5-tuple) table writes a claimed slot's 64-byte key record as two 256-bit stores -- one whole DRAM sector
    each (SASS STG.E.ENL2.256, sm_100 only), so L2 never fetches a sector it is about to overwrite.  ptxas 12.9 silently
    narrows such a store to its first element inside a __noinline__ device function; this pins what the shipped SASS does:
    every claim site (the release store of the slot head is the only MEMBAR.ALL.GPU in these kernels) writes its key record
    either as two 256-bit stores or -- the one out-of-line path -- as four 128-bit stores, never less."""
    import subprocess
    from collections import Counter

    sass = subprocess.run(["cuobjdump", "-sass", fp.lib_path()], capture_output=True, text=True).stdout
    fn, claims, st256, st128 = None, Counter(), Counter(), Counter()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
        elif "MEMBAR.ALL.GPU" in line:
            claims[fn] += 1
        elif "STG.E.ENL2.256" in line:
            st256[fn] += 1
        elif re.search(r"\bSTG\.E\.128\b", line):
            st128[fn] += 1
    wide = [f for f in claims if "AggConsumerILi4E" in f or "k_add_rowsILi11E" in f or "k_aggregate_columnsILi4E" in f]
    assert len(wide) >= 10 and set(claims) == set(wide), sorted(set(claims) ^ set(wide))
    for f in wide:
        assert st256[f] % 2 == 0 and st128[f] % 4 == 0 and st256[f] // 2 + st128[f] // 4 == claims[f], (f, claims[f], st256[f], st128[f])
        assert st256[f] >= 2, f                                  # the in-line (hot) site is the 256-bit one


def test_no_gpu_means_loud_failure_not_fallback(fp):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(fp.FlowAggError) as e:
        fp.FlowAgg()
    assert "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "flow-pipeline_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                bad = re.search(r"(^\s*(from|import)\s+oracle\b)|(#include\s+\"[^\"]*oracle)|liboracle|fo_[a-z_]+\s*\(", src, re.M)
                assert not bad, f"{f} links or calls the oracle: {bad.group(0)}"


def test_mocker_host_is_bytewise_proto_marshal(fp, mocker_10k):
    # golden blob = upb SerializeToString of the same fields (tests/golden/make_golden.py)
    cfg = fp.FaMockerConfig.make(seed=1, flows_per_second=20, framed=False)
    buf, offs = fp.mocker_host(cfg, 0, 10000)
    assert np.array_equal(buf, mocker_10k["blob"]) and np.array_equal(offs, mocker_10k["offsets"])


def test_mocker_framed_and_modes_decode_with_oracle(fp, oracle):
    for mode, n_as in ((0, 3), (1, 256), (2, 256)):
        cfg = fp.FaMockerConfig.make(seed=7, flows_per_second=1000, n_src_as=n_as, n_dst_as=n_as, addr_mode=mode, framed=True)
        buf, offs = fp.mocker_host(cfg, 12345, 4000)
        cols = oracle.decode_columns(buf, offs, framed=True)
        assert cols["valid"].all()
        assert (cols["sampling_rate"] == 1).all() and (cols["etype"] == 0x86DD).all() and (cols["proto"] == 0).all()
        assert cols["bytes"].max() < 1500 and cols["packets"].max() < 100            # mocker.go:59-60
        assert cols["src_as"].min() >= 65000 and cols["src_as"].max() < 65000 + n_as  # mocker.go:61,79
        assert np.array_equal(cols["sequence_num"], np.arange(12345, 12345 + 4000, dtype=np.uint32))
        assert (cols["time_received"] == 1584912398 + np.arange(12345, 16345) // 1000).all()
        assert (cols["src_addr"][:, :8] == np.frombuffer(bytes.fromhex("20010db800000001"), dtype=np.uint8)).all()
        if mode == 2:  # unique 5-tuples
            assert len({bytes(a) for a in cols["src_addr"]}) == 4000
        # a different window of the same stream is the same bytes (counter-based generator)
        buf2, offs2 = fp.mocker_host(cfg, 12345 + 1000, 100)
        assert np.array_equal(buf2, buf[offs[1000]:offs[1100]])


def test_topk_merge_is_pure_host_arithmetic(fp):
    a = np.zeros(3, dtype=fp.HH_DTYPE)
    b = np.zeros(3, dtype=fp.HH_DTYPE)
    a["key"][:, 0] = [1, 2, 3]; a["estimate"] = [50, 40, 10]
    b["key"][:, 0] = [2, 9, 8]; b["estimate"] = [40, 45, 45]
    out = fp.FlowAgg.topk_merge([a, b], 1, 4)
    assert list(out["estimate"]) == [50, 45, 45, 40]
    assert list(out["key"][:, 0]) == [1, 8, 9, 2]  # ties by key ascending, duplicate key 2 kept once


def test_sum_rows_by_key(fp):
    import importlib

    par = importlib.import_module("flow-pipeline_b200.parallel")
    r = np.zeros(4, dtype=fp.ROW_DTYPE)
    r["key"][:, 0] = [5, 1, 5, 1]; r["key"][:, 1] = [0, 2, 0, 3]
    r["bytes"] = [1, 2, 3, 2**64 - 1]; r["packets"] = 1; r["count"] = 1
    out = par.sum_rows_by_key(r, 2)
    assert [tuple(k[:2]) for k in out["key"]] == [(1, 2), (1, 3), (5, 0)]
    assert list(out["bytes"]) == [2, 2**64 - 1, 4] and list(out["count"]) == [1, 1, 2]


def test_oracle_threads_agree(oracle, fp):
    cfg = fp.FaMockerConfig.make(seed=3, flows_per_second=100, n_src_as=16, n_dst_as=16, framed=True)
    buf, offs = fp.mocker_host(cfg, 0, 50000)
    r1, c1, _ = oracle.run_batch(buf, offs, key_mode="flows5m", cms=(4, 10), threads=1)
    r8, c8, _ = oracle.run_batch(buf, offs, key_mode="flows5m", cms=(4, 10), threads=8)
    assert np.array_equal(r1, r8) and np.array_equal(c1, c8)


def test_host_inserter_mirror_flags_and_consume_loop(fp, tmp_path):
    """The C++ mirror of inserter.go: reference flag names parse, one ConsumeClaim per partition
    walks its claim and marks every message (dry run: no GPU here)."""
    import subprocess

    exe = os.path.join(ROOT, "flow-pipeline_b200", "host", "flowagg-inserter")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.dirname(exe)], check=True)
    cfg = fp.FaMockerConfig.make(seed=1, flows_per_second=20, framed=True)
    files = []
    for p in range(2):
        buf, _ = fp.mocker_host(cfg, p * 3000, 3000)
        f = tmp_path / f"claim{p}.bin"
        f.write_bytes(buf.tobytes())
        files.append(str(f))
    ref_flags = ["-loglevel", "info", "-metrics.addr", ":8081", "-metrics.path", "/metrics", "-kafka.version", "2.1.1",
                 "-kafka.topic", "flows", "-kafka.brokers", "kafka:9092", "-kafka.group", "postgres-inserter",
                 "-flush.dur", "5s", "-flush.count", "1000", "-postgres.user", "postgres", "-postgres.pass", "x",
                 "-postgres.host", "127.0.0.1", "-postgres.port", "5432", "-postgres.dbname", "postgres"]
    r = subprocess.run([exe, *ref_flags, "-claim.file", ",".join(files), "-dry-run"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "6000 messages marked over 2 partitions" in r.stderr
    assert subprocess.run([exe, "-no.such.flag", "1"], capture_output=True).returncode == 2
    # without a GPU the real path must fail loudly, not fall back
    import torch

    if not torch.cuda.is_available():
        r = subprocess.run([exe, "-claim.file", files[0]], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr


def test_row_owner_is_a_balanced_deterministic_partition(fp, oracle):
    """fa_row_owner (pure host function of the library): the hash partition of the box-wide exchange.  Same key ->
    same owner whatever else the row holds, owners cover [0, n) evenly, invalid arguments are rejected."""
    import importlib

    par = importlib.import_module("flow-pipeline_b200.parallel")
    cfg = fp.FaMockerConfig.make(seed=5, flows_per_second=100, addr_mode=1, framed=True)
    buf, offs = fp.mocker_host(cfg, 0, 40000)
    for mode, kw in (("srcaddr", 4), ("5tuple", 11), ("aspair", 2), ("srcport", 1)):
        rows, _, _ = oracle.run_batch(buf, offs, key_mode=mode)
        rows = rows.view(fp.ROW_DTYPE)
        for n in (2, 3, 8):
            own = fp.row_owner(mode, rows, n)
            assert own.max() < n
            if len(rows) > 1000:
                share = np.bincount(own, minlength=n) / len(rows)
                assert share.min() > 0.6 / n and share.max() < 1.4 / n, (mode, n, share)
            again = rows.copy()
            again["bytes"] += 1                      # values do not matter, only the key words of the mode
            again["key"][:, kw:] = 0xABCD
            assert np.array_equal(fp.row_owner(mode, again, n), own)
        grouped, counts = par.partition_rows(rows, mode, 4)
        assert counts.sum() == len(rows) and np.array_equal(np.sort(fp.row_owner(mode, grouped, 4)), fp.row_owner(mode, grouped, 4))
    L = fp.load_library()
    assert L.fa_row_owner(99, None, 0, 2, None) != 0 and L.fa_row_owner(1, None, 0, 0, None) != 0
    assert L.fa_row_owner(1, None, 0, 2, None) == 0


def test_bench_reference_arm_contract_on_cpu():
    """`bench.py --impl reference` needs no GPU: one JSON line on stdout, the arm's keys, same metric/config naming as
    the GPU arm; under torchrun every rank but 0 exits 0 without work."""
    import json
    import subprocess
    import sys

    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--flows", "300000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "flows/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "flows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("configs[1]") and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    # same config tag as the GPU arm (the driver compares the two lines)
    sys.path.insert(0, ROOT)
    import bench

    assert d["config"] == bench.bench_config() and d["cpu_baseline"]["isa"] in ("native", "portable")
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_never_maps_the_product_library():
    """The CPU arm's input comes from oracle/libmocker_ref.so: the product library must not be in the process's maps."""
    import subprocess
    import sys

    code = ("import sys; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','1','--flows','200000'];"
            "import runpy\n"
            "try:\n    runpy.run_path(%r, run_name='__main__')\nexcept SystemExit: pass\n"
            "maps=open('/proc/self/maps').read(); sys.stderr.write('MAPS_HAS_FLOWAGG=%%d\\n' %% ('libflowagg' in maps)); "
            "sys.stderr.write('MAPS_HAS_ORACLE=%%d\\n' %% ('liboracle' in maps))") % os.path.join(ROOT, "bench.py")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "MAPS_HAS_FLOWAGG=0" in r.stderr and "MAPS_HAS_ORACLE=1" in r.stderr, r.stderr[-1500:]


def test_oracle_side_mocker_is_bytewise_the_product_mocker(fp, oracle):
    """oracle/libmocker_ref.so (the CPU arms' producer) and fa_mocker_host emit the same bytes for the same (config, index)."""
    for kw in (dict(seed=1, flows_per_second=250_000, n_src_as=256, n_dst_as=256, framed=True),
               dict(seed=9, flows_per_second=0, n_src_as=3, n_dst_as=3, framed=False),
               dict(seed=4, flows_per_second=7, n_src_as=16, n_dst_as=5, framed=True, addr_mode=1),
               dict(seed=5, flows_per_second=7, n_src_as=16, n_dst_as=5, framed=True, addr_mode=2)):
        for first in (0, 2 ** 28 - 50, 2 ** 32 - 100):
            cfg = fp.FaMockerConfig.make(**kw)
            a, ao = fp.mocker_host(cfg, first, 3000)
            b, bo = oracle.mocker_host(first=first, n=3000, **kw)
            assert np.array_equal(a, b) and np.array_equal(ao, bo)


def test_bench_gpu_arm_fails_loudly_without_a_gpu():
    import subprocess
    import sys

    import torch

    if torch.cuda.is_available():
        pytest.skip("CPU-container check")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--flows", "100000"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stdout + r.stderr)


def test_bench_cuts_the_stream_at_record_boundaries():
    """bench.py's e2e leg without shipped offsets hands fa_submit(offsets = NULL) pieces of at most 64 MiB; every piece must start
    and end on a record boundary and the pieces must tile the stream (include/flowagg.h)."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    rng = np.random.default_rng(5)
    sizes = rng.integers(0, 300, 20000)
    sizes[[17, 4000]] = [5000, 1200]                       # two records larger than the limit
    offs = np.concatenate([[7], 7 + np.cumsum(sizes)])     # a stream that does not start at byte 0
    for limit in (1000, 4096, 1 << 20):
        pieces = bench.cut_pieces(offs, limit)
        assert pieces[0][0] == offs[0] and pieces[-1][1] == offs[-1]
        assert all(a1 == b0 for (_, a1), (b0, _) in zip(pieces, pieces[1:]))            # contiguous, no overlap
        bounds = set(offs.tolist())
        assert all(b0 in bounds and b1 in bounds and b1 > b0 for b0, b1 in pieces)
        big = {int(offs[i]) for i in range(len(sizes)) if sizes[i] > limit}
        assert all(b1 - b0 <= limit or b0 in big for b0, b1 in pieces)                   # only a lone oversized record exceeds it
    assert bench.cut_pieces(np.array([0]), 10) == []
