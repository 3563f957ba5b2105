"""CPU: the oracle against the golden vectors (tests/golden/, made from the reference's
embedded descriptor through upb) -- this is what pins the oracle."""
import numpy as np
import pytest

from conftest import GOLD2ORACLE, concat_records, frame

U64 = ("TimeReceived", "SamplingRate", "TimeFlowStart", "Bytes", "Packets")


def test_edge_cases_match_protobuf_go_verdict(oracle, edge_cases):
    n_diff = 0
    for c in edge_cases["cases"]:
        rc, d = oracle.decode(bytes.fromhex(c["hex"]))
        assert (rc == 0) == c["go_ok"], (c["name"], rc, c["note"])
        if c["go_ok"] != c["upb_ok"]:
            n_diff += 1
            continue
        if c["upb_ok"]:
            for g, o in GOLD2ORACLE.items():
                want = c["fields"][g]
                if g in ("SrcAddr", "DstAddr", "SamplerAddress"):
                    raw = bytes.fromhex(want)
                    assert d[o] == raw[:16].ljust(16, b"\0"), (c["name"], g)
                    assert d[o + "_len"] == len(raw), (c["name"], g)
                else:
                    assert d[o] == want, (c["name"], g, d[o], want)
    assert n_diff == 3  # the three documented upb / protobuf-go divergences


def test_survey_golden_message_bytes(oracle, edge_cases):
    c = [c for c in edge_cases["cases"] if c["name"] == "survey_golden_message"][0]
    assert len(bytes.fromhex(c["hex"])) == 80
    rc, d = oracle.decode(bytes.fromhex(c["hex"]))
    assert rc == 0 and d["bytes"] == 1499 and d["packets"] == 99 and d["src_as"] == 65001 and d["dst_as"] == 65002
    assert d["etype"] == 0x86DD and d["time_received"] == 1584912398 and d["src_port"] == 443 and d["dst_port"] == 51234
    assert oracle.ip_string(d["src_addr"]) == "2001:db8:0:1::80"  # README.md:155


def _check_columns(cols, g, valid=None):
    n = len(g["offsets"]) - 1
    valid = np.ones(n, dtype=bool) if valid is None else valid.astype(bool)
    assert np.array_equal(cols["valid"].astype(bool), valid)
    for gk, ok in GOLD2ORACLE.items():
        if gk in ("SrcAddr", "DstAddr", "SamplerAddress"):
            assert np.array_equal(cols[ok][valid], g[gk][valid]), gk
            assert np.array_equal(cols[ok + "_len"][valid], g[gk + "Len"][valid]), gk
        else:
            assert np.array_equal(cols[ok][valid].astype(np.uint64), g[gk][valid].astype(np.uint64)), gk


def test_mocker_10k_columns(oracle, mocker_10k):
    g = mocker_10k
    cols = oracle.decode_columns(g["blob"], g["offsets"], framed=False)
    _check_columns(cols, g)


def test_mocker_10k_rollup_equals_pandas(oracle, mocker_10k):
    g = mocker_10k
    for threads in (1, 4):
        rows, _, res = oracle.run_batch(g["blob"], g["offsets"], framed=False, key_mode="flows5m", threads=threads)
        assert res["n_bad"] == 0
        assert np.array_equal(rows["key"][:, :4], g["rollup_key"])
        assert np.array_equal(np.stack([rows["bytes"], rows["packets"], rows["count"]], axis=1), g["rollup_val"])
        assert rows["count"].sum() == 10000


def test_fuzz_2k_columns(oracle, fuzz_2k):
    g = fuzz_2k
    cols = oracle.decode_columns(g["blob"], g["offsets"], framed=False)
    _check_columns(cols, g, g["valid"])


def test_framed_equals_unframed(oracle, fuzz_2k):
    g = fuzz_2k
    msgs = [bytes(g["blob"][g["offsets"][i]:g["offsets"][i + 1]]) for i in range(300)]
    blob, offs = concat_records(frame(msgs))
    a = oracle.decode_columns(blob, offs, framed=True)
    b = oracle.decode_columns(g["blob"], g["offsets"][:301], framed=False)
    for k in a:
        if k != "rc":
            assert np.array_equal(a[k], b[k]), k
    n, walked = oracle.frame_walk(blob)
    assert n == 300 and np.array_equal(walked[:301], offs)


def test_bad_framing_is_a_bad_record(oracle):
    msg = bytes.fromhex("4805")
    blob, offs = concat_records([b"\x03" + msg, b"\x02" + msg, b"\x01" + msg, b""])
    cols = oracle.decode_columns(blob, offs, framed=True)
    assert list(cols["valid"]) == [0, 1, 0, 0]


@pytest.mark.parametrize("addr,want", [
    (b"", "0.0.0.0"), (bytes([10, 1, 2, 3]), "10.1.2.3"), (bytes(16), "::"), (bytes(15) + b"\x01", "::1"),
    (bytes.fromhex("20010db8000000010000000000000080"), "2001:db8:0:1::80"),
    (bytes.fromhex("20010db8000000010000000000000000"), "2001:db8:0:1::"),
    (bytes.fromhex("00000000000000000000ffffc0a80101"), "192.168.1.1"),
    (bytes.fromhex("0101a8c0000000000000000000000000"), "101:a8c0::"),  # README.md:191-202
    (bytes.fromhex("20010000000000010000000000000001"), "2001:0:0:1::1"),
    (bytes.fromhex("00010000000000010000000100000001"), "1::1:0:1:0:1"),
    (bytes.fromhex("00010001000000010001000100010001"), "1:1:0:1:1:1:1:1"),
    (b"abc", "?616263"),
])
def test_ip_string_matches_go_net_ip(oracle, addr, want):
    assert oracle.ip_string(addr) == want


def test_key_modes_and_cms(oracle, mocker_10k):
    g = mocker_10k
    rows, cms, res = oracle.run_batch(g["blob"], g["offsets"], framed=False, key_mode="srcaddr", cms=(4, 12))
    assert len(rows) == 256 and rows["count"].sum() == 10000
    # CMS never under-estimates, and with 256 keys in 4096 columns it is usually exact
    top = oracle.topk(cms, 4, 12, 4, rows, 10)
    exact = {bytes(r["key"][:4].tobytes()): int(r["bytes"]) for r in rows}
    for h in top:
        assert int(h["estimate"]) >= exact[bytes(h["key"][:4].tobytes())]
    assert list(top["estimate"]) == sorted(top["estimate"], reverse=True)
    # linearity: sketch(a) + sketch(b) == sketch(a ++ b)
    half = 5000
    _, c1, _ = oracle.run_batch(g["blob"], g["offsets"][: half + 1], framed=False, key_mode="srcaddr", cms=(4, 12))
    sub = g["offsets"][half:] - g["offsets"][half]
    _, c2, _ = oracle.run_batch(g["blob"][g["offsets"][half]:], sub, framed=False, key_mode="srcaddr", cms=(4, 12))
    assert np.array_equal(c1 + c2, cms)
    for mode, groups in (("aspair", 9), ("dstaddr", 256), ("5tuple", 10000)):
        r, _, _ = oracle.run_batch(g["blob"], g["offsets"], framed=False, key_mode=mode)
        assert len(r) == groups, mode
        assert r["bytes"].sum() == g["Bytes"].sum() and r["packets"].sum() == g["Packets"].sum()


def _flow(t, src_as, dst_as, nbytes, pkts, etype=0x86DD):
    """A mocker-shaped FlowMessage by hand: TimeReceived(2), SamplingRate(3), Bytes(9), Packets(10), SrcAS(14), DstAS(15), Etype(30)."""
    def v(x):
        out = bytearray()
        while True:
            b = x & 0x7F
            x >>= 7
            out.append(b | 0x80 if x else b)
            if not x:
                return bytes(out)
    return (b"\x10" + v(t) + b"\x18\x01" + b"\x48" + v(nbytes) + b"\x50" + v(pkts) + b"\x70" + v(src_as) + b"\x78" + v(dst_as) +
            b"\xf0\x01" + v(etype))


def test_rollup_reproduces_the_readme_sample_rows(oracle):
    """The only published output of the Clickhouse half (README.md:143-183): flows received at 2020-03-22 21:26:38/39 land in
    Date 2020-03-22, Timeslot 21:25:00; `SELECT * FROM flows_5m WHERE SrcAS = 65001` shows one row per (SrcAS, DstAS) with
    ETypeMap.EType [34525] and the nested sums equal to the row's own.  Flows chosen to add up to the README's three rows."""
    import time

    t0 = 1584912398  # 2020-03-22 21:26:38 UTC, README.md:155
    assert time.strftime("%Y-%m-%d %H:%M:%S", time.gmtime(t0)) == "2020-03-22 21:26:38"
    parts = {(65001, 65000): [(1000, 50), (1000, 50), (900, 50), (30, 2)],          # Bytes 2930, Packets 152, Count 4
             (65001, 65001): [(1000, 90), (900, 90), (35, 10)],                      # 1935, 190, 3
             (65001, 65002): [(1000, 48)] * 4 + [(400, 48), (420, 48)],              # 4820, 288, 6
             (65000, 65002): [(7, 1)]}                                               # not selected by WHERE SrcAS = 65001
    msgs = []
    for (s, d), flows in parts.items():
        for i, (b, p) in enumerate(flows):
            msgs.append(_flow(t0 + (i % 2), s, d, b, p))
    blob, offs = concat_records(msgs)
    rows, _, res = oracle.run_batch(blob, offs, key_mode="flows5m", framed=False)
    assert res["n_bad"] == 0
    sel = rows[rows["key"][:, 1] == 65001]
    got = [(time.strftime("%Y-%m-%d", time.gmtime(int(r["key"][0]))), time.strftime("%Y-%m-%d %H:%M:%S", time.gmtime(int(r["key"][0]))),
            int(r["key"][1]), int(r["key"][2]), int(r["key"][3]), int(r["bytes"]), int(r["packets"]), int(r["count"])) for r in sel]
    assert got == [("2020-03-22", "2020-03-22 21:25:00", 65001, 65000, 34525, 2930, 152, 4),
                   ("2020-03-22", "2020-03-22 21:25:00", 65001, 65001, 34525, 1935, 190, 3),
                   ("2020-03-22", "2020-03-22 21:25:00", 65001, 65002, 34525, 4820, 288, 6)]   # README.md:180-183, in its ORDER BY order


def _rows_as_tuples(rows):
    return [(tuple(int(x) for x in r["key"][:4]), int(r["bytes"]), int(r["packets"]), int(r["count"])) for r in rows]


def test_rollup_boundary_vectors(oracle, rollup_boundaries):
    """The Clickhouse half at its edges (tests/golden/make_rollup_golden.py: expected rows computed there from the documented
    semantics of toStartOfFiveMinute / toDate / DateTime / UInt64 sum, independently of the oracle): five-minute slot edges,
    day rollover, the DateTime's 32-bit wrap, UInt64 wrap of the sums, EType never merging, extreme AS numbers, the all-ones key."""
    from conftest import rollup_case_rows

    for case in rollup_boundaries["cases"]:
        msgs = [bytes.fromhex(h) for h in case["messages_hex"]]
        for framed in (False, True):
            blob, offs = concat_records(frame(msgs) if framed else msgs)
            rows, _, res = oracle.run_batch(blob, offs, framed=framed, key_mode="flows5m")
            assert res["n_bad"] == 0 and res["n_records"] == len(msgs), case["name"]
            assert _rows_as_tuples(rows) == rollup_case_rows(case), case["name"]
            for r, w in zip(rows, case["rows"]):
                assert int(r["key"][0]) // 86400 == w["Date"], case["name"]          # toDate(TimeReceived): fa_row_date / create.sh:66
        # a second pass over the same flows doubles every sum modulo 2^64 (SummingMergeTree merge of two parts, create.sh:88)
        blob, offs = concat_records(msgs + msgs)
        rows, _, _ = oracle.run_batch(blob, offs, framed=False, key_mode="flows5m")
        want = [(k, (2 * b) & (2 ** 64 - 1), (2 * p) & (2 ** 64 - 1), 2 * c) for k, b, p, c in rollup_case_rows(case)]
        assert _rows_as_tuples(rows) == want, case["name"]
