"""CPU: the device decoder itself (flow-pipeline_b200/csrc/decode.cuh, the template every kernel parses records
with), instantiated as host code by tests/decode_host, against the oracle and the golden vectors.  The GPU
parity tests prove the kernels; this proves the parsing logic on every CPU run too, bit for bit, for each
field mask the kernels instantiate."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLD2ORACLE, concat_records, frame

HERE = os.path.dirname(os.path.abspath(__file__))

# bit of fa::F_* (decode.cuh) -> oracle column(s)
NEED_BITS = {0: ["type"], 1: ["time_received"], 2: ["sampling_rate"], 3: ["sequence_num"], 4: ["src_addr"], 5: ["dst_addr"], 6: ["bytes"],
             7: ["packets"], 8: ["sampler_addr"], 9: ["src_as"], 10: ["dst_as"], 11: ["proto"], 12: ["src_port"], 13: ["dst_port"],
             14: ["etype"], 15: ["time_flow_start"]}
DH_NAME = {"src_addr": "src", "dst_addr": "dst", "sampler_addr": "sampler"}


class DhFlow(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("time_received", "sampling_rate", "time_flow_start", "bytes", "packets")] + \
               [(k, C.c_uint32) for k in ("type", "sequence_num", "src_as", "dst_as", "etype", "proto", "src_port", "dst_port",
                                          "src_len", "dst_len", "sampler_len", "pad")] + \
               [("src", C.c_uint8 * 16), ("dst", C.c_uint8 * 16), ("sampler", C.c_uint8 * 16)]


DH_DTYPE = np.dtype([(k, np.uint64) for k in ("time_received", "sampling_rate", "time_flow_start", "bytes", "packets")] +
                    [(k, np.uint32) for k in ("type", "sequence_num", "src_as", "dst_as", "etype", "proto", "src_port", "dst_port",
                                              "src_len", "dst_len", "sampler_len", "pad")] +
                    [("src", np.uint8, 16), ("dst", np.uint8, 16), ("sampler", np.uint8, 16)])


@pytest.fixture(scope="module")
def dh():
    assert C.sizeof(DhFlow) == DH_DTYPE.itemsize
    d = os.path.join(HERE, "decode_host")
    r = subprocess.run(["make", "-C", d], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    L = C.CDLL(os.path.join(d, "_build", "libdecode_host.so"))
    L.dh_decode.restype = C.c_int
    L.dh_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.dh_need_count.restype = C.c_int
    L.dh_need_mask.restype = C.c_uint32
    L.dh_need_mask.argtypes = [C.c_int]
    L.dh_learn_shape.restype = C.c_uint32
    L.dh_learn_shape.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p, C.c_uint32]
    L.dh_decode_shaped.restype = C.c_int
    L.dh_decode_shaped.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p,
                                   C.c_void_p, C.POINTER(C.c_uint64)]
    return L


def device_decode(dh, blob, offs, framed, sel):
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    offs = np.ascontiguousarray(offs, dtype=np.uint32)
    n = len(offs) - 1
    out = np.zeros(max(n, 1), dtype=DH_DTYPE)
    valid = np.zeros(max(n, 1), dtype=np.uint8)
    rc = dh.dh_decode(blob.ctypes.data if blob.size else None, blob.size, offs.ctypes.data, n, int(framed), sel, out.ctypes.data, valid.ctypes.data)
    assert rc == 0
    return out[:n], valid[:n].astype(bool)


def learn_shape(dh, blob, offs, framed, stride=1, n_sample=256):
    """The host twin of k_learn_shape: ascending tag values seen in a sample of the batch."""
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    offs = np.ascontiguousarray(offs, dtype=np.uint32)
    tags = np.zeros(1 << 14, dtype=np.uint16)
    m = dh.dh_learn_shape(blob.ctypes.data if blob.size else None, blob.size, offs.ctypes.data, len(offs) - 1, int(framed), stride, n_sample,
                          tags.ctypes.data, tags.size)
    return tags[:m].copy()


def device_decode_shaped(dh, blob, offs, framed, sel, tags):
    """Shape fast path over `tags` first, the order-agnostic decoder for what it does not decide (the kernels' arrangement)."""
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    offs = np.ascontiguousarray(offs, dtype=np.uint32)
    tags = np.ascontiguousarray(tags, dtype=np.uint16)
    n = len(offs) - 1
    out = np.zeros(max(n, 1), dtype=DH_DTYPE)
    valid = np.zeros(max(n, 1), dtype=np.uint8)
    n_fast = C.c_uint64(0)
    rc = dh.dh_decode_shaped(blob.ctypes.data if blob.size else None, blob.size, offs.ctypes.data, n, int(framed), sel,
                             tags.ctypes.data if tags.size else None, tags.size, out.ctypes.data, valid.ctypes.data, C.byref(n_fast))
    assert rc == 0
    return out[:n], valid[:n].astype(bool), n_fast.value


def check_against_oracle(dh, oracle, blob, offs, framed, tags=None):
    """Every instantiated field mask: same accept/reject verdict as the oracle, same values for the fields it keeps.
    tags: run the shape fast path over that table in front of the decoder."""
    want = oracle.decode_columns(blob, offs, framed=framed)
    ok = want["valid"].astype(bool)
    for sel in range(dh.dh_need_count()):
        if tags is None:
            got, valid = device_decode(dh, blob, offs, framed, sel)
        else:
            got, valid, _ = device_decode_shaped(dh, blob, offs, framed, sel, tags)
        assert np.array_equal(valid, ok), f"mask {sel}: verdicts differ at {np.flatnonzero(valid != ok)[:10]}"
        mask = dh.dh_need_mask(sel)
        for bit, cols in NEED_BITS.items():
            if not (mask >> bit) & 1:
                continue
            for col in cols:
                if col.endswith("_addr"):
                    assert np.array_equal(got[DH_NAME[col]][ok], want[col][ok]), (sel, col)
                    assert np.array_equal(got[DH_NAME[col] + "_len"][ok], want[col + "_len"][ok]), (sel, col)
                else:
                    assert np.array_equal(got[col][ok].astype(np.uint64), want[col][ok].astype(np.uint64)), (sel, col)
    return ok


def test_edge_cases_through_the_device_decoder(dh, oracle, edge_cases):
    msgs = [bytes.fromhex(c["hex"]) for c in edge_cases["cases"]]
    blob, offs = concat_records(msgs)
    ok = check_against_oracle(dh, oracle, blob, offs, framed=False)
    assert [bool(x) for x in ok] == [c["go_ok"] for c in edge_cases["cases"]]      # protobuf-go's verdicts
    blob, offs = concat_records(frame(msgs))
    check_against_oracle(dh, oracle, blob, offs, framed=True)


def test_golden_sets_through_the_device_decoder(dh, oracle, fuzz_2k, mocker_10k):
    for g, framed in ((fuzz_2k, False), (mocker_10k, False)):
        ok = check_against_oracle(dh, oracle, g["blob"], g["offsets"], framed)
        got, valid = device_decode(dh, g["blob"], g["offsets"], framed, 0)
        # and straight against the upb-made golden columns
        gv = g["valid"].astype(bool) if "valid" in g else np.ones(len(ok), dtype=bool)
        both = ok & gv
        for gk, okey in GOLD2ORACLE.items():
            if gk in ("SrcAddr", "DstAddr", "SamplerAddress"):
                assert np.array_equal(got[DH_NAME[okey]][both], g[gk][both]), gk
            else:
                assert np.array_equal(got[okey][both].astype(np.uint64), g[gk][both].astype(np.uint64)), gk


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5])
def test_schemaless_wire_fuzz_through_the_device_decoder(dh, oracle, seed):
    from test_gpu_parity import _random_wire_messages

    msgs = _random_wire_messages(seed, 4000)
    blob, offs = concat_records(msgs)
    ok = check_against_oracle(dh, oracle, blob, offs, framed=False)
    assert 0.1 < ok.mean() < 0.9
    blob, offs = concat_records(frame(msgs))
    check_against_oracle(dh, oracle, blob, offs, framed=True)


def test_framing_errors_and_ragged_spans(dh, oracle, mocker_10k):
    """Length prefixes that disagree with the span, empty spans, spans cut short."""
    g = mocker_10k
    msgs = [bytes(g["blob"][g["offsets"][i]: g["offsets"][i + 1]]) for i in range(200)]
    framed = frame(msgs)
    bad = [framed[0][:-1], framed[1] + b"\x00", b"", b"\x00", b"\x80", b"\xff" * 10 + b"\x01", framed[2][1:], b"\x05" + msgs[3]]
    blob, offs = concat_records(framed + bad)
    ok = check_against_oracle(dh, oracle, blob, offs, framed=True)
    assert ok[:200].all()
    assert bool(ok[203])            # b"\x00": a framed empty message is a valid all-zero flow
    assert not ok[200] and not ok[201] and not ok[202] and not ok[204] and not ok[205]


# ---- the shape fast path (decode.cuh: decode_record_shape) in front of the decoder --------------------------------

MOCKER_TAGS = [2 << 3, 3 << 3, 4 << 3, (6 << 3) | 2, (7 << 3) | 2, 9 << 3, 10 << 3, 14 << 3, 15 << 3, 21 << 3, 22 << 3, 30 << 3, 38 << 3]


def test_shape_learned_from_mocker_records_is_the_mocker_field_list(dh, mocker_10k):
    g = mocker_10k
    tags = learn_shape(dh, g["blob"], g["offsets"], framed=False, stride=37)
    assert (tags & 0x3FFF).tolist() == MOCKER_TAGS           # mocker/mocker.go:75-90: the 13 fields it sets, ascending
    # the varint lengths seen ride in bits 14 (1..4 bytes) and 15 (5 bytes): the two timestamps are always 5 bytes long
    saw4, saw5 = (tags >> 14) & 1, (tags >> 15) & 1
    for t, a, b in zip(tags & 0x3FFF, saw4, saw5):
        if t in (2 << 3, 38 << 3):
            assert (a, b) == (0, 1)
        elif t & 7 == 0:
            assert (a, b) == (1, 0)
        else:
            assert (a, b) == (0, 0)
    framed_blob, framed_offs = concat_records(frame([bytes(g["blob"][g["offsets"][i]: g["offsets"][i + 1]]) for i in range(300)]))
    assert (learn_shape(dh, framed_blob, framed_offs, framed=True) & 0x3FFF).tolist() == MOCKER_TAGS


def test_shape_fast_path_takes_every_mocker_record_and_equals_the_oracle(dh, oracle, mocker_10k):
    g = mocker_10k
    tags = learn_shape(dh, g["blob"], g["offsets"], framed=False)
    check_against_oracle(dh, oracle, g["blob"], g["offsets"], False, tags=tags)
    for sel in range(dh.dh_need_count()):
        _, valid, n_fast = device_decode_shaped(dh, g["blob"], g["offsets"], False, sel, tags)
        assert valid.all() and n_fast == len(valid)          # zero-valued (absent) fields included: they skip their step
    msgs = [bytes(g["blob"][g["offsets"][i]: g["offsets"][i + 1]]) for i in range(2000)]
    blob, offs = concat_records(frame(msgs))
    check_against_oracle(dh, oracle, blob, offs, True, tags=tags)
    assert device_decode_shaped(dh, blob, offs, True, 1, tags)[2] == 2000


@pytest.mark.parametrize("table", ["learned", "mocker", "empty", "all_small", "reversed_subset", "huge"])
def test_results_never_depend_on_the_shape_table(dh, oracle, edge_cases, fuzz_2k, table):
    """Whatever the table says, fast path + fallback == the oracle: edge cases (both framings), the 67-field fuzz set."""
    msgs = [bytes.fromhex(c["hex"]) for c in edge_cases["cases"]]
    sets = [concat_records(msgs) + (False,), concat_records(frame(msgs)) + (True,), (fuzz_2k["blob"], fuzz_2k["offsets"], False)]
    for blob, offs, framed in sets:
        if table == "learned":
            tags = learn_shape(dh, blob, offs, framed, n_sample=4096)
        elif table == "mocker":
            tags = np.array(MOCKER_TAGS, dtype=np.uint16)
        elif table == "empty":
            tags = np.zeros(0, dtype=np.uint16)
        elif table == "all_small":                           # every field number 1..15 with every wire type: 120 > kShapeMax -> no fast path
            tags = np.array([(n << 3) | w for n in range(1, 16) for w in range(8)], dtype=np.uint16)
        elif table == "reversed_subset":                     # not ascending, with illegal entries: still only ever accepts what it verified
            tags = np.array([(38 << 3), (15 << 3), (14 << 3) | 2, 0, 7, (100 << 3) | 2, (9 << 3), (2 << 3) | 5, (6 << 3) | 2], dtype=np.uint16)
        else:
            tags = np.arange(8, 8 + 40, dtype=np.uint16)
        check_against_oracle(dh, oracle, blob, offs, framed, tags=tags)


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_schemaless_wire_fuzz_with_a_learned_shape(dh, oracle, seed):
    from test_gpu_parity import _random_wire_messages

    msgs = _random_wire_messages(seed, 3000)
    for framed in (False, True):
        blob, offs = concat_records(frame(msgs) if framed else msgs)
        tags = learn_shape(dh, blob, offs, framed, stride=11)
        check_against_oracle(dh, oracle, blob, offs, framed, tags=tags)


def test_shape_boundaries_five_byte_varints_long_varints_and_lengths(dh, oracle):
    """Values at the edges of what the fast path takes (1..5-byte varints, 1-byte lengths) and just beyond (6+ bytes,
    2-byte lengths, 17-byte addresses): same rows as the oracle either way."""
    def vi(x):
        out = bytearray()
        while True:
            b = x & 0x7f
            x >>= 7
            out.append(b | (0x80 if x else 0))
            if not x:
                return bytes(out)

    msgs = []
    for v in [0, 1, 127, 128, 16383, 16384, 2 ** 21 - 1, 2 ** 21, 2 ** 28 - 1, 2 ** 28, 2 ** 32 - 1, 2 ** 32, 2 ** 35 - 1, 2 ** 35, 2 ** 63, 2 ** 64 - 1]:
        for num in (2, 9, 14, 21, 38):                        # u64 and u32 fields, 1- and 2-byte tags
            msgs.append(vi(num << 3) + vi(v) + b"\x78\x01")
    for ln in (0, 1, 4, 15, 16, 17, 127, 128, 200):
        msgs.append(b"\x32" + vi(ln) + bytes(range(ln)) + b"\x48\x05")
    msgs.append(b"\x48\x85\x80\x80\x80\x00")                # over-long 5-byte encoding of 5
    msgs.append(b"\x48\x80\x80\x80\x80\x80\x00")            # 6 bytes
    msgs.append(b"\xc8\x00\x05")                              # field 9 with an over-long 2-byte tag
    for framed in (False, True):
        blob, offs = concat_records(frame(msgs) if framed else msgs)
        tags = learn_shape(dh, blob, offs, framed, n_sample=10000)
        check_against_oracle(dh, oracle, blob, offs, framed, tags=tags)
        check_against_oracle(dh, oracle, blob, offs, framed, tags=np.array(sorted(set(MOCKER_TAGS + [(14 << 3), (21 << 3)])), dtype=np.uint16))
