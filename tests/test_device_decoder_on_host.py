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


def check_against_oracle(dh, oracle, blob, offs, framed):
    """Every instantiated field mask: same accept/reject verdict as the oracle, same values for the fields it keeps."""
    want = oracle.decode_columns(blob, offs, framed=framed)
    ok = want["valid"].astype(bool)
    for sel in range(dh.dh_need_count()):
        got, valid = device_decode(dh, blob, offs, framed, sel)
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
