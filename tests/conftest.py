import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o

    o.lib()  # builds liboracle.so if it is not there yet
    return o


@pytest.fixture(scope="session")
def fp():
    import flow_pipeline_b200 as pkg

    if not os.path.exists(pkg.lib_path()):
        pkg.build()
    pkg.load_library()
    return pkg


@pytest.fixture(scope="session")
def edge_cases():
    with open(os.path.join(GOLDEN, "edge_cases.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def rollup_boundaries():
    with open(os.path.join(GOLDEN, "rollup_boundaries.json")) as f:
        return json.load(f)


def rollup_case_rows(case):
    """Expected flows_5m rows of one tests/golden/rollup_boundaries.json case as (key words, bytes, packets, count) tuples in
    the table's ORDER BY order (create.sh:88-90); Date is derived from Timeslot and checked separately."""
    return [((r["Timeslot"], r["SrcAS"], r["DstAS"], r["EType"]), r["Bytes"], r["Packets"], r["Count"]) for r in case["rows"]]


@pytest.fixture(scope="session")
def mocker_10k():
    return dict(np.load(os.path.join(GOLDEN, "mocker_10k.npz")))


@pytest.fixture(scope="session")
def fuzz_2k():
    return dict(np.load(os.path.join(GOLDEN, "fuzz_2k.npz")))


GOLD2ORACLE = {
    "TimeReceived": "time_received", "SamplingRate": "sampling_rate", "TimeFlowStart": "time_flow_start", "Bytes": "bytes",
    "Packets": "packets", "Type": "type", "SequenceNum": "sequence_num", "SrcAS": "src_as", "DstAS": "dst_as", "Etype": "etype",
    "Proto": "proto", "SrcPort": "src_port", "DstPort": "dst_port", "SrcAddr": "src_addr", "DstAddr": "dst_addr",
    "SamplerAddress": "sampler_addr",
}


def concat_records(msgs):
    """list of bytes -> (uint8 blob, uint32 offsets[n+1])"""
    offs = np.zeros(len(msgs) + 1, dtype=np.uint32)
    offs[1:] = np.cumsum([len(m) for m in msgs])
    blob = np.frombuffer(b"".join(msgs), dtype=np.uint8).copy() if msgs else np.zeros(0, dtype=np.uint8)
    return blob, offs


def frame(msgs):
    """varint(len) || msg framing of mocker/mocker.go:98-101"""
    out = []
    for m in msgs:
        n = len(m)
        p = b""
        while True:
            b = n & 0x7F
            n >>= 7
            if n:
                p += bytes([b | 0x80])
            else:
                p += bytes([b])
                break
        out.append(p + m)
    return out
