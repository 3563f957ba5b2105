"""CPU: property tests of the decode semantics (hypothesis).  proto.Unmarshal as called at inserter/inserter.go:124 does
not care about field order, keeps the last of repeated scalar tags, treats absent fields as zero and skips unknown
ones; the roll-up (create.sh:92-110) is a commutative, associative sum.  Checked on the oracle AND on the device decoder
(decode.cuh compiled for the host -- tests/decode_host)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from conftest import concat_records, frame
from test_device_decoder_on_host import device_decode, dh  # noqa: F401  (dh is a fixture)

# field name -> (number, is_bytes, value strategy); the 16 fields the kernels keep (pb-ext/flow.pb.go:58-143)
U64 = st.one_of(st.integers(0, 2 ** 64 - 1), st.integers(0, 300), st.sampled_from([2 ** 28 - 1, 2 ** 28, 2 ** 32 - 1, 2 ** 32, 2 ** 35, 2 ** 63]))
U32 = st.one_of(st.integers(0, 2 ** 32 - 1), st.integers(0, 70000))
ADDR = st.one_of(st.binary(min_size=16, max_size=16), st.binary(min_size=4, max_size=4), st.binary(min_size=0, max_size=20))
FIELDS = {
    "type": (1, False, st.integers(0, 5)), "time_received": (2, False, U64), "sampling_rate": (3, False, U64), "sequence_num": (4, False, U32),
    "src_addr": (6, True, ADDR), "dst_addr": (7, True, ADDR), "bytes": (9, False, U64), "packets": (10, False, U64),
    "sampler_addr": (11, True, ADDR), "src_as": (14, False, U32), "dst_as": (15, False, U32), "proto": (20, False, U32),
    "src_port": (21, False, U32), "dst_port": (22, False, U32), "etype": (30, False, U32), "time_flow_start": (38, False, U64),
}
UNKNOWN = st.lists(st.tuples(st.sampled_from([5, 8, 12, 13, 16, 19, 23, 29, 31, 39, 60, 99, 102, 300, 2047, 2048, 70000]),
                             st.sampled_from([0, 1, 2, 5]), st.binary(min_size=0, max_size=12), st.integers(0, 2 ** 64 - 1)), max_size=4)


def varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def enc_field(num, is_bytes, v):
    return varint((num << 3) | 2) + varint(len(v)) + v if is_bytes else varint(num << 3) + varint(v)


def enc_unknown(num, wt, blob, v):
    if wt == 0:
        return varint(num << 3) + varint(v)
    if wt == 1:
        return varint((num << 3) | 1) + (v & (2 ** 64 - 1)).to_bytes(8, "little")
    if wt == 5:
        return varint((num << 3) | 5) + (v & (2 ** 32 - 1)).to_bytes(4, "little")
    return varint((num << 3) | 2) + varint(len(blob)) + blob


flows = st.fixed_dictionaries({}, optional={k: s for k, (_, _, s) in FIELDS.items()})


def expected(fields):
    """What proto.Unmarshal leaves in the struct: absent = zero value."""
    d = {}
    for k, (_, is_bytes, _) in FIELDS.items():
        v = fields.get(k, b"" if is_bytes else 0)
        if is_bytes:
            d[k] = v[:16].ljust(16, b"\0")
            d[k + "_len"] = len(v)
        else:
            d[k] = v
    return d


def check_decodes_to(oracle, dh, msg, want):
    rc, got = oracle.decode(msg)
    assert rc == 0
    for k, v in want.items():
        assert got[k] == v, (k, got[k], v, msg.hex())
    # the device decoder, all 16 fields kept
    for framed in (False, True):
        blob, offs = concat_records(frame([msg]) if framed else [msg])
        out, valid = device_decode(dh, blob, offs, framed, 0)
        assert valid[0]
        for k, v in want.items():
            name = {"src_addr": "src", "dst_addr": "dst", "sampler_addr": "sampler", "src_addr_len": "src_len", "dst_addr_len": "dst_len",
                    "sampler_addr_len": "sampler_len"}.get(k, k)
            g = out[name][0]
            assert (bytes(g) == v) if isinstance(v, bytes) else (int(g) == v), (k, msg.hex())


SET = settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


@SET
@given(fields=flows, order=st.randoms(use_true_random=False), unknown=UNKNOWN)
def test_field_order_absence_and_unknown_fields_do_not_matter(oracle, dh, fields, order, unknown):
    parts = [enc_field(FIELDS[k][0], FIELDS[k][1], v) for k, v in fields.items()] + [enc_unknown(*u) for u in unknown]
    want = expected(fields)
    check_decodes_to(oracle, dh, b"".join(parts), want)         # as given
    order.shuffle(parts)
    check_decodes_to(oracle, dh, b"".join(parts), want)         # any order
    canon = [enc_field(FIELDS[k][0], FIELDS[k][1], v) for k, v in sorted(fields.items(), key=lambda kv: FIELDS[kv[0]][0])
             if v not in (0, b"")]                             # proto3 writers leave zero values out
    check_decodes_to(oracle, dh, b"".join(canon), want)


@SET
@given(fields=flows, earlier=flows, order=st.randoms(use_true_random=False))
def test_last_value_wins(oracle, dh, fields, earlier, order):
    """A repeated scalar tag keeps its last value, bytes are replaced not appended."""
    first = [enc_field(FIELDS[k][0], FIELDS[k][1], v) for k, v in earlier.items()]
    order.shuffle(first)
    last = [enc_field(FIELDS[k][0], FIELDS[k][1], v) for k, v in fields.items()]
    merged = dict(earlier)
    merged.update(fields)
    check_decodes_to(oracle, dh, b"".join(first + last), expected(merged))


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(batch=st.lists(flows, min_size=1, max_size=40), order=st.randoms(use_true_random=False), cut=st.integers(0, 40))
def test_rollup_is_order_free_and_merges_like_a_summing_merge_tree(oracle, batch, order, cut):
    """GROUP BY is a commutative sum: any order of the flows, and any split into two parts merged afterwards
    (SummingMergeTree, create.sh:88-90), gives the same rows."""
    msgs = [b"".join(enc_field(FIELDS[k][0], FIELDS[k][1], v) for k, v in f.items()) for f in batch]
    blob, offs = concat_records(msgs)
    rows, _, _ = oracle.run_batch(blob, offs, framed=False, key_mode="flows5m")
    shuffled = list(msgs)
    order.shuffle(shuffled)
    b2, o2 = concat_records(shuffled)
    rows2, _, _ = oracle.run_batch(b2, o2, framed=False, key_mode="flows5m")
    assert np.array_equal(rows, rows2)
    cut = min(cut, len(msgs))
    parts = []
    for part in (msgs[:cut], msgs[cut:]):
        pb, po = concat_records(part)
        parts.append(oracle.run_batch(pb, po, framed=False, key_mode="flows5m")[0])
    merged = {}
    for part in parts:
        for r in part:
            k = tuple(int(x) for x in r["key"])
            b, p, c = merged.get(k, (0, 0, 0))
            merged[k] = ((b + int(r["bytes"])) % 2 ** 64, (p + int(r["packets"])) % 2 ** 64, c + int(r["count"]))
    got = {tuple(int(x) for x in r["key"]): (int(r["bytes"]), int(r["packets"]), int(r["count"])) for r in rows}
    assert got == merged
