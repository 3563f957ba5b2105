#!/usr/bin/env python
"""Generate tests/golden/rollup_boundaries.json: boundary vectors for the CLICKHOUSE half of the path (the flows_5m
roll-up of compose/clickhouse/create.sh:70-110), which the reference never tests and which cannot be executed here.

Every case is a handful of FlowMessages (serialised by upb from the reference's own embedded descriptor, as in
make_golden.py) and the rows the fully merged flows_5m table must hold for them, computed HERE in plain Python
integers from the documented semantics of the functions the DDL calls -- independently of oracle/ and of the kernels:

  Timeslot = toStartOfFiveMinute(TimeReceived)   create.sh:96    t - t mod 300 on the DateTime (UInt32 seconds, UTC)
  Date     = toDate(TimeReceived)                create.sh:66    floor(t / 86400) days since 1970-01-01 (UTC server)
  TimeReceived is a DateTime column              create.sh:39    a UInt32: the message's uint64 is stored modulo 2^32
  Bytes/Packets = sum(...), Count = count()      create.sh:105-107   UInt64 arithmetic, wrapping modulo 2^64
  one row per (Date, Timeslot, SrcAS, DstAS, ETypeMap.EType)   create.sh:88-90 (the SummingMergeTree ORDER BY),
  rows in that order; the kernels' key words are (Timeslot, SrcAS, DstAS, EType), Date being derived from Timeslot.

Run in the CPU container only (/root/reference is read by this script, never by the tests).
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_flow_message  # noqa: E402

M64 = (1 << 64) - 1
M32 = (1 << 32) - 1


def expected_rows(flows):
    groups = {}
    for f in flows:
        t = f.get("TimeReceived", 0) & M32          # DateTime column
        slot = t - t % 300                          # toStartOfFiveMinute
        key = (slot // 86400, slot, f.get("SrcAS", 0) & M32, f.get("DstAS", 0) & M32, f.get("Etype", 0) & M32)
        b, p, c = groups.get(key, (0, 0, 0))
        groups[key] = ((b + f.get("Bytes", 0)) & M64, (p + f.get("Packets", 0)) & M64, (c + 1) & M64)
    return [{"Date": k[0], "Timeslot": k[1], "SrcAS": k[2], "DstAS": k[3], "EType": k[4], "Bytes": v[0], "Packets": v[1], "Count": v[2]}
            for k, v in sorted(groups.items())]


def main():
    FM = load_flow_message()
    base = dict(SamplingRate=1, SrcAS=65001, DstAS=65002, Etype=0x86DD, Bytes=100, Packets=2)
    day = 18343 * 86400                             # 2020-03-22 00:00:00 UTC (README.md:155's day)
    cases = {
        "slot_edges": [dict(base, TimeReceived=day + s) for s in (0, 1, 299, 300, 301, 599, 600)],
        "day_rollover": [dict(base, TimeReceived=day + s) for s in (86399, 86400, 86400 + 299, 86400 - 300, 2 * 86400)],
        "epoch_and_first_slots": [dict(base, TimeReceived=s) for s in (0, 1, 299, 300, 86399, 86400)],
        "time_received_absent_is_epoch": [dict(base), dict(base, TimeReceived=0), dict(base, TimeReceived=5)],
        "u32_boundary_of_datetime": [dict(base, TimeReceived=s) for s in (M32, M32 - 295, M32 - 296, (1 << 32), (1 << 32) + 300, (1 << 40) + 7, M64)],
        "u64_wrap_of_sums": [dict(base, TimeReceived=day, Bytes=M64, Packets=M64), dict(base, TimeReceived=day + 1, Bytes=2, Packets=1),
                             dict(base, TimeReceived=day + 2, Bytes=1 << 63, Packets=1 << 63), dict(base, TimeReceived=day + 3, Bytes=1 << 63, Packets=1 << 63)],
        "etype_never_merges": [dict(base, TimeReceived=day, Etype=e) for e in (0x0800, 0x86DD, 0, 0x0800, M32)],
        "as_extremes": [dict(base, TimeReceived=day, SrcAS=a, DstAS=b) for a, b in ((0, 0), (M32, M32), (0, M32), (M32, 0), (M32, M32), (4200000000, 1))],
        "all_ones_key": [dict(base, TimeReceived=M32, SrcAS=M32, DstAS=M32, Etype=M32), dict(base, TimeReceived=M32 - 3, SrcAS=M32, DstAS=M32, Etype=M32)],
    }
    out = {"comment": __doc__.strip().split("\n\n")[1], "cases": []}
    for name, flows in cases.items():
        msgs = []
        for f in flows:
            m = FM()
            for k, v in f.items():
                setattr(m, k, v)
            msgs.append(m.SerializeToString().hex())
        out["cases"].append({"name": name, "flows": flows, "messages_hex": msgs, "rows": expected_rows(flows)})
    with open(os.path.join(HERE, "rollup_boundaries.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
