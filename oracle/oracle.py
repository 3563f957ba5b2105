"""ctypes wrapper of the CPU ORACLE (oracle/flow_oracle.c).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package never
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_LIB_KIND = None  # "portable" (-march=x86-64-v2, built in the CPU container) or "native" (-march=native, built on this host)
_MOCKER = None

FO_MAX_KEY_WORDS = 12
KEY_MODES = {"flows5m": 0, "aspair": 1, "srcaddr": 2, "dstaddr": 3, "5tuple": 4, "srcport": 5, "dstport": 6}
KEY_WORDS = [4, 2, 4, 4, 11, 1, 1]


class FoFlow(C.Structure):
    _fields_ = [
        ("time_received", C.c_uint64), ("sampling_rate", C.c_uint64), ("time_flow_start", C.c_uint64),
        ("bytes", C.c_uint64), ("packets", C.c_uint64),
        ("type", C.c_uint32), ("sequence_num", C.c_uint32), ("src_as", C.c_uint32), ("dst_as", C.c_uint32),
        ("etype", C.c_uint32), ("proto", C.c_uint32), ("src_port", C.c_uint32), ("dst_port", C.c_uint32),
        ("src_addr_len", C.c_uint32), ("dst_addr_len", C.c_uint32), ("sampler_addr_len", C.c_uint32),
        ("src_addr_off", C.c_uint32), ("dst_addr_off", C.c_uint32),
        ("src_addr", C.c_uint8 * 16), ("dst_addr", C.c_uint8 * 16), ("sampler_addr", C.c_uint8 * 16),
    ]


ROW_DTYPE = np.dtype([("key", "<u4", (FO_MAX_KEY_WORDS,)), ("bytes", "<u8"), ("packets", "<u8"), ("count", "<u8")])
HH_DTYPE = np.dtype([("key", "<u4", (FO_MAX_KEY_WORDS,)), ("estimate", "<u8")])


class FoBatchResult(C.Structure):
    _fields_ = [("n_records", C.c_uint64), ("n_bad", C.c_uint64), ("n_nokey", C.c_uint64), ("seconds", C.c_double)]


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def use_native():
    """Build oracle/_native/liboracle.so with -march=native ON THIS HOST and use it from now on (bench.py's CPU arms: the
    baseline gets the box's own ISA).  Falls back to the portable build when there is no compiler.
    Must be called before the first oracle call of the process.  Returns the kind in use."""
    global _LIB_KIND
    if _LIB is not None:
        return _LIB_KIND
    r = subprocess.run(["make", "-s", "-C", _HERE, "native"], capture_output=True, text=True)
    _LIB_KIND = "native" if r.returncode == 0 and os.path.exists(os.path.join(_HERE, "_native", "liboracle.so")) else "portable"
    return _LIB_KIND


def lib_kind():
    return _LIB_KIND or "portable"


class MockerConfig(C.Structure):
    """include/flowagg.h: fa_mocker_config (the producer's parameters; mocker/mocker.go:57-102)."""
    _fields_ = [("seed", C.c_uint64), ("t0", C.c_uint64), ("flows_per_second", C.c_uint64), ("n_src_as", C.c_uint32),
                ("n_dst_as", C.c_uint32), ("addr_mode", C.c_uint32), ("framed", C.c_uint32)]


def mocker_host(seed=1, t0=1584912398, flows_per_second=0, n_src_as=3, n_dst_as=3, addr_mode=0, framed=True, first=0, n=0):
    """The synthetic producer's records [first, first+n) from oracle/libmocker_ref.so (host only; never maps libflowagg.so)."""
    global _MOCKER
    if _MOCKER is None:
        path = os.path.join(_HERE, "libmocker_ref.so")
        if not os.path.exists(path):
            build()
        _MOCKER = C.CDLL(path)
        _MOCKER.fo_mocker_host.restype = C.c_int
        _MOCKER.fo_mocker_host.argtypes = [C.POINTER(MockerConfig), C.c_uint64, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p,
                                           C.POINTER(C.c_size_t)]
    cfg = MockerConfig(seed, t0, flows_per_second, n_src_as, n_dst_as, addr_mode, 1 if framed else 0)
    nbytes = C.c_size_t(0)
    _MOCKER.fo_mocker_host(C.byref(cfg), first, n, None, 0, None, C.byref(nbytes))
    buf = np.empty(nbytes.value, dtype=np.uint8)
    offs = np.empty(n + 1, dtype=np.uint32)
    rc = _MOCKER.fo_mocker_host(C.byref(cfg), first, n, buf.ctypes.data, buf.size, offs.ctypes.data, C.byref(nbytes))
    if rc != 0:
        raise RuntimeError(f"fo_mocker_host failed: {rc}")
    return buf, offs


def lib():
    global _LIB, _LIB_KIND
    if _LIB is None:
        path = os.path.join(_HERE, "_native", "liboracle.so") if _LIB_KIND == "native" else os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB_KIND = _LIB_KIND or "portable"
        L = C.CDLL(path)
        L.fo_decode.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(FoFlow)]
        L.fo_decode.restype = C.c_int
        L.fo_decode_record.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.POINTER(FoFlow)]
        L.fo_decode_record.restype = C.c_int
        L.fo_frame_walk.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.fo_frame_walk.restype = C.c_long
        L.fo_ip_string.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p]
        L.fo_ip_string.restype = None
        L.fo_key_words.argtypes = [C.c_int]
        L.fo_make_key.argtypes = [C.c_int, C.POINTER(FoFlow), C.c_void_p]
        L.fo_hash64.argtypes = [C.c_void_p, C.c_int]
        L.fo_hash64.restype = C.c_uint64
        L.fo_agg_new.argtypes = [C.c_int, C.c_int]
        L.fo_agg_new.restype = C.c_void_p
        L.fo_agg_free.argtypes = [C.c_void_p]
        L.fo_agg_add.argtypes = [C.c_void_p, C.POINTER(FoFlow)]
        L.fo_agg_add_row.argtypes = [C.c_void_p, C.c_void_p]
        L.fo_agg_size.argtypes = [C.c_void_p]
        L.fo_agg_size.restype = C.c_size_t
        L.fo_agg_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.fo_agg_rows.restype = C.c_size_t
        L.fo_cms_add.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint64]
        L.fo_cms_estimate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.fo_cms_estimate.restype = C.c_uint64
        L.fo_topk.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        L.fo_topk.restype = C.c_size_t
        L.fo_run_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_int, C.c_int, C.POINTER(FoBatchResult)]
        L.fo_run_batch.restype = C.c_int
        L.fo_run_slabs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                   C.c_int, C.POINTER(FoBatchResult)]
        L.fo_run_slabs.restype = C.c_int
        _LIB = L
    return _LIB


FLOW_FIELDS = ["time_received", "sampling_rate", "time_flow_start", "bytes", "packets", "type", "sequence_num",
               "src_as", "dst_as", "etype", "proto", "src_port", "dst_port"]


def decode(msg: bytes):
    """proto.Unmarshal restatement -> (rc, dict).  rc 0 = ok, <0 = error class."""
    f = FoFlow()
    buf = (C.c_uint8 * max(len(msg), 1)).from_buffer_copy(msg or b"\0")
    rc = lib().fo_decode(buf, len(msg), C.byref(f))
    return rc, flow_dict(f)


def flow_dict(f):
    d = {k: int(getattr(f, k)) for k in FLOW_FIELDS}
    d["src_addr"] = bytes(f.src_addr)
    d["dst_addr"] = bytes(f.dst_addr)
    d["sampler_addr"] = bytes(f.sampler_addr)
    d["src_addr_len"] = int(f.src_addr_len)
    d["dst_addr_len"] = int(f.dst_addr_len)
    d["sampler_addr_len"] = int(f.sampler_addr_len)
    return d


def decode_columns(buf: np.ndarray, offsets: np.ndarray, framed: bool):
    """Decode every record of a batch; returns dict of numpy columns (+ 'valid', 'rc')."""
    L = lib()
    n = len(offsets) - 1
    cols = {k: np.zeros(n, dtype=np.uint64 if k in ("time_received", "sampling_rate", "time_flow_start", "bytes", "packets")
                        else np.uint32) for k in FLOW_FIELDS}
    for k in ("src_addr", "dst_addr", "sampler_addr"):
        cols[k] = np.zeros((n, 16), dtype=np.uint8)
        cols[k + "_len"] = np.zeros(n, dtype=np.uint32)
    cols["valid"] = np.zeros(n, dtype=np.uint8)
    cols["rc"] = np.zeros(n, dtype=np.int32)
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    ptr = buf.ctypes.data if buf.size else None
    f = FoFlow()
    for i in range(n):
        rc = L.fo_decode_record(ptr, int(offsets[i]), int(offsets[i + 1]), int(framed), C.byref(f))
        cols["rc"][i] = rc
        if rc != 0:
            continue
        cols["valid"][i] = 1
        for k in FLOW_FIELDS:
            cols[k][i] = getattr(f, k)
        for k in ("src_addr", "dst_addr", "sampler_addr"):
            cols[k][i] = np.frombuffer(bytes(getattr(f, k)), dtype=np.uint8)
            cols[k + "_len"][i] = getattr(f, k + "_len")
    return cols


def frame_walk(buf: np.ndarray):
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    cap = buf.size + 2
    offs = np.zeros(cap, dtype=np.uint32)
    n = lib().fo_frame_walk(buf.ctypes.data if buf.size else None, buf.size, offs.ctypes.data, cap)
    return n, offs


def ip_string(addr: bytes) -> str:
    out = C.create_string_buffer(64 + 2 * len(addr))
    b = (C.c_uint8 * max(len(addr), 1)).from_buffer_copy(addr or b"\0")
    lib().fo_ip_string(b, len(addr), out)
    return out.value.decode()


def hash64(key_words) -> int:
    k = np.ascontiguousarray(key_words, dtype=np.uint32)
    return int(lib().fo_hash64(k.ctypes.data, len(k)))


def run_batch(buf, offsets, framed=True, key_mode="flows5m", scale=False, cms=None, threads=1, aggregate=True):
    """Decode + roll-up a whole batch.  Returns (rows ndarray ROW_DTYPE sorted, cms ndarray|None, result dict)."""
    L = lib()
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
    n = len(offsets) - 1
    mode = KEY_MODES[key_mode] if isinstance(key_mode, str) else int(key_mode)
    agg = L.fo_agg_new(mode, int(scale)) if aggregate else None
    cms_arr = None
    depth = wlog2 = 0
    if cms:
        depth, wlog2 = cms
        cms_arr = np.zeros(depth << wlog2, dtype=np.uint64)
    res = FoBatchResult()
    # without a table the C driver defaults the sketch key to SrcAddr; pass an agg to pick the mode
    tmp_agg = agg
    if agg is None and cms:
        tmp_agg = L.fo_agg_new(mode, int(scale))
    L.fo_run_batch(buf.ctypes.data if buf.size else None, offsets.ctypes.data, n, int(framed), tmp_agg,
                   cms_arr.ctypes.data if cms_arr is not None else None, depth, wlog2, int(threads), C.byref(res))
    rows = np.zeros(0, dtype=ROW_DTYPE)
    if tmp_agg:
        sz = L.fo_agg_size(tmp_agg)
        rows = np.zeros(sz, dtype=ROW_DTYPE)
        if sz:
            L.fo_agg_rows(tmp_agg, rows.ctypes.data, sz)
        L.fo_agg_free(tmp_agg)
    return rows, cms_arr, {"n_records": res.n_records, "n_bad": res.n_bad, "n_nokey": res.n_nokey, "seconds": res.seconds}


def run_slabs(slabs, framed=True, key_mode="aspair", threads=1):
    """Decode + roll-up several (buf, offsets) slabs as one stream (the CPU baseline leg of bench.py)."""
    L = lib()
    mode = KEY_MODES[key_mode] if isinstance(key_mode, str) else int(key_mode)
    bufs = [np.ascontiguousarray(b, dtype=np.uint8) for b, _ in slabs]
    offs = [np.ascontiguousarray(o, dtype=np.uint32) for _, o in slabs]
    k = len(slabs)
    pb = (C.c_void_p * k)(*[b.ctypes.data for b in bufs])
    po = (C.c_void_p * k)(*[o.ctypes.data for o in offs])
    pn = (C.c_size_t * k)(*[len(o) - 1 for o in offs])
    agg = L.fo_agg_new(mode, 0)
    res = FoBatchResult()
    L.fo_run_slabs(pb, po, pn, k, int(framed), agg, None, 0, 0, int(threads), C.byref(res))
    sz = L.fo_agg_size(agg)
    rows = np.zeros(sz, dtype=ROW_DTYPE)
    if sz:
        L.fo_agg_rows(agg, rows.ctypes.data, sz)
    L.fo_agg_free(agg)
    return rows, {"n_records": res.n_records, "n_bad": res.n_bad, "n_nokey": res.n_nokey, "seconds": res.seconds}


def topk(cms_arr, depth, wlog2, n_words, cand_rows, k):
    out = np.zeros(k, dtype=HH_DTYPE)
    cand_rows = np.ascontiguousarray(cand_rows)
    n = lib().fo_topk(cms_arr.ctypes.data, depth, wlog2, n_words, cand_rows.ctypes.data if len(cand_rows) else None,
                      len(cand_rows), k, out.ctypes.data)
    return out[:n]
