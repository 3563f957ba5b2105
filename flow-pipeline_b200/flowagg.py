"""ctypes binding of libflowagg.so (include/flowagg.h).

This is the same C ABI a cgo shim inside the reference's ConsumeClaim
(inserter/inserter.go:176) would bind; Python drives it for the tests and the
benchmark.  There is no CPU fallback here: every compute call goes to the
sm_100a kernels, and construction fails loudly when the library or a B200 is
missing.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FA_ABI_VERSION = 2
FA_MAX_KEY_WORDS = 12
KEY_MODES = {"flows5m": 0, "aspair": 1, "srcaddr": 2, "dstaddr": 3, "5tuple": 4, "srcport": 5, "dstport": 6}
KEY_WORDS = [4, 2, 4, 4, 11, 1, 1]

FA_CFG_CMS = 0x1
FA_CFG_SCALE_SAMPLING = 0x2
FA_CFG_COLUMNS = 0x4
FA_CFG_NO_AGGREGATE = 0x8
FA_CFG_CALLER_STREAM = 0x10
FA_CFG_TOPK_ONLY = 0x20
FA_FRAMED = 0x1
FA_FLUSH_KEEP = 0x1
FA_FLUSH_UNSORTED = 0x2
FA_CMS_LOCAL, FA_CMS_GLOBAL = 0, 1
FA_ADDR_MOCKER, FA_ADDR_ZIPF24, FA_ADDR_UNIQUE = 0, 1, 2

ROW_DTYPE = np.dtype([("key", "<u4", (FA_MAX_KEY_WORDS,)), ("bytes", "<u8"), ("packets", "<u8"), ("count", "<u8")])
HH_DTYPE = np.dtype([("key", "<u4", (FA_MAX_KEY_WORDS,)), ("estimate", "<u8")])

COLUMNS = {
    "valid": np.uint8, "time_received": np.uint64, "time_flow_start": np.uint64, "sampling_rate": np.uint64,
    "bytes": np.uint64, "packets": np.uint64, "type": np.uint32, "sequence_num": np.uint32, "src_as": np.uint32,
    "dst_as": np.uint32, "etype": np.uint32, "proto": np.uint32, "src_port": np.uint32, "dst_port": np.uint32,
    "src_addr": (np.uint8, 16), "dst_addr": (np.uint8, 16), "sampler_addr": (np.uint8, 16),
    "src_addr_len": np.uint8, "dst_addr_len": np.uint8, "sampler_addr_len": np.uint8,
}


class FlowAggError(RuntimeError):
    def __init__(self, status, what, detail=""):
        self.status = status
        super().__init__(f"{what}: status {status}" + (f" ({detail})" if detail else ""))


class FaConfig(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("device", C.c_int32), ("key_mode", C.c_uint32), ("flags", C.c_uint32),
                ("table_capacity", C.c_uint64), ("cms_depth", C.c_uint32), ("cms_width_log2", C.c_uint32),
                ("max_batch_bytes", C.c_uint64), ("max_batch_records", C.c_uint32), ("topk_k", C.c_uint32),
                ("stream", C.c_void_p)]


class FaStats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("n_records", "n_bad", "n_nokey", "n_dropped", "n_groups", "n_submits", "bytes_in", "n_kernels", "gpu_busy_us")]


class FaMockerConfig(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("t0", C.c_uint64), ("flows_per_second", C.c_uint64), ("n_src_as", C.c_uint32),
                ("n_dst_as", C.c_uint32), ("addr_mode", C.c_uint32), ("framed", C.c_uint32)]

    @classmethod
    def make(cls, seed=1, t0=1584912398, flows_per_second=0, n_src_as=3, n_dst_as=3, addr_mode=FA_ADDR_MOCKER, framed=True):
        return cls(seed, t0, flows_per_second, n_src_as, n_dst_as, addr_mode, 1 if framed else 0)


def lib_path():
    # FLOWAGG_LIB: another build of the same library (kernel experiments: profiles/r01/experiments)
    return os.environ.get("FLOWAGG_LIB") or os.path.join(_HERE, "libflowagg.so")


def build(verbose=False):
    """Compile libflowagg.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libflowagg.so failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)
    return lib_path()


_PROTOS = {
    "fa_create": (C.c_int, [C.POINTER(FaConfig), C.POINTER(C.c_void_p)]),
    "fa_destroy": (None, [C.c_void_p]),
    "fa_strerror": (C.c_char_p, [C.c_int]),
    "fa_last_error": (C.c_char_p, [C.c_void_p]),
    "fa_host_buffer": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p),
                                 C.POINTER(C.c_size_t)]),
    "fa_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_uint32]),
    "fa_submit_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_uint32]),
    "fa_sync": (C.c_int, [C.c_void_p]),
    "fa_stats_get": (C.c_int, [C.c_void_p, C.POINTER(FaStats)]),
    "fa_flush": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_uint32]),
    "fa_flush_begin": (C.c_int, [C.c_void_p, C.c_uint32]),
    "fa_flush_end": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "fa_merge_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32]),
    "fa_row_owner": (C.c_int, [C.c_int, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]),
    "fa_flush_box": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_uint32]),
    "fa_reset": (C.c_int, [C.c_void_p]),
    "fa_cms_read": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "fa_cms_device": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "fa_topk_local": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t)]),
    "fa_topk_merge": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t)]),
    "fa_topk": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t)]),
    "fa_columns": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fa_columns_read": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]),
    "fa_timer_start": (C.c_int, [C.c_void_p]),
    "fa_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "fa_mocker_host": (C.c_int, [C.POINTER(FaMockerConfig), C.c_uint64, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p,
                                 C.POINTER(C.c_size_t)]),
    "fa_mocker_device": (C.c_int, [C.c_void_p, C.POINTER(FaMockerConfig), C.c_uint64, C.c_uint32, C.c_void_p, C.c_size_t,
                                   C.c_void_p, C.POINTER(C.c_size_t)]),
    "fa_build_info": (C.c_char_p, []),
}


def load_library():
    """dlopen libflowagg.so and declare every symbol of include/flowagg.h.  Raises if the
    library is missing: there is no Python or CPU substitute for it."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise FlowAggError(-2, "libflowagg.so not built",
                               f"{path} is missing; run __graft_entry__.build() (nvcc, sm_100a). No CPU fallback exists.")
        L = C.CDLL(path)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)  # AttributeError if the header and the library diverge
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def exported_symbols():
    return list(_PROTOS.keys())


def mocker_host(cfg: FaMockerConfig, first: int, n: int):
    """mocker/mocker.go:57-102 on the host: returns (bytes ndarray, offsets ndarray[n+1])."""
    L = load_library()
    need = C.c_size_t(0)
    L.fa_mocker_host(C.byref(cfg), first, n, None, 0, None, C.byref(need))
    buf = np.zeros(max(need.value, 1) + 16, dtype=np.uint8)
    offs = np.zeros(n + 1, dtype=np.uint32)
    rc = L.fa_mocker_host(C.byref(cfg), first, n, buf.ctypes.data, need.value, offs.ctypes.data, C.byref(need))
    if rc:
        raise FlowAggError(rc, "fa_mocker_host")
    return buf[: need.value], offs


def row_owner(key_mode, rows, n_owners):
    """Owner in [0, n_owners) of every row's key: the hash partition of the box-wide exchange (fa_row_owner)."""
    L = load_library()
    mode = KEY_MODES[key_mode] if isinstance(key_mode, str) else int(key_mode)
    rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
    out = np.zeros(len(rows), dtype=np.uint32)
    rc = L.fa_row_owner(mode, rows.ctypes.data, len(rows), n_owners, out.ctypes.data)
    if rc:
        raise FlowAggError(rc, "fa_row_owner")
    return out


def _ptr(x):
    """Device or host address of a torch tensor / numpy array / int."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()  # torch tensor


class FlowAgg:
    """One fa_ctx: the state of one ConsumeClaim goroutine (inserter.go:75-88,176) on one GPU."""

    def __init__(self, key_mode="flows5m", device=0, cms=False, scale_sampling=False, columns=False, aggregate=True,
                 table_capacity=0, cms_depth=0, cms_width_log2=0, max_batch_bytes=0, max_batch_records=0, stream=None,
                 topk_only=False, topk_k=0):
        self._L = load_library()
        self.key_mode = KEY_MODES[key_mode] if isinstance(key_mode, str) else int(key_mode)
        self.kw = KEY_WORDS[self.key_mode]
        flags = (FA_CFG_CMS if cms else 0) | (FA_CFG_SCALE_SAMPLING if scale_sampling else 0) | (FA_CFG_COLUMNS if columns else 0)
        if topk_only:
            flags |= FA_CFG_TOPK_ONLY | FA_CFG_CMS
        if not aggregate:
            flags |= FA_CFG_NO_AGGREGATE
        if stream is not None:
            flags |= FA_CFG_CALLER_STREAM  # also when the handle is 0: torch's default stream
        self.cfg = FaConfig(FA_ABI_VERSION, device, self.key_mode, flags, table_capacity, cms_depth, cms_width_log2,
                            max_batch_bytes, max_batch_records, topk_k, stream)
        self.cms_depth = cms_depth or 4
        self.cms_width_log2 = cms_width_log2 or 20
        self._h = C.c_void_p()
        rc = self._L.fa_create(C.byref(self.cfg), C.byref(self._h))
        if rc:
            detail = self._L.fa_last_error(self._h).decode() if self._h else ""
            if self._h:
                self._L.fa_destroy(self._h)
                self._h = C.c_void_p()
            raise FlowAggError(rc, "fa_create", detail)

    # -- plumbing
    def _check(self, rc, what, ok=(0,)):
        if rc not in ok:
            raise FlowAggError(rc, what, self._L.fa_last_error(self._h).decode())
        return rc

    def close(self):
        if self._h:
            self._L.fa_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- ingest (inserter.go:113-165)
    def host_buffer(self, slot=0):
        buf, off = C.c_void_p(), C.c_void_p()
        cb, cr = C.c_size_t(), C.c_size_t()
        self._check(self._L.fa_host_buffer(self._h, slot, C.byref(buf), C.byref(cb), C.byref(off), C.byref(cr)), "fa_host_buffer")
        b = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_uint8)), shape=(cb.value,))
        o = np.ctypeslib.as_array(C.cast(off, C.POINTER(C.c_uint32)), shape=(cr.value + 1,))
        return b, o

    def submit(self, buf, offsets, framed=True, n_records=None, nbytes=None):
        """Host buffers (numpy arrays or pinned torch tensors).  offsets=None (framed streams of at most max_batch_bytes): the record
        boundaries are found on the GPU, and the call runs one behind -- it stages this buffer and launches the previously staged one;
        stats / flush / sync finish the pending one (include/flowagg.h)."""
        n = (len(offsets) - 1) if offsets is not None and n_records is None else (n_records or 0)
        ln = nbytes if nbytes is not None else (buf.nbytes if isinstance(buf, np.ndarray) else buf.numel() * buf.element_size())
        self._check(self._L.fa_submit(self._h, _ptr(buf), ln, _ptr(offsets), n, FA_FRAMED if framed else 0), "fa_submit")

    def submit_device(self, d_buf, d_offsets, n_records, nbytes, framed=True):
        """Device buffers (torch CUDA tensors or raw addresses)."""
        self._check(self._L.fa_submit_device(self._h, _ptr(d_buf), nbytes, _ptr(d_offsets), n_records, FA_FRAMED if framed else 0),
                    "fa_submit_device")

    def sync(self):
        self._check(self._L.fa_sync(self._h), "fa_sync")

    def stats(self):
        s = FaStats()
        self._check(self._L.fa_stats_get(self._h, C.byref(s)), "fa_stats_get")
        return {k: int(getattr(s, k)) for k, _ in FaStats._fields_}

    # -- emit (inserter.go:90-111 -> flows_5m rows, create.sh:70-110)
    def flush(self, keep=False, sort=True, allow_full=False, out=None):
        """Roll-up rows in ORDER BY order.  out: a caller-owned ROW_DTYPE array to fill and reuse across flushes,
        the way a C/Go caller reuses its row slice (pinned memory receives the rows without a staging copy);
        the returned array is then a view of it."""
        flags = (FA_FLUSH_KEEP if keep else 0) | (0 if sort else FA_FLUSH_UNSORTED)
        ok = (0, -5) if allow_full else (0,)
        n = C.c_size_t()
        cap = len(out) if out is not None else getattr(self, "_rows_cap", 1 << 16)
        while True:
            rows = out if out is not None and len(out) >= cap else np.empty(cap, dtype=ROW_DTYPE)
            rc = self._L.fa_flush(self._h, rows.ctypes.data, len(rows), C.byref(n), flags)
            if rc == -4 and n.value > len(rows):  # FA_ERR_CAPACITY: nothing was reset, retry with the size it reported
                cap = n.value
                out = None
                continue
            self._check(rc, "fa_flush", ok)
            self._rows_cap = max(len(rows), 1 << 16)
            return rows[: n.value]

    def flush_begin(self, keep=False, sort=True):
        """First half of an asynchronous flush: swap tables and enqueue the drain on the side stream; returns at once.
        Submits issued from now on already fill the fresh table."""
        flags = (FA_FLUSH_KEEP if keep else 0) | (0 if sort else FA_FLUSH_UNSORTED)
        self._check(self._L.fa_flush_begin(self._h, flags), "fa_flush_begin")

    def flush_end(self, allow_full=False, out=None):
        """Second half: wait for the drain (blocking, not spinning) and return the rows (same conventions as flush)."""
        ok = (0, -5) if allow_full else (0,)
        n = C.c_size_t()
        cap = len(out) if out is not None else getattr(self, "_rows_cap", 1 << 16)
        while True:
            rows = out if out is not None and len(out) >= cap else np.empty(cap, dtype=ROW_DTYPE)
            rc = self._L.fa_flush_end(self._h, rows.ctypes.data, len(rows), C.byref(n))
            if rc == -4 and n.value > len(rows):  # FA_ERR_CAPACITY: the rows are kept, retry with the size it reported
                cap = n.value
                out = None
                continue
            self._check(rc, "fa_flush_end", ok)
            self._rows_cap = max(len(rows), 1 << 16)
            return rows[: n.value]

    def merge_rows(self, rows, n=None, owner=0, n_owners=1):
        """SummingMergeTree's merge (create.sh:88-90): add rows that are already aggregates (another context's or an
        earlier window's flush output).  rows: a ROW_DTYPE array, or a device tensor / address with n rows.
        n_owners > 1: only the rows whose key this context owns (row_owner(...) == owner)."""
        if isinstance(rows, np.ndarray):
            rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
            n = len(rows)
        self._check(self._L.fa_merge_rows(self._h, _ptr(rows), int(n), owner, n_owners), "fa_merge_rows")

    @staticmethod
    def flush_box(ctxs, sort=True, allow_full=False):
        """Exact roll-up over every context of the box (one per GPU): hash-partitioned exchange over peer memory,
        then every context's share, in ORDER BY order.  Resets every context."""
        L = load_library()
        arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
        flags = 0 if sort else FA_FLUSH_UNSORTED
        n = C.c_size_t()
        cap = 1 << 16
        while True:
            rows = np.empty(cap, dtype=ROW_DTYPE)
            rc = L.fa_flush_box(arr, len(ctxs), rows.ctypes.data, len(rows), C.byref(n), flags)
            if rc == -4 and n.value > len(rows):
                cap = n.value
                continue
            if rc and not (allow_full and rc == -5):
                raise FlowAggError(rc, "fa_flush_box", "; ".join(L.fa_last_error(c._h).decode() for c in ctxs))
            return rows[: n.value]

    def reset(self):
        self._check(self._L.fa_reset(self._h), "fa_reset")

    # -- sketch / heavy hitters
    def cms_read(self):
        words = self.cms_depth << self.cms_width_log2
        out = np.zeros(words, dtype=np.uint64)
        self._check(self._L.fa_cms_read(self._h, out.ctypes.data, words), "fa_cms_read")
        return out

    def cms_device(self, which=FA_CMS_LOCAL):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self._L.fa_cms_device(self._h, which, C.byref(p), C.byref(n)), "fa_cms_device")
        return p.value, n.value

    def topk_local(self, k, which=FA_CMS_LOCAL):
        out = np.zeros(k, dtype=HH_DTYPE)
        n = C.c_size_t()
        self._check(self._L.fa_topk_local(self._h, which, k, out.ctypes.data, C.byref(n)), "fa_topk_local")
        return out[: n.value]

    @staticmethod
    def topk_merge(lists, key_words, k):
        L = load_library()
        allv = np.ascontiguousarray(np.concatenate(lists)) if len(lists) else np.zeros(0, dtype=HH_DTYPE)
        out = np.zeros(k, dtype=HH_DTYPE)
        n = C.c_size_t()
        rc = L.fa_topk_merge(allv.ctypes.data if len(allv) else None, len(allv), key_words, k, out.ctypes.data, C.byref(n))
        if rc:
            raise FlowAggError(rc, "fa_topk_merge")
        return out[: n.value]

    @staticmethod
    def topk(ctxs, k):
        """Single-process box-wide top-K: NCCL all-reduce of the sketches when len(ctxs) > 1."""
        L = load_library()
        arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
        out = np.zeros(k, dtype=HH_DTYPE)
        n = C.c_size_t()
        rc = L.fa_topk(arr, len(ctxs), k, out.ctypes.data, C.byref(n))
        if rc:
            raise FlowAggError(rc, "fa_topk", L.fa_last_error(ctxs[0]._h).decode())
        return out[: n.value]

    # -- kernel-1 output
    def columns(self, names=None):
        st = self.stats()
        out = {}
        # n of the last submit: ask for 'valid' with a generous buffer
        cap = int(self.cfg.max_batch_records or (4 << 20))
        for name in (names or COLUMNS.keys()):
            dt = COLUMNS[name]
            width = 1
            if isinstance(dt, tuple):
                dt, width = dt
            arr = np.zeros(cap * width, dtype=dt)
            self._check(self._L.fa_columns_read(self._h, name.encode(), arr.ctypes.data, arr.nbytes), "fa_columns_read")
            out[name] = arr
        return out, st

    # -- timing on the context's stream
    def timer_start(self):
        self._check(self._L.fa_timer_start(self._h), "fa_timer_start")

    def timer_stop(self):
        ms = C.c_float()
        self._check(self._L.fa_timer_stop(self._h, C.byref(ms)), "fa_timer_stop")
        return ms.value

    # -- synthetic input on the device (mocker/mocker.go:57-102)
    def mocker_device(self, cfg: FaMockerConfig, first, n, d_buf, cap, d_offsets):
        nb = C.c_size_t()
        self._check(self._L.fa_mocker_device(self._h, C.byref(cfg), first, n, _ptr(d_buf), cap, _ptr(d_offsets), C.byref(nb)),
                    "fa_mocker_device")
        return nb.value
