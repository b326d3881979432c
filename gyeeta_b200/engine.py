"""ctypes binding of libgysketch.so (the C ABI of include/gysketch.h). Fails loudly when the CUDA library is missing."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgysketch.so")

EVENT_DTYPE = np.dtype([("svc_id", "<u8"), ("flow_key", "<u8"), ("value", "<u4"), ("host_idx", "<u4"),
                        ("tsec", "<u4"), ("type", "<u2"), ("flags", "<u2")], align=True)
assert EVENT_DTYPE.itemsize == 32
SERIAL_DTYPE = np.dtype([("count", "<u8"), ("sum", "<i8")])
FLOW_EST_DTYPE = np.dtype([("flow_key", "<u8"), ("count", "<u4"), ("kbytes", "<u4")])

EV_CONNECT, EV_ACCEPT, EV_CLOSE_CLI, EV_CLOSE_SER, EV_RESP, EV_TASK, EV_ACTIVE = 1, 2, 3, 4, 5, 6, 7
EVF_CLI_ERROR, EVF_SER_ERROR = 1, 2
HIST_RESP_CUR, HIST_RESP_LAST, HIST_RESP_ALL, HIST_TASK_CPU_PCT, HIST_TASK_CPU_DELAY, HIST_TASK_BLKIO_DELAY, HIST_RESP_5MIN, HIST_RESP_5DAY, HIST_QPS, HIST_ACTIVE_CONN = range(10)
STATE_IDLE, STATE_GOOD, STATE_OK, STATE_BAD, STATE_SEVERE, STATE_DOWN = range(6)
ISSUE_NONE, ISSUE_LISTENER_TASKS, ISSUE_QPS_HIGH, ISSUE_ACTIVE_CONN_HIGH, ISSUE_SERVER_ERRORS, ISSUE_OS_CPU, ISSUE_OS_MEMORY, ISSUE_DEPENDENT, ISSUE_UNKNOWN = range(9)
TOPN_QPS, TOPN_CONNS, TOPN_NET, TOPN_ISSUE = range(4)
RAW_EVENT32, RAW_TCP_IPV4_EVENT, RAW_TCP_IPV4_RESP, RAW_TCP_IPV6_EVENT, RAW_TCP_IPV6_RESP, RAW_API_TRAN, RAW_RESP16, RAW_TCP24, RAW_TASK24 = range(9)
RESP16_DTYPE = np.dtype([("svc_id", "<u8"), ("usec", "<u4"), ("host_idx", "<u2"), ("cli_port", "u1"), ("flags", "u1")])
TCP24_DTYPE = np.dtype([("svc_id", "<u8"), ("flow_key", "<u8"), ("bytes", "<u4"), ("host_idx", "<u2"), ("type", "u1"), ("pad", "u1")])
TASK24_DTYPE = np.dtype([("aggr_task_id", "<u8"), ("cpu_pct", "<u4"), ("cpu_delay_msec", "<u4"), ("blkio_delay_msec", "<u4"), ("host_idx", "<u2"), ("pad", "<u2")])
PROC_SAMPLE_DTYPE = np.dtype([("aggr_task_id", "<u8"), ("pid", "<i4"), ("cpu_pct", "<f4"), ("rss_mb", "<u4"), ("cpu_delay_msec", "<u4"),
                              ("vm_delay_msec", "<u4"), ("blkio_delay_msec", "<u4"), ("tcp_kbytes", "<u4"), ("tcp_conns", "<u4"), ("state", "u1"),
                              ("issue", "u1"), ("issue_bit_hist", "u1"), ("severe_issue_bit_hist", "u1"), ("is_issue", "u1"), ("pad", "u1", 3),
                              ("comm", "S16")])
assert PROC_SAMPLE_DTYPE.itemsize == 64
assert RESP16_DTYPE.itemsize == 16 and TCP24_DTYPE.itemsize == 24 and TASK24_DTYPE.itemsize == 24
NOTIFY_LISTENER_STATE, NOTIFY_TCP_CONN, NOTIFY_AGGR_TASK_STATE, NOTIFY_ACTIVE_CONN_STATS = 0x309, 0x30C, 0x310, 0x312
(HOSTTOP_SVC_ISSUE, HOSTTOP_SVC_QPS, HOSTTOP_SVC_CONNS, HOSTTOP_SVC_NET, HOSTTOP_TASK_ISSUE, HOSTTOP_TASK_NET, HOSTTOP_TASK_CPU, HOSTTOP_TASK_RSS,
 HOSTTOP_TASK_CPU_DELAY, HOSTTOP_TASK_VM_DELAY, HOSTTOP_TASK_BLKIO_DELAY) = range(11)
FLAG_AUTO_REGISTER = 1
TD_CAP = 256


class GyskError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"gysketch error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("max_svcs", C.c_uint32), ("max_tasks", C.c_uint32),
                ("cms_depth", C.c_uint32), ("cms_log2_width", C.c_uint32), ("hll_p", C.c_uint32),
                ("td_compression", C.c_uint32), ("max_batch", C.c_uint32), ("flags", C.c_uint32), ("rank", C.c_uint32),
                ("world", C.c_uint32), ("stage_batch", C.c_uint32), ("idle_evict_secs", C.c_uint32), ("reserved", C.c_uint32 * 2)]


class SvcSummary(C.Structure):
    _fields_ = [("glob_id", C.c_uint64), ("found", C.c_int32), ("nqrys_5s", C.c_uint32), ("total_resp_5sec", C.c_uint64),
                ("p95_5s_resp_ms", C.c_int64), ("p99_5s_resp_ms", C.c_int64), ("p25_5s_resp_ms", C.c_int64),
                ("p95_5min_resp_ms", C.c_int64), ("p99_5min_resp_ms", C.c_int64), ("nqrys_5min", C.c_uint64),
                ("p95_5day_resp_ms", C.c_int64), ("nqrys_5day", C.c_uint64),
                ("p95_all_resp_ms", C.c_int64), ("p99_all_resp_ms", C.c_int64), ("nqrys_all", C.c_uint64),
                ("max_resp_ms", C.c_int64), ("nconns_5s", C.c_uint32), ("kbytes_5s", C.c_uint32), ("nconns_all", C.c_uint64),
                ("kbytes_all", C.c_uint64), ("distinct_clients", C.c_double), ("td_p50_us", C.c_double),
                ("td_p95_us", C.c_double), ("td_p99_us", C.c_double), ("td_count", C.c_uint64),
                ("nconns_active", C.c_uint32), ("active_kbytes", C.c_uint32), ("max_rtt_msec", C.c_float),
                ("cli_errors", C.c_uint32), ("ser_errors", C.c_uint32), ("curr_state", C.c_uint8), ("curr_issue", C.c_uint8),
                ("issue_bit_hist", C.c_uint8), ("high_resp_bit_hist", C.c_uint8)]

    def asdict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class ListenerStateIn(C.Structure):
    """gysk_listener_state_in: the inputs of TCP_LISTENER::get_curr_state (include/gysketch.h)"""
    _fields_ = [(n, C.c_int64) for n in ("r5p95", "r5p99", "r300p95", "r300p99", "r5dp95", "r5dp99", "r5dp25", "rallp95", "rallp99")] + \
               [(n, C.c_uint64) for n in ("nqrys_5s", "total_resp_msec", "tcount_5d")] + \
               [(n, C.c_double) for n in ("mean5", "mean300", "mean5d", "meanall")] + \
               [(n, C.c_int64) for n in ("qps_p95", "qps_p25", "act_p95", "act_p25", "secs_5d")] + \
               [("last_qps_count", C.c_int32), ("nconn", C.c_int32), ("curr_active_conn", C.c_int32), ("ser_errors", C.c_uint32),
                ("nactive_conn_arr", C.c_uint8 * 16), ("task_issue", C.c_uint8), ("task_severe", C.c_uint8), ("task_delay", C.c_uint8),
                ("cpu_issue", C.c_uint8), ("mem_issue", C.c_uint8), ("pad0", C.c_uint8 * 3), ("ntasks_issue", C.c_int32),
                ("ntasks_noissue", C.c_int32), ("tasks_delay_msec", C.c_uint64), ("nserdepends", C.c_uint32), ("pad1", C.c_uint32)]


def classify_listener(inp, high_resp_bit_hist=0):
    """gysk_classify_listener: -> (state, issue, new high_resp_bit_hist)"""
    L = load_library()
    hb, st, iss = C.c_uint8(high_resp_bit_hist), C.c_uint8(), C.c_uint8()
    rc = L.gysk_classify_listener(C.byref(inp), C.byref(hb), C.byref(st), C.byref(iss))
    if rc:
        raise GyskError(rc, "gysk_classify_listener")
    return st.value, iss.value, hb.value


class HostSummary(C.Structure):
    _fields_ = [("nstates", C.c_int32 * 8)] + [(n, C.c_int32) for n in ("tot_qps", "tot_act_conn", "tot_kb_inbound", "tot_kb_outbound",
                                                                         "tot_ser_errors", "nlisteners", "nactive", "pad")]


class ClusterState(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("nhosts", "nsvc_issue", "nsvcissue_hosts", "nsvc", "total_qps", "svc_net_mb")] + [("pad", C.c_uint32 * 2)]


class TopnEntry(C.Structure):
    _fields_ = [("glob_id", C.c_uint64), ("score", C.c_uint64), ("host_idx", C.c_uint32), ("pad", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("events_in", "events_dropped", "events_resp", "events_tcp", "events_task", "nsvcs",
                                          "ntasks", "batches", "kernel_launches", "wire_msgs_ok", "wire_msgs_bad", "svcs_evicted")]

    def asdict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class BufferDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("dptr", C.c_void_p), ("nbytes", C.c_uint64), ("redop", C.c_int32), ("pad", C.c_int32)]


_lib = None


def load_library(path=None):
    """dlopen libgysketch.so. Raises when it has not been built: there is no fallback implementation."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise GyskError(-19, f"{path} is missing: build it with `python -m gyeeta_b200.build` "
                             "(CUDA library is the product; no CPU fallback exists)")
    L = C.CDLL(path)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    sig = {
        "gysk_abi_version": (i32, []),
        "gysk_config_default": (None, [vp]),
        "gysk_create": (i32, [vp, vp]),
        "gysk_destroy": (None, [vp]),
        "gysk_last_error": (C.c_char_p, [vp]),
        "gysk_get_stats": (i32, [vp, vp]),
        "gysk_hot_rows_in_use": (C.c_int64, [vp]),
        "gysk_last_batch_keys": (C.c_int64, [vp]),
        "gysk_register_ids": (i32, [vp, vp, u32, i32]),
        "gysk_ingest": (i32, [vp, vp, u32, u32, vp, u32, vp]),
        "gysk_ingest_msg": (i32, [vp, vp, u32, vp, u32]),
        "gysk_ingest_raw": (i32, [vp, vp, u32, u32, vp, u32]),
        "gysk_ingest_pinned": (i32, [vp, vp, u64]),
        "gysk_ingest_device": (i32, [vp, vp, u64]),
        "gysk_sync": (i32, [vp]),
        "gysk_flush": (i32, [vp, u32]),
        "gysk_evicted_ids": (i32, [vp, vp, u32, vp]),
        "gysk_query_svcs": (i32, [vp, vp, u32, vp]),
        "gysk_query_flows": (i32, [vp, vp, u32, i32, vp]),
        "gysk_query_host_summary": (i32, [vp, u32, vp]),
        "gysk_topn_svcs": (i32, [vp, i32, C.c_int32, u32, vp, vp]),
        "gysk_topn_tasks": (i32, [vp, i32, u32, vp, vp]),
        "gysk_topn_host": (i32, [vp, i32, C.c_int32, u32, vp, vp]),
        "gysk_query_cluster_state": (i32, [vp, vp, u32, vp]),
        "gysk_export_hist": (i32, [vp, u64, i32, vp, vp, vp]),
        "gysk_export_task_hist": (i32, [vp, u64, i32, vp, vp, vp]),
        "gysk_export_hll": (i32, [vp, u64, vp]),
        "gysk_export_conn_bitmap": (i32, [vp, u64, i32, vp, vp]),
        "gysk_export_tdigest": (i32, [vp, u64, vp, vp, u32, vp, vp, vp]),
        "gysk_query_quantiles": (i32, [vp, u64, vp, u32, vp]),
        "gysk_tdigest_to_pgtext": (i32, [vp, vp, u32, u32, vp, u32]),
        "gysk_encode_listener_state": (i32, [vp, u32, vp, u32, vp, vp]),
        "gysk_export_tdigest_pgtext": (i32, [vp, u64, vp, u32]),
        "gysk_export_cms": (i32, [vp, i32, vp]),
        "gysk_hist_nbuckets": (i32, [i32]),
        "gysk_hist_bucket": (i32, [i32, C.c_int64]),
        "gysk_hist_percentiles": (i32, [i32, i32, vp, u64, vp, u32, vp]),
        "gysk_classify_listener": (i32, [vp, vp, vp, vp]),
        "gysk_task_groupby": (i32, [vp, vp, u32, vp, u32, vp]),
        "gysk_hll_estimate": (C.c_double, [vp, u32]),
        "gysk_tdigest_quantile": (C.c_double, [vp, vp, u32, C.c_double, C.c_double, C.c_double]),
        "gysk_uint64_hash": (u32, [u64]),
        "gysk_set_logical_map": (i32, [vp, vp, vp, u32]),
        "gysk_merge_prepare": (i32, [vp]),
        "gysk_merge_buffers": (i32, [vp, vp, u32, vp]),
        "gysk_merge_tdigest_slab": (i32, [vp, vp, vp]),
        "gysk_merge_finish": (i32, [vp, vp, u32]),
        "gysk_query_logical": (i32, [vp, vp, u32, vp]),
        "gysk_query_flows_global": (i32, [vp, vp, u32, i32, vp]),
        "gysk_nccl_unique_id": (i32, [vp]),
        "gysk_nccl_comm_init": (i32, [vp, vp, u32, u32]),
        "gysk_merge_global": (i32, [vp, vp]),
        "gysk_stream": (vp, [vp]),
        "gysk_profile_enable": (i32, [vp, i32]),
        "gysk_profile_read": (i32, [vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)          # AttributeError here = the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    if path == LIB_PATH:
        _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Engine:
    """One engine = one GPU. Mirrors the C ABI one to one."""

    def __init__(self, device=0, max_svcs=1 << 14, max_tasks=1 << 12, cms_depth=4, cms_log2_width=20, hll_p=12,
                 td_compression=200, max_batch=1 << 20, auto_register=True, rank=0, world=1, stage_batch=0, idle_evict_secs=0):
        self.L = load_library()
        cfg = Config()
        self.L.gysk_config_default(C.byref(cfg))
        cfg.device, cfg.max_svcs, cfg.max_tasks = device, max_svcs, max_tasks
        cfg.cms_depth, cfg.cms_log2_width, cfg.hll_p, cfg.td_compression = cms_depth, cms_log2_width, hll_p, td_compression
        cfg.max_batch = max_batch
        cfg.stage_batch = stage_batch
        cfg.idle_evict_secs = idle_evict_secs
        cfg.flags = FLAG_AUTO_REGISTER if auto_register else 0
        cfg.rank, cfg.world = rank, world
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = self.L.gysk_create(C.byref(cfg), C.byref(self.h))
        if rc:
            raise GyskError(rc, (self.L.gysk_last_error(None) or b"").decode())
        self._host_id = (C.c_uint8 * 16)()

    def _chk(self, rc):
        if rc:
            raise GyskError(rc, (self.L.gysk_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.gysk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- ingest ----
    def register_ids(self, ids, is_task=False):
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        self._chk(self.L.gysk_register_ids(self.h, _p(ids), len(ids), int(is_task)))

    def ingest_events(self, ev, host_idx=0):
        assert ev.dtype == EVENT_DTYPE
        ev = np.ascontiguousarray(ev)
        self._chk(self.L.gysk_ingest_raw(self.h, self._host_id, host_idx, RAW_EVENT32, _p(ev), len(ev)))

    def ingest_raw(self, kind, buf, n, host_idx=0):
        self._chk(self.L.gysk_ingest_raw(self.h, self._host_id, host_idx, kind, _p(buf), n))

    def ingest_raw_ptr(self, kind, ptr, n, host_idx=0):
        """ingest_raw on a raw host address (e.g. a page-locked torch tensor's data_ptr())"""
        self._chk(self.L.gysk_ingest_raw(self.h, self._host_id, host_idx, kind, C.c_void_p(ptr), n))

    def ingest_pinned_ptr(self, ptr, n):
        self._chk(self.L.gysk_ingest_pinned(self.h, C.c_void_p(ptr), n))

    def ingest_device_ptr(self, dptr, n):
        self._chk(self.L.gysk_ingest_device(self.h, C.c_void_p(dptr), n))

    def ingest_msg(self, msg, host_idx=0):
        """msg: writable bytes-like holding COMM_HEADER + EVENT_NOTIFY + records"""
        buf = np.frombuffer(msg, dtype=np.uint8)
        return self.L.gysk_ingest_msg(self.h, self._host_id, host_idx, _p(buf), len(buf))

    def sync(self):
        self._chk(self.L.gysk_sync(self.h))

    def flush(self, tsec=0):
        self._chk(self.L.gysk_flush(self.h, tsec))

    def evicted_ids(self, cap=1 << 16):
        """ids evicted by the most recent flush (LISTEN_FLAG_DELETE notifications)"""
        out = np.zeros(cap, dtype=np.uint64)
        n = C.c_uint32()
        self._chk(self.L.gysk_evicted_ids(self.h, _p(out), cap, C.byref(n)))
        return out[:min(n.value, cap)].copy()

    def stream(self):
        return self.L.gysk_stream(self.h)

    def profile_enable(self, on=True):
        self._chk(self.L.gysk_profile_enable(self.h, int(on)))

    def profile_read(self):
        a, b, n = C.c_double(), C.c_double(), C.c_uint64()
        self._chk(self.L.gysk_profile_read(self.h, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    # ---- queries ----
    def stats(self):
        s = Stats()
        self._chk(self.L.gysk_get_stats(self.h, C.byref(s)))
        return s.asdict()

    def hot_rows_in_use(self):
        """rows of dense value bins handed out to hot services so far (diagnostic; results never depend on it)"""
        n = self.L.gysk_hot_rows_in_use(self.h)
        if n < 0:
            self._chk(int(n))
        return int(n)

    def last_batch_keys(self):
        """response samples of the last device batch that travelled as sort keys (diagnostic)"""
        n = self.L.gysk_last_batch_keys(self.h)
        if n < 0:
            self._chk(int(n))
        return int(n)

    def query_svcs(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        out = (SvcSummary * len(ids))()
        self._chk(self.L.gysk_query_svcs(self.h, _p(ids), len(ids), out))
        return [o.asdict() for o in out]

    def listener_state_records(self, ids):
        """query_svcs + gysk_encode_listener_state: the LISTENER_STATE_NOTIFY records (bytes) of the known ids, <= 512"""
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        out = (SvcSummary * len(ids))()
        self._chk(self.L.gysk_query_svcs(self.h, _p(ids), len(ids), out))
        buf = C.create_string_buffer(88 * 512)
        nrecs, nbytes = C.c_uint32(), C.c_uint32()
        self._chk(self.L.gysk_encode_listener_state(out, len(ids), buf, len(buf), C.byref(nrecs), C.byref(nbytes)))
        return nrecs.value, buf.raw[: nbytes.value]

    def query_flows(self, keys, last_window=False):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        out = np.zeros(len(keys), dtype=FLOW_EST_DTYPE)
        self._chk(self.L.gysk_query_flows(self.h, _p(keys), len(keys), int(last_window), _p(out)))
        return out

    def topn(self, metric, n=10, host_idx=-1):
        out = (TopnEntry * n)()
        k = C.c_uint32()
        self._chk(self.L.gysk_topn_svcs(self.h, metric, host_idx, n, out, C.byref(k)))
        return [(o.glob_id, o.score, o.host_idx) for o in out[: k.value]]

    def topn_tasks(self, metric, n=10):
        out = (TopnEntry * n)()
        k = C.c_uint32()
        self._chk(self.L.gysk_topn_tasks(self.h, metric, n, out, C.byref(k)))
        return [(o.glob_id, o.score) for o in out[: k.value]]

    def task_groupby(self, samples, cap=None):
        """gysk_task_groupby: PROC_SAMPLE_DTYPE array -> (wire.TASK records of the groups in order of first appearance, number of groups)"""
        from .wire import TASK
        samples = np.ascontiguousarray(samples, dtype=PROC_SAMPLE_DTYPE)
        cap = len(samples) if cap is None else cap
        out = np.zeros(max(cap, 1), dtype=TASK)
        ng = C.c_uint32()
        self._chk(self.L.gysk_task_groupby(self.h, _p(samples), len(samples), _p(out), cap, C.byref(ng)))
        return out[: min(ng.value, cap)].copy(), ng.value

    def topn_host(self, what, n=10, host_idx=-1):
        out = (TopnEntry * n)()
        k = C.c_uint32()
        self._chk(self.L.gysk_topn_host(self.h, what, host_idx, n, out, C.byref(k)))
        return [(o.glob_id, o.score, o.host_idx) for o in out[: k.value]]

    def cluster_state(self, host_idxs=None):
        cs = ClusterState()
        if host_idxs is None:
            self._chk(self.L.gysk_query_cluster_state(self.h, None, 0, C.byref(cs)))
        else:
            h = np.ascontiguousarray(host_idxs, dtype=np.uint32)
            self._chk(self.L.gysk_query_cluster_state(self.h, _p(h), len(h), C.byref(cs)))
        return {f: getattr(cs, f) for f, _ in cs._fields_ if f != "pad"}

    def host_summary(self, host_idx):
        hs = HostSummary()
        rc = self.L.gysk_query_host_summary(self.h, host_idx, C.byref(hs))
        if rc == -2:
            return None
        self._chk(rc)
        d = {f: getattr(hs, f) for f, _ in hs._fields_ if f not in ("nstates", "pad")}
        d["nstates"] = list(hs.nstates)
        return d

    def export_hist(self, id_, which):
        out = np.zeros(15, dtype=SERIAL_DTYPE)
        total, mx = C.c_uint64(), C.c_int64()
        rc = self.L.gysk_export_hist(self.h, int(id_), which, _p(out), C.byref(total), C.byref(mx))
        if rc == -2:
            return None
        self._chk(rc)
        return out, total.value, mx.value

    def export_hll(self, id_):
        regs = np.zeros(1 << self.cfg.hll_p, dtype=np.uint8)
        rc = self.L.gysk_export_hll(self.h, int(id_), _p(regs))
        if rc == -2:
            return None
        self._chk(rc)
        return regs

    def export_conn_bitmap(self, id_, last_window=False):
        masks = np.zeros(15, dtype=np.uint32)
        cnt = np.zeros(15, dtype=np.uint8)
        rc = self.L.gysk_export_conn_bitmap(self.h, int(id_), int(last_window), _p(masks), _p(cnt))
        if rc == -2:
            return None
        self._chk(rc)
        return masks, cnt

    def export_tdigest(self, id_):
        means = np.zeros(TD_CAP, dtype=np.float64)
        weights = np.zeros(TD_CAP, dtype=np.uint64)
        n, mn, mx = C.c_uint32(), C.c_double(), C.c_double()
        rc = self.L.gysk_export_tdigest(self.h, int(id_), _p(means), _p(weights), TD_CAP, C.byref(n), C.byref(mn), C.byref(mx))
        if rc == -2:
            return None
        self._chk(rc)
        return means[: n.value].copy(), weights[: n.value].copy(), mn.value, mx.value

    def export_tdigest_pgtext(self, id_):
        buf = C.create_string_buffer(8192)
        rc = self.L.gysk_export_tdigest_pgtext(self.h, int(id_), buf, len(buf))
        if rc == -2:
            return None
        if rc < 0:
            self._chk(rc)
        return buf.value.decode()

    def quantiles(self, id_, qs):
        qs = np.ascontiguousarray(qs, dtype=np.float64)
        out = np.zeros(len(qs), dtype=np.float64)
        self._chk(self.L.gysk_query_quantiles(self.h, int(id_), _p(qs), len(qs), _p(out)))
        return out

    # ---- multi-GPU merge ----
    def set_logical_map(self, glob_ids, logical_ids):
        g = np.ascontiguousarray(glob_ids, dtype=np.uint64)
        l = np.ascontiguousarray(logical_ids, dtype=np.uint64)
        assert len(g) == len(l)
        self._chk(self.L.gysk_set_logical_map(self.h, _p(g), _p(l), len(g)))

    def merge_prepare(self):
        self._chk(self.L.gysk_merge_prepare(self.h))

    def merge_buffers(self):
        descs = (BufferDesc * 8)()
        n = C.c_uint32()
        self._chk(self.L.gysk_merge_buffers(self.h, descs, 8, C.byref(n)))
        return [(d.name.decode(), d.dptr, d.nbytes, d.redop) for d in descs[: n.value]]

    def merge_tdigest_slab(self):
        p, nb = C.c_void_p(), C.c_uint64()
        self._chk(self.L.gysk_merge_tdigest_slab(self.h, C.byref(p), C.byref(nb)))
        return p.value, nb.value

    def merge_finish(self, gathered_ptr=None, world=1):
        self._chk(self.L.gysk_merge_finish(self.h, C.c_void_p(gathered_ptr), world))

    def nccl_unique_id(self):
        buf = (C.c_uint8 * 128)()
        self._chk(self.L.gysk_nccl_unique_id(buf))
        return bytes(buf)

    def nccl_comm_init(self, uid, nranks, rank):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._chk(self.L.gysk_nccl_comm_init(self.h, buf, nranks, rank))

    def merge_global(self, comm=None):
        """fold + one grouped NCCL launch + merge-compress, all inside the library (gysk_merge_global)"""
        self._chk(self.L.gysk_merge_global(self.h, C.c_void_p(comm)))

    def query_logical(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        out = (SvcSummary * len(ids))()
        self._chk(self.L.gysk_query_logical(self.h, _p(ids), len(ids), out))
        return [o.asdict() for o in out]

    def query_flows_global(self, keys, last_window=False):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        out = np.zeros(len(keys), dtype=FLOW_EST_DTYPE)
        self._chk(self.L.gysk_query_flows_global(self.h, _p(keys), len(keys), int(last_window), _p(out)))
        return out

    def export_cms(self, last_window=False):
        out = np.zeros(self.cfg.cms_depth << self.cfg.cms_log2_width, dtype=np.uint64)
        self._chk(self.L.gysk_export_cms(self.h, int(last_window), _p(out)))
        return out
