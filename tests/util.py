"""shared helpers for the GPU parity tests: run the same seeded event stream through the CUDA engine (via the C ABI)
and through the CPU oracle, and compare."""
import numpy as np

from gyeeta_b200 import engine as ge
from oracle import pyoracle as po

assert ge.EVENT_DTYPE == po.EVENT_DTYPE


def make_pair(**kw):
    """(Engine, OracleEngine) with identical configuration"""
    eng = ge.Engine(**kw)
    idle = kw.get("idle_evict_secs", 0)
    ocfg = dict(max_svcs=kw.get("max_svcs", 1 << 14), max_tasks=kw.get("max_tasks", 1 << 12), cms_depth=kw.get("cms_depth", 4),
                cms_log2_width=kw.get("cms_log2_width", 20), hll_p=kw.get("hll_p", 12), td_compression=kw.get("td_compression", 200),
                flags=1 if kw.get("auto_register", True) else 0, rank=kw.get("rank", 0), world=kw.get("world", 1))
    orc = po.OracleEngine(**ocfg)
    if idle:
        orc.set_idle_evict(idle)
    return eng, orc


def feed_both(eng, orc, ev, batch):
    """identical device batches on both sides (the t-digest update is per batch)"""
    for off in range(0, len(ev), batch):
        chunk = ev[off: off + batch]
        eng.ingest_events(chunk)
        eng.sync()
        orc.ingest(chunk)


def assert_hist_equal(eng, orc, id_, which):
    a = eng.export_hist(id_, which)
    b = orc.export_hist(id_, which)
    assert (a is None) == (b is None), (hex(id_), which)
    if a is None:
        return 0
    assert np.array_equal(a[0]["count"], b[0]["count"]), (hex(id_), which, a[0]["count"], b[0]["count"])
    assert np.array_equal(a[0]["sum"], b[0]["sum"]), (hex(id_), which)
    assert a[1] == b[1] and a[2] == b[2], (hex(id_), which, a[1:], b[1:])
    return a[1]


def exact_quantile(vals, q):
    v = np.sort(np.asarray(vals, dtype=np.float64))
    return float(v[min(len(v) - 1, max(0, int(np.ceil(q * len(v))) - 1))])


def threaded_oracle(ev, nthreads, **ocfg):
    """The CPU oracle over a large stream: events pre-sharded by host_idx % nthreads (every per-service / per-task sketch lives
    wholly in one shard), one oracle engine per thread, each shard ingested as ONE device batch. Returns
    (engines, owner) with owner(id, host_idx) -> engine; the count-min table of the whole stream is the sum of the shard tables."""
    import ctypes as C
    L = po.lib()
    owner = (ev["host_idx"] % nthreads).astype(np.int32)
    order = np.argsort(owner, kind="stable")
    cuts = np.searchsorted(owner[order], np.arange(1, nthreads))
    shards = [np.ascontiguousarray(a) for a in np.split(ev[order], cuts)]
    engines = [po.OracleEngine(**ocfg) for _ in range(nthreads)]
    eh = (C.c_void_p * nthreads)(*[e.h for e in engines])
    sp = (C.c_void_p * nthreads)(*[s_.ctypes.data for s_ in shards])
    cn = (C.c_uint64 * nthreads)(*[len(s_) for s_ in shards])
    L.gyo_bench_ingest(eh, sp, cn, nthreads, 0)               # batch 0 = the whole shard in one gyo_ingest call
    return engines, shards


def td_p99_tolerance(n, eps=0.01):
    """value tolerance of the digest's p99 against the EXACT sample quantile on the sigma <= 1.5 log-normal streams.
    Two parts: the systematic interpolation error of a delta = 200 K_1 digest (~0.3 %) — and the order-statistic noise of the
    sample itself: the p99 cluster holds n pi sqrt(.0099) / 200 samples whose individual positions the digest does not keep; the
    exact quantile, one of them, deviates from the cluster's straight line by ~ sqrt(cluster / 4) sample spacings of
    sigma / (phi(z_99) n) each (1 sigma = 0.5 % at n = 47 K, falling as 1 / sqrt(n)). 1 % holds from n = 150 K on; below that
    no sketch of a few hundred centroids can promise it, and the bound is 3.5 sigma of that noise."""
    return max(eps, 3.5 * 0.005 * float(np.sqrt(47_000.0 / n)))
