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
