"""BASELINE.json configs[4] in miniature (the sustained stream: 5-s windows, rolling levels, eviction of idle services, a fixed
list of global queries replayed every window): two engines as rank 0 / rank 1 of a 2-way host shard with idle eviction on, the
merge step after every flush, logical-service answers checked against the oracle every window."""
import numpy as np
import pytest

from gyeeta_b200 import engine as ge
from gyeeta_b200 import synth
from oracle import pyoracle as po
from tests.test_gpu_merge import _emulate_collectives

pytestmark = pytest.mark.gpu


def test_windows_eviction_and_global_query_replay():
    import torch
    rng = np.random.default_rng(55)
    nsvc, nhosts = 240, 32
    kw = dict(max_svcs=512, max_tasks=32, max_batch=1 << 15, cms_log2_width=12, idle_evict_secs=20)
    shards = [ge.Engine(rank=r, world=2, **kw) for r in range(2)]
    orc = po.OracleEngine(max_svcs=512, max_tasks=32, cms_log2_width=12)
    orc.set_idle_evict(20)
    ids = synth.service_ids(nsvc)
    logical = (np.arange(nsvc, dtype=np.uint64) // np.uint64(8)) + np.uint64(500)
    lids = np.unique(logical)
    busy = set(int(i) for i in ids[: nsvc // 2])                      # the second half goes silent after the second window
    evicted_total = set()
    for w, t in enumerate([5, 10, 15, 20, 25, 30, 35, 40, 45, 50, 55, 60]):
        ev = synth.gen_mixed(rng, 20_000, nsvc, ntask=8, nhosts=nhosts, nclients=2000)
        if w >= 2:
            ev = ev[(ev["type"] == ge.EV_TASK) | np.isin(ev["svc_id"], np.fromiter(busy, dtype=np.uint64))]
        ev["tsec"] = t
        for e in shards:
            e.ingest_events(ev); e.sync()
        orc.ingest(ev)
        for e in shards:
            e.flush(t)
        orc.flush(t)
        got = np.sort(np.concatenate([e.evicted_ids() for e in shards]))
        want, _ = orc.evicted_ids()
        assert np.array_equal(got, np.sort(want)), (t, len(got), len(want))
        evicted_total |= set(int(i) for i in got)
        if w == 0:
            for e in shards:
                e.set_logical_map(ids, logical)
        _emulate_collectives(torch, shards)
        a, b = shards[0].query_logical(lids), shards[1].query_logical(lids)
        for x, y, lid in zip(a, b, lids):
            members = ids[logical == lid]
            hs = [orc.export_hist(int(m), 1) for m in members]
            ha = [orc.export_hist(int(m), 2) for m in members]
            cs = [orc.export_conn(int(m)) for m in members]
            assert x["nqrys_5s"] == y["nqrys_5s"] == sum(h[1] for h in hs if h is not None), (t, int(lid))
            assert x["nqrys_all"] == y["nqrys_all"] == sum(h[1] for h in ha if h is not None), (t, int(lid))
            assert x["nconns_5s"] == sum(c[1] & 0xFFFFFFFF for c in cs if c is not None)
            regs = np.zeros(4096, dtype=np.uint8)
            for m in members:
                r = orc.export_hll(int(m))
                if r is not None:
                    regs = np.maximum(regs, r)
            assert x["distinct_clients"] == y["distinct_clients"] == po.lib().gyo_hll_estimate(po._p(regs), 12)
    # first seen t = 5, last active t = 10, idle 20 s: gone at the first flush with t > 45
    assert evicted_total == set(int(i) for i in ids) - busy
    st = [e.stats() for e in shards]
    assert sum(s["svcs_evicted"] for s in st) == len(evicted_total) and sum(s["nsvcs"] for s in st) == orc.nsvcs()
