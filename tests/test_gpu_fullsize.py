"""BASELINE.json configurations at FULL size: bit-exact against the threaded CPU oracle (events pre-sharded by host, one oracle
engine per host core), and size-independent properties: conservation (every event lands in exactly one counter / bucket), the count-min checksum of checksums (every row sums to
the number of TCP events and to the total kbytes), idempotence of HLL under replay and exact doubling of the additive state,
t-digest weight conservation and rank error, sortedness of centroids."""
import numpy as np
import pytest

from gyeeta_b200 import engine as ge
from gyeeta_b200 import synth

pytestmark = pytest.mark.gpu


def _ingest(eng, ev, chunk=1 << 22):
    for off in range(0, len(ev), chunk):
        eng.ingest_events(ev[off: off + chunk])
    eng.sync()


def test_config1_one_million_samples_one_service():
    rng = np.random.default_rng(1)
    ev = synth.gen_resp_config1(rng, 1_000_000)
    id_ = int(ev["svc_id"][0])
    eng = ge.Engine(max_svcs=64, max_tasks=8, max_batch=1 << 20)
    _ingest(eng, ev)
    hist, total, mx = eng.export_hist(id_, ge.HIST_RESP_CUR)
    ms = ev["value"] // 1000
    assert total == 1_000_000 and int(hist["sum"].sum()) == int(ms.sum()) and mx == int(ms.max())
    # bucket counts against numpy's own histogram over the reference thresholds (gy_statistics.h:1677)
    thr = np.array([1, 10, 30, 60, 100, 150, 200, 300, 450, 700, 1000, 3000, 15000])
    want = np.bincount(1 + np.searchsorted(thr, ms, side="left"), minlength=15)
    assert np.array_equal(hist["count"], want.astype(np.uint64))
    means, weights, mn, mxv = eng.export_tdigest(id_)
    assert int(weights.sum()) == 1_000_000 and np.all(np.diff(means) >= 0) and mn == ev["value"].min() and mxv == ev["value"].max()
    sv = np.sort(ev["value"])
    for q, g in zip((0.5, 0.95, 0.99), eng.quantiles(id_, [0.5, 0.95, 0.99])):
        ex = float(sv[int(np.ceil(q * len(sv))) - 1])
        assert abs(np.searchsorted(sv, g) / len(sv) - q) < 0.001
        assert abs(g - ex) / ex < 0.01, (q, g, ex)


def test_config2_ten_million_tcp_events_10k_services():
    rng = np.random.default_rng(2)
    ev = synth.gen_tcp(rng, 10_000_000, 10_000, zipf_s=1.1, nclients=1_000_000, nhosts=512)
    eng = ge.Engine(max_svcs=1 << 14, max_tasks=8, max_batch=1 << 22)
    _ingest(eng, ev)
    st = eng.stats()
    assert st["events_in"] == st["events_tcp"] == len(ev) and st["events_dropped"] == 0
    cms = eng.export_cms().reshape(4, -1)
    kb = int((ev["value"] >> 10).astype(np.uint64).sum())
    for r in range(4):                                   # checksum of checksums: each row holds every event exactly once
        assert int((cms[r] & np.uint64(0xFFFFFFFF)).sum()) == len(ev)
        assert int((cms[r] >> np.uint64(32)).sum()) == kb
    # count-min never underestimates; overshoot bounded by e/w * N with high probability
    keys, cnt = np.unique(ev["flow_key"], return_counts=True)
    pick = rng.choice(len(keys), 5000, replace=False)
    est = eng.query_flows(keys[pick])
    assert np.all(est["count"] >= cnt[pick])
    assert np.mean(est["count"] - cnt[pick] <= np.e / (1 << 20) * len(ev)) > 0.98
    # HLL per service within 3 sigma of the exact distinct count (p = 12), on the 20 most popular services
    ids, c = np.unique(ev["svc_id"], return_counts=True)
    top = ids[np.argsort(-c)[:20]]
    summ = eng.query_svcs(top)
    before = [eng.export_hll(int(i)).copy() for i in top[:5]]
    for s_, id_ in zip(summ, top):
        exact = len(np.unique(ev["flow_key"][ev["svc_id"] == id_]))
        assert abs(s_["distinct_clients"] - exact) / exact < 3 * 1.04 / 64, (exact, s_["distinct_clients"])
    # replay the same stream: HLL registers are idempotent, the additive state doubles exactly
    _ingest(eng, ev)
    for i, b in zip(top[:5], before):
        assert np.array_equal(eng.export_hll(int(i)), b)
    cms2 = eng.export_cms().reshape(4, -1)
    assert np.array_equal(cms2, cms * np.uint64(2))


def test_config3_hundred_million_mixed_events_properties():
    import torch
    import bench
    dev = torch.device("cuda", 0)
    n = 100_000_000
    evd = bench.gen_events_gpu(torch, n, 99, 0, 1, dev)
    torch.cuda.synchronize()
    eng = ge.Engine(max_svcs=1 << 17, max_tasks=1 << 15, max_batch=1 << 24)
    eng.ingest_device_ptr(evd.data_ptr(), n)
    eng.sync()
    st = eng.stats()
    etype = (evd[:, 3] >> 32) & 0xFFFF
    n_resp, n_task = int((etype == 5).sum()), int((etype == 6).sum())
    n_tcp = n - n_resp - n_task
    assert st["events_in"] == n and st["events_dropped"] == 0
    assert (st["events_resp"], st["events_tcp"], st["events_task"]) == (n_resp, n_tcp, n_task)
    cms = eng.export_cms().reshape(4, -1)
    vals = evd[:, 2] & 0xFFFFFFFF
    kb = int((vals[(etype >= 1) & (etype <= 4)] >> 10).sum())
    for r in range(4):
        assert int((cms[r] & np.uint64(0xFFFFFFFF)).sum()) == n_tcp and int((cms[r] >> np.uint64(32)).sum()) == kb
    # conservation over all services: histogram totals and t-digest weights both add up to the RESP events, msec sums match
    ids = synth.service_ids(bench.NSVC)
    tot_hist = tot_td = tot_sum = tot_conn = 0
    eng.flush(5)
    for off in range(0, len(ids), 8192):
        for s_ in eng.query_svcs(ids[off: off + 8192]):
            if s_["found"]:
                tot_hist += s_["nqrys_5s"]; tot_td += s_["td_count"]; tot_sum += s_["total_resp_5sec"]; tot_conn += s_["nconns_5s"]
    assert tot_hist == tot_td == n_resp and tot_conn == n_tcp
    assert tot_sum == int((vals[etype == 5] // 1000).sum())
    # one hot service: rank error of the digest against the exact empirical distribution
    u, c = torch.unique(evd[:4_000_000, 0][etype[:4_000_000] == 5], return_counts=True)
    sid = int(u[torch.argmax(c)])
    sv = torch.sort(vals[(evd[:, 0] == sid) & (etype == 5)]).values
    got = eng.quantiles(sid & 0xFFFFFFFFFFFFFFFF, [0.5, 0.95, 0.99])
    for q, g in zip((0.5, 0.95, 0.99), got):
        rank = int(torch.searchsorted(sv, torch.tensor([int(g)], device=dev))[0]) / sv.numel()
        assert abs(rank - q) < 0.001, (q, g, rank)


def test_config2_bit_exact_vs_threaded_oracle():
    """configs[1] at full size (10 M tcp_conn events, 10 K services): the WHOLE count-min table, the HLL registers and the exact
    {count, kbytes} cell of EVERY service, bit for bit against the CPU oracle"""
    import os
    from tests.util import threaded_oracle
    rng = np.random.default_rng(2)
    ev = synth.gen_tcp(rng, 10_000_000, 10_000, zipf_s=1.1, nclients=1_000_000, nhosts=512)
    eng = ge.Engine(max_svcs=1 << 14, max_tasks=8, max_batch=1 << 24)
    _ingest(eng, ev, chunk=1 << 24)
    nthr = max(1, min(os.cpu_count() or 1, 64))
    orcs, _ = threaded_oracle(ev, nthr, max_svcs=1 << 14, max_tasks=8)
    cms = np.zeros(4 << 20, dtype=np.uint64)
    for o in orcs:
        cms += o.cms()
    assert np.array_equal(eng.export_cms(), cms)
    ids, first = np.unique(ev["svc_id"], return_index=True)
    assert len(ids) > 9000
    for id_, i0 in zip(ids, first):
        o = orcs[int(ev["host_idx"][i0]) % nthr]
        assert np.array_equal(eng.export_hll(int(id_)), o.export_hll(int(id_))), hex(int(id_))
    summ = eng.query_svcs(ids)
    eng.flush(5)
    for o in orcs:
        o.flush(5)
    summ = eng.query_svcs(ids)
    for s_, id_, i0 in zip(summ, ids, first):
        c = orcs[int(ev["host_idx"][i0]) % nthr].export_conn(int(id_))
        assert (s_["nconns_5s"], s_["kbytes_5s"]) == (c[1] & 0xFFFFFFFF, c[1] >> 32)


def test_config3_bit_exact_sampled_services_vs_threaded_oracle():
    """configs[2] at full size (100 M mixed events, 100 K services, one device batch): count-min table, and on 1 500 sampled
    services (hot, lukewarm and cold) the RESP histogram, CONN_BITMAP, HLL registers and the t-digest — centroid for centroid —
    plus 300 sampled task histograms, bit for bit against the CPU oracle"""
    import os
    import torch
    import bench
    from tests.util import threaded_oracle
    dev = torch.device("cuda", 0)
    n = 100_000_000
    evd = bench.gen_events_gpu(torch, n, 77, 0, 1, dev)
    torch.cuda.synchronize()
    eng = ge.Engine(max_svcs=1 << 17, max_tasks=1 << 15, max_batch=(1 << 27) - 1)
    eng.ingest_device_ptr(evd.data_ptr(), n)
    eng.sync()
    ev = evd.cpu().numpy().view(np.uint8).reshape(-1).view(ge.EVENT_DTYPE)
    del evd
    torch.cuda.empty_cache()
    nthr = max(1, min(os.cpu_count() or 1, 128))
    orcs, _ = threaded_oracle(ev, nthr, max_svcs=bench.NSVC + 16, max_tasks=bench.NTASK + 16)
    cms = np.zeros(4 << 20, dtype=np.uint64)
    for o in orcs:
        cms += o.cms()
    assert np.array_equal(eng.export_cms(), cms)
    st = eng.stats()
    tot = {k: sum(o.counters()[k] for o in orcs) for k in ("in", "dropped", "resp", "tcp", "task", "nsvcs")}
    assert (st["events_in"], st["events_dropped"], st["events_resp"], st["events_tcp"], st["events_task"], st["nsvcs"]) == \
        tuple(tot[k] for k in ("in", "dropped", "resp", "tcp", "task", "nsvcs"))
    head = ev[:20_000_000]
    svc = head[head["type"] != ge.EV_TASK]
    ids, first, cnt = np.unique(svc["svc_id"], return_index=True, return_counts=True)
    order = np.argsort(-cnt)
    rng = np.random.default_rng(0)
    pick = np.concatenate([order[:200], order[2000:2300], rng.choice(order[5000:], 1000, replace=False)])
    ntd = 0
    for j in pick:
        id_ = int(ids[j])
        o = orcs[int(svc["host_idx"][first[j]]) % nthr]
        a, b = eng.export_hist(id_, ge.HIST_RESP_CUR), o.export_hist(id_, ge.HIST_RESP_CUR)
        assert (a is None) == (b is None)
        if a is not None:
            assert np.array_equal(a[0], b[0]) and a[1:] == b[1:], hex(id_)
        assert np.array_equal(eng.export_hll(id_), o.export_hll(id_)), hex(id_)
        g, w = eng.export_conn_bitmap(id_), o.export_conn_bitmap(id_)
        assert np.array_equal(g[0], w[0])
        td = o.export_tdigest(id_)
        means, weights, mn, mx = eng.export_tdigest(id_)
        om, ow = td.centroids()
        assert np.array_equal(weights, ow) and np.array_equal(means, om), hex(id_)
        if len(ow):
            assert mn == td.minv and mx == td.maxv
            ntd += 1
    assert ntd > 1000
    tasks = head[head["type"] == ge.EV_TASK]
    tids, tfirst = np.unique(tasks["svc_id"], return_index=True)
    # the synthetic stream spreads a task's samples over hosts: its histograms are the sum over the oracle shards
    for j in rng.choice(len(tids), 300, replace=False):
        for which in (ge.HIST_TASK_CPU_PCT, ge.HIST_TASK_CPU_DELAY, ge.HIST_TASK_BLKIO_DELAY):
            a = eng.export_hist(int(tids[j]), which)
            cnt_, sum_, tot_, max_ = np.zeros(15, dtype=np.uint64), np.zeros(15, dtype=np.int64), 0, -(1 << 31)
            for o in orcs:
                b = o.export_hist(int(tids[j]), which)
                if b is not None:
                    cnt_ += b[0]["count"]; sum_ += b[0]["sum"]; tot_ += b[1]; max_ = max(max_, b[2])
            assert np.array_equal(a[0]["count"], cnt_) and np.array_equal(a[0]["sum"], sum_) and a[1] == tot_ and a[2] == max_
