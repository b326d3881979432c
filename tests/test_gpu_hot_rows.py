"""Hot services take their response samples as direct updates of a dense, L2-resident row of value bins instead of sort keys
(DESIGN.md §4). That is routing only: every number must be the one the oracle (and the key path) produces. The parity tests are
re-run with the switch-over forced early (8 samples in a batch), with almost no rows (the overflow stays on the key path), and
with the path off."""
import numpy as np
import pytest

from gyeeta_b200 import engine as ge
from gyeeta_b200 import synth
from tests import test_gpu_parity as tp
from tests.util import assert_hist_equal, feed_both, make_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,hmin", [(2048, 8), (5, 8), (0, 4096)])
def test_hot_rows_do_not_change_any_number(monkeypatch, rows, hmin):
    monkeypatch.setenv("GYSK_HOT_ROWS", str(rows))
    monkeypatch.setenv("GYSK_HOT_MIN", str(hmin))
    tp.test_mixed_stream_bit_exact(3000, 300_000, 1 << 15)
    tp.test_tdigest_many_services_skewed()
    tp.test_idle_service_eviction_and_slot_reuse()
    tp.test_window_membership_is_by_arrival()
    tp.test_full_value_range_keys()


@pytest.mark.parametrize("rows,hmin,want", [(2048, 8, None), (5, 8, 5), (0, 8, 0), (2048, 1 << 26, 0)])
def test_hot_rows_are_taken_and_every_batch_is_bit_exact(monkeypatch, rows, hmin, want):
    """eight batches; after each one histograms and centroids of hot and cold services equal the oracle's"""
    monkeypatch.setenv("GYSK_HOT_ROWS", str(rows))
    monkeypatch.setenv("GYSK_HOT_MIN", str(hmin))
    rng = np.random.default_rng(77)
    nsvc = 400
    eng, orc = make_pair(max_svcs=512, max_tasks=64, max_batch=1 << 16, cms_log2_width=12)
    ids_seen = set()
    for b in range(8):
        ev = synth.gen_mixed(rng, 60_000, nsvc, ntask=8, nhosts=8, nclients=2000, zipf_s=1.05)
        if b == 5:
            ev = ev[ev["type"] != ge.EV_RESP]                 # a batch without a single response sample: hot rows stay empty
        feed_both(eng, orc, ev, 1 << 16)
        resp = ev[ev["type"] == ge.EV_RESP]
        ids, counts = np.unique(resp["svc_id"], return_counts=True)
        ids_seen |= set(int(i) for i in ids)
        order = np.argsort(-counts) if len(ids) else []
        pick = [int(ids[j]) for j in list(order[:12]) + list(order[-12:])] + sorted(ids_seen)[:6]
        for id_ in pick:
            assert_hist_equal(eng, orc, id_, ge.HIST_RESP_CUR)
            (means, weights, mn, mx), td = eng.export_tdigest(id_), orc.export_tdigest(id_)
            om, ow = td.centroids()
            assert np.array_equal(weights, ow) and np.array_equal(means, om) and mn == td.minv and mx == td.maxv, (b, hex(id_))
        if b == 3:
            eng.flush(5); orc.flush(5)
    got = eng.hot_rows_in_use()
    if want is None:
        assert 20 <= got <= nsvc, got                        # the head of the Zipf distribution turned hot
    else:
        assert got == want, got
    s, o = eng.stats(), orc.counters()
    assert s["events_resp"] == o["resp"] and s["events_dropped"] == o["dropped"]
