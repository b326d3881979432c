"""drop-in boundary: whole wire messages (COMM_HEADER + EVENT_NOTIFY + variable-stride records, common/gy_comm_proto.h) through
gysk_ingest_msg, the validators' accept / reject behaviour (common/gy_comm_proto.cc:840-996), and the raw eBPF record kinds."""
import numpy as np
import pytest

from gyeeta_b200 import engine as ge
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

from gyeeta_b200.wire import HDR, IP_PORT, TCP_CONN, TASK, PM_MAGIC, COMM_EVENT_NOTIFY, build_msg


def test_tcp_conn_and_task_messages_roundtrip():
    rng = np.random.default_rng(12)
    eng = ge.Engine(max_svcs=256, max_tasks=64, max_batch=4096, cms_log2_width=12)
    orc = po.OracleEngine(max_svcs=256, max_tasks=64, cms_log2_width=12)
    exp = []
    recs = []
    for i in range(300):
        r = np.zeros(1, dtype=TCP_CONN)
        r["ser_glob_id"] = 1000 + int(rng.integers(0, 20))
        r["cli_task_aggr_id"] = 5000 + int(rng.integers(0, 50)) if i % 17 else 0        # some records fail the :9143 guard
        acc = bool(rng.integers(0, 2))
        r["is_accept"], r["is_connect"] = acc, not acc
        closed = bool(rng.integers(0, 2))
        r["tusec_close"] = 7_000_000 if closed else 0
        r["tusec_start"] = 3_000_000
        r["bytes_sent"], r["bytes_rcvd"] = int(rng.integers(0, 1 << 20)), int(rng.integers(0, 1 << 33))
        tail = bytes(rng.integers(65, 90, int(rng.integers(0, 40)), dtype=np.uint8))
        recs.append((r, tail))
        if r["cli_task_aggr_id"][0]:
            e = np.zeros(1, dtype=ge.EVENT_DTYPE)
            e["svc_id"], e["flow_key"] = r["ser_glob_id"], r["cli_task_aggr_id"]
            e["value"] = min(int(r["bytes_sent"][0]) + int(r["bytes_rcvd"][0]), 0xFFFFFFFF) if closed else 0
            e["type"] = (4 if closed else 2) if acc else (3 if closed else 1)
            exp.append(e)
    msg = build_msg(ge.NOTIFY_TCP_CONN, recs)
    assert eng.ingest_msg(msg, host_idx=3) == 0
    trecs = []
    for i in range(100):
        t = np.zeros(1, dtype=TASK)
        t["aggr_task_id"] = 9000 + i % 13
        t["total_cpu_pct"] = float(rng.random() * 500)
        t["cpu_delay_msec"], t["blkio_delay_msec"] = int(rng.integers(0, 70000)), int(rng.integers(0, 1 << 32))
        trecs.append((t, b"issue text" if i % 3 == 0 else b""))
        e = np.zeros(1, dtype=ge.EVENT_DTYPE)
        e["svc_id"] = t["aggr_task_id"]
        e["value"] = np.uint32(int(t["total_cpu_pct"][0]))
        e["flow_key"] = int(t["cpu_delay_msec"][0]) | (int(t["blkio_delay_msec"][0]) << 32)
        e["type"] = 6
        exp.append(e)
    assert eng.ingest_msg(build_msg(ge.NOTIFY_AGGR_TASK_STATE, trecs), host_idx=3) == 0
    eng.sync()
    orc.ingest(np.concatenate(exp))
    assert np.array_equal(eng.export_cms(), orc.cms())
    for sid in range(1000, 1020):
        assert np.array_equal(eng.export_hll(sid), orc.export_hll(sid))
    for tid in range(9000, 9013):
        for which in (3, 4, 5):
            a, b = eng.export_hist(tid, which), orc.export_hist(tid, which)
            assert np.array_equal(a[0], b[0]) and a[1:] == b[1:]
    st = eng.stats()
    assert st["wire_msgs_ok"] == 2 and st["events_tcp"] == orc.counters()["tcp"] and st["events_task"] == 100


def test_validators_reject_bad_messages():
    eng = ge.Engine(max_svcs=64, max_tasks=16, max_batch=2048, cms_log2_width=10)
    r = np.zeros(1, dtype=TCP_CONN)
    r["ser_glob_id"], r["cli_task_aggr_id"], r["is_accept"] = 1, 2, 1
    good = build_msg(ge.NOTIFY_TCP_CONN, [(r, b"abc")] * 3)
    assert eng.ingest_msg(bytearray(good)) == 0
    bad = bytearray(good); np.frombuffer(bad, dtype=HDR, count=1)["magic"] = 0x1234                 # wrong magic
    assert eng.ingest_msg(bad) == -22
    bad = bytearray(good); np.frombuffer(bad, dtype=HDR, count=1)["nevents"] = 4                    # claims one record too many
    assert eng.ingest_msg(bad) == -22
    bad = bytearray(good); np.frombuffer(bad, dtype=HDR, count=1)["nevents"] = 2049                 # > MAX_NUM_CONNS
    assert eng.ingest_msg(bad) == -22
    bad = bytearray(good); bad[24 + 279] = 3                                                      # padding_len breaks 8-byte alignment
    assert eng.ingest_msg(bad) == -22
    bad = bytearray(good); np.frombuffer(bad, dtype=HDR, count=1)["total_sz"] = len(good) + 64      # longer than the buffer
    assert eng.ingest_msg(bad) == -22
    unk = bytearray(good); np.frombuffer(unk, dtype=HDR, count=1)["subtype"] = 0x301                # not on the hot path
    assert eng.ingest_msg(unk) == -95
    eng.sync()
    st = eng.stats()
    assert st["wire_msgs_ok"] == 1 and st["wire_msgs_bad"] == 5 and st["events_tcp"] == 3
    # the validator NUL-forces the trailing string in place like the reference (:866)
    m = bytearray(good)
    eng.ingest_msg(m)
    assert m[24 + 280 + 2] == 0


def test_raw_ebpf_records():
    eng = ge.Engine(max_svcs=64, max_tasks=16, max_batch=2048, cms_log2_width=10)
    resp = np.zeros(1000, dtype=np.dtype([("saddr", "<u4"), ("daddr", "<u4"), ("netns", "<u4"), ("sport", "<u2"), ("dport", "<u2"),
                                          ("lsndtime", "<u4"), ("lrcvtime", "<u4")]))
    rng = np.random.default_rng(3)
    # inet_sport / skc_dport arrive in NETWORK byte order (the reference applies ntohs, gy_socket_stat.cc:1526-1527)
    cport = rng.integers(16000, 60000, len(resp)).astype(np.uint16)
    resp["saddr"], resp["netns"], resp["sport"] = 0x0A000001, 4026531840, np.uint16(8080).byteswap()
    resp["daddr"] = rng.integers(1, 1 << 31, len(resp)); resp["dport"] = cport.byteswap()
    resp["lrcvtime"] = rng.integers(0, 1 << 31, len(resp))
    ms = rng.integers(0, 20000, len(resp)).astype(np.uint32)
    ms[::50] = 2_000_000                                   # beyond the 1 000 000 msec validity rule: dropped on the host
    resp["lsndtime"] = resp["lrcvtime"] + ms
    eng.ingest_raw(ge.RAW_TCP_IPV4_RESP, resp, len(resp)); eng.sync()
    st = eng.stats()
    kept = int((ms <= 1_000_000).sum())
    assert st["events_resp"] == kept and st["nsvcs"] == 1
    # reference histogram over the same msec values
    want = po.hist_run(po.lib(), "gyo_hist_run", 0, 0, ms[ms <= 1_000_000].astype(np.int64))
    L = eng.L
    import ctypes as C
    # the service id is derived on the host; find it through the only registered listener: query via a second raw batch
    conn = np.zeros(10, dtype=np.dtype([("ts_ns", "<u8"), ("bytes_received", "<u8"), ("bytes_acked", "<u8"), ("pid", "<u4"), ("tid", "<u4"),
                                        ("comm", "S16"), ("saddr", "<u4"), ("daddr", "<u4"), ("netns", "<u4"), ("sport", "<u2"),
                                        ("dport", "<u2"), ("ipver", "u1"), ("type", "u1")], align=True))
    assert conn.dtype.itemsize == 72
    conn["saddr"], conn["netns"], conn["sport"], conn["type"], conn["bytes_acked"] = 0x0A000001, 4026531840, np.uint16(8080).byteswap(), 4, 10240
    conn["daddr"] = np.arange(10) + 100
    eng.ingest_raw(ge.RAW_TCP_IPV4_EVENT, conn, len(conn)); eng.sync()
    st = eng.stats()
    assert st["nsvcs"] == 1 and st["events_tcp"] == 10      # same (netns, ip, port) -> same service id as the resp events
    assert int(eng.export_cms().sum() & 0xFFFFFFFF) == 40   # 10 events x 4 rows, count halves
    assert want["total"] == kept
    # CONN_BITMAP index = HOST-order client port & 0x1F (gy_socket_stat.h:403-410): per response bucket the set of port slots seen
    JL = po.lib()
    seed = 0xceedfead
    sid = (JL.gyo_jhash_2words(0x0A000001, 4026531840, seed) << 32) | JL.gyo_jhash_2words(8080, 4026531840, seed ^ 0x0A000001)
    masks, cnts = eng.export_conn_bitmap(sid)
    ok = ms <= 1_000_000
    want_masks = np.zeros(15, dtype=np.uint32)
    for m_, p_ in zip(ms[ok], cport[ok]):
        want_masks[JL.gyo_bucket(0, int(m_))] |= np.uint32(1 << (int(p_) & 31))
    assert np.array_equal(masks, want_masks)
    assert [bin(int(x)).count("1") for x in want_masks] == list(cnts)
    h = eng.export_hist(sid, ge.HIST_RESP_CUR)
    assert np.array_equal(h[0]["count"], want["stats"]["count"][:15]) and h[1] == kept


def test_listener_state_host_summary():
    """NOTIFY_LISTENER_STATE: validated like LISTENER_STATE_NOTIFY::validate and rolled up per host exactly as
    LISTEN_SUMM_STATS::update (server/gy_msocket.h:854-866) does inside partha_listener_state (gy_mconnhdlr.cc:11251)."""
    LSN = np.dtype([("glob_id", "<u8"), ("nqrys_5s", "<u4"), ("total_resp_5sec", "<u4"), ("nconns", "<u4"), ("nconns_active", "<u4"),
                    ("ntasks", "<u4"), ("p95_5s", "<u4"), ("p95_5min", "<u4"), ("kb_in", "<u4"), ("kb_out", "<u4"), ("ser_errors", "<u4"),
                    ("cli_errors", "<u4"), ("t1", "<u4"), ("t2", "<u4"), ("t3", "<u4"), ("t4", "<u4"), ("t5", "<u4"), ("t6", "<u4"),
                    ("ntasks_issue", "<u2"), ("is_http", "u1"), ("curr_state", "u1"), ("curr_issue", "u1"), ("issue_bit_hist", "u1"),
                    ("high_resp_bit_hist", "u1"), ("last_issue_subsrc", "u1"), ("query_flags", "u1"), ("issue_string_len", "u1"),
                    ("padding_len", "u1"), ("pad", "u1")])
    assert LSN.itemsize == 88
    rng = np.random.default_rng(4)
    eng = ge.Engine(max_svcs=64, max_tasks=16, max_batch=2048, cms_log2_width=10)
    recs, want = [], dict(nstates=[0] * 8, tot_qps=0, tot_act_conn=0, tot_kb_inbound=0, tot_kb_outbound=0, tot_ser_errors=0, nlisteners=0, nactive=0)
    for i in range(200):
        r = np.zeros(1, dtype=LSN)
        r["glob_id"] = 100 + i
        r["nqrys_5s"] = int(rng.integers(0, 5000)) if i % 4 else 0
        r["nconns_active"], r["kb_in"], r["kb_out"], r["ser_errors"] = rng.integers(0, 1000, 4)
        r["curr_state"] = int(rng.integers(0, 8))
        r["query_flags"] = 0xC0 if i % 11 == 0 else (1 if i % 13 == 0 else 0)      # LISTEN_FLAG_DELETE / LISTEN_FLAG_ALERT / none
        recs.append((r, b"some issue" if i % 5 == 0 else b""))
        # partha_listener_state (gy_mconnhdlr.cc:11183-11251): DELETE records and states beyond STATE_DOWN never reach summstats.update
        if int(r["query_flags"][0]) == 0xC0 or int(r["curr_state"][0]) > 5:
            continue
        want["nstates"][int(r["curr_state"][0])] += 1
        want["tot_qps"] += int(r["nqrys_5s"][0]) // 5
        want["tot_act_conn"] += int(r["nconns_active"][0]); want["tot_kb_inbound"] += int(r["kb_in"][0])
        want["tot_kb_outbound"] += int(r["kb_out"][0]); want["tot_ser_errors"] += int(r["ser_errors"][0])
        want["nlisteners"] += 1; want["nactive"] += int(r["nqrys_5s"][0] != 0)
    assert eng.ingest_msg(build_msg(ge.NOTIFY_LISTENER_STATE, recs), host_idx=9) == 0
    assert eng.host_summary(9) == want
    assert eng.host_summary(10) is None
    # cluster roll-up of the host summaries (MS_CLUSTER_STATE::STATE_ONE service fields, gy_mconnhdlr.cc:16036-16046)
    assert eng.ingest_msg(build_msg(ge.NOTIFY_LISTENER_STATE, recs[:50]), host_idx=11) == 0
    h11 = eng.host_summary(11)
    issues = lambda h: h["nstates"][3] + h["nstates"][4] + h["nstates"][5]
    cs = eng.cluster_state()
    assert cs == dict(nhosts=2, nsvc_issue=issues(want) + issues(h11), nsvcissue_hosts=int(issues(want) > 0) + int(issues(h11) > 0),
                      nsvc=want["nlisteners"] + h11["nlisteners"], total_qps=want["tot_qps"] + h11["tot_qps"],
                      svc_net_mb=(want["tot_kb_inbound"] + want["tot_kb_outbound"]) // 1024 + (h11["tot_kb_inbound"] + h11["tot_kb_outbound"]) // 1024)
    assert eng.cluster_state([9, 77])["nhosts"] == 1


def test_listener_state_encoder_roundtrip():
    """summary encoder: the engine's per-service rows leave as LISTENER_STATE_NOTIFY records; fed back as a
    NOTIFY_LISTENER_STATE message they pass LISTENER_STATE_NOTIFY::validate and roll up (LISTEN_SUMM_STATS::update) to the
    totals of the oracle's last window"""
    from gyeeta_b200 import synth
    LSN = np.dtype([("glob_id", "<u8"), ("nqrys_5s", "<u4"), ("total_resp_5sec", "<u4"), ("nconns", "<u4"), ("nconns_active", "<u4"),
                    ("ntasks", "<u4"), ("p95_5s", "<u4"), ("p95_5min", "<u4"), ("kb_in", "<u4"), ("kb_out", "<u4"), ("rest", "u1", 44)])
    assert LSN.itemsize == 88
    rng = np.random.default_rng(8)
    eng = ge.Engine(max_svcs=512, max_tasks=16, max_batch=1 << 15, cms_log2_width=10)
    orc = po.OracleEngine(max_svcs=512, max_tasks=16, cms_log2_width=10)
    ev = synth.gen_mixed(rng, 30_000, 150, ntask=4, nhosts=2, nclients=500)
    eng.ingest_events(ev); orc.ingest(ev)
    eng.flush(5); orc.flush(5)
    ids = np.unique(ev["svc_id"][ev["type"] != ge.EV_TASK])
    ask = np.concatenate([ids, np.array([0xDEAD], dtype=np.uint64)])           # one unknown id: skipped by the encoder
    nrecs, raw = eng.listener_state_records(ask)
    assert nrecs == len(ids) and len(raw) == 88 * nrecs
    recs = np.frombuffer(raw, dtype=LSN)
    tot_qps = nact = 0
    for r in recs:
        h = orc.export_hist(int(r["glob_id"]), 1)
        c = orc.export_conn(int(r["glob_id"]))
        assert r["nqrys_5s"] == h[1] and r["total_resp_5sec"] == int(h[0]["sum"].sum())
        assert r["nconns"] == (c[1] & 0xFFFFFFFF) and r["kb_in"] == (c[1] >> 32)
        tot_qps += int(h[1]) // 5; nact += int(h[1] != 0)
    hdr = np.zeros(1, dtype=HDR)
    hdr["magic"], hdr["data_type"] = PM_MAGIC, COMM_EVENT_NOTIFY
    hdr["total_sz"] = HDR.itemsize + len(raw)
    hdr["subtype"], hdr["nevents"] = ge.NOTIFY_LISTENER_STATE, nrecs
    assert eng.ingest_msg(bytearray(hdr.tobytes() + raw), host_idx=3) == 0
    hs = eng.host_summary(3)
    assert hs["nlisteners"] == nrecs and hs["tot_qps"] == tot_qps and hs["nactive"] == nact
