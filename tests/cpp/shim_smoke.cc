// Compiles the host-side mirror (gyeeta_b200/host/gy_gysk_shim.h) the way gy_mconnhdlr.cc would use it, with stand-in
// PARTHA_INFO / pool types, links libgysketch.so and drives the three handlers. Without a GPU gysk_create() must fail with
// GYSK_ERR_NODEV (no CPU fallback) and the handlers must return false without touching the records.
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "gy_gysk_shim.h"
#include "gysk_wire.h"

struct PARTHA_INFO { uint64_t machine_id_[2] {1, 2}; uint32_t gysk_host_idx_ {7}; };
struct POOL_ALLOC_ARRAY {};
struct PGConnPool {};

int main()
{
	using namespace gysk::wire;

	gysk_config cfg;
	gysk_config_default(&cfg);
	gysk_engine *e = nullptr;
	int rc = gysk_create(&cfg, &e);
	std::printf("gysk_create rc=%d (%s)\n", rc, rc ? gysk_last_error(nullptr) : "ok");

	std::vector<uint8_t> buf(sizeof(TCP_CONN_NOTIFY) * 2 + 16);
	auto *p = reinterpret_cast<TCP_CONN_NOTIFY *>(buf.data());
	std::memset(p, 0, buf.size());
	p[0].ser_glob_id_ = 11; p[0].cli_task_aggr_id_ = 22; p[0].is_tcp_accept_event_ = true; p[0].tusec_close_ = 5; p[0].bytes_sent_ = 4096;

	gysk_shim::GYSK_HANDLER h(e);
	auto partha = std::make_shared<PARTHA_INFO>();
	POOL_ALLOC_ARRAY pool; PGConnPool db;

	bool ok1 = h.partha_tcp_conn_info(partha, p, 1, buf.data() + sizeof(TCP_CONN_NOTIFY), &pool);
	AGGR_TASK_STATE_NOTIFY t {}; t.aggr_task_id_ = 5; t.total_cpu_pct_ = 12.5f;
	bool ok2 = h.partha_aggr_task_state(partha, &t, 1, reinterpret_cast<uint8_t *>(&t + 1), db);
	LISTENER_STATE_NOTIFY l {}; l.glob_id_ = 11;
	bool ok3 = h.partha_listener_state(partha, &l, 1, reinterpret_cast<uint8_t *>(&l + 1), &pool, db);
	ACTIVE_CONN_STATS a {}; a.listener_glob_id_ = 11; a.cli_aggr_task_id_ = 22; a.bytes_sent_ = 10240; a.active_conns_ = 3; a.max_rtt_msec_ = 1.5f;
	bool ok4 = h.handle_partha_active_conns(partha, &a, 1, reinterpret_cast<uint8_t *>(&a + 1), &pool, db);
	tcp_ipv4_resp_event_t r4 {}; r4.saddr = 0x0A000001; r4.sport = 0x5000 /* htons(80) */; r4.netns = 1; r4.lsndtime = 40; r4.lrcvtime = 10;
	bool ok5 = h.handle_resp_events(partha, &r4, 1);
	tcp_ipv6_event_t c6 {}; c6.type = 2; c6.netns = 1; c6.saddr[3] = 1; c6.sport = 0x5000;
	bool ok6 = h.handle_conn_events(partha, &c6, 1);
	std::printf("handlers: %d %d %d\n", ok1, ok2, ok3);
	std::printf("more handlers: %d %d %d\n", ok4, ok5, ok6);

	if (e) {
		gysk_svc_summary s;
		uint64_t id = 11;
		gysk_flush(e, 5);
		rc = gysk_query_svcs(e, &id, 1, &s);
		std::printf("query rc=%d found=%d nconns_5s=%u kbytes_5s=%u\n", rc, s.found, s.nconns_5s, s.kbytes_5s);
		std::printf("active: nconns_active=%u active_kbytes=%u max_rtt=%.1f\n", s.nconns_active, s.active_kbytes, s.max_rtt_msec);
		gysk_stats st;
		gysk_get_stats(e, &st);
		std::printf("stats: resp=%llu tcp=%llu svcs=%llu\n", (unsigned long long)st.events_resp, (unsigned long long)st.events_tcp, (unsigned long long)st.nsvcs);
		bool good = ok1 && ok2 && ok3 && ok4 && ok5 && ok6 && rc == 0 && s.found == 1 && s.nconns_5s == 1 && s.kbytes_5s == 4 &&
				s.nconns_active == 3 && s.active_kbytes == 10 && s.max_rtt_msec == 1.5f && st.events_resp == 1 && st.events_tcp == 3 && st.nsvcs == 3;
		gysk_destroy(e);
		return good ? 0 : 2;
	}
	// no engine: the tick helpers and the record encoder report failure instead of touching anything
	int ndel = 0;
	const bool tick = h.flush_window(5, [&](uint64_t) { ++ndel; });
	uint64_t ids[2] = {11, 12};
	uint8_t recs[2 * sizeof(LISTENER_STATE_NOTIFY)];
	const int nrec = h.listener_state_records(ids, 2, recs, sizeof(recs));
	std::printf("tick: %d deletes: %d records: %d\n", tick, ndel, nrec);
	return (rc == GYSK_ERR_NODEV && !ok1 && !ok2 && !ok3 && !ok4 && !ok5 && !ok6 && !tick && ndel == 0 && nrec == -1) ? 0 : 1;
}
