/*
 * gysketch.h — C ABI of libgysketch.so, the B200-native streaming-sketch aggregation engine that sits
 * behind Gyeeta's madhava ingest path.
 *
 * Every entry point replaces (or is what a cgo/ctypes/C++ shim would bind for) a named piece of the
 * reference; citations are relative to the reference tree:
 *
 *   gysk_ingest()         the per-message dispatch of MCONN_HANDLER::handle_l2_misc  server/gy_mconnhdlr.cc:4745-4800
 *                         (phdr/pevtnot/recs/nevents/pendptr arithmetic) and the record walks of
 *                         partha_tcp_conn_info :9052/:9130, partha_listener_state :10993/:11175,
 *                         partha_aggr_task_state :9959, incl. the L1 validators common/gy_comm_proto.cc:840-996
 *   gysk_ingest_raw()     the per-sample reduction lifted off partha: TCP_SOCK_HANDLER::handle_ipv4_resp_event /
 *                         handle_tcp_resp_event  common/gy_socket_stat.cc:1517-1677, handle_ipv4_conn_event :241,
 *                         SVC_INFO_CAP::upd_stats_on_req  common/gy_proto_parser.cc:2678-2694
 *   gysk_ingest_device()  same reduction for event batches already resident in HBM (bench / device generators)
 *   gysk_flush()          the 5-second reducer TCP_SOCK_HANDLER::listener_stats_update  common/gy_socket_stat.cc:3898-4445
 *   gysk_query_svcs()     readers of MTCP_LISTENER::state_ (server/gy_msocket.h:1304) through SvcStateFields
 *                         (server/gy_mfields.h:1383-1412): qps5s, nqry5s, resp5s, p95resp5s ...
 *   gysk_export_hist()    GY_HISTOGRAM::get_serialized  common/gy_statistics.h:656-673 (HIST_SERIAL byte-compatible)
 *   gysk_hist_percentiles GY_HISTOGRAM::get_percentiles common/gy_statistics.h:707-791
 *   gysk_query_flows()    new capability: count-min point query replacing the exact two-level group-by of
 *                         partha_tcp_conn_info  server/gy_mconnhdlr.cc:9245-9311
 *   gysk_export_hll()     new capability: distinct clients per service, replacing the exact client sets
 *                         (cli_aggr_task_tbl_, server/gy_msocket.h:1335)
 *   gysk_export_tdigest() new capability: response-time quantile sketch; the reference's only t-digest user is the
 *   gysk_query_quantiles  Postgres extension with compression 100 (common/gy_query_common.cc:1805-1858)
 *   gysk_merge_*          additive roll-up  MS_CLUSTER_STATE::STATE_ONE::add_stats common/gy_comm_proto.h:3199-3214 /
 *                         SHCONN_HANDLER::aggregate_cluster_state server/gy_shconnhdlr.cc:4583
 *
 * Conventions: plain pointers and sizes, no exceptions cross the boundary, 0 = ok, negative = -errno style
 * (mirrors the reference handlers' bool/int returns wrapped in GY_CATCH_EXCEPTION, gy_mconnhdlr.cc:4763-4774).
 * CUDA errors are sticky: once a call fails with GYSK_ERR_CUDA every later call fails too; gysk_last_error()
 * returns the text. There is NO CPU fallback: gysk_create() fails when no sm_100 device is usable.
 */
#ifndef GYSKETCH_H
#define GYSKETCH_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GYSK_ABI_VERSION		2

/* ---- error codes ---- */
#define GYSK_OK				0
#define GYSK_ERR_INVAL			(-22)	/* EINVAL  : bad argument / failed wire validation */
#define GYSK_ERR_NOMEM			(-12)	/* ENOMEM  */
#define GYSK_ERR_NOENT			(-2)	/* ENOENT  : unknown service / task id */
#define GYSK_ERR_NOSPC			(-28)	/* ENOSPC  : service table full */
#define GYSK_ERR_NODEV			(-19)	/* ENODEV  : no usable CUDA device (no CPU fallback exists) */
#define GYSK_ERR_CUDA			(-5)	/* EIO     : CUDA runtime error (sticky) */
#define GYSK_ERR_NOTSUP			(-95)	/* EOPNOTSUPP */

/* ---- canonical 32-byte event record ---- */
enum {
	GYSK_EV_CONNECT		= 1,	/* TCP_EVENT_TYPE_CONNECT    common/gy_ebpf_kernel.h:24 */
	GYSK_EV_ACCEPT		= 2,	/* TCP_EVENT_TYPE_ACCEPT */
	GYSK_EV_CLOSE_CLI	= 3,	/* TCP_EVENT_TYPE_CLOSE_CLI */
	GYSK_EV_CLOSE_SER	= 4,	/* TCP_EVENT_TYPE_CLOSE_SER */
	GYSK_EV_RESP		= 5,	/* service response-time sample (tcp_ipv4_resp_event_t / API_TRAN) */
	GYSK_EV_TASK		= 6,	/* per-process 5-s sample (AGGR_TASK_STATE_NOTIFY) */
	GYSK_EV_ACTIVE		= 7,	/* one ACTIVE_CONN_STATS record (common/gy_comm_proto.h:2766): the 15-s inet_diag group-by
					   {ser_glob_id, cli_task_aggr_id} of upd_conn_from_diag (common/gy_socket_stat.cc:6156-6194) */
};

/* gysk_event.flags of a GYSK_EV_RESP event that came from an API_TRAN (SVC_INFO_CAP::upd_stats_on_req, gy_proto_parser.cc:2678-2694) */
#define GYSK_EVF_CLI_ERROR		0x1u	/* stats_.ncli_errors_++ */
#define GYSK_EVF_SER_ERROR		0x2u	/* stats_.nser_errors_++ */

typedef struct gysk_event
{
	uint64_t	svc_id;		/* ser_glob_id_ ; for GYSK_EV_TASK: aggr_task_id_. 0 is invalid (dropped) */
	uint64_t	flow_key;	/* TCP/RESP: cli_task_aggr_id_ or a 64-bit fold of the 5-tuple.
					   TASK: low 32 = cpu_delay_msec_, high 32 = blkio_delay_msec_ */
	uint32_t	value;		/* RESP: response time in usec; TCP: bytes; TASK: (int)total_cpu_pct_; ACTIVE: kbytes sent + received */
	uint32_t	host_idx;	/* dense index of the sending partha (shard key: host_idx % world) */
	uint32_t	tsec;		/* event time, seconds — informational: a sample lands in the window that is open when it ARRIVES,
					   as in the reference (handle_tcp_resp_event stamps samples with time(nullptr) of their
					   processing, common/gy_socket_stat.cc:1560-1579). ACTIVE: IEEE-754 bits of max_rtt_msec_ */
	uint16_t	type;		/* GYSK_EV_* */
	uint16_t	flags;		/* RESP: GYSK_EVF_*; ACTIVE: active_conns_; else 0 */
} gysk_event;

/* ---- histogram classes (bucket thresholds of common/gy_statistics.h:1674-2063) ---- */
enum {
	GYSK_CLS_RESP_TIME	= 0,	/* RESP_TIME_HASH	:1674  (msec) */
	GYSK_CLS_SEMI_LOG	= 1,	/* SEMI_LOG_HASH	:1729 */
	GYSK_CLS_SEMI_LOG_LO	= 2,	/* SEMI_LOG_HASH_LO	:1782 */
	GYSK_CLS_DURATION	= 3,	/* DURATION_HASH	:1835 */
	GYSK_CLS_HASH_10_5000	= 4,	/* HASH_10_5000		:1908 */
	GYSK_CLS_HASH_5_250	= 5,	/* HASH_5_250		:1960 */
	GYSK_CLS_HASH_1_3000	= 6,	/* HASH_1_3000		:2013 */
	GYSK_CLS_PERCENT	= 7,	/* PERCENT_HASH		:1624 */
};

#define GYSK_HIST_MAX_BUCKETS		15	/* largest max_buckets among the classes above */

/* byte-compatible with HIST_SERIAL, common/gy_statistics.h:458-468 */
typedef struct gysk_hist_serial
{
	uint64_t	count;
	int64_t		sum;
} gysk_hist_serial;

/* which histogram of an id */
enum {
	GYSK_HIST_RESP_CUR	= 0,	/* service: response msec, window being filled		(RESP_TIME_HASH, T=int64) */
	GYSK_HIST_RESP_LAST	= 1,	/* service: last closed 5-s window */
	GYSK_HIST_RESP_ALL	= 2,	/* service: since start ("Since Process start" level, gy_statistics.h:1548) */
	GYSK_HIST_TASK_CPU_PCT	= 3,	/* task: MTASK_HIST::cpu_pct_histogram_		(HASH_1_3000, T=int)  server/gy_msocket.h:707 */
	GYSK_HIST_TASK_CPU_DELAY= 4,	/* task: cpu_delay_histogram_			(DURATION_HASH, T=int) */
	GYSK_HIST_TASK_BLKIO_DELAY = 5,	/* task: blkio_delay_histogram_			(DURATION_HASH, T=int) */
	GYSK_HIST_RESP_5MIN	= 6,	/* service: 300-s level    (Level_5s_5min_5days_all, gy_statistics.h:1545-1551; 10 slots, :1105) */
	GYSK_HIST_RESP_5DAY	= 7,	/* service: 432000-s level */
	GYSK_HIST_QPS		= 8,	/* service: TCP_LISTENER::qps_hist_		(SEMI_LOG_HASH_LO, T=int) common/gy_socket_stat.h:548,633: one
					   sample per closed window = queries / 5, common/gy_socket_stat.cc:4111-4121 */
	GYSK_HIST_ACTIVE_CONN	= 9,	/* service: active_conn_hist_			(HASH_1_3000, T=int) :549,635: one sample per closed window =
					   the listener's active connections, :4124-4130 */
};

/* ---- listener state (OBJ_STATE_E, common/gy_json_field_maps.h:242-250) and issue source (LISTENER_ISSUE_SRC, :419-435) ---- */
enum { GYSK_STATE_IDLE = 0, GYSK_STATE_GOOD = 1, GYSK_STATE_OK = 2, GYSK_STATE_BAD = 3, GYSK_STATE_SEVERE = 4, GYSK_STATE_DOWN = 5 };
enum { GYSK_ISSUE_NONE = 0, GYSK_ISSUE_LISTENER_TASKS = 1, GYSK_ISSUE_QPS_HIGH = 2, GYSK_ISSUE_ACTIVE_CONN_HIGH = 3, GYSK_ISSUE_SERVER_ERRORS = 4,
       GYSK_ISSUE_OS_CPU = 5, GYSK_ISSUE_OS_MEMORY = 6, GYSK_ISSUE_DEPENDENT_SERVER_LISTENER = 7, GYSK_ISSUE_SRC_UNKNOWN = 8 };

/* Inputs of TCP_LISTENER::get_curr_state (common/gy_socket_stat.cc:2020-2875), the once-per-window state decision of a listener.
 * gysk_flush() evaluates it on the device for every service from the engine's own state, with the block "outside the path" zero;
 * gysk_classify_listener() is the same code for a caller that has those inputs. Response values in msec. */
typedef struct gysk_listener_state_in
{
	int64_t		r5p95, r5p99;			/* last 5-s window: p95 / p99 bucket thresholds		(:2082-2083) */
	int64_t		r300p95, r300p99;		/* 300-s level						(:2084-2085) */
	int64_t		r5dp95, r5dp99, r5dp25;		/* 5-day level						(:2086-2088) */
	int64_t		rallp95, rallp99;		/* all-time level					(:2089-2090) */
	uint64_t	nqrys_5s;			/* histstat_[n5].tcount_ */
	uint64_t	total_resp_msec;		/* histstat_[n5].tsum_ */
	uint64_t	tcount_5d;			/* histstat_[n5days].tcount_ */
	double		mean5, mean300, mean5d, meanall;	/* tsum / max(tcount, 1), common/gy_statistics.h:1358 */
	int64_t		qps_p95, qps_p25;		/* qps_hist_ percentiles				(:2097) */
	int64_t		act_p95, act_p25;		/* active_conn_hist_ percentiles			(:2098) */
	int64_t		secs_5d;			/* seconds the 5-day level covers so far: min(432000, age)	(:2067-2074) */
	int32_t		last_qps_count;			/* last_qps_count_: queries per second of the window	(:4121) */
	int32_t		nconn;				/* last_chk_nconn_					(:2036) */
	int32_t		curr_active_conn;		/* (:4158-4170) max of nconn_recent_active_ and the CONN_BITMAP bucket counts */
	uint32_t	ser_errors;
	uint8_t		nactive_conn_arr[16];		/* CONN_BITMAP::get_conn_breakup of the window, per response bucket (:4160-4163) */
	/* outside the path: the engine passes zeros */
	uint8_t		task_issue, task_severe, task_delay;	/* TCP_LISTENER::is_task_issue (:1914) verdicts */
	uint8_t		cpu_issue, mem_issue;		/* host state */
	uint8_t		pad0[3];
	int32_t		ntasks_issue, ntasks_noissue;
	uint64_t	tasks_delay_msec;
	uint32_t	nserdepends;			/* related_listen_->id_depends_ count (:2820-2823) */
	uint32_t	pad1;
} gysk_listener_state_in;

/* ---- raw record kinds for gysk_ingest_raw ---- */
enum {
	GYSK_RAW_EVENT32	= 0,	/* gysk_event[] */
	GYSK_RAW_TCP_IPV4_EVENT	= 1,	/* tcp_ipv4_event_t[]      72 B  common/gy_ebpf_kernel.h:37  */
	GYSK_RAW_TCP_IPV4_RESP	= 2,	/* tcp_ipv4_resp_event_t[] 24 B  common/gy_ebpf_kernel.h:106 */
	GYSK_RAW_TCP_IPV6_EVENT	= 3,	/* tcp_ipv6_event_t[]      96 B  common/gy_ebpf_kernel.h:54  (handle_ipv6_conn_event, gy_socket_stat.cc:269) */
	GYSK_RAW_TCP_IPV6_RESP	= 4,	/* tcp_ipv6_resp_event_t[] 64 B  common/gy_ebpf_kernel.h:113 (handle_ipv6_resp_event, gy_socket_stat.cc:1535) */
	GYSK_RAW_API_TRAN	= 5,	/* API_TRAN records, variable stride (common/gy_proto_common.h:140-204; nevents records walked with
					   get_elem_size()); `events` must stay readable for nevents strides */
	GYSK_RAW_RESP16		= 6,	/* gysk_resp16[]  packed response samples */
	GYSK_RAW_TCP24		= 7,	/* gysk_tcp24[]   packed conn events */
	GYSK_RAW_TASK24		= 8,	/* gysk_task24[]  packed process samples */
};

/* Packed per-kind records: what a feeder that already knows the kind of a batch ships instead of the 32-byte canonical record
 * (18.4 bytes per event on the 70 / 20 / 10 mix instead of 32 — the host link is the end-to-end limit). Decoded ON THE DEVICE
 * into the canonical record, as are the fixed-stride eBPF structs above: gysk_ingest_raw copies the raw bytes, a kernel expands them. */
typedef struct gysk_resp16 { uint64_t svc_id; uint32_t usec; uint16_t host_idx; uint8_t cli_port; uint8_t flags; } gysk_resp16;
typedef struct gysk_tcp24 { uint64_t svc_id; uint64_t flow_key; uint32_t bytes; uint16_t host_idx; uint8_t type; uint8_t pad; } gysk_tcp24;
typedef struct gysk_task24 { uint64_t aggr_task_id; uint32_t cpu_pct; uint32_t cpu_delay_msec; uint32_t blkio_delay_msec; uint16_t host_idx; uint16_t pad; } gysk_task24;

/* ---- wire subtypes accepted by gysk_ingest (NOTIFY_TYPE_E, common/gy_comm_proto.h:155-200) ---- */
#define GYSK_NOTIFY_LISTENER_STATE	0x309u
#define GYSK_NOTIFY_TCP_CONN		0x30Cu
#define GYSK_NOTIFY_AGGR_TASK_STATE	0x310u
#define GYSK_NOTIFY_ACTIVE_CONN_STATS	0x312u	/* handle_partha_active_conns, server/gy_mconnhdlr.cc:7705 (dispatched at :5250) */

/* ---- configuration ---- */
#define GYSK_FLAG_AUTO_REGISTER		0x1u	/* unknown svc/task ids are inserted on first sight (device side);
						   without it unknown ids are skipped like a failed
						   listen_tbl_.lookup_single_elem_locked, gy_mconnhdlr.cc:11183 */

typedef struct gysk_config
{
	uint32_t	struct_size;		/* sizeof(gysk_config) */
	int32_t		device;			/* CUDA device ordinal */
	uint32_t	max_svcs;		/* service (listener) capacity */
	uint32_t	max_tasks;		/* aggregated-process capacity */
	uint32_t	cms_depth;		/* rows, 1..8 (default 4) */
	uint32_t	cms_log2_width;		/* columns = 1 << this (default 20) */
	uint32_t	hll_p;			/* registers per service = 1 << p, 4..16 (default 12) */
	uint32_t	td_compression;		/* t-digest delta, 10..220 (default 200: up to 256 centroids kept; exports for Postgres are
						   recompressed to public.tdigest(x, 100), gy_query_common.cc:1855) */
	uint32_t	max_batch;		/* max events per device batch = one ingest + merge pass, < 2^27 (default 1 << 22) */
	uint32_t	flags;			/* GYSK_FLAG_* */
	uint32_t	rank, world;		/* this engine owns events with host_idx % world == rank; world 0/1 = all */
	uint32_t	stage_batch;		/* events per host staging buffer / H2D chunk; 0 = min(max_batch, 1 << 22) */
	uint32_t	idle_evict_secs;	/* a service without events for this long (and older than twice that) is evicted at
						   gysk_flush: TIMEOUT_INET_DIAG_SECS 300, common/gy_socket_stat.h:997, rule of
						   common/gy_socket_stat.cc:3968-3982. 0 (default) = never */
	uint32_t	reserved[2];
} gysk_config;

typedef struct gysk_engine gysk_engine;

/* ---- per-service summary: the fields SvcStateFields exposes + the new sketch answers ---- */
typedef struct gysk_svc_summary
{
	uint64_t	glob_id;
	int32_t		found;			/* 0 if the id is unknown */
	uint32_t	nqrys_5s;		/* LISTENER_STATE_NOTIFY::nqrys_5s_       (last closed window) */
	uint64_t	total_resp_5sec;	/* ::total_resp_5sec_ (msec sum, last closed window) */
	int64_t		p95_5s_resp_ms;		/* ::p95_5s_resp_ms_   = get_percentile(95) of the last window */
	int64_t		p99_5s_resp_ms;
	int64_t		p25_5s_resp_ms;		/* the three percentiles listener_stats_update reads, gy_socket_stat.h:459 */
	int64_t		p95_5min_resp_ms;	/* ::p95_5min_resp_ms_ = get_percentile(95) of the 300-s level */
	int64_t		p99_5min_resp_ms;
	uint64_t	nqrys_5min;
	int64_t		p95_5day_resp_ms;
	uint64_t	nqrys_5day;
	int64_t		p95_all_resp_ms;
	int64_t		p99_all_resp_ms;
	uint64_t	nqrys_all;
	int64_t		max_resp_ms;		/* max_val_seen_ of the all-time histogram */
	uint32_t	nconns_5s;		/* TCP events of the last window */
	uint32_t	kbytes_5s;
	uint64_t	nconns_all;
	uint64_t	kbytes_all;
	double		distinct_clients;	/* HLL estimate */
	double		td_p50_us, td_p95_us, td_p99_us;	/* t-digest quantiles (usec); NaN when empty */
	uint64_t	td_count;
	uint32_t	nconns_active;		/* ACTIVE_CONN_STATS of the last window: sum of active_conns_ (-> LISTENER_STATE_NOTIFY::nconns_active_) */
	uint32_t	active_kbytes;		/* ... sum of (bytes_sent_ + bytes_received_) >> 10 */
	float		max_rtt_msec;		/* ... max of max_rtt_msec_ */
	uint32_t	cli_errors, ser_errors;	/* API_TRAN error counters of the last window (-> ::cli_errors_, ::ser_errors_) */
	uint8_t		curr_state;		/* GYSK_STATE_*: get_curr_state of the last closed window (-> ::curr_state_) */
	uint8_t		curr_issue;		/* GYSK_ISSUE_* (-> ::curr_issue_) */
	uint8_t		issue_bit_hist;		/* one bit per window, 1 = state >= BAD (-> ::issue_bit_hist_, gy_socket_stat.cc:4242-4249) */
	uint8_t		high_resp_bit_hist;	/* one bit per window, 1 = response above the 5-day level (-> ::high_resp_bit_hist_) */
} gysk_svc_summary;

/* per-host roll-up of the listener states of one 5-s tick: LISTEN_SUMM_STATS<int>, server/gy_msocket.h:840-866 */
typedef struct gysk_host_summary
{
	int32_t		nstates[8];		/* per OBJ_STATE_E value (STATE_IDLE .. STATE_DOWN) */
	int32_t		tot_qps;		/* += nqrys_5s_ / 5 */
	int32_t		tot_act_conn;		/* += nconns_active_ */
	int32_t		tot_kb_inbound;
	int32_t		tot_kb_outbound;
	int32_t		tot_ser_errors;
	int32_t		nlisteners;
	int32_t		nactive;		/* += !!nqrys_5s_ */
	int32_t		pad;
} gysk_host_summary;

/* cluster roll-up of the host summaries: the service part of MS_CLUSTER_STATE::STATE_ONE (common/gy_comm_proto.h:3183-3214) as
 * CLUSTER_STATE_ONE::update_from_state fills it from PARTHA_INFO::summstats_ (server/gy_mconnhdlr.cc:16036-16046). The task / cpu /
 * memory issue counters come from the host agent's HOST_STATE_NOTIFY and stay zero here. */
typedef struct gysk_cluster_state
{
	uint32_t	nhosts;			/* hosts with a listener-state summary */
	uint32_t	nsvc_issue;		/* listeners in STATE_BAD / STATE_SEVERE / STATE_DOWN */
	uint32_t	nsvcissue_hosts;	/* hosts with at least one such listener */
	uint32_t	nsvc;			/* += nlisteners */
	uint32_t	total_qps;		/* += tot_qps_ */
	uint32_t	svc_net_mb;		/* += (tot_kb_inbound_ + tot_kb_outbound_) / 1024 */
	uint32_t	pad[2];
} gysk_cluster_state;

/* top-N services of the last closed window (BOUNDED_PRIO_QUEUE users of partha_listener_state, gy_mconnhdlr.cc:11262-11304) */
enum { GYSK_TOPN_QPS = 0, GYSK_TOPN_CONNS = 1, GYSK_TOPN_NET = 2, GYSK_TOPN_ISSUE = 3 /* curr_state > OK, worst first: LISTEN_TOPN::is_comp_issue, server/gy_msocket.h:745 */ };
/* top-N aggregated processes of the last closed window: atask_top_cpu_ / atask_top_cpu_delay_ / atask_top_io_delay_ of
 * partha_aggr_task_state (server/gy_mconnhdlr.cc:10020-10065; a task whose metric is zero never enters a queue).
 * score = the window's sum of cpu_pct / cpu_delay msec / blkio_delay msec samples */
enum { GYSK_TOPN_TASK_CPU = 0, GYSK_TOPN_TASK_CPU_DELAY = 1, GYSK_TOPN_TASK_BLKIO_DELAY = 2 };
typedef struct gysk_topn_entry
{
	uint64_t	glob_id;
	uint64_t	score;			/* nqrys_5s / conn events / kbytes of the last window */
	uint32_t	host_idx;
	uint32_t	pad;
} gysk_topn_entry;

/* per-host rankings of gysk_topn_host: the four listener queues of partha_listener_state (server/gy_mconnhdlr.cc:11262-11304) and the
 * seven process queues of partha_aggr_task_state (:10012-10079), as of the host's last message of that kind (10 entries each,
 * BOUNDED_PRIO_QUEUE::try_emplace_locked semantics). score: SVC_ISSUE = curr_state_ << 32 | tasks_delay_usec_; TASK_ISSUE = severe << 32 |
 * ntasks_issue_ + 1; TASK_CPU = IEEE bits of total_cpu_pct_ (orders like the float); else the raw field */
enum {
	GYSK_HOSTTOP_SVC_ISSUE = 0, GYSK_HOSTTOP_SVC_QPS, GYSK_HOSTTOP_SVC_CONNS, GYSK_HOSTTOP_SVC_NET,
	GYSK_HOSTTOP_TASK_ISSUE, GYSK_HOSTTOP_TASK_NET, GYSK_HOSTTOP_TASK_CPU, GYSK_HOSTTOP_TASK_RSS, GYSK_HOSTTOP_TASK_CPU_DELAY,
	GYSK_HOSTTOP_TASK_VM_DELAY, GYSK_HOSTTOP_TASK_BLKIO_DELAY,
};

typedef struct gysk_flow_est
{
	uint64_t	flow_key;
	uint32_t	count;			/* min over rows of the count halves */
	uint32_t	kbytes;			/* min over rows of the kbytes halves */
} gysk_flow_est;

typedef struct gysk_stats
{
	uint64_t	events_in;		/* events handed to the device */
	uint64_t	events_dropped;		/* svc_id 0, bad type, table full, unknown id without AUTO_REGISTER */
	uint64_t	events_resp, events_tcp, events_task;
	uint64_t	nsvcs, ntasks;
	uint64_t	batches;
	uint64_t	kernel_launches;	/* launches of this library's own kernels so far */
	uint64_t	wire_msgs_ok, wire_msgs_bad;
	uint64_t	svcs_evicted;		/* idle services evicted so far (their slots are recycled) */
} gysk_stats;

/* mergeable device buffers (for the multi-GPU merge step) */
enum { GYSK_RED_SUM_U64 = 0, GYSK_RED_MAX_U8 = 1, GYSK_RED_MAX_I64 = 2 };
typedef struct gysk_buffer_desc
{
	const char	*name;
	void		*dptr;			/* device pointer */
	uint64_t	nbytes;
	int32_t		redop;			/* GYSK_RED_* */
	int32_t		pad;
} gysk_buffer_desc;

/* ---- lifecycle ---- */
int		gysk_abi_version(void);
void		gysk_config_default(gysk_config *cfg);
int		gysk_create(const gysk_config *cfg, gysk_engine **out);
void		gysk_destroy(gysk_engine *e);
const char *	gysk_last_error(gysk_engine *e);		/* e may be NULL: error of the last failed gysk_create */
int		gysk_get_stats(gysk_engine *e, gysk_stats *out);	/* synchronises the ingest stream */
/* diagnostic: rows of dense value bins handed out so far to "hot" services — services that brought GYSK_HOT_MIN (environment, default
 * 4096) response samples in one device batch take their later samples as direct updates of an L2-resident row instead of sort keys
 * (GYSK_HOT_ROWS rows, default 2048, 0 = off). Routing only: no result depends on it. Negative = GYSK_ERR_*. */
int64_t		gysk_hot_rows_in_use(gysk_engine *e);
/* introspection (host only, no device needed): the 64-bit word of value bin `bin` (0 .. 847) inside a hot row's first half
 * {samples | sub-msec remainders}; the bin's usec sum lies GYSK_HOT_ROW_BINS words further on. Neighbouring bins are two 128-byte
 * lines apart. ~0 for a bin the engine does not have. */
#define GYSK_HOT_ROW_BINS	1024u
uint32_t	gysk_hot_row_word(uint32_t bin);
/* diagnostic: response samples of the last device batch that travelled as sort keys (the rest updated hot rows). Negative = GYSK_ERR_*. */
int64_t		gysk_last_batch_keys(gysk_engine *e);

/* ---- registration (control path; mirrors partha_listener_info registering listeners before state arrives) ---- */
int		gysk_register_ids(gysk_engine *e, const uint64_t *ids, uint32_t n, int is_task);

/* ---- ingest ---- */
int		gysk_ingest(gysk_engine *e, const uint8_t host_id[16], uint32_t host_idx, uint32_t subtype,
				void *recs, uint32_t nevents, const void *endptr);
int		gysk_ingest_msg(gysk_engine *e, const uint8_t host_id[16], uint32_t host_idx, void *comm_header_msg, uint32_t msglen);
int		gysk_ingest_raw(gysk_engine *e, const uint8_t host_id[16], uint32_t host_idx, uint32_t kind,
				const void *events, uint32_t nevents);
/* zero-copy variant for GYSK_RAW_EVENT32 in page-locked host memory: the buffer is read asynchronously and must stay
 * valid and unmodified until the next gysk_sync() returns */
int		gysk_ingest_pinned(gysk_engine *e, const gysk_event *pinned_events, uint64_t nevents);
int		gysk_ingest_device(gysk_engine *e, const gysk_event *d_events, uint64_t nevents);
int		gysk_export_task_hist(gysk_engine *e, uint64_t aggr_task_id, int which, gysk_hist_serial out[GYSK_HIST_MAX_BUCKETS],
				uint64_t *total_count, int64_t *max_val);
int		gysk_sync(gysk_engine *e);
int		gysk_flush(gysk_engine *e, uint32_t tsec);
/* ids evicted by the most recent gysk_flush (the LISTEN_FLAG_DELETE notifications of common/gy_socket_stat.cc:4023-4033);
 * synchronises the ingest stream. *n = number of ids (may exceed cap: then only cap are written) */
int		gysk_evicted_ids(gysk_engine *e, uint64_t *out, uint32_t cap, uint32_t *n);

/* ---- queries ---- */
int		gysk_query_svcs(gysk_engine *e, const uint64_t *glob_ids, uint32_t n, gysk_svc_summary *out);
int		gysk_query_flows(gysk_engine *e, const uint64_t *flow_keys, uint32_t n, int last_window, gysk_flow_est *out);
/* LISTEN_SUMM_STATS of the last NOTIFY_LISTENER_STATE message of a host (partha_listener_state, gy_mconnhdlr.cc:11251) */
/* host_idx < 0: over all hosts of this engine; n <= 64 */
int		gysk_topn_svcs(gysk_engine *e, int metric, int32_t host_idx, uint32_t n, gysk_topn_entry *out, uint32_t *nout);
int		gysk_topn_tasks(gysk_engine *e, int metric, uint32_t n, gysk_topn_entry *out, uint32_t *nout);
int		gysk_topn_host(gysk_engine *e, int what /* GYSK_HOSTTOP_* */, int32_t host_idx /* < 0: all hosts */, uint32_t n, gysk_topn_entry *out, uint32_t *nout);
/* roll-up over the given hosts (the hosts of one cluster_name_; host_idxs NULL = every host of this engine): what
 * MCONN_HANDLER::send_cluster_state (server/gy_mconnhdlr.cc:16052) sends to shyama per cluster */
int		gysk_query_cluster_state(gysk_engine *e, const uint32_t *host_idxs, uint32_t n, gysk_cluster_state *out);	/* n <= 64; glob_id = aggr_task_id */
int		gysk_query_host_summary(gysk_engine *e, uint32_t host_idx, gysk_host_summary *out);
int		gysk_export_hist(gysk_engine *e, uint64_t id, int which, gysk_hist_serial out[GYSK_HIST_MAX_BUCKETS],
				uint64_t *total_count, int64_t *max_val);
int		gysk_export_hll(gysk_engine *e, uint64_t glob_id, uint8_t *regs /* 1 << hll_p bytes */);
/* TCP_LISTENER::CONN_BITMAP (common/gy_socket_stat.h:390-455): per response bucket a 32-bit mask over (client port & 0x1F) and its
 * popcount = get_conn_breakup(); bit index = flow_key & 0x1F */
int		gysk_export_conn_bitmap(gysk_engine *e, uint64_t glob_id, int last_window, uint32_t masks[GYSK_HIST_MAX_BUCKETS],
				uint8_t nconn_arr[GYSK_HIST_MAX_BUCKETS]);
int		gysk_export_tdigest(gysk_engine *e, uint64_t glob_id, double *means, uint64_t *weights, uint32_t cap, uint32_t *n,
				double *min_val, double *max_val);
int		gysk_query_quantiles(gysk_engine *e, uint64_t glob_id, const double *qs, uint32_t nq, double *out);
int		gysk_export_cms(gysk_engine *e, int last_window, uint64_t *cells /* depth << log2_width entries */);

/* ---- row a15b: the per-process -> per-aggregate-process group-by in front of partha_aggr_task_state ----
 * One record per process and 5-s tick, holding what TASK_HANDLER's walk has at hand when it folds the process into
 * aggrnotmap.try_emplace(aggr_task_id) (common/gy_task_handler.cc:752-880). */
typedef struct gysk_proc_sample
{
	uint64_t	aggr_task_id;		/* ptask->aggr_task_id_ */
	int32_t		pid;			/* ptask->task_pid */
	float		cpu_pct;		/* avg_cpu_pct / npct: this tick's cpu % averaged with the ticks the server missed (:826-839) */
	uint32_t	rss_mb;
	uint32_t	cpu_delay_msec, vm_delay_msec, blkio_delay_msec;	/* last_*_delay_nsec / GY_NSEC_PER_MSEC (:858-860) */
	uint32_t	tcp_kbytes, tcp_conns;	/* last_sent_tcp_kbytes_ / _conns_: non-zero in the 15-s network ticks only (:791-813) */
	uint8_t		state;			/* pext->issue_hist_[0].state (OBJ_STATE_E) */
	uint8_t		issue;			/* pext->issue_hist_[0].issue */
	uint8_t		issue_bit_hist, severe_issue_bit_hist;
	uint8_t		is_issue;
	uint8_t		pad[3];
	char		comm[16];		/* ptask->task_comm */
} gysk_proc_sample;
/* Folds the samples by aggr_task_id IN ARRAY ORDER with the reference's statement order (the float cpu sum sees the same sequence of
 * additions) and writes one AGGR_TASK_STATE_NOTIFY record (72 bytes, common/gy_comm_proto.h:2114-2170, no issue string) per group,
 * groups in order of first appearance: the body of a NOTIFY_AGGR_TASK_STATE message, ready for gysk_ingest(). *ngroups = number of
 * groups found; at most `cap` records are written. n <= gysk_config.max_batch. */
int		gysk_task_groupby(gysk_engine *e, const gysk_proc_sample *samples, uint32_t n, void *out_records, uint32_t cap, uint32_t *ngroups);

/* ---- pure helpers (host side, no engine): the reference's percentile rule and the sketch estimators ---- */
int		gysk_hist_nbuckets(int cls);
int		gysk_hist_bucket(int cls, int64_t value);	/* RESP_TIME_HASH::get_bucket_from_data & siblings */
int		gysk_hist_percentiles(int cls, int t_is_int, const gysk_hist_serial *stats, uint64_t total_count,
				const float *pcts, uint32_t npct, int64_t *out);
double		gysk_hll_estimate(const uint8_t *regs, uint32_t p);
/* TCP_LISTENER::get_curr_state (common/gy_socket_stat.cc:2020-2875): shifts / sets *high_resp_bit_hist, writes GYSK_STATE_* / GYSK_ISSUE_* */
int		gysk_classify_listener(const gysk_listener_state_in *in, uint8_t *high_resp_bit_hist, uint8_t *state, uint8_t *issue);
/* per-service summaries -> LISTENER_STATE_NOTIFY records (common/gy_comm_proto.h:2183-2254), the body of one
 * NOTIFY_LISTENER_STATE message (<= 512 records, 88 bytes each) that MTCP_LISTENER::set_state / partha_listener_state consume
 * (server/gy_mconnhdlr.cc:11175-11251). Entries with found == 0 are skipped. No engine needed. */
int		gysk_encode_listener_state(const gysk_svc_summary *sums, uint32_t n, void *buf, uint32_t cap, uint32_t *nrecs, uint32_t *nbytes);
/* a digest in the text form of the Postgres `tdigest` type the reference stores and queries (public.tdigest(expr, 100) /
 * tdigest_percentile, common/gy_query_common.cc:1805-1858): "flags 1 count N compression C centroids K (mean, count) ...".
 * Both return the string length, or a negative GYSK_ERR_* */
int		gysk_tdigest_to_pgtext(const double *means, const uint64_t *weights, uint32_t n, uint32_t compression, char *buf, uint32_t cap);
int		gysk_export_tdigest_pgtext(gysk_engine *e, uint64_t glob_id, char *buf, uint32_t cap);
double		gysk_tdigest_quantile(const double *means, const uint64_t *weights, uint32_t n, double min_val, double max_val, double q);
uint32_t	gysk_uint64_hash(uint64_t key);		/* get_uint64_hash, common/gy_common_inc.h:1120 */

/* ---- multi-GPU merge ---- */
/* Which services form a logical service (the reference's svc-mesh cluster id). The same list on every rank; dense logical indices
 * follow first appearance. Ids need not be registered (or owned by this rank): the map keeps every pair and looks the slots up at
 * every gysk_merge_prepare / gysk_merge_global. */
int		gysk_set_logical_map(gysk_engine *e, const uint64_t *glob_ids, const uint64_t *logical_ids, uint32_t n);
int		gysk_merge_prepare(gysk_engine *e);		/* fold per-service sketches into per-logical-service arrays; asynchronous on
							   gysk_stream(e): enqueue the collectives on that stream, or gysk_sync() first */
int		gysk_merge_buffers(gysk_engine *e, gysk_buffer_desc *out, uint32_t cap, uint32_t *n);
int		gysk_merge_tdigest_slab(gysk_engine *e, void **dptr, uint64_t *nbytes);	/* fixed slab to all-gather */
int		gysk_merge_finish(gysk_engine *e, const void *d_gathered_slabs, uint32_t world);
/* The same step with NCCL inside the library (a C++ madhava has no torch.distributed): fold, then ONE grouped NCCL launch — an
 * all-reduce per reduction kind (u64 SUM, i64 MAX, u8 MAX) and the all-gather of the t-digest slabs — then the merge-compress,
 * all enqueued on the engine's stream. libnccl.so.2 is loaded on first use (dlopen; GYSK_ERR_NOTSUP when absent).
 * comm: an ncclComm_t created by the caller over the engines' devices (ncclCommInitAll / ncclCommInitRank), or NULL to use the one
 * gysk_nccl_comm_init() made: rank 0 calls gysk_nccl_unique_id(), ships the 128 bytes to its peers, every rank calls
 * gysk_nccl_comm_init(engine, uid, nranks, rank). */
#define GYSK_NCCL_UNIQUE_ID_BYTES	128
int		gysk_nccl_unique_id(uint8_t out[GYSK_NCCL_UNIQUE_ID_BYTES]);
int		gysk_nccl_comm_init(gysk_engine *e, const uint8_t uid[GYSK_NCCL_UNIQUE_ID_BYTES], uint32_t nranks, uint32_t rank);
int		gysk_merge_global(gysk_engine *e, void *nccl_comm);
int		gysk_query_logical(gysk_engine *e, const uint64_t *logical_ids, uint32_t n, gysk_svc_summary *out);
int		gysk_query_flows_global(gysk_engine *e, const uint64_t *flow_keys, uint32_t n, int last_window, gysk_flow_est *out);

/* per-kernel device timing (CUDA events on the launching stream around the ingest kernel and around the sort +
 * t-digest chain of every device batch). read() synchronises, returns the sums since the last read and resets. */
int		gysk_profile_enable(gysk_engine *e, int on);
int		gysk_profile_read(gysk_engine *e, double *ms_ingest, double *ms_tdigest, uint64_t *nbatches);

/* CUDA stream the engine launches on (cudaStream_t as void*), for callers timing with events */
void *		gysk_stream(gysk_engine *e);

#ifdef __cplusplus
}
#endif

#endif /* GYSKETCH_H */
