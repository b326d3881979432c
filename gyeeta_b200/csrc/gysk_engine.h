// gysk_engine.h — engine object shared by gysk_engine.cu (ingest / query) and gysk_merge.cu (multi-GPU merge).
#pragma once

#include "gysk_kernels.cuh"

#include <algorithm>
#include <atomic>
#include <memory>
#include <map>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

namespace gysk {

constexpr int NBUF = 2;
constexpr uint32_t QCHUNK = 1024;		// ids per query kernel launch
constexpr uint32_t THREAD_STAGE_EVENTS = 1u << 16;	// events per per-thread staging chunk (2 MB page-locked, two chunks per thread)
constexpr uint32_t RAW_BULK_MIN = 16384;		// raw fixed-stride batches from this size on are expanded on the device: below it the
							// per-call copy / launch / event calls under the engine mutex cost more than the
							// ~20 ns per record of expanding on the calling thread (bench.py e2e_wire)

// a calling thread's page-locked staging: filled without the engine mutex (see gysk_engine.cu)
struct ThreadStage
{
	std::mutex		m;
	gysk_event		*buf[2] {};
	cudaEvent_t		copied[2] {};
	int			cur {0};
	uint32_t		fill {0}, cap {0};
	std::atomic<bool>	orphan {false};		// its thread has exited: the next new thread of this engine takes it over
};

// bounded top-N of one host's last message: BOUNDED_PRIO_QUEUE::try_emplace_locked semantics (common/gy_statistics.h:385-414:
// keep the N largest by the comparator; an element enters only when the queue has room or it beats the current minimum)
struct TopEntry { uint64_t id; uint64_t score; };
struct TopQueue
{
	static constexpr size_t N = 10;			// MAX_LISTEN_TOPN / MAX_TASK_TOPN per host, server/gy_mconnhdlr.h:961,975
	std::vector<TopEntry>	v;
	void offer(uint64_t id, uint64_t score)
	{
		if (!score) return;
		if (v.size() < N) { v.push_back({id, score}); std::push_heap(v.begin(), v.end(), [](const TopEntry &a, const TopEntry &b) { return a.score > b.score; }); return; }
		if (score <= v.front().score) return;
		std::pop_heap(v.begin(), v.end(), [](const TopEntry &a, const TopEntry &b) { return a.score > b.score; });
		v.back() = {id, score};
		std::push_heap(v.begin(), v.end(), [](const TopEntry &a, const TopEntry &b) { return a.score > b.score; });
	}
};

// the four listener rankings of partha_listener_state (server/gy_mconnhdlr.cc:11262-11304; comparators LISTEN_TOPN
// server/gy_msocket.h:720-797): by issue (state, then qps), qps, active connections, network kbytes
struct HostTopn
{
	TopQueue	q[4];			// GYSK_HOSTTOP_SVC_ISSUE, _QPS, _CONNS, _NET
	template <typename L> void offer(const L &l)
	{
		if (l.curr_state_ > 2) q[0].offer(l.glob_id_, ((uint64_t)l.curr_state_ << 32) | l.tasks_delay_usec_);	// is_comp_issue: state, then task delay (:745)
		if (l.nqrys_5s_ >= 5) q[1].offer(l.glob_id_, l.nqrys_5s_);						// :11273
		if (l.nconns_active_ >= 1) q[2].offer(l.glob_id_, l.nconns_active_);					// :11284
		q[3].offer(l.glob_id_, (uint64_t)l.curr_kbytes_inbound_ + l.curr_kbytes_outbound_);			// :11295 (offer skips 0)
	}
};
// the seven process rankings of partha_aggr_task_state (server/gy_mconnhdlr.cc:10012-10079; comparators MAGGR_TASK_STATE
// server/gy_msocket.h:454-531, MTASK_ISSUE :602): issue, net, cpu, rss, cpu delay, vm delay, blkio delay
struct HostTaskTopn
{
	TopQueue	q[7];
	template <typename T> void offer(const T &t)
	{
		// is_comp_issue (:602-606): more tasks with issues first; a severe aggregate beats a non-severe one
		if (t.curr_state_ > 2) q[0].offer(t.aggr_task_id_, ((uint64_t)((t.severe_issue_bit_hist_ & 1u) && t.ntasks_issue_) << 32) | ((uint64_t)t.ntasks_issue_ + 1));
		if (t.tcp_kbytes_ > 0) q[1].offer(t.aggr_task_id_, t.tcp_kbytes_);
		if (t.total_cpu_pct_ >= 0.1f) { uint32_t bits; memcpy(&bits, &t.total_cpu_pct_, 4); q[2].offer(t.aggr_task_id_, bits); }	// positive floats order like their bits
		if (t.rss_mb_ >= 5) q[3].offer(t.aggr_task_id_, t.rss_mb_);
		q[4].offer(t.aggr_task_id_, t.cpu_delay_msec_);
		q[5].offer(t.aggr_task_id_, t.vm_delay_msec_);
		q[6].offer(t.aggr_task_id_, t.blkio_delay_msec_);
	}
};

// per-logical-service state of the merge step
struct MergeState
{
	uint32_t		nlogical {0};
	std::vector<uint64_t>	logical_ids;		// dense index -> logical id
	std::unordered_map<uint64_t, uint32_t> index;	// logical id -> dense index
	uint32_t		*d_offsets {nullptr}, *d_members {nullptr};	// CSR: logical -> member slots on this GPU
	unsigned long long	*d_member_ids {nullptr};			// id each member slot held when the map was set
	uint32_t		nmembers {0};
	// one arena so that each reduction kind is a single collective
	uint8_t			*arena {nullptr};
	size_t			arena_bytes {0};
	size_t			off_sum {0}, bytes_sum {0};		// u64 SUM : cms cur/last, hist last/all, conn
	size_t			off_maxi64 {0}, bytes_maxi64 {0};	// i64 MAX : histogram max_val_seen_
	size_t			off_maxu8 {0}, bytes_maxu8 {0};		// u8  MAX : HLL registers
	unsigned long long	*g_cms_cur {nullptr}, *g_cms_last {nullptr};
	HistCell		*l_hist_last {nullptr}, *l_hist_all {nullptr};
	unsigned long long	*l_conn {nullptr};			// [nl][4]: last cnt, last kb, all cnt, all kb
	long long		*l_hmax {nullptr};			// [nl][2]: last, all
	uint8_t			*l_hll {nullptr};
	// t-digest slab: fixed [nl] x {TdHead, Centroid[TD_CAP]}
	uint8_t			*slab {nullptr};
	size_t			slab_bytes {0};
	uint8_t			*final_slab {nullptr};			// merged over ranks
	bool			prepared {false}, finished {false};
	void			*comm {nullptr};			// ncclComm_t of gysk_nccl_comm_init
	uint32_t		comm_world {0};
	bool			comm_owned {false};
	uint8_t			*gathered {nullptr};			// [world] slabs, target of the all-gather
	uint32_t		gathered_world {0};
};

} // namespace gysk

struct gysk_engine
{
	gysk_config		cfg {};
	int			dev {0};
	cudaStream_t		stream {nullptr}, copy_stream {nullptr}, side_stream {nullptr};
	cudaEvent_t		ev_ingested {nullptr}, ev_side_done {nullptr};	// main -> side after ingest_kernel, side -> main at the end of the batch
	gysk::DevState		st {};
	gysk::SortTemp		tmp {};
	std::vector<void *>	dallocs;
	std::vector<void *>	hallocs;

	// staging: per-thread page-locked chunks -> device event buffers (double-buffered) -> kernels
	uint64_t		uid {0};
	std::mutex		tstage_mtx;
	std::vector<std::unique_ptr<gysk::ThreadStage>> tstages;
	gysk_event		*d_events[gysk::NBUF] {};
	cudaEvent_t		ev_copied[gysk::NBUF] {}, ev_done[gysk::NBUF] {};
	uint32_t		stage_fill {0};			// events of d_events[stage_cur] whose copies are enqueued
	int			stage_cur {0};
	uint8_t			*d_raw[gysk::NBUF] {};		// raw records on their way to decode_raw_kernel
	size_t			raw_bytes {0};
	cudaEvent_t		ev_raw_copied[gysk::NBUF] {}, ev_raw_done[gysk::NBUF] {};
	int			raw_cur {0};

	// query scratch
	unsigned long long	*d_qids {nullptr}, *h_qids {nullptr};
	gysk::SvcRaw		*d_svcraw {nullptr}, *h_svcraw {nullptr};
	gysk::TaskRaw		*d_taskraw {nullptr}, *h_taskraw {nullptr};
	uint8_t			*d_hllout {nullptr}, *h_hllout {nullptr};
	int32_t			*d_found {nullptr}, *h_found {nullptr};
	gysk_flow_est		*d_flowout {nullptr}, *h_flowout {nullptr};
	unsigned long long	*h_counters {nullptr};

	// optional per-kernel timing
	bool			profiling {false};
	std::vector<cudaEvent_t> prof_events;		// triples: before ingest, after ingest, after t-digest chain
	size_t			prof_used {0};

	// rolling levels: epoch held by each ring slot (~0 = never written) and the time of the last flush
	uint64_t		ring_epoch[gysk::NLEVELS][gysk::NSLOTS];
	uint32_t		last_flush_tsec {0};

	// idle-service eviction: ids evicted by the last flush arrive lagged (copied behind the flush kernels, read at the next call)
	unsigned long long	*h_evict {nullptr};			// [0] = count, [1..] = ids (pinned)
	cudaEvent_t		ev_evict {nullptr};
	bool			evict_pending {false};
	std::vector<uint64_t>	evicted_ids;				// of the last completed flush
	uint64_t		tombstones {0}, evicted_total {0};
	uint64_t		h_evict_fail {0};			// CTR_INSERT_FAIL as of the last collected flush
	uint64_t		insert_fail_seen {0};			// CTR_INSERT_FAIL at the last table rebuild
	uint32_t		sort_epoch {0};				// radix passes launched so far (tags the look-back status words)

	std::mutex		host_mtx;					// the per-host control-plane state below
	std::unordered_map<uint32_t, gysk_host_summary> host_summ;	// last LISTEN_SUMM_STATS per host (control-plane sized: <= 512 hosts)
	std::unordered_map<uint32_t, gysk::HostTopn> host_topn;		// top-N listeners of each host's last NOTIFY_LISTENER_STATE
	std::unordered_map<uint32_t, gysk::HostTaskTopn> host_task_topn;	// top-N aggregated processes of each host's last NOTIFY_AGGR_TASK_STATE

	gysk::MergeState	mg;

	std::mutex		mtx;
	std::string		err;
	bool			sticky {false};
	std::atomic<uint64_t>	wire_ok {0}, wire_bad {0};
	uint64_t		kernel_launches {0}, batches {0}, merges {0};
};

namespace gysk {

int fail(gysk_engine *e, int code, const char *what, cudaError_t ce = cudaSuccess);
int post_launch(gysk_engine *e, const char *what);
int submit_stage(gysk_engine *e);
int append_chunk(gysk_engine *e, const gysk_event *src_pinned, uint64_t n);
int drain_all(gysk_engine *e);
int sync_locked(gysk_engine *e);
int collect_evicted(gysk_engine *e, bool wait);
void merge_release(gysk_engine *e);
void summarize_raw(const gysk_engine *e, const SvcRaw &r, uint64_t id, gysk_svc_summary &o);

#define CU(e, call) do { cudaError_t ce__ = (call); if (ce__ != cudaSuccess) return gysk::fail((e), GYSK_ERR_CUDA, #call, ce__); } while (0)
// readers: hand every thread's partial chunk to the device, then take the engine
#define GYSK_ENTER(e) { int rc_d__ = gysk::drain_all(e); if (rc_d__) return rc_d__; } std::lock_guard<std::mutex> lk((e)->mtx)
#define CHECK_ENGINE(e) do { if (!(e)) return GYSK_ERR_INVAL; if ((e)->sticky) return GYSK_ERR_CUDA; } while (0)

template <typename T>
int dalloc(gysk_engine *e, T **p, size_t n, bool zero = true)
{
	void *q = nullptr;
	cudaError_t ce = cudaMalloc(&q, n * sizeof(T));

	if (ce != cudaSuccess) return fail(e, GYSK_ERR_NOMEM, "cudaMalloc", ce);
	e->dallocs.push_back(q);
	if (zero) {
		ce = cudaMemsetAsync(q, 0, n * sizeof(T), e->stream);
		if (ce != cudaSuccess) return fail(e, GYSK_ERR_CUDA, "cudaMemsetAsync", ce);
	}
	*p = static_cast<T *>(q);
	return 0;
}

template <typename T>
int halloc(gysk_engine *e, T **p, size_t n)
{
	void *q = nullptr;
	cudaError_t ce = cudaHostAlloc(&q, n * sizeof(T), cudaHostAllocDefault);

	if (ce != cudaSuccess) return fail(e, GYSK_ERR_NOMEM, "cudaHostAlloc", ce);
	e->hallocs.push_back(q);
	*p = static_cast<T *>(q);
	return 0;
}

} // namespace gysk
