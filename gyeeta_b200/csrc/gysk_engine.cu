// gysk_engine.cu — host runtime of libgysketch.so and the C ABI declared in include/gysketch.h.
//
// One engine = one GPU. Ingest calls are serialised by a mutex, copy their input into page-locked staging (the
// reference's handlers never retain caller buffers: DB_WRITE_ARR frees them when the L2 loop iteration ends,
// server/gy_mconnhdlr.h:424-431) and hand full batches to the device: H2D on a copy stream, kernels on the compute
// stream, two device event buffers so the copy of batch k+1 overlaps the kernels of batch k.
#include "gysk_engine.h"
#include "gysk_state.cuh"
#include "gysk_wire.h"

using namespace gysk;

namespace {

thread_local std::string g_create_error;
std::atomic<uint64_t> g_engine_uid {1};

} // namespace

namespace gysk {

int fail(gysk_engine *e, int code, const char *what, cudaError_t ce)
{
	char buf[512];

	if (ce != cudaSuccess) snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(ce));
	else snprintf(buf, sizeof(buf), "%s", what);
	if (e) {
		e->err = buf;
		if (code == GYSK_ERR_CUDA) e->sticky = true;
	}
	else g_create_error = buf;
	return code;
}

int post_launch(gysk_engine *e, const char *what)
{
	cudaError_t ce = cudaGetLastError();
	if (ce != cudaSuccess) return fail(e, GYSK_ERR_CUDA, what, ce);
	return 0;
}

} // namespace gysk

namespace {

uint32_t pow2_at_least(uint64_t v) { uint32_t p = 16; while (p < v) p <<= 1; return p; }

// one device batch: ingest kernel, then the sort + t-digest chain over the keys it emitted
template <typename AfterIngest>
int process_device_batch(gysk_engine *e, const gysk_event *d_ev, uint64_t n, AfterIngest after_ingest)
{
	if (!n) return 0;
	if (n >= (1ull << BIN_CNT_BITS)) return fail(e, GYSK_ERR_INVAL, "device batch holds 2^27 or more events");
	cudaEvent_t *pe = nullptr;
	if (e->profiling) {
		if (e->prof_used + 3 > e->prof_events.size()) {
			for (int i = 0; i < 3; ++i) {
				cudaEvent_t ev;
				CU(e, cudaEventCreate(&ev));
				e->prof_events.push_back(ev);
			}
		}
		pe = &e->prof_events[e->prof_used];
		e->prof_used += 3;
		CU(e, cudaEventRecord(pe[0], e->stream));
	}
	e->kernel_launches += launch_ingest(e->st, e->tmp, d_ev, n, e->cfg.max_svcs, e->stream);
	if (pe) CU(e, cudaEventRecord(pe[1], e->stream));
	const bool side = side_drain_enabled();
	if (side) {
		// the queued connection / process records are applied on the side stream, next to the sort chain of the main stream
		CU(e, cudaEventRecord(e->ev_ingested, e->stream));
		CU(e, cudaStreamWaitEvent(e->side_stream, e->ev_ingested, 0));
		e->kernel_launches += launch_side_drain(e->st, e->tmp, n, e->side_stream);
		CU(e, cudaEventRecord(e->ev_side_done, e->side_stream));
	}
	// the events of this batch are consumed once the ingest kernel has run: callers release / refill the event buffer here,
	// so that the next H2D copy overlaps the merge kernels
	{ int rc_ai = after_ingest(); if (rc_ai) return rc_ai; }
	// No number travels back to the host inside a batch: the list of touched services and its length stay in device memory.
	e->kernel_launches += launch_batch_merge(e->st, e->tmp, n, e->cfg.max_svcs, e->stream);
	if (side) CU(e, cudaStreamWaitEvent(e->stream, e->ev_side_done, 0));		// the batch is complete on the main stream only with its side work
	if (pe) CU(e, cudaEventRecord(pe[2], e->stream));
	e->batches++;
	return post_launch(e, "ingest batch");
}

int process_device_batch(gysk_engine *e, const gysk_event *d_ev, uint64_t n)
{
	return process_device_batch(e, d_ev, n, []() { return 0; });
}

} // namespace

// pick up the eviction list of the last flush once its copy has landed (wait = block until it has)
int gysk::collect_evicted(gysk_engine *e, bool wait)
{
	if (!e->evict_pending) return 0;
	if (wait) CU(e, cudaEventSynchronize(e->ev_evict));
	else if (cudaEventQuery(e->ev_evict) != cudaSuccess) { cudaGetLastError(); CU(e, cudaEventSynchronize(e->ev_evict)); }
	const uint64_t cnt = std::min<uint64_t>(e->h_evict[0], e->cfg.max_svcs);
	e->h_evict_fail = e->h_evict[e->cfg.max_svcs + 1];
	e->evicted_ids.assign(e->h_evict + 1, e->h_evict + 1 + cnt);
	e->tombstones += cnt; e->evicted_total += cnt;
	e->evict_pending = false;
	return 0;
}

// ---- staging: per-thread page-locked buffers -> device event buffer -> kernels -------------------------------------
//
// Up to 16 handle_l2_misc threads call gysk_ingest concurrently (server/gy_mconnhdlr.h:53-63, routing :16252). Each calling
// thread owns a ThreadStage: two page-locked chunks it fills WITHOUT the engine mutex — validation, record walk and compaction of
// the wire records run in parallel across threads. Only a full chunk takes the engine mutex, for as long as it takes to enqueue
// one asynchronous H2D copy into the current device event buffer (and, when that buffer is full, the batch's kernel launches).
// Readers (sync, flush, queries) first drain every thread's partial chunk, so "all events handed in before the call are applied".
// Lock order: ThreadStage::m, then gysk_engine::mtx; drain_all takes tstage_mtx, then each ThreadStage::m in turn.

// device event buffer k holds stage_fill events whose copies are enqueued on copy_stream: run the batch
int gysk::submit_stage(gysk_engine *e)
{
	const int k = e->stage_cur;
	const uint32_t n = e->stage_fill;

	if (!n) return 0;
	CU(e, cudaEventRecord(e->ev_copied[k], e->copy_stream));
	CU(e, cudaStreamWaitEvent(e->stream, e->ev_copied[k], 0));
	int rc = process_device_batch(e, e->d_events[k], n, [&]() -> int { CU(e, cudaEventRecord(e->ev_done[k], e->stream)); return 0; });
	if (rc) return rc;
	e->stage_cur = (k + 1) % NBUF;
	e->stage_fill = 0;
	return 0;
}

// enqueue the copy of n events from page-locked host memory behind what the current device buffer already holds (engine mutex held)
int gysk::append_chunk(gysk_engine *e, const gysk_event *src, uint64_t n)
{
	while (n) {
		const int k = e->stage_cur;
		const uint32_t room = e->cfg.stage_batch - e->stage_fill;
		const uint32_t m = (uint32_t)std::min<uint64_t>(room, n);

		if (e->stage_fill == 0) CU(e, cudaStreamWaitEvent(e->copy_stream, e->ev_done[k], 0));	// the kernels that last read buffer k have run
		CU(e, cudaMemcpyAsync(e->d_events[k] + e->stage_fill, src, (size_t)m * sizeof(gysk_event), cudaMemcpyHostToDevice, e->copy_stream));
		e->stage_fill += m; src += m; n -= m;
		if (e->stage_fill == e->cfg.stage_batch) { int rc = submit_stage(e); if (rc) return rc; }
	}
	return 0;
}

int gysk::sync_locked(gysk_engine *e)
{
	int rc = submit_stage(e);
	if (rc) return rc;
	CU(e, cudaStreamSynchronize(e->copy_stream));
	CU(e, cudaStreamSynchronize(e->stream));
	return 0;
}

namespace {

// engines alive (uid): a thread that exits hands its stages back to the engines that still exist
std::mutex g_live_mtx;
std::vector<uint64_t> g_live;

struct TlsStages
{
	std::vector<std::pair<uint64_t, ThreadStage *>> v;
	~TlsStages()
	{
		std::lock_guard<std::mutex> lk(g_live_mtx);
		for (auto &kv : v) if (std::find(g_live.begin(), g_live.end(), kv.first) != g_live.end()) kv.second->orphan.store(true, std::memory_order_release);
	}
};
thread_local TlsStages tls_stages_holder;
#define tls_stages tls_stages_holder.v

// the calling thread's stage of this engine: its own, else one whose thread is gone (madhava's handler threads live as long as the
// process, but a pool that recycles threads must not pay two cudaHostAlloc calls per new thread), else a new one
ThreadStage *get_stage(gysk_engine *e)
{
	for (auto &kv : tls_stages) if (kv.first == e->uid) return kv.second;
	std::lock_guard<std::mutex> lk(e->tstage_mtx);
	for (auto &old : e->tstages) {
		bool want = true;
		if (old->orphan.compare_exchange_strong(want, false, std::memory_order_acq_rel)) {
			if (tls_stages.size() > 64) tls_stages.erase(tls_stages.begin());
			tls_stages.emplace_back(e->uid, old.get());
			return old.get();
		}
	}
	auto ts = std::make_unique<ThreadStage>();
	ts->cap = std::min<uint32_t>(e->cfg.stage_batch, THREAD_STAGE_EVENTS);
	if (cudaSetDevice(e->dev) != cudaSuccess) return nullptr;
	for (int i = 0; i < 2; ++i) {
		if (cudaHostAlloc((void **)&ts->buf[i], (size_t)ts->cap * sizeof(gysk_event), cudaHostAllocDefault) != cudaSuccess) return nullptr;
		if (cudaEventCreateWithFlags(&ts->copied[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
	}
	ThreadStage *raw = ts.get();
	e->tstages.push_back(std::move(ts));
	if (tls_stages.size() > 64) tls_stages.erase(tls_stages.begin());		// engines long gone
	tls_stages.emplace_back(e->uid, raw);
	return raw;
}

// hand the filled part of the thread's current chunk to the device (ThreadStage::m held by the caller)
int flush_stage(gysk_engine *e, ThreadStage *ts)
{
	if (!ts->fill) return 0;
	{
		std::lock_guard<std::mutex> lk(e->mtx);
		CU(e, cudaSetDevice(e->dev));
		int rc = append_chunk(e, ts->buf[ts->cur], ts->fill);
		if (rc) return rc;
		CU(e, cudaEventRecord(ts->copied[ts->cur], e->copy_stream));
	}
	ts->cur ^= 1; ts->fill = 0;
	CU(e, cudaEventSynchronize(ts->copied[ts->cur]));		// the other chunk's copy (issued a whole chunk ago) has left the host
	return 0;
}

inline gysk_event *stage_slot(gysk_engine *e, ThreadStage *ts, int *rc)
{
	if (ts->fill == ts->cap) { *rc = flush_stage(e, ts); if (*rc) return nullptr; }
	return ts->buf[ts->cur] + ts->fill++;
}

int stage_events(gysk_engine *e, ThreadStage *ts, const gysk_event *ev, uint64_t n)
{
	while (n) {
		if (ts->fill == ts->cap) { int rc = flush_stage(e, ts); if (rc) return rc; }
		const uint32_t m = (uint32_t)std::min<uint64_t>(ts->cap - ts->fill, n);
		memcpy(ts->buf[ts->cur] + ts->fill, ev, (size_t)m * sizeof(gysk_event));
		ts->fill += m; ev += m; n -= m;
	}
	return 0;
}

} // namespace

// every thread's partial chunk goes to the device (called by readers BEFORE they take the engine mutex)
int gysk::drain_all(gysk_engine *e)
{
	std::lock_guard<std::mutex> lk(e->tstage_mtx);
	for (auto &ts : e->tstages) {
		std::lock_guard<std::mutex> l2(ts->m);
		int rc = flush_stage(e, ts.get());
		if (rc) return rc;
	}
	return 0;
}

namespace {

// ---- pure host helpers: the reference's percentile rule and the estimators -------------------------------

struct ClsDesc { int nthr; int64_t thr[16]; int64_t minv, maxv; bool trunc_int; int fixed_diff; };

const ClsDesc g_cls[8] = {
	{13, {1, 10, 30, 60, 100, 150, 200, 300, 450, 700, 1000, 3000, 15000}, 0, 15001, false, 0},		// RESP_TIME_HASH    gy_statistics.h:1677
	{12, {1, 10, 100, 500, 1000, 5000, 25000, 50000, 100000, 300000, 1000000, 5000000}, 0, 5000001, true, 0},	// SEMI_LOG_HASH     :1732
	{13, {1, 10, 50, 200, 500, 1000, 3000, 6000, 10000, 15000, 25000, 60000, 150000}, 0, 150001, true, 0},	// SEMI_LOG_HASH_LO  :1785
	{13, {1, 10, 25, 50, 125, 400, 1000, 3000, 6000, 10000, 25000, 40000, 65000}, 0, 65001, true, 0},	// DURATION_HASH     :1838
	{12, {10, 25, 50, 75, 100, 150, 300, 500, 800, 1000, 2000, 5000}, 0, 5001, true, 0},			// HASH_10_5000      :1911
	{10, {5, 10, 20, 40, 60, 80, 100, 140, 200, 250}, 0, 251, true, 0},					// HASH_5_250        :1963
	{12, {1, 5, 10, 25, 50, 75, 100, 150, 300, 500, 1000, 3000}, 0, 3001, true, 0},				// HASH_1_3000       :2016
	{11, {9, 19, 29, 39, 49, 59, 69, 79, 89, 99, 100}, 0, 101, false, 10},					// PERCENT_HASH      :1624
};

// get_bucket_max_threshold<HashClass, T>, gy_statistics.h:500-515
int64_t bucket_max_threshold(const ClsDesc &d, bool t_is_int, size_t id)
{
	const size_t maxb = (size_t)d.nthr + 2;

	if (id == 0) return d.minv - 1;
	if (id >= maxb - 1) {
		const int64_t maxt = t_is_int ? INT32_MAX : INT64_MAX;
		const int64_t lesst = d.maxv >= INT32_MAX ? INT64_MAX : (d.maxv > (INT16_MAX >> 1) ? INT32_MAX : INT16_MAX);
		return std::min(lesst, maxt);
	}
	return d.thr[id - 1];
}

void hist_from_cells(const HistCell *cells, int nb, gysk_hist_serial *out, uint64_t *total, int64_t *maxv, bool t_is_int)
{
	uint64_t t = 0;

	for (int i = 0; i < GYSK_HIST_MAX_BUCKETS; ++i) {
		if (i < nb) { out[i].count = cells[i].count; out[i].sum = cells[i].sum; t += cells[i].count; }
		else { out[i].count = 0; out[i].sum = 0; }
	}
	*total = t;			// total_count_ always equals the sum of the bucket counts (add_data bumps both)
	int64_t m = cells[HIST_MAX_CELL].sum;
	if (t_is_int && m == INT64_MIN) m = INT32_MIN;		// numeric_limits<int>::min() for GY_HISTOGRAM<int, ...>
	*maxv = m;
}

double td_quantile(const double *means, const uint64_t *weights, uint32_t n, double minv, double maxv, double q)
{
	if (!n) return NAN;
	double total = 0;
	for (uint32_t i = 0; i < n; ++i) total += (double)weights[i];
	if (q <= 0) return minv;
	if (q >= 1) return maxv;

	const double target = q * total;
	double cum = 0, prev_center = 0, prev_mean = minv;

	for (uint32_t i = 0; i < n; ++i) {
		const double center = cum + (double)weights[i] / 2.0;
		if (target < center) {
			const double span = center - prev_center;
			return span > 0 ? prev_mean + (means[i] - prev_mean) * ((target - prev_center) / span) : means[i];
		}
		prev_center = center; prev_mean = means[i];
		cum += (double)weights[i];
	}
	const double span = total - prev_center;
	return span > 0 ? prev_mean + (maxv - prev_mean) * ((target - prev_center) / span) : maxv;
}

double hll_estimate_from_hist(const uint32_t *hist64, uint32_t p)
{
	const uint32_t m = 1u << p;
	double sum = 0, alpha, est;

	for (int r = 63; r >= 0; --r) sum += (double)hist64[r] * ldexp(1.0, -r);
	if (m == 16) alpha = 0.673; else if (m == 32) alpha = 0.697; else if (m == 64) alpha = 0.709;
	else alpha = 0.7213 / (1.0 + 1.079 / (double)m);
	est = alpha * (double)m * (double)m / sum;
	if (est <= 2.5 * (double)m && hist64[0]) est = (double)m * log((double)m / (double)hist64[0]);
	return est;
}

// width of one ring slot per level: Level_5s_5min_5days_all durations {300 s, 432000 s} / 10 slots (gy_statistics.h:1548, :1105)
const uint32_t g_level_width[NLEVELS] = { 30, 43200 };

uint32_t live_mask(const gysk_engine *e, int l)
{
	const uint64_t now_epoch = e->last_flush_tsec / g_level_width[l];
	uint32_t m = 0;

	for (int k = 0; k < NSLOTS; ++k) {
		const uint64_t ep = e->ring_epoch[l][k];
		if (ep != ~0ull && ep + NSLOTS > now_epoch && ep <= now_epoch) m |= 1u << k;
	}
	return m;
}

int gather_svcs(gysk_engine *e, const uint64_t *ids, uint32_t n)		// n <= QCHUNK; results in e->h_svcraw
{
	memcpy(e->h_qids, ids, (size_t)n * sizeof(uint64_t));
	CU(e, cudaMemcpyAsync(e->d_qids, e->h_qids, (size_t)n * sizeof(uint64_t), cudaMemcpyHostToDevice, e->stream));
	e->kernel_launches += launch_gather_svcs(e->st, e->d_qids, n, e->cfg.max_svcs, live_mask(e, 0), live_mask(e, 1), e->d_svcraw, e->stream);
	CU(e, cudaMemcpyAsync(e->h_svcraw, e->d_svcraw, (size_t)n * sizeof(SvcRaw), cudaMemcpyDeviceToHost, e->stream));
	CU(e, cudaStreamSynchronize(e->stream));
	return post_launch(e, "gather_svcs");
}

} // namespace

// SvcRaw (device gather) -> the fields SvcStateFields exposes (server/gy_mfields.h:1383-1412) + the sketch answers
void gysk::summarize_raw(const gysk_engine *e, const SvcRaw &r, uint64_t id, gysk_svc_summary &o)
{
	const float pcts[3] = {95.0f, 99.0f, 25.0f};		// the percentiles listener_stats_update reads, gy_socket_stat.h:459

	memset(&o, 0, sizeof(o));
	o.glob_id = id;
	o.found = r.found;
	o.td_p50_us = o.td_p95_us = o.td_p99_us = NAN;
	if (!r.found) return;

	gysk_hist_serial ser[GYSK_HIST_MAX_BUCKETS];
	uint64_t total; int64_t maxv, p[3];

	hist_from_cells(r.last, 15, ser, &total, &maxv, false);
	gysk_hist_percentiles(GYSK_CLS_RESP_TIME, 0, ser, total, pcts, 3, p);
	o.nqrys_5s = (uint32_t)total;
	for (int b = 0; b < 15; ++b) o.total_resp_5sec += (uint64_t)ser[b].sum;
	o.p95_5s_resp_ms = p[0]; o.p99_5s_resp_ms = p[1]; o.p25_5s_resp_ms = p[2];

	hist_from_cells(r.lvl[0], 15, ser, &total, &maxv, false);
	gysk_hist_percentiles(GYSK_CLS_RESP_TIME, 0, ser, total, pcts, 2, p);
	o.p95_5min_resp_ms = p[0]; o.p99_5min_resp_ms = p[1]; o.nqrys_5min = total;
	hist_from_cells(r.lvl[1], 15, ser, &total, &maxv, false);
	gysk_hist_percentiles(GYSK_CLS_RESP_TIME, 0, ser, total, pcts, 1, p);
	o.p95_5day_resp_ms = p[0]; o.nqrys_5day = total;

	hist_from_cells(r.all, 15, ser, &total, &maxv, false);
	gysk_hist_percentiles(GYSK_CLS_RESP_TIME, 0, ser, total, pcts, 2, p);
	o.p95_all_resp_ms = p[0]; o.p99_all_resp_ms = p[1]; o.nqrys_all = total; o.max_resp_ms = maxv;

	o.nconns_5s = (uint32_t)r.conn_last; o.kbytes_5s = (uint32_t)(r.conn_last >> 32);
	o.nconns_all = r.conn_all_cnt; o.kbytes_all = r.conn_all_kb;
	o.distinct_clients = hll_estimate_from_hist(r.hll_hist, e->cfg.hll_p);
	o.nconns_active = (uint32_t)r.aux.act_last; o.active_kbytes = (uint32_t)(r.aux.act_last >> 32);
	memcpy(&o.max_rtt_msec, &r.aux.rtt_last, 4);
	o.cli_errors = (uint32_t)r.aux.err_last; o.ser_errors = (uint32_t)(r.aux.err_last >> 32);
	o.curr_state = r.sst.state; o.curr_issue = r.sst.issue; o.issue_bit_hist = r.sst.issue_bits; o.high_resp_bit_hist = r.sst.high_bits;

	double means[TD_CAP]; uint64_t w[TD_CAP];
	const uint32_t nc = std::min<uint32_t>(r.td.n, TD_CAP);
	for (uint32_t c = 0; c < nc; ++c) { means[c] = r.cent[c].mean; w[c] = r.cent[c].weight; }
	o.td_count = r.td.total;
	if (nc) {
		o.td_p50_us = td_quantile(means, w, nc, r.td.minv, r.td.maxv, 0.50);
		o.td_p95_us = td_quantile(means, w, nc, r.td.minv, r.td.maxv, 0.95);
		o.td_p99_us = td_quantile(means, w, nc, r.td.minv, r.td.maxv, 0.99);
	}
}

// ============================================================================================================
// C ABI
// ============================================================================================================
extern "C" {

int gysk_abi_version(void) { return GYSK_ABI_VERSION; }

void gysk_config_default(gysk_config *cfg)
{
	if (!cfg) return;
	memset(cfg, 0, sizeof(*cfg));
	cfg->struct_size = sizeof(*cfg);
	cfg->device = 0;
	cfg->max_svcs = 1u << 17;
	cfg->max_tasks = 1u << 16;
	cfg->cms_depth = 4;
	cfg->cms_log2_width = 20;
	cfg->hll_p = 12;
	cfg->td_compression = 200;
	cfg->max_batch = 1u << 22;
	cfg->stage_batch = 0;
	cfg->flags = GYSK_FLAG_AUTO_REGISTER;
	cfg->rank = 0; cfg->world = 1;
}

const char *gysk_last_error(gysk_engine *e)
{
	return e ? e->err.c_str() : g_create_error.c_str();
}

void gysk_destroy(gysk_engine *e)
{
	if (!e) return;
	{ std::lock_guard<std::mutex> lk(g_live_mtx); g_live.erase(std::remove(g_live.begin(), g_live.end(), e->uid), g_live.end()); }
	cudaSetDevice(e->dev);
	if (e->stream) cudaStreamSynchronize(e->stream);
	if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);
	if (e->side_stream) cudaStreamSynchronize(e->side_stream);
	merge_release(e);
	for (int k = 0; k < NBUF; ++k) {
		if (e->ev_copied[k]) cudaEventDestroy(e->ev_copied[k]);
		if (e->ev_done[k]) cudaEventDestroy(e->ev_done[k]);
		if (e->ev_raw_copied[k]) cudaEventDestroy(e->ev_raw_copied[k]);
		if (e->ev_raw_done[k]) cudaEventDestroy(e->ev_raw_done[k]);
	}
	for (auto &ts : e->tstages) {
		for (int i = 0; i < 2; ++i) { if (ts->buf[i]) cudaFreeHost(ts->buf[i]); if (ts->copied[i]) cudaEventDestroy(ts->copied[i]); }
	}
	for (cudaEvent_t ev : e->prof_events) cudaEventDestroy(ev);
	if (e->ev_evict) cudaEventDestroy(e->ev_evict);
	for (void *p : e->dallocs) cudaFree(p);
	for (void *p : e->hallocs) cudaFreeHost(p);
	if (e->stream) cudaStreamDestroy(e->stream);
	if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
	if (e->side_stream) cudaStreamDestroy(e->side_stream);
	if (e->ev_ingested) cudaEventDestroy(e->ev_ingested);
	if (e->ev_side_done) cudaEventDestroy(e->ev_side_done);
	delete e;
}

int gysk_create(const gysk_config *ucfg, gysk_engine **out)
{
	if (!out) return GYSK_ERR_INVAL;
	*out = nullptr;

	gysk_config cfg;
	gysk_config_default(&cfg);
	if (ucfg) {
		if (ucfg->struct_size != sizeof(gysk_config)) return fail(nullptr, GYSK_ERR_INVAL, "gysk_config.struct_size mismatch");
		cfg = *ucfg;
	}
	if (!cfg.world) cfg.world = 1;
	if (!cfg.stage_batch || cfg.stage_batch > cfg.max_batch) cfg.stage_batch = std::min<uint32_t>(cfg.max_batch, 1u << 22);
	if (cfg.max_svcs < 1 || cfg.max_svcs > (1u << 24) || cfg.max_tasks < 1 || cfg.max_tasks > (1u << 24) || cfg.cms_depth < 1 ||
			cfg.cms_depth > 8 || cfg.cms_log2_width < 4 || cfg.cms_log2_width > 28 || cfg.hll_p < 4 || cfg.hll_p > 16 ||
			cfg.td_compression < 10 || cfg.td_compression > (uint32_t)TD_CAP || cfg.max_batch < 1024 || cfg.max_batch >= (1u << 27) ||
			cfg.rank >= cfg.world)
		return fail(nullptr, GYSK_ERR_INVAL, "gysk_config out of range");

	int ndev = 0;
	cudaError_t ce = cudaGetDeviceCount(&ndev);
	if (ce != cudaSuccess || ndev <= 0 || cfg.device < 0 || cfg.device >= ndev)
		return fail(nullptr, GYSK_ERR_NODEV, "no usable CUDA device (libgysketch has no CPU fallback)", ce);

	cudaDeviceProp prop;
	if ((ce = cudaGetDeviceProperties(&prop, cfg.device)) != cudaSuccess) return fail(nullptr, GYSK_ERR_NODEV, "cudaGetDeviceProperties", ce);
	if (prop.major != 10) return fail(nullptr, GYSK_ERR_NODEV, "device is not sm_100 (kernels are built for sm_100a only)");

	gysk_engine *e = new (std::nothrow) gysk_engine;
	if (!e) return fail(nullptr, GYSK_ERR_NOMEM, "new gysk_engine");
	e->cfg = cfg; e->dev = cfg.device; e->uid = g_engine_uid.fetch_add(1);
	{ std::lock_guard<std::mutex> lk(g_live_mtx); g_live.push_back(e->uid); }
	memset(e->ring_epoch, 0xFF, sizeof(e->ring_epoch));

	int rc = 0;
	auto bail = [&](int code) { g_create_error = e->err; gysk_destroy(e); return code; };

	if ((ce = cudaSetDevice(e->dev)) != cudaSuccess) { fail(e, GYSK_ERR_CUDA, "cudaSetDevice", ce); return bail(GYSK_ERR_CUDA); }
	if ((ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)) != cudaSuccess ||
			(ce = cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
			(ce = cudaStreamCreateWithFlags(&e->side_stream, cudaStreamNonBlocking)) != cudaSuccess ||
			(ce = cudaEventCreateWithFlags(&e->ev_ingested, cudaEventDisableTiming)) != cudaSuccess ||
			(ce = cudaEventCreateWithFlags(&e->ev_side_done, cudaEventDisableTiming)) != cudaSuccess) {
		fail(e, GYSK_ERR_CUDA, "cudaStreamCreate", ce); return bail(GYSK_ERR_CUDA);
	}

	DevState &st = e->st;
	const size_t ns = (size_t)cfg.max_svcs + 1, nt = cfg.max_tasks;		// slot max_svcs = the null slot: never handed out, always pristine
	const uint32_t scap = pow2_at_least((uint64_t)ns * 2), tcap = pow2_at_least((uint64_t)nt * 2);

#define A(call) do { if ((rc = (call)) != 0) return bail(rc); } while (0)
	A(dalloc(e, &st.counters, (size_t)CTR_MAX));
	A(dalloc(e, &st.svc_tbl.ent, scap)); st.svc_tbl.mask = scap - 1; st.svc_tbl.max_slots = cfg.max_svcs;
	A(dalloc(e, &st.svc_tbl.count, 1));
	A(dalloc(e, &st.slot_id, ns)); A(dalloc(e, &st.slot_host, ns));
	st.svc_tbl.slot_id = st.slot_id; st.svc_tbl.slot_host = st.slot_host;
	A(dalloc(e, &st.slot_first_seen, ns)); A(dalloc(e, &st.slot_last_active, ns));
	A(dalloc(e, &st.evict_list, ns)); A(dalloc(e, &st.evict_ids, ns));
	A(dalloc(e, &st.svc_tbl.free_n, 1)); A(dalloc(e, &st.svc_tbl.free_slots, ns));
	A(halloc(e, &e->h_evict, ns + 2));
	e->h_evict[0] = 0; e->h_evict[cfg.max_svcs + 1] = 0;
	if ((ce = cudaEventCreateWithFlags(&e->ev_evict, cudaEventDisableTiming)) != cudaSuccess) { fail(e, GYSK_ERR_CUDA, "cudaEventCreate", ce); return bail(GYSK_ERR_CUDA); }
	A(dalloc(e, &st.task_tbl.ent, tcap)); st.task_tbl.mask = tcap - 1; st.task_tbl.max_slots = cfg.max_tasks;
	A(dalloc(e, &st.task_tbl.count, 1));
	A(dalloc(e, &st.hist_cur, ns * HIST_CELLS)); A(dalloc(e, &st.hist_last, ns * HIST_CELLS)); A(dalloc(e, &st.hist_all, ns * HIST_CELLS));
	A(dalloc(e, &st.hist_ring, (size_t)NLEVELS * NSLOTS * cfg.max_svcs * HIST_CELLS));
	A(dalloc(e, &st.conn_cur, ns)); A(dalloc(e, &st.conn_last, ns)); A(dalloc(e, &st.conn_all_cnt, ns)); A(dalloc(e, &st.conn_all_kb, ns));
	A(dalloc(e, &st.bm_cur, ns * HIST_CELLS)); A(dalloc(e, &st.bm_last, ns * HIST_CELLS));
	A(dalloc(e, &st.hll, ns << cfg.hll_p));
	A(dalloc(e, &st.td_cent, ns * TD_CAP)); A(dalloc(e, &st.td_head, ns));
	A(dalloc(e, &st.task_hist, nt * 3 * HIST_CELLS));
	A(dalloc(e, &st.task_prev, nt * 3)); A(dalloc(e, &st.task_last, nt * 3));
	A(dalloc(e, &st.task_slot_id, nt)); A(dalloc(e, &st.task_slot_host, nt));
	st.task_tbl.slot_id = st.task_slot_id; st.task_tbl.slot_host = st.task_slot_host;
	A(dalloc(e, &st.cms_cur, (size_t)cfg.cms_depth << cfg.cms_log2_width)); A(dalloc(e, &st.cms_last, (size_t)cfg.cms_depth << cfg.cms_log2_width));
	st.cms_depth = cfg.cms_depth; st.cms_log2w = cfg.cms_log2_width; st.cms_wmask = (1u << cfg.cms_log2_width) - 1; st.hll_p = cfg.hll_p;
	st.rank = cfg.rank; st.world = cfg.world; st.auto_register = (cfg.flags & GYSK_FLAG_AUTO_REGISTER) ? 1 : 0;
	st.td_delta = (double)cfg.td_compression;
	{
		// the unit grid of the K_1 scale: q_j = (sin(pi (j/delta - 1/2)) + 1)/2 — libm on the host, the same expression as the oracle
		std::vector<double> qtab(cfg.td_compression + 1);
		for (uint32_t j = 0; j <= cfg.td_compression; ++j) qtab[j] = 0.5 * (sin(M_PI * ((double)j / (double)cfg.td_compression - 0.5)) + 1.0);
		qtab[0] = 0.0; qtab[cfg.td_compression] = 1.0;
		double *d_q = nullptr;
		A(dalloc(e, &d_q, qtab.size(), false));
		if ((ce = cudaMemcpy(d_q, qtab.data(), qtab.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) { fail(e, GYSK_ERR_CUDA, "qtab", ce); return bail(GYSK_ERR_CUDA); }
		st.td.qtab = d_q; st.td.delta = cfg.td_compression; st.td.pad = 0;
	}
	A(dalloc(e, &st.slot_batch, ns)); A(dalloc(e, &st.slot_aux, ns));
	{
		// dense value bins of the hot services: GYSK_HOT_ROWS rows of 16 KB (default 2048, 0 switches the path off); a
		// service turns hot with GYSK_HOT_MIN (default 4096) .. GYSK_HOT_MAX samples in one batch unless its fullest bin holds more
		// than GYSK_HOT_BIN_MAX (default 131072) of them. Routing only: results do not depend on any of these.
		auto envl = [](const char *name, long dflt) { const char *v = getenv(name); return v ? atol(v) : dflt; };
		long rows = envl("GYSK_HOT_ROWS", 2048);
		if (rows < 0) rows = 0;
		if (rows > 65536) rows = 65536;
		if ((size_t)rows > ns) rows = (long)ns;
		st.hot_cap = (uint32_t)rows;
		st.hot_min = (uint32_t)std::max(1l, envl("GYSK_HOT_MIN", 4096));
		st.hot_max = (uint32_t)std::min<long>(0xFFFFFFFFl, std::max(1l, envl("GYSK_HOT_MAX", 0xFFFFFFFFl)));
		st.hot_bin_max = (uint32_t)std::min<long>(0xFFFFFFFFl, std::max(1l, envl("GYSK_HOT_BIN_MAX", 131072)));
		st.hot_rows = nullptr; st.hot_slot = nullptr;
		if (rows) { A(dalloc(e, &st.hot_rows, (size_t)rows * HOT_ROW_WORDS)); A(dalloc(e, &st.hot_slot, (size_t)rows)); }
	}
	A(dalloc(e, &st.qps_hist, ns * HIST_CELLS)); A(dalloc(e, &st.act_hist, ns * HIST_CELLS)); A(dalloc(e, &st.slot_state, ns));
	SortTemp &tmp = e->tmp;
	const size_t nsort = std::max<size_t>(std::max<size_t>(ns, nt) + 1, cfg.max_batch);	// RESP keys of a batch; the top-N sorts rank services / tasks
	tmp.max_tiles = (uint32_t)((nsort + SORT_TILE - 1) / SORT_TILE);
	A(dalloc(e, &tmp.keys_a, nsort, false)); A(dalloc(e, &tmp.keys_b, nsort, false));
	A(dalloc(e, &tmp.tile_status, (size_t)RADIX_MAX * tmp.max_tiles));
	tmp.epoch = &e->sort_epoch;
	A(dalloc(e, &tmp.os_ghist, (size_t)8 * RADIX_MAX + 8));
	A(dalloc(e, &tmp.touched, ns));
	{
		const size_t nmw = (size_t)TD_MERGE_MAX_SMS * TD_MERGE_CTAS_PER_SM * 4;		// warps of bins_merge_kernel
		// a batch cannot have more runs than keys, nor than the engine has bins
		const size_t pool_cap = std::min<size_t>((size_t)cfg.max_batch, (size_t)cfg.max_svcs * NBINS) + 8;
		A(dalloc(e, &tmp.pool, pool_cap, false)); A(dalloc(e, &tmp.run_bin, pool_cap, false));
		A(dalloc(e, &tmp.chunk_run, ((size_t)cfg.max_batch >> 7) + 16, false)); A(dalloc(e, &tmp.segs, ns));
		A(dalloc(e, &tmp.items_scratch, nmw * NBINS, false)); A(dalloc(e, &tmp.big_scratch, nmw, false));
		if (side_drain_enabled()) {		// the experiment's record queue: one buffer, filled from both ends
			A(dalloc(e, &tmp.tcpq, (size_t)cfg.max_batch + 64, false)); tmp.taskq = tmp.tcpq + ((size_t)cfg.max_batch + 63);
		}
	}
	st.svc_tbl.insert_fail = st.counters + CTR_INSERT_FAIL; st.task_tbl.insert_fail = nullptr;

	for (int k = 0; k < NBUF; ++k) {
		A(dalloc(e, &e->d_events[k], (size_t)cfg.stage_batch, false));
		e->raw_bytes = (size_t)cfg.stage_batch * 32;				// a raw piece never expands into more than one event buffer
		A(dalloc(e, &e->d_raw[k], e->raw_bytes, false));
		if ((ce = cudaEventCreateWithFlags(&e->ev_copied[k], cudaEventDisableTiming)) != cudaSuccess ||
				(ce = cudaEventCreateWithFlags(&e->ev_raw_copied[k], cudaEventDisableTiming)) != cudaSuccess ||
				(ce = cudaEventCreateWithFlags(&e->ev_raw_done[k], cudaEventDisableTiming)) != cudaSuccess ||
				(ce = cudaEventCreateWithFlags(&e->ev_done[k], cudaEventDisableTiming)) != cudaSuccess) {
			fail(e, GYSK_ERR_CUDA, "cudaEventCreate", ce); return bail(GYSK_ERR_CUDA);
		}
	}
	A(dalloc(e, &e->d_qids, (size_t)QCHUNK)); A(halloc(e, &e->h_qids, (size_t)QCHUNK));
	A(dalloc(e, &e->d_svcraw, (size_t)QCHUNK)); A(halloc(e, &e->h_svcraw, (size_t)QCHUNK));
	A(dalloc(e, &e->d_taskraw, (size_t)QCHUNK)); A(halloc(e, &e->h_taskraw, (size_t)QCHUNK));
	A(dalloc(e, &e->d_hllout, (size_t)1 << cfg.hll_p)); A(halloc(e, &e->h_hllout, (size_t)1 << cfg.hll_p));
	A(dalloc(e, &e->d_found, (size_t)1)); A(halloc(e, &e->h_found, (size_t)1));
	A(dalloc(e, &e->d_flowout, (size_t)QCHUNK)); A(halloc(e, &e->h_flowout, (size_t)QCHUNK));
	A(halloc(e, &e->h_counters, (size_t)CTR_MAX + 2));
#undef A

	e->kernel_launches += launch_init_state(st, cfg.max_svcs + 1, cfg.max_tasks, e->stream);
	if ((ce = cudaStreamSynchronize(e->stream)) != cudaSuccess || (ce = cudaGetLastError()) != cudaSuccess) {
		fail(e, GYSK_ERR_CUDA, "engine init", ce); return bail(GYSK_ERR_CUDA);
	}
	*out = e;
	return GYSK_OK;
}

void *gysk_stream(gysk_engine *e) { return e ? (void *)e->stream : nullptr; }

int gysk_profile_enable(gysk_engine *e, int on)
{
	CHECK_ENGINE(e);
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	int rc = sync_locked(e);
	if (rc) return rc;
	e->profiling = !!on;
	e->prof_used = 0;
	return GYSK_OK;
}

int gysk_profile_read(gysk_engine *e, double *ms_ingest, double *ms_tdigest, uint64_t *nbatches)
{
	CHECK_ENGINE(e);
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	int rc = sync_locked(e);
	if (rc) return rc;
	double a = 0, b = 0;
	for (size_t i = 0; i + 3 <= e->prof_used; i += 3) {
		float t1 = 0, t2 = 0;
		CU(e, cudaEventElapsedTime(&t1, e->prof_events[i], e->prof_events[i + 1]));
		CU(e, cudaEventElapsedTime(&t2, e->prof_events[i + 1], e->prof_events[i + 2]));
		a += t1; b += t2;
	}
	if (ms_ingest) *ms_ingest = a;
	if (ms_tdigest) *ms_tdigest = b;
	if (nbatches) *nbatches = e->prof_used / 3;
	e->prof_used = 0;
	return GYSK_OK;
}

int gysk_get_stats(gysk_engine *e, gysk_stats *out)
{
	CHECK_ENGINE(e);
	if (!out) return GYSK_ERR_INVAL;
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	int rc = sync_locked(e);
	if (rc) return rc;
	CU(e, cudaMemcpyAsync(e->h_counters, e->st.counters, sizeof(unsigned long long) * CTR_MAX, cudaMemcpyDeviceToHost, e->stream));
	CU(e, cudaMemcpyAsync(e->h_counters + CTR_MAX, e->st.svc_tbl.count, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
	CU(e, cudaMemcpyAsync(e->h_counters + CTR_MAX + 1, e->st.task_tbl.count, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
	CU(e, cudaStreamSynchronize(e->stream));
	memset(out, 0, sizeof(*out));
	out->events_in = e->h_counters[CTR_IN]; out->events_dropped = e->h_counters[CTR_DROPPED];
	out->events_resp = e->h_counters[CTR_RESP]; out->events_tcp = e->h_counters[CTR_TCP]; out->events_task = e->h_counters[CTR_TASK];
	if ((rc = collect_evicted(e, true))) return rc;
	int32_t nfree = 0;
	CU(e, cudaMemcpy(&nfree, e->st.svc_tbl.free_n, sizeof(nfree), cudaMemcpyDeviceToHost));
	out->nsvcs = std::min<uint64_t>((uint32_t)e->h_counters[CTR_MAX], e->cfg.max_svcs) - (uint64_t)std::max(nfree, 0);
	out->svcs_evicted = e->evicted_total;
	out->ntasks = std::min<uint64_t>((uint32_t)e->h_counters[CTR_MAX + 1], e->cfg.max_tasks);
	out->batches = e->batches; out->kernel_launches = e->kernel_launches;
	out->wire_msgs_ok = e->wire_ok; out->wire_msgs_bad = e->wire_bad;
	return GYSK_OK;
}

// diagnostic: rows of dense value bins handed out to hot services so far (results never depend on it)
int64_t gysk_hot_rows_in_use(gysk_engine *e)
{
	if (!e) return GYSK_ERR_INVAL;
	GYSK_ENTER(e);
	if (cudaSetDevice(e->dev) != cudaSuccess) return GYSK_ERR_CUDA;
	int rc = sync_locked(e);
	if (rc) return rc;
	if (!e->st.hot_rows) return 0;
	unsigned long long n = 0;
	if (cudaMemcpy(&n, e->st.counters + CTR_NHOT_NEXT, sizeof(n), cudaMemcpyDeviceToHost) != cudaSuccess) return GYSK_ERR_CUDA;
	return (int64_t)n;
}

// introspection (no device needed): word of value bin `bin` inside a hot row's {samples | remainders} half; the usec sum of the bin
// lies GYSK_HOT_ROW_BINS words further on. ~0 for a bin the engine does not have.
uint32_t gysk_hot_row_word(uint32_t bin)
{
	return bin < (uint32_t)NBINS ? hot_word(bin) : 0xFFFFFFFFu;
}

// diagnostic: response samples of the last device batch that travelled as sort keys (the others went to hot rows)
int64_t gysk_last_batch_keys(gysk_engine *e)
{
	if (!e) return GYSK_ERR_INVAL;
	GYSK_ENTER(e);
	if (cudaSetDevice(e->dev) != cudaSuccess) return GYSK_ERR_CUDA;
	int rc = sync_locked(e);
	if (rc) return rc;
	unsigned long long n = 0;
	if (cudaMemcpy(&n, e->st.counters + CTR_NKEYS, sizeof(n), cudaMemcpyDeviceToHost) != cudaSuccess) return GYSK_ERR_CUDA;
	return (int64_t)n;
}

int gysk_register_ids(gysk_engine *e, const uint64_t *ids, uint32_t n, int is_task)
{
	CHECK_ENGINE(e);
	if (!ids && n) return GYSK_ERR_INVAL;
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	for (uint32_t off = 0; off < n; off += QCHUNK) {
		const uint32_t m = std::min(QCHUNK, n - off);
		memcpy(e->h_qids, ids + off, (size_t)m * sizeof(uint64_t));
		CU(e, cudaMemcpyAsync(e->d_qids, e->h_qids, (size_t)m * sizeof(uint64_t), cudaMemcpyHostToDevice, e->stream));
		e->kernel_launches += launch_register(e->st, e->d_qids, m, is_task, e->stream);
		CU(e, cudaStreamSynchronize(e->stream));
	}
	return post_launch(e, "register");
}

// ---- ingest -----------------------------------------------------------------------------------------------

int gysk_ingest_device(gysk_engine *e, const gysk_event *d_events, uint64_t n)
{
	CHECK_ENGINE(e);
	if (!d_events && n) return GYSK_ERR_INVAL;
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	int rc = submit_stage(e);					// keep arrival order
	if (rc) return rc;
	for (uint64_t off = 0; off < n; off += e->cfg.max_batch) {
		const uint64_t m = std::min<uint64_t>(e->cfg.max_batch, n - off);
		if ((rc = process_device_batch(e, d_events + off, m))) return rc;
	}
	return GYSK_OK;
}

int gysk_ingest_pinned(gysk_engine *e, const gysk_event *pinned, uint64_t n)
{
	CHECK_ENGINE(e);
	if (!pinned && n) return GYSK_ERR_INVAL;
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	// zero copy on the host: the H2D copies read the caller's page-locked buffer directly, chunk by chunk into the two device
	// event buffers; the copy of chunk c+1 is enqueued as soon as the ingest kernel of chunk c is, so the link never idles
	return append_chunk(e, pinned, n);
}

} // extern "C"

// ---- raw records: decoded into the canonical 32-byte event — by the same inline functions on the host (a few records: the
// calling thread's stage) and on the device (bulk: the raw bytes cross the link, a kernel expands them) ----------
namespace gysk {

__host__ __device__ inline uint32_t bswap16(uint32_t v) { return ((v & 0xFFu) << 8) | ((v >> 8) & 0xFFu); }

// listener id of a raw eBPF record: the reference keys listeners by NS_IP_PORT {ip, port, netns} (gy_socket_stat.cc:1529) and
// derives glob_id_ with CityHash (:1824); the id is opaque to the engine, so any deterministic 64-bit fold of the same triple
// serves: two lookup2 words. IPv6 addresses are folded to 32 bits first.
__host__ __device__ inline uint64_t raw_svc_id(uint32_t ip, uint32_t netns, uint32_t port)
{
	uint64_t id = ((uint64_t)jhash_2words(ip, netns, GY_SEED) << 32) | jhash_2words(port, netns, GY_SEED ^ ip);
	return id ? id : 1;
}
__host__ __device__ inline uint32_t fold_ip6(const uint32_t w[4]) { return jhash_2words(w[2], w[3], jhash_2words(w[0], w[1], GY_SEED)); }

__host__ __device__ inline void ev_pad(gysk_event &o) { o.svc_id = 0; o.flow_key = 0; o.value = 0; o.host_idx = 0; o.tsec = 0; o.type = 0xFFFF; o.flags = 0; }

// TCP_SOCK_HANDLER::handle_ipv4_resp_event / handle_ipv6_resp_event, common/gy_socket_stat.cc:1517-1552: tresp = lsndtime - lrcvtime
// (msec), dropped when (uint32_t)tresp > 1 000 000; ports arrive in network byte order (ntohs :1526-1527); the client port keys
// CONN_BITMAP (gy_socket_stat.h:403-410)
__host__ __device__ inline void decode_resp(uint32_t sip, uint32_t cip, uint32_t netns, uint32_t sport_be, uint32_t dport_be, uint32_t lsnd, uint32_t lrcv,
		uint32_t host_idx, gysk_event &o)
{
	const uint32_t tresp = lsnd - lrcv;
	if (tresp > 1000000u) { ev_pad(o); return; }
	const uint32_t sport = bswap16(sport_be), dport = bswap16(dport_be);
	o.svc_id = raw_svc_id(sip, netns, sport);
	o.flow_key = ((uint64_t)cip << 32) | dport;
	o.value = tresp * 1000u; o.host_idx = host_idx; o.tsec = 0; o.type = GYSK_EV_RESP; o.flags = 0;
}

// TCP_SOCK_HANDLER::handle_ipv4_conn_event / handle_ipv6_conn_event, common/gy_socket_stat.cc:241-294: type 1..4; on close the
// byte counters are added to the listener totals (handle_bpf_close_ser :850)
__host__ __device__ inline void decode_conn(uint32_t saddr, uint32_t daddr, uint32_t netns, uint32_t sport_be, uint32_t dport_be, uint32_t type,
		uint64_t bytes, uint64_t ts_ns, uint32_t host_idx, gysk_event &o)
{
	if (type < GYSK_EV_CONNECT || type > GYSK_EV_CLOSE_SER) { ev_pad(o); return; }
	const bool ser_side = (type == GYSK_EV_ACCEPT || type == GYSK_EV_CLOSE_SER);
	const uint32_t hs = bswap16(sport_be), hd = bswap16(dport_be);
	const uint32_t sip = ser_side ? saddr : daddr, sport = ser_side ? hs : hd;
	const uint32_t cip = ser_side ? daddr : saddr, cport = ser_side ? hd : hs;
	o.svc_id = raw_svc_id(sip, netns, sport);
	o.flow_key = ((uint64_t)cip << 32) | cport;
	o.value = bytes > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)bytes;
	o.host_idx = host_idx; o.tsec = (uint32_t)(ts_ns / 1000000000ull); o.type = (uint16_t)type; o.flags = 0;
}

__host__ __device__ inline uint32_t raw_stride(uint32_t kind)
{
	switch (kind) {
	case GYSK_RAW_EVENT32 : return 32;
	case GYSK_RAW_TCP_IPV4_EVENT : return sizeof(wire::tcp_ipv4_event_t);
	case GYSK_RAW_TCP_IPV4_RESP : return sizeof(wire::tcp_ipv4_resp_event_t);
	case GYSK_RAW_TCP_IPV6_EVENT : return sizeof(wire::tcp_ipv6_event_t);
	case GYSK_RAW_TCP_IPV6_RESP : return sizeof(wire::tcp_ipv6_resp_event_t);
	case GYSK_RAW_RESP16 : return sizeof(gysk_resp16);
	case GYSK_RAW_TCP24 : return sizeof(gysk_tcp24);
	case GYSK_RAW_TASK24 : return sizeof(gysk_task24);
	default : return 0;
	}
}

__host__ __device__ inline void decode_raw(uint32_t kind, const void *rec, uint32_t host_idx, gysk_event &o)
{
	switch (kind) {
	case GYSK_RAW_TCP_IPV4_RESP : {
		const wire::tcp_ipv4_resp_event_t &p = *static_cast<const wire::tcp_ipv4_resp_event_t *>(rec);
		decode_resp(p.saddr, p.daddr, p.netns, p.sport, p.dport, p.lsndtime, p.lrcvtime, host_idx, o);
		break;
	}
	case GYSK_RAW_TCP_IPV6_RESP : {
		const wire::tcp_ipv6_resp_event_t &p = *static_cast<const wire::tcp_ipv6_resp_event_t *>(rec);
		decode_resp(fold_ip6(p.saddr), fold_ip6(p.daddr), p.netns, p.sport, p.dport, p.lsndtime, p.lrcvtime, host_idx, o);
		break;
	}
	case GYSK_RAW_TCP_IPV4_EVENT : {
		const wire::tcp_ipv4_event_t &p = *static_cast<const wire::tcp_ipv4_event_t *>(rec);
		decode_conn(p.saddr, p.daddr, p.netns, p.sport, p.dport, p.type, p.bytes_received + p.bytes_acked, p.ts_ns, host_idx, o);
		break;
	}
	case GYSK_RAW_TCP_IPV6_EVENT : {
		const wire::tcp_ipv6_event_t &p = *static_cast<const wire::tcp_ipv6_event_t *>(rec);
		decode_conn(fold_ip6(p.saddr), fold_ip6(p.daddr), p.netns, p.sport, p.dport, p.type, p.bytes_received + p.bytes_acked, p.ts_ns, host_idx, o);
		break;
	}
	case GYSK_RAW_RESP16 : {
		const gysk_resp16 &p = *static_cast<const gysk_resp16 *>(rec);
		o.svc_id = p.svc_id; o.flow_key = p.cli_port; o.value = p.usec; o.host_idx = p.host_idx; o.tsec = 0; o.type = GYSK_EV_RESP; o.flags = p.flags;
		break;
	}
	case GYSK_RAW_TCP24 : {
		const gysk_tcp24 &p = *static_cast<const gysk_tcp24 *>(rec);
		o.svc_id = p.svc_id; o.flow_key = p.flow_key; o.value = p.bytes; o.host_idx = p.host_idx; o.tsec = 0; o.type = p.type; o.flags = 0;
		break;
	}
	case GYSK_RAW_TASK24 : {
		const gysk_task24 &p = *static_cast<const gysk_task24 *>(rec);
		o.svc_id = p.aggr_task_id; o.flow_key = (uint64_t)p.cpu_delay_msec | ((uint64_t)p.blkio_delay_msec << 32); o.value = p.cpu_pct;
		o.host_idx = p.host_idx; o.tsec = 0; o.type = GYSK_EV_TASK; o.flags = 0;
		break;
	}
	default : ev_pad(o); break;
	}
}

// one thread per raw record: 16-byte stores of the expanded event
__global__ void __launch_bounds__(256) decode_raw_kernel(uint32_t kind, const uint8_t *__restrict__ raw, uint32_t n, uint32_t host_idx, gysk_event *__restrict__ out)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	gysk_event o;
	decode_raw(kind, raw + (size_t)i * raw_stride(kind), host_idx, o);
	uint4 *d = reinterpret_cast<uint4 *>(out + i);
	d[0] = make_uint4((uint32_t)o.svc_id, (uint32_t)(o.svc_id >> 32), (uint32_t)o.flow_key, (uint32_t)(o.flow_key >> 32));
	d[1] = make_uint4(o.value, o.host_idx, o.tsec, (uint32_t)o.type | ((uint32_t)o.flags << 16));
}

} // namespace gysk

namespace {

// bulk path of a fixed-stride raw kind: raw bytes H2D (straight from the caller's buffer when that is page-locked, else through
// the thread's stage), expansion on the device into the current device event buffer, batch kernels when it is full
int ingest_raw_bulk(gysk_engine *e, ThreadStage *ts, uint32_t kind, uint32_t host_idx, const uint8_t *src, uint64_t n)
{
	const uint32_t stride = raw_stride(kind);
	cudaPointerAttributes pa {};
	const bool pinned = cudaPointerGetAttributes(&pa, src) == cudaSuccess && pa.type == cudaMemoryTypeHost;
	cudaGetLastError();
	int rc = pinned ? 0 : flush_stage(e, ts);		// pageable input bounces through the thread's chunk: hand over what it holds first
	if (rc) return rc;
	const uint64_t per_stage = (uint64_t)ts->cap * sizeof(gysk_event) / stride;		// records per thread chunk (as bytes)
	while (n) {
		std::unique_lock<std::mutex> lk(e->mtx);
		CU(e, cudaSetDevice(e->dev));
		const int k = e->stage_cur;
		const uint64_t room = e->cfg.stage_batch - e->stage_fill;
		uint64_t m = std::min<uint64_t>(room, n);
		if (m * stride > e->raw_bytes) m = e->raw_bytes / stride;
		if (!pinned) m = std::min<uint64_t>(m, per_stage);
		const int r = e->raw_cur;
		const uint8_t *hsrc = src;
		if (!pinned) {
			lk.unlock();
			CU(e, cudaEventSynchronize(ts->copied[ts->cur]));
			memcpy(ts->buf[ts->cur], src, (size_t)m * stride);
			hsrc = reinterpret_cast<const uint8_t *>(ts->buf[ts->cur]);
			lk.lock();
			if (e->stage_cur != k || e->raw_cur != r || e->cfg.stage_batch - e->stage_fill < m) continue;	// another thread moved on: size the piece again
		}
		if (e->stage_fill == 0) CU(e, cudaStreamWaitEvent(e->copy_stream, e->ev_done[k], 0));
		CU(e, cudaStreamWaitEvent(e->copy_stream, e->ev_raw_done[r], 0));		// the decode kernel that last read raw buffer r has run
		CU(e, cudaMemcpyAsync(e->d_raw[r], hsrc, (size_t)m * stride, cudaMemcpyHostToDevice, e->copy_stream));
		if (!pinned) { CU(e, cudaEventRecord(ts->copied[ts->cur], e->copy_stream)); ts->cur ^= 1; }
		CU(e, cudaEventRecord(e->ev_raw_copied[r], e->copy_stream));
		// the expansion runs on the COPY stream too: buffer k's events are complete in copy-stream order, as append_chunk's are
		decode_raw_kernel<<<(uint32_t)((m + 255) / 256), 256, 0, e->copy_stream>>>(kind, e->d_raw[r], (uint32_t)m, host_idx, e->d_events[k] + e->stage_fill);
		CU(e, cudaEventRecord(e->ev_raw_done[r], e->copy_stream));
		e->kernel_launches++;
		e->raw_cur = (r + 1) % NBUF;
		e->stage_fill += (uint32_t)m; src += m * stride; n -= m;
		if (e->stage_fill == e->cfg.stage_batch && (rc = submit_stage(e))) return rc;
	}
	return post_launch(e, "raw decode");
}

// API_TRAN, common/gy_proto_common.h:140-204 (probe-confirmed offsets: response_usec_ 48, glob_id_ 120, errorcode_ 152, cliport_ 166,
// request_len_ 170, lenext_ 172, padlen_ 174, sizeof 176; get_elem_size() = sizeof + request_len_ + lenext_ + padlen_).
// SVC_INFO_CAP::upd_stats_on_req (gy_proto_parser.cc:2678-2694): nrequests_++, resp_cache_.add_cache(response_usec_ / 1000),
// error counters. is_error / is_serv_err are the parser's verdict; the record carries errorcode_: != 0 is an error, >= 500 a
// server error (HTTP status convention of the reference's http parser).
struct ApiTranView
{
	static constexpr size_t SIZE = 176;
	static uint64_t u64(const uint8_t *p, size_t off) { uint64_t v; memcpy(&v, p + off, 8); return v; }
	static uint32_t u32(const uint8_t *p, size_t off) { uint32_t v; memcpy(&v, p + off, 4); return v; }
	static uint16_t u16(const uint8_t *p, size_t off) { uint16_t v; memcpy(&v, p + off, 2); return v; }
};

} // namespace

extern "C" {

int gysk_ingest_raw(gysk_engine *e, const uint8_t host_id[16], uint32_t host_idx, uint32_t kind, const void *events, uint32_t n)
{
	CHECK_ENGINE(e);
	(void)host_id;
	if (!events && n) return GYSK_ERR_INVAL;
	ThreadStage *ts = get_stage(e);
	if (!ts) return fail(e, GYSK_ERR_NOMEM, "thread stage");
	std::lock_guard<std::mutex> tl(ts->m);
	int rc = 0;

	if (kind == GYSK_RAW_EVENT32) return stage_events(e, ts, static_cast<const gysk_event *>(events), n);

	if (kind == GYSK_RAW_API_TRAN) {
		const uint8_t *p = static_cast<const uint8_t *>(events);
		for (uint32_t i = 0; i < n; ++i) {
			gysk_event *o = stage_slot(e, ts, &rc);
			if (!o) return rc;
			const uint64_t usec = ApiTranView::u64(p, 48);
			const int32_t err = (int32_t)ApiTranView::u32(p, 152);
			o->svc_id = ApiTranView::u64(p, 120);
			o->flow_key = ApiTranView::u16(p, 166);				// cliport_ (host order): CONN_BITMAP index
			o->value = usec > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)usec;	// beyond 1 000 000 msec: dropped by the validity rule on the device
			o->host_idx = host_idx; o->tsec = (uint32_t)(ApiTranView::u64(p, 16) / 1000000ull);	// tupd_usec_
			o->type = GYSK_EV_RESP;
			o->flags = err == 0 ? 0 : (err >= 500 ? GYSK_EVF_SER_ERROR : GYSK_EVF_CLI_ERROR);
			p += ApiTranView::SIZE + ApiTranView::u16(p, 170) + ApiTranView::u16(p, 172) + p[174];
		}
		return GYSK_OK;
	}

	const uint32_t stride = raw_stride(kind);
	if (!stride) return fail(e, GYSK_ERR_INVAL, "gysk_ingest_raw: unknown kind");
	if (n >= RAW_BULK_MIN) return ingest_raw_bulk(e, ts, kind, host_idx, static_cast<const uint8_t *>(events), n);
	// a handful of records (one perf-buffer wake-up): expanded by the calling thread
	const uint8_t *p = static_cast<const uint8_t *>(events);
	for (uint32_t i = 0; i < n; ++i, p += stride) {
		gysk_event tmp;
		decode_raw(kind, p, host_idx, tmp);
		if (tmp.type == 0xFFFF) continue;
		gysk_event *o = stage_slot(e, ts, &rc);
		if (!o) return rc;
		*o = tmp;
	}
	return GYSK_OK;
}

// One wire message body, exactly the arguments handle_l2_misc hands to partha_<kind>() (gy_mconnhdlr.cc:4745-4800).
int gysk_ingest(gysk_engine *e, const uint8_t host_id[16], uint32_t host_idx, uint32_t subtype, void *recs, uint32_t nevents, const void *endptr)
{
	CHECK_ENGINE(e);
	(void)host_id;
	if (!recs || !endptr || (const uint8_t *)endptr < (const uint8_t *)recs) return GYSK_ERR_INVAL;
	ThreadStage *ts = get_stage(e);
	if (!ts) return fail(e, GYSK_ERR_NOMEM, "thread stage");
	std::lock_guard<std::mutex> tl(ts->m);
	const uint8_t *pend = static_cast<const uint8_t *>(endptr);
	int rc = 0;

	switch (subtype) {

	case GYSK_NOTIFY_TCP_CONN : {
		using T = wire::TCP_CONN_NOTIFY;
		T *pone = static_cast<T *>(recs);
		if (!wire::validate_batch<T>(pone, nevents, pend, T::MAX_NUM_CONNS, [](const T & t) -> size_t { return t.cli_cmdline_len_; })) {
			e->wire_bad++;
			return fail(e, GYSK_ERR_INVAL, "TCP_CONN_NOTIFY::validate failed");
		}
		// record walk of partha_tcp_conn_info (gy_mconnhdlr.cc:9130): i < nconns && ptr < pendptr, stride get_elem_size()
		for (uint32_t i = 0; i < nevents && (const uint8_t *)pone < pend; ++i, pone = (T *)((uint8_t *)pone + pone->get_elem_size())) {
			if (!pone->ser_glob_id_ || !pone->cli_task_aggr_id_) continue;		// :9143 guard of the group-by
			const bool closed = !!pone->tusec_close_;
			uint16_t type;
			if (pone->is_tcp_accept_event_) type = closed ? GYSK_EV_CLOSE_SER : GYSK_EV_ACCEPT;
			else if (pone->is_tcp_connect_event_) type = closed ? GYSK_EV_CLOSE_CLI : GYSK_EV_CONNECT;
			else continue;
			gysk_event *o = stage_slot(e, ts, &rc);
			if (!o) return rc;
			const uint64_t bytes = closed ? pone->bytes_sent_ + pone->bytes_rcvd_ : 0;
			o->svc_id = pone->ser_glob_id_; o->flow_key = pone->cli_task_aggr_id_;
			o->value = bytes > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)bytes;
			o->host_idx = host_idx; o->tsec = (uint32_t)((closed ? pone->tusec_close_ : pone->tusec_start_) / 1000000ull);
			o->type = type; o->flags = 0;
		}
		e->wire_ok++;
		return GYSK_OK;
	}

	case GYSK_NOTIFY_AGGR_TASK_STATE : {
		using T = wire::AGGR_TASK_STATE_NOTIFY;
		T *pone = static_cast<T *>(recs);
		if (!wire::validate_batch<T>(pone, nevents, pend, T::MAX_NUM_TASKS, [](const T & t) -> size_t { return t.issue_string_len_; })) {
			e->wire_bad++;
			return fail(e, GYSK_ERR_INVAL, "AGGR_TASK_STATE_NOTIFY::validate failed");
		}
		// partha_aggr_task_state (gy_mconnhdlr.cc:9959) -> MAGGR_TASK::set_local_task_state (gy_msocket.h:1009)
		{
			HostTaskTopn tt;						// the seven per-host rankings of :10012-10079
			const T *q = pone;
			for (uint32_t i = 0; i < nevents && (const uint8_t *)q < pend; ++i, q = (const T *)((const uint8_t *)q + q->get_elem_size())) tt.offer(*q);
			std::lock_guard<std::mutex> lk(e->host_mtx);
			e->host_task_topn[host_idx] = std::move(tt);
		}
		for (uint32_t i = 0; i < nevents && (const uint8_t *)pone < pend; ++i, pone = (T *)((uint8_t *)pone + pone->get_elem_size())) {
			if (!pone->aggr_task_id_) continue;
			gysk_event *o = stage_slot(e, ts, &rc);
			if (!o) return rc;
			o->svc_id = pone->aggr_task_id_;
			o->flow_key = (uint64_t)pone->cpu_delay_msec_ | ((uint64_t)pone->blkio_delay_msec_ << 32);
			o->value = (uint32_t)(int)pone->total_cpu_pct_;				// (int)ptask->total_cpu_pct_ , gy_msocket.h:1014
			o->host_idx = host_idx; o->tsec = 0; o->type = GYSK_EV_TASK; o->flags = 0;
		}
		e->wire_ok++;
		return GYSK_OK;
	}

	case GYSK_NOTIFY_ACTIVE_CONN_STATS : {
		// ACTIVE_CONN_STATS::validate (common/gy_comm_proto.h:2806): fixed stride, nevents <= MAX_NUM_CONNS, all records inside the message
		using T = wire::ACTIVE_CONN_STATS;
		const T *pone = static_cast<const T *>(recs);
		if (nevents > T::MAX_NUM_CONNS || (const uint8_t *)(pone + nevents) > pend) {
			e->wire_bad++;
			return fail(e, GYSK_ERR_INVAL, "ACTIVE_CONN_STATS::validate failed");
		}
		// handle_partha_active_conns (gy_mconnhdlr.cc:7705) -> insert_active_conns (:7788): one record per {listener, client process}
		for (uint32_t i = 0; i < nevents; ++i, ++pone) {
			if (!pone->listener_glob_id_) continue;
			gysk_event *o = stage_slot(e, ts, &rc);
			if (!o) return rc;
			const uint64_t kb = (pone->bytes_sent_ + pone->bytes_received_) >> 10;
			o->svc_id = pone->listener_glob_id_; o->flow_key = pone->cli_aggr_task_id_;
			o->value = kb > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)kb;
			o->host_idx = host_idx;
			const float rtt = pone->max_rtt_msec_ > 0 ? pone->max_rtt_msec_ : 0.0f;
			memcpy(&o->tsec, &rtt, 4);
			o->type = GYSK_EV_ACTIVE; o->flags = pone->active_conns_;
		}
		e->wire_ok++;
		return GYSK_OK;
	}

	case GYSK_NOTIFY_LISTENER_STATE : {
		using T = wire::LISTENER_STATE_NOTIFY;
		T *pone = static_cast<T *>(recs);
		if (!wire::validate_batch<T>(pone, nevents, pend, T::MAX_NUM_LISTENERS, [](const T & t) -> size_t { return t.issue_string_len_; })) {
			e->wire_bad++;
			return fail(e, GYSK_ERR_INVAL, "LISTENER_STATE_NOTIFY::validate failed");
		}
		// Pre-aggregated 5-s listener state. With the per-sample reduction lifted onto the GPU the per-listener fields are
		// derived by the engine itself; what these records still feed is the per-host roll-up of partha_listener_state
		// (gy_mconnhdlr.cc:11175-11251): summstats.update(*pone) per record == LISTEN_SUMM_STATS::update, gy_msocket.h:854-866,
		// and the per-host top-N queues (:11262-11304). <= 512 records per host per 5 s: host-side integer work.
		gysk_host_summary hs;
		memset(&hs, 0, sizeof(hs));
		HostTopn topn;
		for (uint32_t i = 0; i < nevents && (const uint8_t *)pone < pend; ++i, pone = (T *)((uint8_t *)pone + pone->get_elem_size())) {
			// gy_mconnhdlr.cc:11183-11251: LISTEN_FLAG_DELETE records only delete the listener, records with
			// curr_state_ > STATE_DOWN count as errors; neither reaches summstats.update()
			if (pone->query_flags_ == wire::LISTEN_FLAG_DELETE || pone->curr_state_ > wire::STATE_DOWN) continue;
			hs.nstates[pone->curr_state_]++;
			hs.tot_qps += (int32_t)(pone->nqrys_5s_ / 5);
			hs.tot_act_conn += (int32_t)pone->nconns_active_;
			hs.tot_kb_inbound += (int32_t)pone->curr_kbytes_inbound_;
			hs.tot_kb_outbound += (int32_t)pone->curr_kbytes_outbound_;
			hs.tot_ser_errors += (int32_t)pone->ser_errors_;
			hs.nlisteners++;
			hs.nactive += !!pone->nqrys_5s_;
			topn.offer(*pone);
		}
		{
			std::lock_guard<std::mutex> lk(e->host_mtx);
			e->host_summ[host_idx] = hs;
			e->host_topn[host_idx] = topn;
		}
		e->wire_ok++;
		return GYSK_OK;
	}

	default :
		return fail(e, GYSK_ERR_NOTSUP, "gysk_ingest: subtype not on the hot path");
	}
}

// Whole message: COMM_HEADER + EVENT_NOTIFY + records, the pointer arithmetic of handle_l2_misc (gy_mconnhdlr.cc:4745-4760)
int gysk_ingest_msg(gysk_engine *e, const uint8_t host_id[16], uint32_t host_idx, void *msg, uint32_t msglen)
{
	CHECK_ENGINE(e);
	if (!msg || msglen < sizeof(wire::COMM_HEADER) + sizeof(wire::EVENT_NOTIFY)) return GYSK_ERR_INVAL;
	wire::COMM_HEADER *phdr = static_cast<wire::COMM_HEADER *>(msg);

	if (phdr->magic_ != wire::PM_HDR_MAGIC || phdr->data_type_ != wire::COMM_EVENT_NOTIFY || phdr->total_sz_ > msglen ||
			phdr->total_sz_ >= wire::MAX_COMM_DATA_SZ || phdr->padding_sz_ > phdr->total_sz_ ||
			phdr->get_act_len() < sizeof(wire::COMM_HEADER) + sizeof(wire::EVENT_NOTIFY)) {
		e->wire_bad++;
		return fail(e, GYSK_ERR_INVAL, "COMM_HEADER::validate failed");
	}
	uint8_t *pendptr = static_cast<uint8_t *>(msg) + phdr->get_act_len();
	wire::EVENT_NOTIFY *pevtnot = reinterpret_cast<wire::EVENT_NOTIFY *>(phdr + 1);

	return gysk_ingest(e, host_id, host_idx, pevtnot->subtype_, pevtnot + 1, pevtnot->nevents_, pendptr);
}

int gysk_sync(gysk_engine *e)
{
	CHECK_ENGINE(e);
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	return sync_locked(e);
}

int gysk_flush(gysk_engine *e, uint32_t tsec)
{
	CHECK_ENGINE(e);
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	int rc = submit_stage(e);
	if (rc) return rc;

	// rolling levels: the closing window goes to slot (tsec / width) % 10; a slot still holding an older epoch is cleared first
	const size_t plane = (size_t)e->cfg.max_svcs * HIST_CELLS;
	HistCell *planes[NLEVELS];
	for (int l = 0; l < NLEVELS; ++l) {
		const uint64_t epoch = tsec / g_level_width[l];
		const int k = (int)(epoch % NSLOTS);
		planes[l] = e->st.hist_ring + ((size_t)l * NSLOTS + k) * plane;
		if (e->ring_epoch[l][k] != epoch) {
			CU(e, cudaMemsetAsync(planes[l], 0, plane * sizeof(HistCell), e->stream));
			e->ring_epoch[l][k] = epoch;
		}
	}
	e->last_flush_tsec = tsec;

	if ((rc = collect_evicted(e, false))) return rc;		// list of the previous flush (normally long complete)
	// tombstones lengthen probe chains: once they fill an eighth of the table, rebuild it from the live slots
	const uint64_t dead = e->h_evict_fail > e->insert_fail_seen ? e->h_evict_fail - e->insert_fail_seen : 0;
	if (e->tombstones + dead > ((uint64_t)e->st.svc_tbl.mask + 1) / 8) {
		e->kernel_launches += launch_rebuild_table(e->st, e->cfg.max_svcs, e->stream);
		e->tombstones = 0; e->insert_fail_seen = e->h_evict_fail;
	}
	e->kernel_launches += launch_flush(e->st, e->cfg.max_svcs, planes[0], planes[1], tsec, e->cfg.idle_evict_secs, live_mask(e, 0), live_mask(e, 1), e->stream);
	e->kernel_launches += launch_task_flush(e->st, e->cfg.max_tasks, e->stream);
	if (e->cfg.idle_evict_secs) {
		// count + ids travel to the host behind the kernels; nobody waits for them here
		CU(e, cudaMemcpyAsync(e->h_evict, e->st.counters + CTR_NEVICT, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
		CU(e, cudaMemcpyAsync(e->h_evict + e->cfg.max_svcs + 1, e->st.counters + CTR_INSERT_FAIL, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
		CU(e, cudaMemcpyAsync(e->h_evict + 1, e->st.evict_ids, (size_t)e->cfg.max_svcs * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
		CU(e, cudaEventRecord(e->ev_evict, e->stream));
		e->evict_pending = true;
	}
	std::swap(e->st.cms_cur, e->st.cms_last);
	CU(e, cudaMemsetAsync(e->st.cms_cur, 0, sizeof(unsigned long long) * ((size_t)e->cfg.cms_depth << e->cfg.cms_log2_width), e->stream));
	return post_launch(e, "flush");
}

int gysk_evicted_ids(gysk_engine *e, uint64_t *out, uint32_t cap, uint32_t *n)
{
	CHECK_ENGINE(e);
	if (!n || (!out && cap)) return GYSK_ERR_INVAL;
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	int rc = collect_evicted(e, true);
	if (rc) return rc;
	*n = (uint32_t)e->evicted_ids.size();
	for (uint32_t i = 0; i < *n && i < cap; ++i) out[i] = e->evicted_ids[i];
	return GYSK_OK;
}

// ---- queries ------------------------------------------------------------------------------------------------

int gysk_query_svcs(gysk_engine *e, const uint64_t *ids, uint32_t n, gysk_svc_summary *out)
{
	CHECK_ENGINE(e);
	if ((!ids || !out) && n) return GYSK_ERR_INVAL;
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	int rc = submit_stage(e);
	if (rc) return rc;

	for (uint32_t off = 0; off < n; off += QCHUNK) {
		const uint32_t m = std::min(QCHUNK, n - off);
		if ((rc = gather_svcs(e, ids + off, m))) return rc;

		for (uint32_t i = 0; i < m; ++i) {
			summarize_raw(e, e->h_svcraw[i], ids[off + i], out[off + i]);
		}
	}
	return GYSK_OK;
}

int gysk_export_hist(gysk_engine *e, uint64_t id, int which, gysk_hist_serial out[GYSK_HIST_MAX_BUCKETS], uint64_t *total, int64_t *maxv)
{
	CHECK_ENGINE(e);
	if (!out || !total || !maxv) return GYSK_ERR_INVAL;
	if (which >= GYSK_HIST_TASK_CPU_PCT && which <= GYSK_HIST_TASK_BLKIO_DELAY) return gysk_export_task_hist(e, id, which, out, total, maxv);
	if (which > GYSK_HIST_ACTIVE_CONN) return GYSK_ERR_INVAL;
	if (which < 0) return GYSK_ERR_INVAL;
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	int rc = submit_stage(e);
	if (rc) return rc;
	if ((rc = gather_svcs(e, &id, 1))) return rc;
	const SvcRaw &r = e->h_svcraw[0];
	if (!r.found) return GYSK_ERR_NOENT;
	if (which == GYSK_HIST_RESP_5MIN || which == GYSK_HIST_RESP_5DAY) {
		hist_from_cells(r.lvl[which - GYSK_HIST_RESP_5MIN], 15, out, total, maxv, false);
		if (*total == 0) *maxv = INT64_MIN;
		return GYSK_OK;
	}
	if (which == GYSK_HIST_QPS) { hist_from_cells(r.qps, 15, out, total, maxv, true); return GYSK_OK; }
	if (which == GYSK_HIST_ACTIVE_CONN) { hist_from_cells(r.act, 14, out, total, maxv, true); return GYSK_OK; }
	hist_from_cells(which == GYSK_HIST_RESP_CUR ? r.cur : (which == GYSK_HIST_RESP_LAST ? r.last : r.all), 15, out, total, maxv, false);
	return GYSK_OK;
}

int gysk_export_task_hist(gysk_engine *e, uint64_t id, int which, gysk_hist_serial out[GYSK_HIST_MAX_BUCKETS], uint64_t *total, int64_t *maxv)
{
	CHECK_ENGINE(e);
	if (!out || !total || !maxv || which < GYSK_HIST_TASK_CPU_PCT || which > GYSK_HIST_TASK_BLKIO_DELAY) return GYSK_ERR_INVAL;
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	int rc = submit_stage(e);
	if (rc) return rc;
	e->h_qids[0] = id;
	CU(e, cudaMemcpyAsync(e->d_qids, e->h_qids, sizeof(uint64_t), cudaMemcpyHostToDevice, e->stream));
	e->kernel_launches += launch_gather_tasks(e->st, e->d_qids, 1, e->d_taskraw, e->stream);
	CU(e, cudaMemcpyAsync(e->h_taskraw, e->d_taskraw, sizeof(TaskRaw), cudaMemcpyDeviceToHost, e->stream));
	CU(e, cudaStreamSynchronize(e->stream));
	if ((rc = post_launch(e, "gather_tasks"))) return rc;
	if (!e->h_taskraw[0].found) return GYSK_ERR_NOENT;
	const int h = which - GYSK_HIST_TASK_CPU_PCT;
	hist_from_cells(e->h_taskraw[0].h[h], h == 0 ? 14 : 15, out, total, maxv, true);
	return GYSK_OK;
}

// CONN_BITMAP::get_conn_breakup (common/gy_socket_stat.h:412-433): per response bucket the number of (client port & 31) slots seen
int gysk_export_conn_bitmap(gysk_engine *e, uint64_t id, int last_window, uint32_t masks[GYSK_HIST_MAX_BUCKETS], uint8_t nconn_arr[GYSK_HIST_MAX_BUCKETS])
{
	CHECK_ENGINE(e);
	if (!masks || !nconn_arr) return GYSK_ERR_INVAL;
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	int rc = submit_stage(e);
	if (rc) return rc;
	if ((rc = gather_svcs(e, &id, 1))) return rc;
	const SvcRaw &r = e->h_svcraw[0];
	if (!r.found) return GYSK_ERR_NOENT;
	const uint32_t *bm = last_window ? r.bm_last : r.bm_cur;
	for (int j = 0; j < GYSK_HIST_MAX_BUCKETS; ++j) { masks[j] = bm[j]; nconn_arr[j] = (uint8_t)__builtin_popcount(bm[j]); }
	return GYSK_OK;
}

int gysk_export_hll(gysk_engine *e, uint64_t id, uint8_t *regs)
{
	CHECK_ENGINE(e);
	if (!regs) return GYSK_ERR_INVAL;
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	int rc = submit_stage(e);
	if (rc) return rc;
	e->kernel_launches += launch_gather_hll(e->st, id, e->d_hllout, e->d_found, e->stream);
	CU(e, cudaMemcpyAsync(e->h_hllout, e->d_hllout, (size_t)1 << e->cfg.hll_p, cudaMemcpyDeviceToHost, e->stream));
	CU(e, cudaMemcpyAsync(e->h_found, e->d_found, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
	CU(e, cudaStreamSynchronize(e->stream));
	if ((rc = post_launch(e, "gather_hll"))) return rc;
	if (!*e->h_found) return GYSK_ERR_NOENT;
	memcpy(regs, e->h_hllout, (size_t)1 << e->cfg.hll_p);
	return GYSK_OK;
}

int gysk_export_tdigest(gysk_engine *e, uint64_t id, double *means, uint64_t *weights, uint32_t cap, uint32_t *n, double *minv, double *maxv)
{
	CHECK_ENGINE(e);
	if (!means || !weights || !n) return GYSK_ERR_INVAL;
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	int rc = submit_stage(e);
	if (rc) return rc;
	if ((rc = gather_svcs(e, &id, 1))) return rc;
	const SvcRaw &r = e->h_svcraw[0];
	if (!r.found) return GYSK_ERR_NOENT;
	const uint32_t nc = std::min<uint32_t>(std::min<uint32_t>(r.td.n, TD_CAP), cap);
	for (uint32_t c = 0; c < nc; ++c) { means[c] = r.cent[c].mean; weights[c] = r.cent[c].weight; }
	*n = nc;
	if (minv) *minv = r.td.minv;
	if (maxv) *maxv = r.td.maxv;
	return r.td.n > cap ? GYSK_ERR_NOSPC : GYSK_OK;
}

// Summary encoder (output side): per-service summaries -> one NOTIFY_LISTENER_STATE message body, i.e. the records
// MTCP_LISTENER::set_state / the listener-state DB insert read (LISTENER_STATE_NOTIFY, common/gy_comm_proto.h:2183-2254; consumer
// partha_listener_state, server/gy_mconnhdlr.cc:11175-11251). Fields the engine computes are filled (nqrys_5s_, total_resp_5sec_,
// p95_5s_resp_ms_, p95_5min_resp_ms_, nconns_ = TCP events of the window, curr_kbytes_inbound_ = their kbytes); what only the host
// agent knows (task counters, errors, http flag) is zero; curr_state_ is STATE_IDLE (0) without queries, else STATE_OK (2) — the
// state classifier is out of scope. No issue string: every record is sizeof(LISTENER_STATE_NOTIFY) = 88 bytes, 8-byte aligned.
int gysk_encode_listener_state(const gysk_svc_summary *sums, uint32_t n, void *buf, uint32_t cap, uint32_t *nrecs, uint32_t *nbytes)
{
	if ((!sums && n) || !buf || !nrecs || !nbytes) return GYSK_ERR_INVAL;
	auto clamp32 = [](uint64_t v) -> uint32_t { return v > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)v; };
	auto clampms = [](int64_t v) -> uint32_t { return v < 0 ? 0u : (v > 0xFFFFFFFFll ? 0xFFFFFFFFu : (uint32_t)v); };
	uint8_t *p = static_cast<uint8_t *>(buf);
	uint32_t k = 0;

	for (uint32_t i = 0; i < n; ++i) {
		if (!sums[i].found) continue;
		if (k >= wire::LISTENER_STATE_NOTIFY::MAX_NUM_LISTENERS) break;			// one message holds at most 512 records (:2222)
		if ((size_t)(k + 1) * sizeof(wire::LISTENER_STATE_NOTIFY) > cap) return GYSK_ERR_NOSPC;
		wire::LISTENER_STATE_NOTIFY r;
		memset(&r, 0, sizeof(r));
		r.glob_id_ = sums[i].glob_id;
		r.nqrys_5s_ = sums[i].nqrys_5s;
		r.total_resp_5sec_ = clamp32(sums[i].total_resp_5sec);
		r.p95_5s_resp_ms_ = clampms(sums[i].p95_5s_resp_ms);
		r.p95_5min_resp_ms_ = clampms(sums[i].p95_5min_resp_ms);
		r.nconns_ = sums[i].nconns_5s;
		r.nconns_active_ = sums[i].nconns_active;
		r.ser_errors_ = sums[i].ser_errors; r.cli_errors_ = sums[i].cli_errors;
		r.curr_kbytes_inbound_ = sums[i].kbytes_5s;
		r.curr_state_ = sums[i].curr_state; r.curr_issue_ = sums[i].curr_issue;		// the device-side get_curr_state of the window
		r.issue_bit_hist_ = sums[i].issue_bit_hist; r.high_resp_bit_hist_ = sums[i].high_resp_bit_hist;
		memcpy(p + (size_t)k * sizeof(r), &r, sizeof(r));
		k++;
	}
	*nrecs = k; *nbytes = k * (uint32_t)sizeof(wire::LISTENER_STATE_NOTIFY);
	return GYSK_OK;
}

// Text form of the Postgres `tdigest` type (extension tvondra/tdigest, loaded by the reference with `create extension if not
// exists tdigest`, common/gy_query_common.cc:3385-3387; version unpinned there). tdigest_out prints
//   "flags %d count %ld compression %d centroids %d" followed by " (%lf, %ld)" per centroid, flags = 1 (TDIGEST_STORES_MEAN),
// and tdigest_in parses the same with sscanf: a row built from this string answers tdigest_percentile(col, p) like the rows the
// reference aggregates with public.tdigest(expr, 100) (gy_query_common.cc:1805-1858). Means are printed with 17 significant
// digits (sscanf %lf reads them back exactly). Returns the string length (without NUL), or GYSK_ERR_NOSPC.
int gysk_tdigest_to_pgtext(const double *means, const uint64_t *weights, uint32_t n, uint32_t compression, char *buf, uint32_t cap)
{
	if ((!means || !weights) && n) return GYSK_ERR_INVAL;
	if (!buf || !cap) return GYSK_ERR_INVAL;
	uint64_t total = 0;
	for (uint32_t i = 0; i < n; ++i) total += weights[i];
	int off = snprintf(buf, cap, "flags 1 count %llu compression %u centroids %u", (unsigned long long)total, compression, n);
	if (off < 0 || (uint32_t)off >= cap) return GYSK_ERR_NOSPC;
	for (uint32_t i = 0; i < n; ++i) {
		const int k = snprintf(buf + off, cap - off, " (%.17g, %llu)", means[i], (unsigned long long)weights[i]);
		if (k < 0 || (uint32_t)(off + k) >= cap) return GYSK_ERR_NOSPC;
		off += k;
	}
	return off;
}

// The engine keeps delta = 200 internally; the Postgres side of the reference aggregates with public.tdigest(expr, 100)
// (common/gy_query_common.cc:1855) and the extension refuses to combine digests of different compression, so the export is one
// more fixed-grid K_1 compress at compression 100 (host side, same rule as the device).
static uint32_t host_td_compress(const double *means, const uint64_t *w, uint32_t n, uint32_t delta, double *om, uint64_t *ow)
{
	if (!n) return 0;
	std::vector<double> qtab(delta + 1);
	for (uint32_t j = 0; j <= delta; ++j) qtab[j] = 0.5 * (sin(M_PI * ((double)j / (double)delta - 0.5)) + 1.0);
	qtab[0] = 0.0; qtab[delta] = 1.0;
	uint64_t W = 0, pref = 0, cw = 0;
	for (uint32_t i = 0; i < n; ++i) W += w[i];
	uint32_t nout = 0, cur = 0;
	double csum = 0.0;
	for (uint32_t i = 0; i < n; ++i) {
		uint32_t lo = cur;
		while (lo + 1 < delta && (uint64_t)(qtab[lo + 1] * (double)W) <= pref) ++lo;		// cell j starts at weight (uint64) (q_j W)
		if (i && lo != cur) { om[nout] = csum / (double)cw; ow[nout++] = cw; cw = 0; csum = 0.0; }
		cur = lo;
		csum += means[i] * (double)w[i]; cw += w[i]; pref += w[i];
	}
	om[nout] = csum / (double)cw; ow[nout++] = cw;
	return nout;
}

int gysk_export_tdigest_pgtext(gysk_engine *e, uint64_t id, char *buf, uint32_t cap)
{
	double means[TD_CAP], minv = 0, maxv = 0, om[TD_CAP];
	uint64_t w[TD_CAP], ow[TD_CAP];
	uint32_t n = 0;
	int rc = gysk_export_tdigest(e, id, means, w, TD_CAP, &n, &minv, &maxv);
	if (rc) return rc;
	const uint32_t no = host_td_compress(means, w, n, 100, om, ow);
	return gysk_tdigest_to_pgtext(om, ow, no, 100, buf, cap);
}

int gysk_query_quantiles(gysk_engine *e, uint64_t id, const double *qs, uint32_t nq, double *out)
{
	double means[TD_CAP], minv = 0, maxv = 0;
	uint64_t w[TD_CAP];
	uint32_t n = 0;
	int rc = gysk_export_tdigest(e, id, means, w, TD_CAP, &n, &minv, &maxv);

	if (rc) return rc;
	if ((!qs || !out) && nq) return GYSK_ERR_INVAL;
	for (uint32_t i = 0; i < nq; ++i) out[i] = td_quantile(means, w, n, minv, maxv, qs[i]);
	return GYSK_OK;
}

int gysk_query_flows(gysk_engine *e, const uint64_t *keys, uint32_t n, int last_window, gysk_flow_est *out)
{
	CHECK_ENGINE(e);
	if ((!keys || !out) && n) return GYSK_ERR_INVAL;
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	int rc = submit_stage(e);
	if (rc) return rc;
	for (uint32_t off = 0; off < n; off += QCHUNK) {
		const uint32_t m = std::min(QCHUNK, n - off);
		memcpy(e->h_qids, keys + off, (size_t)m * sizeof(uint64_t));
		CU(e, cudaMemcpyAsync(e->d_qids, e->h_qids, (size_t)m * sizeof(uint64_t), cudaMemcpyHostToDevice, e->stream));
		e->kernel_launches += launch_query_flows(e->st, e->d_qids, m, last_window, e->d_flowout, e->stream);
		CU(e, cudaMemcpyAsync(e->h_flowout, e->d_flowout, (size_t)m * sizeof(gysk_flow_est), cudaMemcpyDeviceToHost, e->stream));
		CU(e, cudaStreamSynchronize(e->stream));
		memcpy(out + off, e->h_flowout, (size_t)m * sizeof(gysk_flow_est));
	}
	return post_launch(e, "query_flows");
}

int gysk_topn_svcs(gysk_engine *e, int metric, int32_t host_idx, uint32_t n, gysk_topn_entry *out, uint32_t *nout)
{
	CHECK_ENGINE(e);
	if (!out || !nout || n == 0 || n > 64 || metric < GYSK_TOPN_QPS || metric > GYSK_TOPN_ISSUE) return GYSK_ERR_INVAL;
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	int rc = sync_locked(e);
	if (rc) return rc;
	uint32_t nslots = 0;
	CU(e, cudaMemcpy(&nslots, e->st.svc_tbl.count, sizeof(uint32_t), cudaMemcpyDeviceToHost));
	nslots = std::min(nslots, e->cfg.max_svcs);
	gysk_topn_entry *d_out = reinterpret_cast<gysk_topn_entry *>(e->d_flowout);		// QCHUNK * 16 B >= 64 * 24 B
	CU(e, cudaMemsetAsync(d_out, 0, sizeof(gysk_topn_entry) * n, e->stream));
	{
		const int nl = launch_topn(e->st, e->tmp, nslots, metric, host_idx, n, d_out, e->stream);
		if (nl < 0) return fail(e, GYSK_ERR_INVAL, "gysk_topn_svcs: sort failed");
		e->kernel_launches += nl;
	}
	gysk_topn_entry *h_out = reinterpret_cast<gysk_topn_entry *>(e->h_flowout);
	CU(e, cudaMemcpyAsync(h_out, d_out, sizeof(gysk_topn_entry) * n, cudaMemcpyDeviceToHost, e->stream));
	CU(e, cudaStreamSynchronize(e->stream));
	if ((rc = post_launch(e, "topn"))) return rc;
	uint32_t k = 0;
	for (uint32_t i = 0; i < n; ++i) if (h_out[i].glob_id && h_out[i].score) out[k++] = h_out[i];
	*nout = k;
	return GYSK_OK;
}

int gysk_topn_tasks(gysk_engine *e, int metric, uint32_t n, gysk_topn_entry *out, uint32_t *nout)
{
	CHECK_ENGINE(e);
	if (!out || !nout || n == 0 || n > 64 || metric < GYSK_TOPN_TASK_CPU || metric > GYSK_TOPN_TASK_BLKIO_DELAY) return GYSK_ERR_INVAL;
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	int rc = sync_locked(e);
	if (rc) return rc;
	uint32_t ntasks = 0;
	CU(e, cudaMemcpy(&ntasks, e->st.task_tbl.count, sizeof(uint32_t), cudaMemcpyDeviceToHost));
	ntasks = std::min(ntasks, e->cfg.max_tasks);
	gysk_topn_entry *d_out = reinterpret_cast<gysk_topn_entry *>(e->d_flowout);		// QCHUNK * 16 B >= 64 * 24 B
	CU(e, cudaMemsetAsync(d_out, 0, sizeof(gysk_topn_entry) * n, e->stream));
	{
		const int nl = launch_topn_tasks(e->st, e->tmp, ntasks, metric, n, d_out, e->stream);
		if (nl < 0) return fail(e, GYSK_ERR_INVAL, "gysk_topn_tasks: sort failed");
		e->kernel_launches += nl;
	}
	gysk_topn_entry *h_out = reinterpret_cast<gysk_topn_entry *>(e->h_flowout);
	CU(e, cudaMemcpyAsync(h_out, d_out, sizeof(gysk_topn_entry) * n, cudaMemcpyDeviceToHost, e->stream));
	CU(e, cudaStreamSynchronize(e->stream));
	if ((rc = post_launch(e, "topn_tasks"))) return rc;
	uint32_t k = 0;
	for (uint32_t i = 0; i < n; ++i) if (h_out[i].glob_id && h_out[i].score) out[k++] = h_out[i];
	*nout = k;
	return GYSK_OK;
}

int gysk_query_host_summary(gysk_engine *e, uint32_t host_idx, gysk_host_summary *out)
{
	CHECK_ENGINE(e);
	if (!out) return GYSK_ERR_INVAL;
	std::lock_guard<std::mutex> lk(e->host_mtx);
	auto it = e->host_summ.find(host_idx);
	if (it == e->host_summ.end()) return GYSK_ERR_NOENT;
	*out = it->second;
	return GYSK_OK;
}

int gysk_query_cluster_state(gysk_engine *e, const uint32_t *host_idxs, uint32_t n, gysk_cluster_state *out)
{
	CHECK_ENGINE(e);
	if (!out || (!host_idxs && n)) return GYSK_ERR_INVAL;
	std::lock_guard<std::mutex> lk(e->host_mtx);
	memset(out, 0, sizeof(*out));
	auto add = [&](const gysk_host_summary & hs) {
		const uint32_t issues = (uint32_t)(hs.nstates[3] + hs.nstates[4] + hs.nstates[5]);	// STATE_BAD, STATE_SEVERE, STATE_DOWN
		out->nhosts++;
		out->nsvc_issue += issues; out->nsvcissue_hosts += !!issues;
		out->nsvc += (uint32_t)hs.nlisteners;
		out->total_qps += (uint32_t)hs.tot_qps;
		out->svc_net_mb += (uint32_t)((hs.tot_kb_inbound + hs.tot_kb_outbound) / 1024);
	};
	if (!host_idxs) { for (const auto & kv : e->host_summ) add(kv.second); }
	else {
		for (uint32_t i = 0; i < n; ++i) {
			auto it = e->host_summ.find(host_idxs[i]);
			if (it != e->host_summ.end()) add(it->second);
		}
	}
	return GYSK_OK;
}

// the per-host rankings the reference keeps in PARTHA_INFO (BOUNDED_PRIO_QUEUE, 10 entries per host, gy_mconnhdlr.h:961,975), as
// of each host's last NOTIFY_LISTENER_STATE / NOTIFY_AGGR_TASK_STATE message; host_idx < 0 merges the hosts of this engine
int gysk_topn_host(gysk_engine *e, int what, int32_t host_idx, uint32_t n, gysk_topn_entry *out, uint32_t *nout)
{
	CHECK_ENGINE(e);
	if (!out || !nout || !n || n > 64 || what < 0 || what > GYSK_HOSTTOP_TASK_BLKIO_DELAY) return GYSK_ERR_INVAL;
	std::lock_guard<std::mutex> lk(e->host_mtx);
	std::vector<gysk_topn_entry> all;
	auto take = [&](uint32_t h, const TopQueue &q) { for (const TopEntry &t : q.v) all.push_back(gysk_topn_entry {t.id, t.score, h, 0}); };
	if (what <= GYSK_HOSTTOP_SVC_NET) {
		for (const auto &kv : e->host_topn) if (host_idx < 0 || kv.first == (uint32_t)host_idx) take(kv.first, kv.second.q[what]);
	}
	else {
		for (const auto &kv : e->host_task_topn) if (host_idx < 0 || kv.first == (uint32_t)host_idx) take(kv.first, kv.second.q[what - GYSK_HOSTTOP_TASK_ISSUE]);
	}
	std::sort(all.begin(), all.end(), [](const gysk_topn_entry &a, const gysk_topn_entry &b) { return a.score != b.score ? a.score > b.score : a.glob_id < b.glob_id; });
	*nout = (uint32_t)std::min<size_t>(all.size(), n);
	for (uint32_t i = 0; i < *nout; ++i) out[i] = all[i];
	return GYSK_OK;
}

int gysk_export_cms(gysk_engine *e, int last_window, uint64_t *cells)
{
	CHECK_ENGINE(e);
	if (!cells) return GYSK_ERR_INVAL;
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	int rc = sync_locked(e);
	if (rc) return rc;
	CU(e, cudaMemcpy(cells, last_window ? e->st.cms_last : e->st.cms_cur, sizeof(uint64_t) * ((size_t)e->cfg.cms_depth << e->cfg.cms_log2_width),
			cudaMemcpyDeviceToHost));
	return GYSK_OK;
}

// ---- pure helpers ------------------------------------------------------------------------------------------------

int gysk_hist_nbuckets(int cls)
{
	if (cls < 0 || cls > GYSK_CLS_PERCENT) return GYSK_ERR_INVAL;
	return g_cls[cls].nthr + 2;
}

int gysk_hist_bucket(int cls, int64_t value)
{
	if (cls < 0 || cls > GYSK_CLS_PERCENT) return GYSK_ERR_INVAL;
	const ClsDesc &d = g_cls[cls];
	int64_t data = d.trunc_int ? (int64_t)(int32_t)value : value;

	if (data < d.minv) return 0;
	if (data >= d.maxv) return d.nthr + 1;
	if (d.fixed_diff) return (int)(1 + (data - d.minv) / d.fixed_diff);
	int b = 1;
	for (int i = 0; i < d.nthr; ++i) b += (data > d.thr[i]);
	return b;
}

// GY_HISTOGRAM::get_percentiles, common/gy_statistics.h:707-791: float multiplier, size_t * float cut-off,
// first bucket whose cumulative count reaches it, answer = that bucket's upper threshold cast to T
int gysk_hist_percentiles(int cls, int t_is_int, const gysk_hist_serial *stats, uint64_t total_count, const float *pcts, uint32_t npct, int64_t *out)
{
	if (cls < 0 || cls > GYSK_CLS_PERCENT || !stats || (!pcts && npct) || (!out && npct)) return GYSK_ERR_INVAL;
	const ClsDesc &d = g_cls[cls];
	const size_t nb = (size_t)d.nthr + 2;

	for (uint32_t n = 0; n < npct; ++n) {
		const float multiplier = (float)(pcts[n] / 100.0);
		const size_t ncutoff = (size_t)((float)total_count * multiplier);
		size_t i, total = 0;

		for (i = 0; i < nb; ++i) {
			total += stats[i].count;
			if (total >= ncutoff) break;
		}
		int64_t v = bucket_max_threshold(d, !!t_is_int, i < nb ? i : (total_count > 0 ? nb : 0));
		out[n] = t_is_int ? (int64_t)(int32_t)v : v;
	}
	return GYSK_OK;
}

// TCP_LISTENER::get_curr_state for a caller that has the inputs from outside the path (process status, host state, dependencies)
int gysk_classify_listener(const gysk_listener_state_in *in, uint8_t *high_resp_bit_hist, uint8_t *state, uint8_t *issue)
{
	if (!in || !high_resp_bit_hist || !state || !issue) return GYSK_ERR_INVAL;
	classify_listener(*in, *high_resp_bit_hist, *state, *issue);
	return GYSK_OK;
}

double gysk_hll_estimate(const uint8_t *regs, uint32_t p)
{
	if (!regs || p < 4 || p > 16) return NAN;
	uint32_t hist[64] = {0};
	for (uint32_t i = 0; i < (1u << p); ++i) hist[regs[i] > 63 ? 63 : regs[i]]++;
	return hll_estimate_from_hist(hist, p);
}

double gysk_tdigest_quantile(const double *means, const uint64_t *weights, uint32_t n, double minv, double maxv, double q)
{
	if ((!means || !weights) && n) return NAN;
	return td_quantile(means, weights, n, minv, maxv, q);
}

uint32_t gysk_uint64_hash(uint64_t key) { return uint64_hash(key); }

// ---- multi-GPU merge: implemented in gysk_merge.cu ------------------------------------------------------------

} // extern "C"
