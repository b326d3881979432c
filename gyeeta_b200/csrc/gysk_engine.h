// gysk_engine.h — engine object shared by gysk_engine.cu (ingest / query) and gysk_merge.cu (multi-GPU merge).
#pragma once

#include "gysk_kernels.cuh"

#include <algorithm>
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
	cudaStream_t		stream {nullptr}, copy_stream {nullptr};
	gysk::DevState		st {};
	gysk::SortTemp		tmp {};
	std::vector<void *>	dallocs;
	std::vector<void *>	hallocs;

	// staging
	gysk_event		*h_stage[gysk::NBUF] {};
	gysk_event		*d_events[gysk::NBUF] {};
	cudaEvent_t		ev_copied[gysk::NBUF] {}, ev_done[gysk::NBUF] {};
	uint32_t		stage_fill {0};
	int			stage_cur {0};

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

	std::unordered_map<uint32_t, gysk_host_summary> host_summ;	// last LISTEN_SUMM_STATS per host (control-plane sized: <= 512 hosts)

	gysk::MergeState	mg;

	std::mutex		mtx;
	std::string		err;
	bool			sticky {false};
	uint64_t		kernel_launches {0}, batches {0}, wire_ok {0}, wire_bad {0}, merges {0};
};

namespace gysk {

int fail(gysk_engine *e, int code, const char *what, cudaError_t ce = cudaSuccess);
int post_launch(gysk_engine *e, const char *what);
int submit_stage(gysk_engine *e);
int sync_locked(gysk_engine *e);
int collect_evicted(gysk_engine *e, bool wait);
void merge_release(gysk_engine *e);
void summarize_raw(const gysk_engine *e, const SvcRaw &r, uint64_t id, gysk_svc_summary &o);

#define CU(e, call) do { cudaError_t ce__ = (call); if (ce__ != cudaSuccess) return gysk::fail((e), GYSK_ERR_CUDA, #call, ce__); } while (0)
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
