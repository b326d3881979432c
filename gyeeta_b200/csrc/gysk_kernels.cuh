// gysk_kernels.cuh — launch interface between the engine runtime (gysk_engine.cu) and the kernels.
#pragma once

#include "gysk_device.cuh"
#include "gysk_tdigest.cuh"
#include "../../include/gysketch.h"

namespace gysk {

struct DevState
{
	// id tables
	IdTable			svc_tbl, task_tbl;
	// per-service state, indexed by slot
	HistCell		*hist_cur, *hist_last, *hist_all;	// [max_svcs][16]
	HistCell		*hist_ring;				// [NLEVELS][NSLOTS][max_svcs][16] rolling 300-s / 5-day levels
	unsigned long long	*conn_cur, *conn_last;			// packed {count:32, kbytes:32}
	unsigned long long	*slot_id;				// [max_svcs] slot -> glob_id (written by the inserter)
	uint32_t		*slot_host;				// [max_svcs] slot -> host_idx of the first event seen
	uint32_t		*slot_first_seen, *slot_last_active;	// [max_svcs] tsec of the first flush that saw the slot / of the last window with events
	uint32_t		*evict_list;				// [max_svcs] slots evicted by the last flush; evict_ids: their ids
	unsigned long long	*evict_ids;
	unsigned long long	*conn_all_cnt, *conn_all_kb;
	uint32_t		*bm_cur, *bm_last;			// [max_svcs][16] CONN_BITMAP transposed: per bucket a mask over (client port & 31)
	SlotBatch		*slot_batch;				// [max_svcs] exact extremes of the batch's RESP samples, hot row
	unsigned long long	*hot_rows;				// [hot_cap][HOT_ROW_WORDS] dense value bins of the hot services (nullptr: feature off)
	uint32_t		*hot_slot;				// [hot_cap] row -> slot
	uint32_t		hot_cap, hot_min, hot_max, hot_bin_max;	// rows; a service turns hot with hot_min..hot_max samples in one batch, its fullest bin <= hot_bin_max
	SlotAux			*slot_aux;				// [max_svcs] active-conn roll-up, error counters
	HistCell		*qps_hist, *act_hist;			// [max_svcs][16] TCP_LISTENER::qps_hist_ / active_conn_hist_: one sample per closed window
	SlotState		*slot_state;				// [max_svcs] listener state of the last evaluated window
	uint8_t			*hll;					// [max_svcs][1 << hll_p]
	Centroid		*td_cent;				// [max_svcs][TD_CAP]
	TdHead			*td_head;				// [max_svcs]
	// per-task state
	HistCell		*task_hist;				// [max_tasks][3][16]
	HistCell		*task_prev, *task_last;			// [max_tasks][3] {count, sum}: totals at the last flush / of the last closed window
	unsigned long long	*task_slot_id;				// [max_tasks] slot -> aggr_task_id
	uint32_t		*task_slot_host;
	// flow sketch
	unsigned long long	*cms_cur, *cms_last;			// [depth][1 << log2w]
	uint32_t		cms_depth, cms_wmask, cms_log2w, hll_p;
	uint32_t		rank, world, auto_register;
	double			td_delta;
	TdParams		td;
	unsigned long long	*counters;				// [CTR_MAX]
};

struct SortTemp
{
	unsigned long long	*keys_a, *keys_b;	// [max_batch] RESP sort keys of the batch (also the top-N sort keys)
	unsigned long long	*tile_status;		// [max_tiles][512] look-back status words {pass epoch | state | count}: never cleared
	uint32_t		*epoch;			// HOST counter of radix passes launched (tags the status words)
	uint32_t		*os_ghist;		// [8][512] global digit histograms of the one-sweep passes + [8] tile tickets
	uint32_t		*touched;		// [max_svcs] services with RESP samples in the batch
	ulonglong2		*pool;			// [pool_cap] RunRec: one record per run (non-empty bin of a service) of the batch
	uint16_t		*run_bin;		// [pool_cap] bin index of each run
	uint32_t		*chunk_run;		// [max_batch / 128] run of the first key of every 128-key chunk
	uint4			*segs;			// [max_svcs] BatchSeg of each touched service
	Centroid		*items_scratch;		// [merge warps][NBINS] a warp's list of batch items
	TdWorkBig		*big_scratch;		// [merge warps] work arrays for merged lists beyond 2 x TD_CAP entries
	uint4			*tcpq, *taskq;		// ONE buffer of max_batch records {slot, value, flow key}: connection records from the front
							// (tcpq), process records from the back (taskq = last entry, growing down); ingest_kernel
							// resolves the ids and queues them, side_drain_kernel applies them next to the sort chain
	uint32_t		max_tiles;
};

// raw per-id record gathered for queries / exports
struct SvcRaw
{
	unsigned long long	id;
	int32_t			found;
	uint32_t		slot;
	HistCell		cur[HIST_CELLS], last[HIST_CELLS], all[HIST_CELLS];
	HistCell		lvl[2][HIST_CELLS];			// sums of the live slots of the rolling levels
	unsigned long long	conn_cur, conn_last, conn_all_cnt, conn_all_kb;
	uint32_t		bm_cur[HIST_CELLS], bm_last[HIST_CELLS];
	uint32_t		hll_hist[64];
	TdHead			td;
	SlotAux			aux;
	SlotState		sst;
	HistCell		qps[HIST_CELLS], act[HIST_CELLS];
	Centroid		cent[TD_CAP];
};

struct TaskRaw
{
	unsigned long long	id;
	int32_t			found;
	uint32_t		slot;
	HistCell		h[3][HIST_CELLS];
};

static constexpr int NLEVELS = 2;			// rolling levels beyond the 5-s window: 300 s, 432000 s (gy_statistics.h:1548)
static constexpr int NSLOTS = 10;			// slots per level (gy_statistics.h:1105)
static constexpr int SORT_TILE = 4096;		// keys per CTA tile in the radix passes
static constexpr int RADIX_MAX_BITS = 9;
static constexpr int RADIX_MAX = 1 << RADIX_MAX_BITS;
static constexpr int OS_MAX_PASSES_VK = 5;		// {slot : <= 24 | bin : 10} = <= 34 bits in digits of <= 8 bits
static constexpr int TD_MERGE_CTAS_PER_SM = 7, TD_MERGE_MAX_SMS = 192;	// bins_merge_kernel grid (<= 4 warps per CTA)

// every launcher returns the number of kernel launches it issued
int launch_init_state(const DevState &st, uint32_t max_svcs, uint32_t max_tasks, cudaStream_t s);
int launch_register(const DevState &st, const unsigned long long *d_ids, uint32_t n, int is_task, cudaStream_t s);
int launch_ingest(const DevState &st, const SortTemp &tmp, const gysk_event *d_ev, uint64_t n, uint32_t max_svcs, cudaStream_t s);
int launch_batch_merge(const DevState &st, const SortTemp &tmp, uint64_t n_events, uint32_t max_svcs, cudaStream_t s);
bool side_drain_enabled();
int launch_side_drain(const DevState &st, const SortTemp &tmp, uint64_t n_events, cudaStream_t s);
int launch_radix_sort(const SortTemp &tmp, const unsigned long long *d_n, uint64_t n_max, int lo1, int hi1, int lo2, int hi2, int *which, cudaStream_t s);
int radix_sort_plan(int lo1, int hi1, int lo2, int hi2, int out[][4], int cap);
int launch_topn(const DevState &st, const SortTemp &tmp, uint32_t nslots, int metric, int host_filter, uint32_t want, gysk_topn_entry *d_out, cudaStream_t s);
int launch_topn_tasks(const DevState &st, const SortTemp &tmp, uint32_t ntasks, int metric, uint32_t want, gysk_topn_entry *d_out, cudaStream_t s);
int launch_task_flush(const DevState &st, uint32_t max_tasks, cudaStream_t s);
int launch_flush(const DevState &st, uint32_t max_svcs, HistCell *ring_plane0, HistCell *ring_plane1, uint32_t tsec, uint32_t idle_secs,
		uint32_t live_mask0, uint32_t live_mask1, cudaStream_t s);
int launch_rebuild_table(const DevState &st, uint32_t max_svcs, cudaStream_t s);
int launch_gather_svcs(const DevState &st, const unsigned long long *d_ids, uint32_t n, uint32_t max_svcs, uint32_t live_mask0, uint32_t live_mask1,
		SvcRaw *d_out, cudaStream_t s);
int launch_gather_tasks(const DevState &st, const unsigned long long *d_ids, uint32_t n, TaskRaw *d_out, cudaStream_t s);
int launch_gather_hll(const DevState &st, unsigned long long id, uint8_t *d_out, int32_t *d_found, cudaStream_t s);
int launch_query_flows(const DevState &st, const unsigned long long *d_keys, uint32_t n, int last_window, gysk_flow_est *d_out, cudaStream_t s);

} // namespace gysk
