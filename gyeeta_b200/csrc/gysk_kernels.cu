// gysk_kernels.cu — hand-written sm_100a kernels of the streaming-sketch engine.
//
//   ingest_kernel        one pass over a batch of 32-byte events: id -> slot, then per event type
//                          RESP : one 64-bit sort key {slot | value bin | usec}; CONN_BITMAP bit, batch min / max when they change
//                          TCP  : count-min cell adds, HLL register max, per-service exact cell
//                          TASK : MAGGR_TASK::set_local_task_state (3 histograms)     server/gy_msocket.h:1009-1018
//                          ACTIVE : ACTIVE_CONN_STATS records                          server/gy_mconnhdlr.cc:7705
//   os_pass_kernel       stable one-sweep LSD radix pass (RESP keys of a batch on {slot, bin}; the top-N rankings)
//   runs_mark_kernel     sorted keys -> runs (one per non-empty value bin of a service), service segments, touched list
//   runs_sum_kernel      per run: samples, exact usec sum
//   bins_merge_kernel    per touched service: runs -> GY_HISTOGRAM::add_data for every sample (RESP_TIME_HASH, common/gy_statistics.h:
//                        596-623, :1698) and -> the merging t-digest (K_1 scale)               DESIGN.md §2
//   flush_kernel         5-s window roll                                               common/gy_socket_stat.cc:3898
//   state_kernel         listener state of the closed window (get_curr_state)          common/gy_socket_stat.cc:2020-2875
//   evict_kernel         idle listeners leave, slots recycled                          common/gy_socket_stat.cc:3968-4037
//   gather_* / query_* / topn_*   read side
#include "gysk_kernels.cuh"
#include "gysk_state.cuh"

#include <cfloat>
#include <climits>
#include <cmath>
#include <algorithm>
#include <type_traits>
#include <cstdlib>
#include <cstring>

namespace gysk {

// RESP sort key = {slot : 24 | bin index : 10 | usec : 30}. Bin index = td_code(usec) + RESP_TIME_HASH bucket of usec / 1000: both
// terms are monotone in usec, so the index is too and no bin straddles a histogram bucket. The radix passes sort on
// the top 10 + slot bits only: the samples of one (service, bin) end up as one contiguous RUN, in no particular order inside it.
static constexpr int KEY_GROUP_SHIFT = 30;				// key >> 30 = {slot, bin}
static constexpr int KEY_SLOT_SHIFT = KEY_GROUP_SHIFT + TD_CODE_BITS;
__device__ __forceinline__ uint32_t key_usec(unsigned long long k) { return (uint32_t)k & 0x3FFFFFFFu; }
__device__ __forceinline__ uint32_t key_slot(unsigned long long k) { return (uint32_t)(k >> KEY_SLOT_SHIFT); }
__device__ __forceinline__ uint32_t key_bin(unsigned long long k) { return (uint32_t)(k >> KEY_GROUP_SHIFT) & ((1u << TD_CODE_BITS) - 1u); }

struct SortPlan { int np; int shift[OS_MAX_PASSES_VK]; int bits[OS_MAX_PASSES_VK]; int exp; };	// digit p = (key >> shift[p]) & ((1 << bits[p]) - 1)
// exp: ABLATION switches for timing runs only (GYSK_EXP_ABLATE; results are wrong when set): 1 = no batch-extreme / CONN_BITMAP loads and
// atomics, 2 = no digit histograms, 4 = no TCP drain, 8 = no TASK drain

// ---------------------------------------------------------------------------------------------------
// state init / registration
// ---------------------------------------------------------------------------------------------------
__global__ void init_state_kernel(DevState st, uint32_t max_svcs, uint32_t max_tasks)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;

	if (i < max_svcs) {
		st.hist_cur[(size_t)i * HIST_CELLS + HIST_MAX_CELL].sum = LLONG_MIN;	// max_val_seen_{numeric_limits<T>::min()} gy_statistics.h:560
		st.hist_last[(size_t)i * HIST_CELLS + HIST_MAX_CELL].sum = LLONG_MIN;
		st.hist_all[(size_t)i * HIST_CELLS + HIST_MAX_CELL].sum = LLONG_MIN;
		st.td_head[i].minv = INFINITY;
		st.td_head[i].maxv = -INFINITY;
		st.slot_batch[i].minv = 0xFFFFFFFFu;
		st.qps_hist[(size_t)i * HIST_CELLS + HIST_MAX_CELL].sum = INT_MIN;	// GY_HISTOGRAM<int, ...>: numeric_limits<int>::min()
		st.act_hist[(size_t)i * HIST_CELLS + HIST_MAX_CELL].sum = INT_MIN;
		st.slot_state[i] = SlotState {GYSK_STATE_OK, GYSK_ISSUE_NONE, 0, 0, 0};
	}
	if (i < max_tasks) {
		for (int h = 0; h < 3; ++h) st.task_hist[((size_t)i * 3 + h) * HIST_CELLS + HIST_MAX_CELL].sum = LLONG_MIN;
	}
}

__global__ void register_kernel(DevState st, const unsigned long long *ids, uint32_t n, int is_task)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;

	if (i < n && ids[i] && ids[i] != KEY_TOMBSTONE) table_lookup(is_task ? st.task_tbl : st.svc_tbl, ids[i], true);
}

// ---------------------------------------------------------------------------------------------------
// ingest
// ---------------------------------------------------------------------------------------------------
// Service popularity is Zipf-skewed: a few {count, sum} cells would take millions of same-address L2 atomics per batch and
// serialise the kernel. Two levels take that pressure off L2:
//   warp : lanes updating the same cell are grouped with match.any, reduced with a shuffle loop and represented by one leader;
//   CTA  : a direct-mapped shared-memory table privatises hot cells for the lifetime of the CTA (a cell is admitted when it
//          shows up at least twice inside one warp), and is flushed with one RED pair per entry when the CTA retires.
// Everything stays exact integer arithmetic, so the result is independent of grouping and order.
template <int BITS_>
struct HotTableT				// structure of arrays, 20 B per entry
{
	static constexpr int BITS = BITS_;
	static constexpr int N = 1 << BITS;
	uint32_t		tag[N];			// cell id + 1, 0 = free
	uint32_t		count[N];
	unsigned long long	sum[N];
	int			vmax[N];
};
static constexpr uint32_t CELL_TASK = 1u << 30;			// cell ids: conn = slot, task = CELL_TASK | (tslot*48 + hist*16 + bucket)

// one RED per field, nothing is read back: {count, sum} (+ max_val_seen_ for histogram cells)
__device__ __forceinline__ void cell_add_global(const DevState &st, uint32_t cell, uint32_t cnt, unsigned long long sum, int vmax)
{
	if (cell & CELL_TASK) {
		HistCell *c = st.task_hist + (cell & ~CELL_TASK);
		red_add_u64(&c->count, cnt); red_add_u64((unsigned long long *)&c->sum, sum);
		// max_val_seen_ of the histogram: a fire-and-forget RED.MAX — looking first (to skip the atomic) made every process record wait
		// for an L2 round trip (7.7 % of the kernel's stall samples, profiles/r02_ncu_full_raw_final.csv); the busy processes' cells
		// live in the CTA's hot table and reach this point once per CTA
		red_max_s64(&st.task_hist[(cell & ~CELL_TASK) | 15u].sum, (long long)vmax);
	}
	else red_add_u64(st.conn_cur + cell, (unsigned long long)cnt + (sum << 32));	// packed {count, kbytes}
}

// all 32 lanes call this; lanes with active == false only take part in the collectives.
// Group sums use a shuffle loop bounded by the largest group of the warp (typically 1-4): redux with per-lane masks would
// make the compiler iterate over every distinct group. No global load anywhere: the updates are fire-and-forget REDs.
template <typename HotTable>
__device__ __forceinline__ void cell_add(const DevState &st, HotTable &hot, bool active, uint32_t cell, int data)
{
	const int lane = threadIdx.x & 31;
	const uint32_t id = active ? cell : (0x80000000u | (uint32_t)lane);
	const uint32_t m = __match_any_sync(0xffffffffu, id);
	const uint32_t cnt = __popc(m);
	const uint32_t maxcnt = __reduce_max_sync(0xffffffffu, cnt);
	long long sum = data;
	int gmax = data;
	uint32_t rest = m & ~(1u << lane);

	for (uint32_t t = 1; t < maxcnt; ++t) {
		const int src = rest ? (__ffs(rest) - 1) : lane;
		const int other = __shfl_sync(0xffffffffu, data, src);
		if (rest) { sum += other; gmax = max(gmax, other); rest &= rest - 1; }
	}

	if (!active || (m & ((1u << lane) - 1u))) return;		// group leader = lowest lane

	// two candidate entries per cell, hashed with a per-CTA seed: which hot cells collide differs from CTA to CTA, so no cell
	// loses its privatisation everywhere at once (slot numbers, hence cell ids, depend on registration order)
	const uint32_t seed = blockIdx.x * 0x9E3779B9u;
	uint32_t h = ((cell ^ seed) * 2654435761u) >> (32 - HotTable::BITS);
	uint32_t tag = *((volatile uint32_t *)&hot.tag[h]);
	bool hit = tag == cell + 1;

	if (!hit) {
		const uint32_t h2 = ((cell ^ ~seed) * 0x85EBCA6Bu) >> (32 - HotTable::BITS);
		const uint32_t tag2 = *((volatile uint32_t *)&hot.tag[h2]);
		if (tag2 == cell + 1) { hit = true; h = h2; }
		else if (cnt >= 2) {
			// admission: a cell that shows up twice in one warp takes a free candidate entry
			if (tag == 0) { tag = atomicCAS(&hot.tag[h], 0u, cell + 1); hit = tag == 0 || tag == cell + 1; }
			if (!hit && tag2 == 0) { const uint32_t t2 = atomicCAS(&hot.tag[h2], 0u, cell + 1); if (t2 == 0 || t2 == cell + 1) { hit = true; h = h2; } }
		}
	}
	if (hit) { atomicAdd(&hot.count[h], cnt); atomicAdd(&hot.sum[h], (unsigned long long)sum); atomicMax(&hot.vmax[h], gmax); }
	else cell_add_global(st, cell, cnt, (unsigned long long)sum, gmax);
}

// HLL register update in two halves, so that the caller can put other work between the load of the register word and its use
__device__ __forceinline__ uint32_t hll_peek(const uint8_t *regs, uint32_t idx)
{
	return __ldca(reinterpret_cast<const uint32_t *>(regs) + (idx >> 2));	// stale is harmless: the CAS re-validates
}

__device__ __forceinline__ void hll_raise(uint8_t *regs, uint32_t idx, uint32_t rank, uint32_t w)
{
	uint32_t *wp = reinterpret_cast<uint32_t *>(regs) + (idx >> 2);
	const uint32_t sh = (idx & 3u) * 8u;

	while (((w >> sh) & 0xFFu) < rank) {
		const uint32_t nw = (w & ~(0xFFu << sh)) | (rank << sh);
		const uint32_t old = atomicCAS(wp, w, nw);
		if (old == w) break;
		w = old;
	}
}

__device__ __forceinline__ void hll_update(uint8_t *regs, uint32_t idx, uint32_t rank)
{
	hll_raise(regs, idx, rank, hll_peek(regs, idx));
}

// ---- TMA (bulk async copy) of a warp's next event chunk into shared memory, completion on the warp's own mbarrier ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, unsigned long long *bar)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			:: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
	asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}"
			:: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// The kernel is WARP-AUTONOMOUS: no block barrier anywhere in the event loop. A warp takes a chunk of 32 x EPT events and
//   (1) decodes them fully converged: 2 x 128-bit load per event, shard filter, id -> slot lookup with the first table probe
//       of all EPT events in flight together;
//   (2) a RESP sample (70 % of the stream) becomes ONE 64-bit sort key {slot : 24 | value bin : 10 | usec : 30} in the warp's key
//       queue (ballot + popc placement; the queue leaves as a coalesced run, and its digits are counted into the CTA's radix
//       histograms on the way out); beside it only what would change state: the batch extremes of the slot and the CONN_BITMAP bit,
//       each behind a load so that the atomic is issued only while the value still moves;
//   (3) TCP / TASK events join one of the warp's two private shared-memory queues (ballot + popc, no atomics: the queue
//       lengths are warp-uniform registers) and a queue is drained only in whole multiples of 32 entries, so every lane works in
//       every iteration whatever the mix of the stream is: TCP = two lookup2 hashes per flow key -> four count-min REDs + HLL
//       register + the service's exact {count, kbytes} cell; TASK = the three histograms of
//       MAGGR_TASK::set_local_task_state with one (event, histogram) pair per lane; the < 32 left-over entries move to the front.
// The histogram cells and the t-digest of a service are produced from its bins by bins_merge_kernel after the batch.
struct alignas(16) IngestRec { uint32_t slot; uint32_t value; unsigned long long flow_key; };		// moved as one 128-bit word
static_assert(sizeof(IngestRec) == 16, "IngestRec travels as one uint4");

// m queued connection records (all 32 lanes call; q in shared or global memory): two lookup2 hashes per flow key -> four count-min
// REDs + the HLL register + the service's exact {count, kbytes} cell
template <typename HotTable>
__device__ __forceinline__ void drain_tcp_recs(const DevState &st, HotTable &hot, const IngestRec *q, uint32_t m, int lane)
{
	for (uint32_t i = lane; i < ((m + 31u) & ~31u); i += 32) {
		const bool act = i < m;
		uint32_t cell = 0, idx = 0, rank = 0, hw = 0; int kb = 0;
		if (act) {
			const IngestRec r = q[i];
			uint32_t h1, h2;
			flow_hashes(r.flow_key, h1, h2);
			// the HLL register word is asked for first and looked at last: the count-min REDs and the cell update hide its latency
			hll_idx_rank2(h1, h2, st.hll_p, idx, rank);
			hw = hll_peek(st.hll + ((size_t)r.slot << st.hll_p), idx);
			const unsigned long long inc = cms_increment(r.value);
			for (uint32_t row = 0; row < st.cms_depth; ++row)
				red_add_u64(st.cms_cur + ((size_t)row << st.cms_log2w) + cms_index2(h1, h2, row, st.cms_wmask), inc);
			cell = r.slot;
			kb = (int)(r.value >> 10);
		}
		cell_add(st, hot, act, cell, kb);
		if (act) hll_raise(st.hll + ((size_t)cell << st.hll_p), idx, rank, hw);
	}
}

// m queued process records: the three histograms of MAGGR_TASK::set_local_task_state with one (record, histogram) pair per lane
template <typename HotTable, bool BACKWARDS = false>
__device__ __forceinline__ void drain_task_recs(const DevState &st, HotTable &hot, const IngestRec *q, uint32_t m, int lane)
{
	const uint32_t ntrip = m * 3u;
	for (uint32_t p = lane; p < ((ntrip + 31u) & ~31u); p += 32) {
		const bool act = p < ntrip;
		uint32_t cell = 0; int d = 0;
		if (act) {
			const uint32_t e = p / 3u, h = p - e * 3u;
			const IngestRec r = BACKWARDS ? *(q - (long long)e) : q[e];
			// GY_HISTOGRAM<int, ...>::add_data(int): the three values narrow to int (server/gy_msocket.h:1014-1016)
			d = h == 0 ? (int)r.value : (h == 1 ? (int)(uint32_t)r.flow_key : (int)(uint32_t)(r.flow_key >> 32));
			const uint32_t b = h == 0 ? (uint32_t)bucket_hash_1_3000(d) : (uint32_t)bucket_duration(d);
			cell = CELL_TASK | (r.slot * 3u * HIST_CELLS + h * HIST_CELLS + b);
		}
		cell_add(st, hot, act, cell, d);
	}
}


template <int WARPS, int EPT, bool TMA, int DH, bool SIDE>
struct IngestSharedT
{
	static constexpr int CHUNK = 32 * EPT;			// events per warp and round
	static constexpr int KQ_CAP = EPT <= 2 ? 192 : 256, KQ_FLUSH = KQ_CAP - CHUNK;	// a flush leaves room for a whole chunk of RESP events
	static constexpr int RQ_CAP = 32 + CHUNK;			// < 32 left over + one chunk
	static_assert(CHUNK <= 128, "key queue sized for chunks of at most 128 events");
	using HotTable = HotTableT<SIDE ? 1 : 9>;			// with the side drain the cells are privatised there
	struct Warp { unsigned long long kq[KQ_CAP]; IngestRec tcp[RQ_CAP], task[RQ_CAP]; };
	alignas(128) uint4	evbuf[TMA ? WARPS * CHUNK * 2 : 1];	// per warp: its next chunk of 32-byte events, filled by cp.async.bulk
	unsigned long long	mbar[TMA ? WARPS : 1];
	Warp		w[WARPS];
	HotTable	hot;
	uint32_t	dhist[OS_MAX_PASSES_VK][DH];			// digit histograms of this CTA's keys, one per radix pass (DH = 256 unless a pass has 9-bit digits)
};

template <int WARPS, int MIN_CTAS, int EPT, bool TMA, int DH, bool SIDE>
__global__ void __launch_bounds__(WARPS * 32, MIN_CTAS) ingest_kernel(DevState st, const gysk_event *__restrict__ ev, uint64_t n,
		unsigned long long *__restrict__ keys, uint32_t *__restrict__ ghist, SortPlan plan, uint4 *__restrict__ tcpq, uint4 *__restrict__ taskq)
{
	using Shared = IngestSharedT<WARPS, EPT, TMA, DH, SIDE>;
	using HotTable = typename Shared::HotTable;
	constexpr int CHUNK = Shared::CHUNK;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	Shared &S = *reinterpret_cast<Shared *>(smem_raw);
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	typename Shared::Warp &W = S.w[wid];
	const uint32_t lt = (1u << lane) - 1u;
	uint32_t c_in = 0, c_foreign = 0, n_resp = 0, n_active = 0;	// per thread: < 2^32 events per launch
	uint32_t nk = 0, ntcp = 0, ntask = 0;				// queue lengths (warp-uniform)
	unsigned long long t_tcp = 0, t_task = 0;			// queued in total (warp-uniform)

	for (int i = threadIdx.x; i < HotTable::N; i += WARPS * 32) { S.hot.tag[i] = 0; S.hot.count[i] = 0; S.hot.sum[i] = 0; S.hot.vmax[i] = INT_MIN; }
	for (int i = threadIdx.x; i < OS_MAX_PASSES_VK * DH; i += WARPS * 32) (&S.dhist[0][0])[i] = 0;
	if (TMA && lane == 0) mbar_init(&S.mbar[wid], 1);
	__syncthreads();						// the only block barriers: here and before the retire step
	if (TMA) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");	// mbarrier init visible to the async proxy

	const uint64_t nchunks = (n + CHUNK - 1) / CHUNK;
	const uint64_t nwarps = (uint64_t)gridDim.x * WARPS;
	const uint64_t gwarp = (uint64_t)blockIdx.x * WARPS + wid;
	uint4 *evw = S.evbuf + (TMA ? wid * CHUNK * 2 : 0);
	uint32_t tma_phase = 0;
	auto tma_issue = [&](uint64_t chunk) {
		const uint64_t b0 = chunk * CHUNK;
		const uint64_t cnt = n - b0 < (uint64_t)CHUNK ? n - b0 : (uint64_t)CHUNK;
		// every lane's generic-proxy reads of the buffer happened before the __syncwarp that precedes this call
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		tma_load_1d(evw, ev + b0, (uint32_t)cnt * 32u, &S.mbar[wid]);
	};
	if (TMA && lane == 0 && gwarp < nchunks) tma_issue(gwarp);

	// ---- queue drains (all 32 lanes, m = multiple of 32 except in the final drain): applied here, or — SIDE — handed as one
	//      coalesced run to the batch's record queues, which side_drain_kernel works off while the radix passes run ----
	// one buffer holds both queues: connection records grow from its front, process records from its back (taskq points at the last
	// entry; together they never exceed the batch's event count)
	auto hand_over = [&](const IngestRec *q, uint32_t m, uint4 *gq, int ctr, bool backwards) {
		unsigned long long base = 0;
		if (lane == 0) base = atomicAdd(st.counters + ctr, (unsigned long long)m);
		base = __shfl_sync(0xffffffffu, base, 0);
		for (uint32_t i = lane; i < m; i += 32) {
			const uint4 v = *reinterpret_cast<const uint4 *>(q + i);
			if (backwards) __stcs(gq - (long long)(base + i), v); else __stcs(gq + base + i, v);
		}
	};
	auto drain_tcp = [&](uint32_t m) {
		if (plan.exp & 4) return;
		if (SIDE) hand_over(W.tcp, m, tcpq, CTR_NTCPQ, false); else drain_tcp_recs(st, S.hot, W.tcp, m, lane);
	};
	auto drain_task = [&](uint32_t m) {
		if (plan.exp & 8) return;
		if (SIDE) hand_over(W.task, m, taskq, CTR_NTASKQ, true); else drain_task_recs(st, S.hot, W.task, m, lane);
	};
	auto keep_rest = [&](IngestRec *q, uint32_t m, uint32_t total) {		// entries [m, total) move to the front (total - m < 32)
		IngestRec r;
		const bool mv = m + lane < total;
		if (mv) r = q[m + lane];
		__syncwarp();
		if (mv) q[lane] = r;
		__syncwarp();
	};
	auto flush_keys = [&]() {
		unsigned long long base = 0;
		if (lane == 0) base = atomicAdd(st.counters + CTR_NKEYS, (unsigned long long)nk);
		base = __shfl_sync(0xffffffffu, base, 0);
		for (uint32_t q = lane; q < nk; q += 32) {
			const unsigned long long k = W.kq[q];
#pragma unroll
			for (int p = 0; p < OS_MAX_PASSES_VK; ++p)
				if (p < plan.np && !(plan.exp & 2)) atomicAdd(&S.dhist[p][(uint32_t)(k >> plan.shift[p]) & ((1u << plan.bits[p]) - 1u)], 1u);
			__stcs(keys + base + q, k);
		}
		nk = 0;
		__syncwarp();
	};

	for (uint64_t chunk = gwarp; chunk < nchunks; chunk += nwarps) {
		const uint64_t cbase = chunk * CHUNK;
		uint4 ra[EPT], rb[EPT];
		if (TMA) mbar_wait(&S.mbar[wid], tma_phase), tma_phase ^= 1u;
#pragma unroll
		for (int k = 0; k < EPT; ++k) {
			const uint64_t i = cbase + (uint64_t)k * 32 + lane;
			if (i < n) {
				if (TMA) { ra[k] = evw[2 * (k * 32 + lane)]; rb[k] = evw[2 * (k * 32 + lane) + 1]; }
				else {
					ra[k] = __ldcs(reinterpret_cast<const uint4 *>(ev + i));		// streamed once: evict-first, keep L2 for
					rb[k] = __ldcs(reinterpret_cast<const uint4 *>(ev + i) + 1);	// the id table / histogram / count-min lines
				}
			}
			else { ra[k] = make_uint4(0, 0, 0, 0); rb[k] = make_uint4(0, 0, 0, 0xFFFFu); }	// type 0xFFFF: padding, not counted
		}
		if (TMA) {
			__syncwarp();								// the whole warp has copied its events out of the buffer
			if (lane == 0 && chunk + nwarps < nchunks) tma_issue(chunk + nwarps);	// next chunk flies in while this one is processed
		}

		// decode; put the first id-table probe of all EPT events in flight before any of them is resolved
		uint4 praw[EPT];
		uint32_t ppos[EPT];
		uint32_t kind[EPT];		// 0 none, GYSK_EV_RESP, GYSK_EV_ACCEPT (= any TCP type), GYSK_EV_TASK
#pragma unroll
		for (int k = 0; k < EPT; ++k) {
			const unsigned long long svc = ((unsigned long long)ra[k].y << 32) | ra[k].x;
			const uint32_t value = rb[k].x, host_idx = rb[k].y;
			const uint32_t type = rb[k].w & 0xFFFFu;
			const bool is_resp = type == GYSK_EV_RESP, is_task = type == GYSK_EV_TASK;
			const bool is_tcp = type >= GYSK_EV_CONNECT && type <= GYSK_EV_CLOSE_SER, is_active = type == GYSK_EV_ACTIVE;
			bool mine = type != 0xFFFFu;

			kind[k] = 0; ppos[k] = 0; praw[k] = make_uint4(0, 0, 0, 0);
			if (mine && st.world > 1 && (host_idx % st.world) != st.rank) { c_foreign++; mine = false; }
			if (mine) {
				c_in++;
				// usec -> msec as SVC_INFO_CAP::upd_stats_on_req (gy_proto_parser.cc:2678); validity rule of
				// handle_ipv4_resp_event (gy_socket_stat.cc:1519-1524): drop beyond 1 000 000 msec
				if (svc + 1ull > 1ull && (is_tcp || is_task || is_active || (is_resp && value < 1000001000u))) {	// id not 0 / ~0 (tombstone); msec <= 1 000 000
					kind[k] = is_resp ? (uint32_t)GYSK_EV_RESP : (is_task ? (uint32_t)GYSK_EV_TASK : (is_active ? (uint32_t)GYSK_EV_ACTIVE : (uint32_t)GYSK_EV_ACCEPT));
					praw[k] = table_probe_first(is_task ? st.task_tbl : st.svc_tbl, svc, ppos[k]);
				}
			}
		}
		// resolve every slot, then put the second round of loads (the slot's batch record and CONN_BITMAP word) of all EPT events
		// in flight together, and fire the bin REDs — nothing below waits for them
		int slotv[EPT];
		uint4 sbv[EPT];
		uint32_t mwv[EPT], bkt[EPT];
#pragma unroll
		for (int k = 0; k < EPT; ++k) {
			const bool is_resp = kind[k] == GYSK_EV_RESP, is_task = kind[k] == GYSK_EV_TASK;
			int slot = -1;
			if (kind[k]) {
				// one id lookup for all three event kinds (services and tasks live in separate tables)
				slot = table_resolve(is_task ? st.task_tbl : st.svc_tbl, ((unsigned long long)ra[k].y << 32) | ra[k].x, st.auto_register, rb[k].y, ppos[k], praw[k]);
			}
			slotv[k] = slot; sbv[k] = make_uint4(0, 0, 0, 0); mwv[k] = 0; bkt[k] = 0;
			if (slot >= 0 && is_resp) {
				const uint32_t v = rb[k].x, ms = v / 1000u;		// usec -> msec as SVC_INFO_CAP::upd_stats_on_req (gy_proto_parser.cc:2678)
				const uint32_t b = (uint32_t)bucket_resp_time((long long)ms);
				bkt[k] = b;
				if (!(plan.exp & 1)) {
					sbv[k] = ld_cg_v4(st.slot_batch + slot);
					mwv[k] = __ldcg(st.bm_cur + (size_t)slot * HIST_CELLS + b);
				}
				else { sbv[k] = make_uint4(0u, 0xFFFFFFFFu, 0u, 0u); mwv[k] = 0xFFFFFFFFu; }
				n_resp++;
			}
		}
#pragma unroll
		for (int k = 0; k < EPT; ++k) {
			const bool is_resp = kind[k] == GYSK_EV_RESP, is_task = kind[k] == GYSK_EV_TASK, is_tcp = kind[k] == GYSK_EV_ACCEPT;
			const int slot = slotv[k];
			const bool ok = slot >= 0;
			// a hot service (sbv.w = 1 + its row of dense value bins, handed out by bins_merge_kernel after an earlier batch) takes its
			// sample as two REDs into the L2-resident row; everybody else's sample becomes a sort key
			const uint32_t hotrow = (ok && is_resp) ? sbv[k].w : 0u;
			const uint32_t m_resp = __ballot_sync(0xffffffffu, ok && is_resp && !hotrow), m_tcp = __ballot_sync(0xffffffffu, ok && is_tcp),
					m_task = __ballot_sync(0xffffffffu, ok && is_task);
			if (ok) {
				if (is_resp) {
					const uint32_t bin = td_code(rb[k].x) + bkt[k];
					if (hotrow) {
						unsigned long long *hb = st.hot_rows + (size_t)(hotrow - 1u) * HOT_ROW_WORDS + hot_word(bin);
						const uint32_t us = rb[k].x;
						red_add_u64(hb, 1ull | ((unsigned long long)(us - (us / 1000u) * 1000u) << BIN_CNT_BITS));
						red_add_u64(hb + HOT_ROW_BINS, (unsigned long long)us);
					}
					else W.kq[nk + __popc(m_resp & lt)] = ((unsigned long long)(uint32_t)slot << KEY_SLOT_SHIFT) |
							((unsigned long long)bin << KEY_GROUP_SHIFT) | rb[k].x;
					// the rest only when it changes something: batch extremes (minv != ~0 also marks the slot as touched) and the
					// CONN_BITMAP bit — TCP_LISTENER::CONN_BITMAP::add_response (common/gy_socket_stat.h:403-410), transposed: per
					// bucket a mask over client port & 31
					const uint32_t v = rb[k].x;
					if (v < sbv[k].x) atomicMin(&st.slot_batch[slot].minv, v);
					if (v > sbv[k].y) atomicMax(&st.slot_batch[slot].maxv, v);
					const uint32_t bit = 1u << (ra[k].z & 0x1Fu);
					if (!(mwv[k] & bit)) atomicOr(st.bm_cur + (size_t)slot * HIST_CELLS + bkt[k], bit);
					const uint32_t ef = rb[k].w >> 16;			// API_TRAN error flags: rare
					if (ef & 3u) red_add_u64(&st.slot_aux[slot].err_cur, (unsigned long long)(ef & 1u) | ((unsigned long long)((ef >> 1) & 1u) << 32));
				}
				else if (kind[k] == GYSK_EV_ACTIVE) {
					// one pre-aggregated {listener, client process} record of the 15-s inet_diag scan (gy_socket_stat.cc:6156-6194): a few
					// per flow and minute — handled on the spot. The flow sketch takes its connections and kbytes, the service its totals.
					const unsigned long long fk = ((unsigned long long)ra[k].w << 32) | ra[k].z;
					const unsigned long long inc = (unsigned long long)(rb[k].w >> 16) | ((unsigned long long)rb[k].x << 32);
					uint32_t h1, h2, idx, rank;
					flow_hashes(fk, h1, h2);
					for (uint32_t row = 0; row < st.cms_depth; ++row)
						red_add_u64(st.cms_cur + ((size_t)row << st.cms_log2w) + cms_index2(h1, h2, row, st.cms_wmask), inc);
					hll_idx_rank2(h1, h2, st.hll_p, idx, rank);
					hll_update(st.hll + ((size_t)slot << st.hll_p), idx, rank);
					red_add_u64(&st.slot_aux[slot].act_cur, inc);
					atomicMax(&st.slot_aux[slot].rtt_cur, rb[k].z);		// non-negative floats order like their bit patterns
					n_active++;
				}
				else {
					IngestRec r; r.slot = (uint32_t)slot; r.value = rb[k].x; r.flow_key = ((unsigned long long)ra[k].w << 32) | ra[k].z;
					if (is_tcp) W.tcp[ntcp + __popc(m_tcp & lt)] = r;
					else W.task[ntask + __popc(m_task & lt)] = r;
				}
			}
			nk += __popc(m_resp); ntcp += __popc(m_tcp); ntask += __popc(m_task);
		}
		__syncwarp();

		if (ntcp >= 32) { const uint32_t m = ntcp & ~31u; drain_tcp(m); keep_rest(W.tcp, m, ntcp); t_tcp += m; ntcp -= m; }
		if (ntask >= 32) { const uint32_t m = ntask & ~31u; drain_task(m); keep_rest(W.task, m, ntask); t_task += m; ntask -= m; }
		if (nk > (uint32_t)Shared::KQ_FLUSH) flush_keys();
	}
	// what is left in the queues
	if (ntcp) { drain_tcp(ntcp); t_tcp += ntcp; }
	if (ntask) { drain_task(ntask); t_task += ntask; }
	if (nk) flush_keys();

	__syncthreads();
	// retire: one RED group per privatised cell, one RED per digit this CTA saw
	if (!SIDE) for (int i = threadIdx.x; i < HotTable::N; i += WARPS * 32) {
		if (S.hot.tag[i] && S.hot.count[i]) cell_add_global(st, S.hot.tag[i] - 1, S.hot.count[i], S.hot.sum[i], S.hot.vmax[i]);
	}
	for (int i = threadIdx.x; i < plan.np * DH; i += WARPS * 32) {
		const uint32_t c = (&S.dhist[0][0])[i];
		if (c) atomicAdd(ghist + (i / DH) * RADIX_MAX + (i % DH), c);		// the passes read their histogram at stride RADIX_MAX
	}

	// statsmap-style counters (gy_mconnhdlr.cc:4708-4715): warp-reduce, one atomic per warp and counter
	c_in = __reduce_add_sync(0xffffffffu, c_in);
	c_foreign = __reduce_add_sync(0xffffffffu, c_foreign);
	const unsigned long long t_resp = __reduce_add_sync(0xffffffffu, n_resp);
	t_tcp += __reduce_add_sync(0xffffffffu, n_active);		// counted with the connection events
	if (lane == 0) {
		if (c_in) atomicAdd(st.counters + CTR_IN, (unsigned long long)c_in);
		if (c_foreign) atomicAdd(st.counters + CTR_FOREIGN, (unsigned long long)c_foreign);
		if (t_resp) atomicAdd(st.counters + CTR_RESP, t_resp);
		if (t_tcp) atomicAdd(st.counters + CTR_TCP, t_tcp);
		if (t_task) atomicAdd(st.counters + CTR_TASK, t_task);
		// dropped = taken in but not queued (svc_id 0, bad type or value, table full, unknown id); two's complement arithmetic
		const unsigned long long q = t_resp + t_tcp + t_task;
		if (c_in != q) atomicAdd(st.counters + CTR_DROPPED, (unsigned long long)c_in - q);
	}
}

// ---------------------------------------------------------------------------------------------------
// EXPERIMENT (GYSK_SIDE_DRAIN=1, off by default): the connection and process records ingest_kernel queued, applied next to the sort
// chain on the engine's side stream. Idea: the radix passes are bound by instruction issue and leave the L2 atomic units idle, the ~85 M
// REDs of the two drains (ablation: 0.8 + 0.75 ms inside ingest_kernel, profiles/r02_ablation.json) are bound by exactly those units.
// Outcome: correct (the whole GPU suite passes with it on) but slower, see side_drain_enabled(). Persistent grid; a warp takes 32
// records at a time; hot cells are privatised per CTA as in ingest_kernel.
// ---------------------------------------------------------------------------------------------------
static constexpr int SD_WARPS = 8;

__global__ void __launch_bounds__(SD_WARPS * 32) side_drain_kernel(DevState st, const uint4 *__restrict__ tcpq, const uint4 *__restrict__ taskq)
{
	using HotTable = HotTableT<9>;
	__shared__ HotTable hot;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	for (int i = threadIdx.x; i < HotTable::N; i += SD_WARPS * 32) { hot.tag[i] = 0; hot.count[i] = 0; hot.sum[i] = 0; hot.vmax[i] = INT_MIN; }
	__syncthreads();
	const unsigned long long ntcp = st.counters[CTR_NTCPQ], ntask = st.counters[CTR_NTASKQ];
	const unsigned long long gw = (unsigned long long)blockIdx.x * SD_WARPS + wid, nw = (unsigned long long)gridDim.x * SD_WARPS;

	for (unsigned long long base = gw * 32; base < ntcp; base += nw * 32)
		drain_tcp_recs(st, hot, reinterpret_cast<const IngestRec *>(tcpq) + base, (uint32_t)(ntcp - base < 32 ? ntcp - base : 32), lane);
	for (unsigned long long base = gw * 32; base < ntask; base += nw * 32)
		drain_task_recs<HotTable, true>(st, hot, reinterpret_cast<const IngestRec *>(taskq) - (long long)base, (uint32_t)(ntask - base < 32 ? ntask - base : 32), lane);

	__syncthreads();
	for (int i = threadIdx.x; i < HotTable::N; i += SD_WARPS * 32) {
		if (hot.tag[i] && hot.count[i]) cell_add_global(st, hot.tag[i] - 1, hot.count[i], hot.sum[i], hot.vmax[i]);
	}
}

// ---------------------------------------------------------------------------------------------------
// stable LSD radix sort, 8- or 9-bit digits, tile = SORT_TILE keys per CTA of 256 threads
// ---------------------------------------------------------------------------------------------------
// (used by the top-N rankings: {score | slot} keys over the registered services / tasks)
// A pass sorts on a digit made of up to two bit fields of the key: digit = ((k >> s1) & m1) | (((k >> s2) & m2) << b1).
struct DigitSpec { int s1, b1, s2, b2; };
__device__ __forceinline__ uint32_t key_digit(unsigned long long k, const DigitSpec &D)
{
	return ((uint32_t)(k >> D.s1) & ((1u << D.b1) - 1u)) | (((uint32_t)(k >> D.s2) & ((1u << D.b2) - 1u)) << D.b1);
}

// ---------------------------------------------------------------------------------------------------
// one-sweep radix pass: 16 B of HBM traffic per key and pass (read once, write once)
//
//   os_hist_kernel     one read of the keys fills the GLOBAL digit histograms of every pass (a stable pass does not change how
//                      many keys carry a digit value);
//   os_pass_kernel     a CTA takes the next tile (ticket from an atomic counter, so every predecessor tile is already running),
//                      ranks its keys per digit, publishes the tile's digit counts and obtains the number of keys with the same
//                      digit in all earlier tiles by decoupled look-back over the status words of its predecessors
//                      (status word = pass epoch : 32 | state : 2 | count : 30; state 1 = this tile's count, 2 = inclusive prefix
//                      up to this tile; a word of another epoch reads as "not there yet", so the array is never cleared),
//                      reorders the tile by digit in shared memory and writes every digit's run to its final place.
// The number of keys comes from DEVICE memory: the grid is sized for the largest possible count and surplus CTAs leave at once.
// Stability: tiles are ordered by ticket = tile index, ranks inside a tile follow the input order (warp, round, lane).
// ---------------------------------------------------------------------------------------------------
static constexpr int OS_THREADS = 256;			// thread t owns digits t, t + 256 in the per-digit steps
static constexpr int OS_WARPS = OS_THREADS / 32;
static constexpr int OS_KPT = SORT_TILE / OS_THREADS;	// 16 keys per thread
static constexpr int OS_MAX_PASSES = 8;
static constexpr uint32_t OS_FLAG_AGG = 1u << 30, OS_FLAG_PREFIX = 2u << 30, OS_COUNT_MASK = (1u << 30) - 1u;

struct DigitSpecs { DigitSpec d[OS_MAX_PASSES]; int np; };

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p)
{
	unsigned long long v;
	asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
	return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v)
{
	asm volatile("st.volatile.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// lane-privatised histogram copies (lane & (copies - 1)), skewed by one bank each: 8 copies of 257 words per pass for 8-bit
// digits, 4 copies of 513 words when a pass has 9 bits
template <int copies, int stride>
__global__ void __launch_bounds__(512) os_hist_kernel(const unsigned long long *__restrict__ keys, const unsigned long long *__restrict__ d_n, DigitSpecs P,
		uint32_t *__restrict__ ghist /* [np][RADIX_MAX] */)
{
	extern __shared__ __align__(16) unsigned char osh_smem[];
	uint32_t *h = reinterpret_cast<uint32_t *>(osh_smem);		// [np][copies][stride]
	const int copy = threadIdx.x & (copies - 1);
	constexpr int pstride = copies * stride;
	const uint64_t n = *d_n;

	for (int i = threadIdx.x; i < P.np * pstride; i += blockDim.x) h[i] = 0;
	__syncthreads();

	const uint64_t gstride = (uint64_t)gridDim.x * blockDim.x * 2;
	for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += gstride) {
		unsigned long long k0, k1 = 0;
		const bool two = i + 1 < n;
		if (two) { const ulonglong2 v = __ldcs(reinterpret_cast<const ulonglong2 *>(keys + i)); k0 = v.x; k1 = v.y; }
		else k0 = keys[i];
#pragma unroll
		for (int p = 0; p < OS_MAX_PASSES; ++p) {
			if (p < P.np) {
				uint32_t *hp = h + p * pstride + copy * stride;
				atomicAdd(hp + key_digit(k0, P.d[p]), 1u);
				if (two) atomicAdd(hp + key_digit(k1, P.d[p]), 1u);
			}
		}
	}
	__syncthreads();
	for (int j = threadIdx.x; j < P.np * RADIX_MAX; j += blockDim.x) {
		const int p = j >> RADIX_MAX_BITS, d = j & (RADIX_MAX - 1);
		if (d >= stride - 1) continue;
		uint32_t s = 0;
		for (int c = 0; c < copies; ++c) s += h[p * pstride + c * stride + d];
		if (s) atomicAdd(&ghist[j], s);
	}
}

// exclusive scan over the 256 threads of the CTA of a packed pair {hi: < 2^32, lo: < 2^16 summed}; smem >= 8 u64.
// *total = sum over all threads (same value in every thread)
__device__ __forceinline__ unsigned long long os_block_exclusive_scan(unsigned long long v, unsigned long long *smem, unsigned long long *total)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	unsigned long long incl = v;

#pragma unroll
	for (int off = 1; off < 32; off <<= 1) {
		const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, off);
		if (lane >= off) incl += t;
	}
	if (lane == 31) smem[wid] = incl;
	__syncthreads();
	unsigned long long woff = 0, tot = 0;
#pragma unroll
	for (int w = 0; w < OS_WARPS; ++w) { const unsigned long long x = smem[w]; if (w < wid) woff += x; if (total) tot += x; }
	if (total) *total = tot;
	return woff + incl - v;
}

template <int RBITS>
struct OneSweepSharedT
{
	static constexpr int RADIX = 1 << RBITS;
	unsigned long long	keys[SORT_TILE];		// tile reordered by digit
	uint32_t		whist[OS_WARPS][RADIX];		// per-warp digit counts, then exclusive prefix over the warps
	uint32_t		dstart[RADIX];			// tile-local start of each digit
	uint32_t		goff[RADIX];			// output index of tile-local position 0 of each digit's run (mod 2^32)
	unsigned long long	scan[2][OS_WARPS];
	float			fscan[OS_WARPS];
	uint32_t		tile;
};

template <int RBITS>
__global__ void __launch_bounds__(OS_THREADS, 4) os_pass_kernel(const unsigned long long *__restrict__ in, unsigned long long *__restrict__ out,
		const unsigned long long *__restrict__ d_n, DigitSpec D, const uint32_t *__restrict__ ghist /* [RADIX] of this pass */,
		unsigned long long *__restrict__ status /* [ntiles][RADIX] */, uint32_t *__restrict__ ticket, uint32_t epoch,
		int rank_mode /* 0 auto, 1 match.any, 2 ballots */)
{
	constexpr int RADIX = 1 << RBITS;
	constexpr int DPT = RADIX >= OS_THREADS ? RADIX / OS_THREADS : 1;	// digits per thread in the per-digit steps (threads >= RADIX idle there)
	extern __shared__ __align__(16) unsigned char os_smem[];
	OneSweepSharedT<RBITS> &S = *reinterpret_cast<OneSweepSharedT<RBITS> *>(os_smem);
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t lt_mask = (1u << lane) - 1u;
	const uint32_t n = (uint32_t)*d_n;
	const unsigned long long etag = (unsigned long long)epoch << 32;

	// persistent CTAs: the grid fills the machine once and every CTA keeps taking tile tickets until the keys are used up. The host
	// sizes nothing by the key count (it never reads it back): with most samples on the hot rows a batch may leave a fraction of
	// the tiles its event count would allow, and a CTA per POSSIBLE tile would spend more time starting and leaving than sorting.
	for (;;) {
	if (threadIdx.x == 0) S.tile = atomicAdd(ticket, 1u);
	for (int i = threadIdx.x; i < OS_WARPS * RADIX; i += OS_THREADS) (&S.whist[0][0])[i] = 0;
	__syncthreads();
	const uint32_t tile = S.tile;
	if ((uint64_t)tile * SORT_TILE >= n) return;		// no tile left
	const uint32_t wbase = tile * (uint32_t)SORT_TILE + (uint32_t)wid * (OS_KPT * 32);

	unsigned long long k[OS_KPT];
	uint32_t rk[OS_KPT / 2];			// two 16-bit ranks per word
#pragma unroll
	for (int r = 0; r < OS_KPT; ++r) {
		const uint32_t i = wbase + (uint32_t)r * 32 + lane;
		k[r] = i < n ? __ldcs(in + i) : 0ull;
	}

	// Lanes holding the same digit form a group. match.any finds the groups in one instruction, but the hardware walks the
	// distinct values of the warp one by one (ADU pipe: 73 % busy on a pass whose digits are uniform, ncu r01); one ballot per
	// digit bit costs the same whatever the data. The CTA picks per pass: expected number of distinct digits among 32 keys,
	// from the global histogram of the pass.
	bool use_ballot;
	{
		float distinct = 0.f;
#pragma unroll
		for (int j = 0; j < DPT; ++j) {
			if (threadIdx.x + j * OS_THREADS >= RADIX) continue;
			const float pd = (float)ghist[threadIdx.x + j * OS_THREADS] / (float)n;
			float q = 1.f - pd; q *= q; q *= q; q *= q; q *= q; q *= q;		// (1 - p)^32
			distinct += 1.f - q;
		}
#pragma unroll
		for (int off = 16; off > 0; off >>= 1) distinct += __shfl_xor_sync(0xffffffffu, distinct, off);
		if (lane == 0) S.fscan[wid] = distinct;
		__syncthreads();
		float tot = 0.f;
#pragma unroll
		for (int w = 0; w < OS_WARPS; ++w) tot += S.fscan[w];
		use_ballot = rank_mode == 2 || (rank_mode == 0 && tot > 10.f);
	}
	const bool partial = (tile + 1) * (uint32_t)SORT_TILE > n;

	// rank of every key among the keys of its digit inside this warp's chunk (rounds in order, lanes in order); the group
	// leader bumps the warp's digit counter and hands the previous value to its group
	uint32_t dg[OS_KPT / 2];			// two 16-bit digits per word (the VK digit costs a clz + shifts: computed once)
#pragma unroll
	for (int r = 0; r < OS_KPT; ++r) {
		const bool valid = wbase + (uint32_t)r * 32 + lane < n;
		const uint32_t d = valid ? key_digit(k[r], D) : ((uint32_t)RADIX + lane);
		if (r & 1) dg[r >> 1] |= d << 16; else dg[r >> 1] = d;
		uint32_t m;
		if (use_ballot) {
			m = 0xffffffffu;
#pragma unroll
			for (int b = 0; b < RBITS; ++b) {
				const bool bit = (d >> b) & 1u;
				const uint32_t bal = __ballot_sync(0xffffffffu, bit);
				m &= bit ? bal : ~bal;
			}
			if (partial) m &= __ballot_sync(0xffffffffu, valid);
		}
		else m = __match_any_sync(0xffffffffu, d);
		const int leader = __ffs(m) - 1;
		uint32_t old = 0;
		if (valid && lane == leader) { old = S.whist[wid][d]; S.whist[wid][d] = old + __popc(m); }
		__syncwarp();
		old = __shfl_sync(0xffffffffu, old, leader & 31);
		const uint32_t rank = old + __popc(m & lt_mask);
		if (r & 1) rk[r >> 1] |= rank << 16; else rk[r >> 1] = rank;
	}
	__syncthreads();

	// thread t owns digits t (+ 256): prefix over the warps, the tile's count of the digit -> published at once, so successors
	// can look back through it; then the scans over the digits, with a carry between the two halves of a 9-bit pass
	uint32_t dtotal[DPT];
#pragma unroll
	for (int j = 0; j < DPT; ++j) {
		const uint32_t d = threadIdx.x + j * OS_THREADS;
		uint32_t run = 0;
		dtotal[j] = 0;
		if (d >= RADIX) continue;
#pragma unroll
		for (int w = 0; w < OS_WARPS; ++w) { const uint32_t t = S.whist[w][d]; S.whist[w][d] = run; run += t; }
		dtotal[j] = run;
		st_volatile_u64(status + (size_t)tile * RADIX + d, etag | (tile == 0 ? OS_FLAG_PREFIX : OS_FLAG_AGG) | run);
	}
	unsigned long long carry = 0;
#pragma unroll
	for (int j = 0; j < DPT; ++j) {
		const uint32_t d = threadIdx.x + j * OS_THREADS;
		unsigned long long tot = 0;
		// {global count of digit d, tile count of digit d} -> exclusive scans over the digits in one go
		const bool own = d < RADIX;
		const unsigned long long sc = carry + os_block_exclusive_scan(own ? (((unsigned long long)ghist[d] << 16) | dtotal[j]) : 0ull, S.scan[j], DPT > 1 ? &tot : nullptr);
		carry += tot;
		if (!own) continue;
		const uint32_t gexcl = (uint32_t)(sc >> 16), dstart = (uint32_t)(sc & 0xFFFFu);
		// decoupled look-back: keys with digit d in the tiles before this one
		uint32_t excl = 0;
		if (tile > 0) {
			uint32_t p = tile - 1;
			for (;;) {
				const unsigned long long v = ld_volatile_u64(status + (size_t)p * RADIX + d);
				if ((v >> 32) != epoch || !((uint32_t)v >> 30)) continue;	// predecessor has its ticket, so it is running: its count will come
				excl += (uint32_t)v & OS_COUNT_MASK;
				if ((uint32_t)v & OS_FLAG_PREFIX) break;
				--p;
			}
			st_volatile_u64(status + (size_t)tile * RADIX + d, etag | OS_FLAG_PREFIX | (excl + dtotal[j]));
		}
		S.dstart[d] = dstart;
		S.goff[d] = gexcl + excl - dstart;
	}
	__syncthreads();

	// reorder the tile by digit in shared memory
#pragma unroll
	for (int r = 0; r < OS_KPT; ++r) {
		if (wbase + (uint32_t)r * 32 + lane < n) {
			const uint32_t dd = (r & 1) ? (dg[r >> 1] >> 16) : (dg[r >> 1] & 0xFFFFu);
			const uint32_t rank = (r & 1) ? (rk[r >> 1] >> 16) : (rk[r >> 1] & 0xFFFFu);
			S.keys[S.dstart[dd] + S.whist[wid][dd] + rank] = k[r];
		}
	}
	__syncthreads();

	// consecutive threads write consecutive keys: every digit's run leaves as full sectors
	const uint32_t tbase = tile * (uint32_t)SORT_TILE;
	const uint32_t nvalid = n - tbase < (uint32_t)SORT_TILE ? n - tbase : (uint32_t)SORT_TILE;
#pragma unroll 4
	for (uint32_t i = threadIdx.x; i < nvalid; i += OS_THREADS) {
		const unsigned long long key = S.keys[i];
		out[S.goff[key_digit(key, D)] + i] = key;
	}
	__syncthreads();		// the tile's shared state is free for the next ticket
	}
}

// ---------------------------------------------------------------------------------------------------
// per batch: sorted RESP keys -> runs (one per non-empty value bin of a service) -> histogram cells + t-digest
// ---------------------------------------------------------------------------------------------------
// (1) runs_mark_kernel   one sweep over the sorted keys: where do runs ({slot, bin} changes) and service segments start. A CTA
//     covers 1024 keys; the number of runs before it comes from a decoupled look-back over one status word per CTA (as in the
//     radix passes). Per run start: its pool entry is zeroed and its bin index recorded; per 128-key chunk: the index of the run
//     its first key belongs to; per service: its key segment, its run range, and its place in the list of touched services.
// (2) runs_sum_kernel    every 128-key chunk adds its samples into the pool entries of its runs: {samples | sub-msec remainders,
//     usec sum} — consecutive chunks hit consecutive entries, so the REDs stay in L2.
// (3) bins_merge_kernel  one warp per touched service: runs -> items {mean, weight} + GY_HISTOGRAM::add_data for every sample
//     of the run (count and exact msec sum per bucket), then the merging t-digest step.
struct RunRec { unsigned long long cw; unsigned long long us; };		// {samples : 27 | remainders : 37}, usec sum — as Bin
struct BatchSeg { uint32_t run0, nruns; uint32_t key0, nkeys; };		// a touched service's runs in the pool / keys in the sorted array

static constexpr int RM_V = 8;	// keys per thread: positions wbase + t * 32 + lane; a CTA of RM_THREADS threads covers RM_THREADS * RM_V keys

template <int RM_THREADS>
__global__ void __launch_bounds__(RM_THREADS) runs_mark_kernel(const unsigned long long *__restrict__ keys, const unsigned long long *__restrict__ d_n,
		unsigned long long *__restrict__ status, uint32_t epoch, RunRec *__restrict__ pool, uint16_t *__restrict__ run_bin,
		uint32_t *__restrict__ chunk_run, BatchSeg *__restrict__ segs /* [slot] */, uint32_t *__restrict__ touched, unsigned long long *ntouched,
		unsigned long long *nruns_total)
{
	constexpr int RM_TILE = RM_THREADS * RM_V;
	__shared__ uint32_t wsum[RM_THREADS / 32], s_excl;
	const uint64_t n = *d_n;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint64_t tbase = (uint64_t)blockIdx.x * RM_TILE;
	if (tbase >= n) return;
	const uint64_t wbase = tbase + (uint64_t)wid * (32 * RM_V);		// 128 consecutive keys per warp
	const unsigned long long etag = (unsigned long long)epoch << 32;

	unsigned long long g[RM_V + 2];		// {slot, bin} of key wbase - 1 (lane 0 only), own 4 keys, successor of the last (lane 31 only)
#pragma unroll
	for (int t = 0; t < RM_V; ++t) {
		const uint64_t i = wbase + (uint64_t)t * 32 + lane;
		g[1 + t] = i < n ? (keys[i] >> KEY_GROUP_SHIFT) : ~0ull;
	}
	g[0] = (lane == 0 && wbase && wbase < n) ? (keys[wbase - 1] >> KEY_GROUP_SHIFT) : ~0ull;
	g[RM_V + 1] = (lane == 31 && wbase + 32 * RM_V < n) ? (keys[wbase + 32 * RM_V] >> KEY_GROUP_SHIFT) : ~0ull;

	uint32_t words[RM_V];			// run-start ballots
	uint32_t isrun = 0, isseg = 0, isend = 0;	// bit t: own key t starts a run / starts / ends a service segment
	uint32_t wcount = 0;
#pragma unroll
	for (int t = 0; t < RM_V; ++t) {
		const uint64_t i = wbase + (uint64_t)t * 32 + lane;
		unsigned long long prev = __shfl_up_sync(0xffffffffu, g[1 + t], 1);
		const unsigned long long prev0 = __shfl_sync(0xffffffffu, g[t], 31);		// key t - 1 of lane 31 (t >= 1)
		if (lane == 0) prev = t == 0 ? g[0] : prev0;
		unsigned long long next = __shfl_down_sync(0xffffffffu, g[1 + t], 1);
		const unsigned long long next0 = __shfl_sync(0xffffffffu, g[2 + (t < RM_V - 1 ? t : 0)], 0);	// key t + 1 of lane 0
		if (lane == 31) next = t == RM_V - 1 ? g[RM_V + 1] : next0;
		const bool valid = i < n;
		const bool run = valid && g[1 + t] != prev;			// i == 0: prev is the sentinel
		const bool seg = valid && (g[1 + t] >> TD_CODE_BITS) != (prev >> TD_CODE_BITS);
		const bool end = valid && (g[1 + t] >> TD_CODE_BITS) != (next >> TD_CODE_BITS);	// last key: next is the sentinel
		words[t] = __ballot_sync(0xffffffffu, run);
		wcount += __popc(words[t]);
		isrun |= (run ? 1u : 0u) << t; isseg |= (seg ? 1u : 0u) << t; isend |= (end ? 1u : 0u) << t;
	}
	if (lane == 0) wsum[wid] = wcount;
	__syncthreads();
	uint32_t wexcl = 0, ctotal = 0;
#pragma unroll
	for (int w = 0; w < RM_THREADS / 32; ++w) { const uint32_t c = wsum[w]; if (w < wid) wexcl += c; ctotal += c; }
	// decoupled look-back over the CTAs before this one (warp 0, 32 predecessors per step): number of run starts before the tile
	if (wid == 0) {
		const uint32_t tile = blockIdx.x;
		if (lane == 0) st_volatile_u64(status + tile, etag | (tile == 0 ? OS_FLAG_PREFIX : OS_FLAG_AGG) | ctotal);
		uint32_t excl = 0;
		int p = (int)tile - 1;
		while (p >= 0) {
			const int idx = p - lane;
			const unsigned long long v = idx >= 0 ? ld_volatile_u64(status + idx) : (etag | OS_FLAG_PREFIX);	// before tile 0: prefix 0
			const bool ok = (v >> 32) == epoch && ((uint32_t)v >> 30) != 0;
			const uint32_t ready = __ballot_sync(0xffffffffu, ok), pfx = __ballot_sync(0xffffffffu, ok && ((uint32_t)v & OS_FLAG_PREFIX));
			const uint32_t upto = pfx ? (uint32_t)__ffs((int)pfx) : 32u;				// lanes 0 .. upto - 1 count
			const uint32_t need = upto == 32 ? 0xffffffffu : ((1u << upto) - 1u);
			if ((ready & need) != need) continue;							// a predecessor has its ticket, its count will come
			excl += __reduce_add_sync(0xffffffffu, (uint32_t)lane < upto ? ((uint32_t)v & OS_COUNT_MASK) : 0u);
			if (pfx) break;
			p -= 32;
		}
		if (lane == 0) {
			if (tile > 0) st_volatile_u64(status + tile, etag | OS_FLAG_PREFIX | (excl + ctotal));
			s_excl = excl;
			if (tbase + RM_TILE >= n) *nruns_total = (unsigned long long)excl + ctotal;
		}
	}
	__syncthreads();
	const uint32_t before_warp = s_excl + wexcl;		// run starts at positions < wbase

	// the run the first key of each 128-key chunk belongs to: #starts at positions <= the chunk's first, minus one
	if (lane == 0) {
		uint32_t before = before_warp;
#pragma unroll
		for (int c = 0; c < RM_V / 4; ++c) {
			chunk_run[(wbase >> 7) + c] = before + (words[4 * c] & 1u) - 1u;
			before += __popc(words[4 * c]) + __popc(words[4 * c + 1]) + __popc(words[4 * c + 2]) + __popc(words[4 * c + 3]);
		}
	}

	// one cursor bump per warp for all the service segments that start in it
	uint32_t nstart = __popc(isseg);
	uint32_t incl = nstart;
#pragma unroll
	for (int off = 1; off < 32; off <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += v; }
	const uint32_t wtotal = __shfl_sync(0xffffffffu, incl, 31);
	unsigned long long tb = 0;
	if (wtotal && lane == 0) tb = atomicAdd(ntouched, (unsigned long long)wtotal);
	tb = __shfl_sync(0xffffffffu, tb, 0) + (incl - nstart);

	uint32_t acc = before_warp;
#pragma unroll
	for (int t = 0; t < RM_V; ++t) {
		const uint64_t i = wbase + (uint64_t)t * 32 + lane;
		const uint32_t incl_here = acc + __popc(words[t] & (0xFFFFFFFFu >> (31 - lane)));	// run starts at positions <= i
		const uint32_t slot = (uint32_t)(g[1 + t] >> TD_CODE_BITS);
		if ((isrun >> t) & 1u) {
			pool[incl_here - 1] = RunRec {0, 0};
			run_bin[incl_here - 1] = (uint16_t)(g[1 + t] & ((1u << TD_CODE_BITS) - 1u));
		}
		if ((isseg >> t) & 1u) { segs[slot].run0 = incl_here - 1; segs[slot].key0 = (uint32_t)i; touched[tb++] = slot; }
		if ((isend >> t) & 1u) { segs[slot].nruns = incl_here; segs[slot].nkeys = (uint32_t)(i + 1); }	// ends for now: the merge kernel subtracts
		acc += __popc(words[t]);
	}
}

static constexpr int RS_V = 4;			// consecutive sorted samples per lane

__global__ void __launch_bounds__(256) runs_sum_kernel(const unsigned long long *__restrict__ keys, const unsigned long long *__restrict__ d_n,
		const uint32_t *__restrict__ chunk_run, RunRec *__restrict__ pool)
{
	const uint64_t n = *d_n;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * RS_V;
	const int lane = threadIdx.x & 31;

	for (uint64_t base = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x - lane) * RS_V; base < n; base += stride) {
		const uint64_t i0 = base + (uint64_t)lane * RS_V;
		unsigned long long kk[RS_V];
		if (i0 + RS_V <= n) {
			const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(keys + i0), c = *reinterpret_cast<const ulonglong2 *>(keys + i0 + 2);
			kk[0] = a.x; kk[1] = a.y; kk[2] = c.x; kk[3] = c.y;
		}
		else {
#pragma unroll
			for (int t = 0; t < RS_V; ++t) kk[t] = i0 + t < n ? keys[i0 + t] : ~0ull;
		}
		const uint32_t run_first = chunk_run[base >> 7];

		// whole chunk inside one run (a popular bin of a hot service): one RED pair for 128 samples
		const unsigned long long gfirst = __shfl_sync(0xffffffffu, kk[0], 0) >> KEY_GROUP_SHIFT, glast = __shfl_sync(0xffffffffu, kk[RS_V - 1], 31) >> KEY_GROUP_SHIFT;
		if (gfirst == glast && glast != (~0ull >> KEY_GROUP_SHIFT)) {
			unsigned long long us = 0;
			uint32_t rem = 0;
#pragma unroll
			for (int t = 0; t < RS_V; ++t) { const uint32_t v = key_usec(kk[t]); us += v; rem += v - (v / 1000u) * 1000u; }
			const unsigned long long gsum = (unsigned long long)__reduce_add_sync(0xffffffffu, (uint32_t)us & 0xFFFFFu) +
					((unsigned long long)__reduce_add_sync(0xffffffffu, (uint32_t)(us >> 20)) << 20);	// us < 2^32: 20 + 12 bits, x 32 lanes fits
			rem = __reduce_add_sync(0xffffffffu, rem);
			if (lane == 0) {
				red_add_u64(&pool[run_first].cw, (unsigned long long)(32 * RS_V) | ((unsigned long long)rem << BIN_CNT_BITS));
				red_add_u64(&pool[run_first].us, gsum);
			}
			continue;
		}

		// run index of every own sample: run of the chunk's first key + run starts at later positions up to the sample
		unsigned long long gprev = __shfl_up_sync(0xffffffffu, kk[RS_V - 1], 1) >> KEY_GROUP_SHIFT;		// last key of the previous lane
		uint32_t heads = 0, nh = 0;
#pragma unroll
		for (int t = 0; t < RS_V; ++t) {
			const unsigned long long gcur = kk[t] >> KEY_GROUP_SHIFT;
			const bool head = kk[t] != ~0ull && !(lane == 0 && t == 0) && gcur != gprev;
			heads |= (head ? 1u : 0u) << t; nh += head ? 1u : 0u;
			gprev = gcur;
		}
		uint32_t incl = nh;
#pragma unroll
		for (int off = 1; off < 32; off <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += v; }
		uint32_t run = run_first + (incl - nh);

		// own samples -> lane-local runs; sample t carries the totals of its run so far, only the last one of a run is emitted
		uint32_t pr[RS_V], pc[RS_V], prem[RS_V];
		unsigned long long ps[RS_V];
		bool tail[RS_V];
#pragma unroll
		for (int t = 0; t < RS_V; ++t) {
			const bool valid = kk[t] != ~0ull;
			tail[t] = valid;
			if ((heads >> t) & 1u) ++run;
			const uint32_t v = key_usec(kk[t]);
			pr[t] = valid ? run : (0x80000000u | (uint32_t)lane); pc[t] = valid ? 1u : 0u; ps[t] = valid ? v : 0u; prem[t] = valid ? v - (v / 1000u) * 1000u : 0u;
			if (t > 0 && valid && pr[t] == pr[t - 1]) { pc[t] += pc[t - 1]; ps[t] += ps[t - 1]; prem[t] += prem[t - 1]; tail[t - 1] = false; }
		}
#pragma unroll
		for (int t = 0; t < RS_V; ++t) {
			const bool act = tail[t];
			if (!__any_sync(0xffffffffu, act)) continue;
			const uint32_t rid = act ? pr[t] : (0x80000000u | (uint32_t)lane);
			const unsigned long long v = act ? ps[t] : 0ull;
			uint32_t cnt = act ? pc[t] : 0u, rem = act ? prem[t] : 0u;
			const uint32_t m = __match_any_sync(0xffffffffu, rid);
			unsigned long long gsum = v;
			const uint32_t maxcnt = __reduce_max_sync(0xffffffffu, (uint32_t)__popc(m));
			uint32_t rest = m & ~(1u << lane);
			const uint32_t cnt0 = cnt, rem0 = rem;
			for (uint32_t u = 1; u < maxcnt; ++u) {
				const int src = rest ? (__ffs(rest) - 1) : lane;
				const unsigned long long ov = __shfl_sync(0xffffffffu, v, src);
				const uint32_t oc = __shfl_sync(0xffffffffu, cnt0, src), orr = __shfl_sync(0xffffffffu, rem0, src);
				if (rest) { gsum += ov; cnt += oc; rem += orr; rest &= rest - 1; }
			}
			if (act && (m & ((1u << lane) - 1u)) == 0) {
				red_add_u64(&pool[rid].cw, (unsigned long long)cnt | ((unsigned long long)rem << BIN_CNT_BITS));
				red_add_u64(&pool[rid].us, gsum);
			}
		}
	}
}

static constexpr int TD_WARPS = 3;		// warps (= services in flight) per CTA; work area per warp: 10.4 KB at 384 entries, 13.8 KB at 512

// One warp per touched service. Its runs become the batch's items {mean = exact usec sum / samples, weight = samples} — in value
// order, because the bin index is monotone — and every run adds {samples, exact msec sum} to its bucket of the window histogram:
// GY_HISTOGRAM::add_data for each of its samples (common/gy_statistics.h:596-623); max_val_seen_ and the digest's ends come from
// the batch's exact extremes. The items are then merged with the old centroids (old first on equal means) and the greedy K_1
// pass cuts the list to at most TD_CAP clusters (warp_merge_compress). Lists of up to 2 x TD_CAP entries work in shared memory;
// longer ones (a first batch can fill several hundred bins) in the warp's L2-resident scratch — same code, same result.
// SMEM_N = longest merged list (old centroids + batch items) that works in shared memory: the smaller the work area, the more
// services a SM has in flight (the kernel is latency-bound: ~2 600 dependent warp instructions per service, ncu profiles/).
template <int SMEM_N>
__global__ void __launch_bounds__(TD_WARPS * 32) bins_merge_kernel(DevState st, const uint32_t *__restrict__ touched, const unsigned long long *__restrict__ ntouched_p,
		const RunRec *__restrict__ pool, const uint16_t *__restrict__ run_bin, const BatchSeg *__restrict__ segs,
		Centroid *__restrict__ items_scratch /* [nwarps][NBINS] */, TdWorkBig *__restrict__ big_scratch /* [nwarps] */)
{
	__shared__ TdWorkT<SMEM_N> work[TD_WARPS];
	// window histogram of the service's batch: 32-bit shared-memory atomics (native; a 64-bit shared atomicAdd is a CAS loop). A bucket's
	// sample count of one batch fits 32 bits (max_batch < 2^27); the msec sum is kept as {low word, carries + high words}
	__shared__ uint32_t hcnt[TD_WARPS][16], hsum_lo[TD_WARPS][16], hsum_hi[TD_WARPS][16];
	__shared__ uint16_t first_idx[16];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t gw = blockIdx.x * TD_WARPS + wid, nwarps = gridDim.x * TD_WARPS;

	Centroid *items = items_scratch + (size_t)gw * NBINS;
	const uint32_t ntouched = (uint32_t)*ntouched_p;

	if (threadIdx.x < 16) {
		// thresholds of RESP_TIME_HASH in msec (gy_statistics.h:1677): bucket b >= 2 starts at (thr[b-2] + 1) msec, bucket 1 at 0;
		// bin index = td_code(usec) + bucket, so bucket b's bins start at td_code(first usec of the bucket) + b
		constexpr uint32_t thr[13] = {1, 10, 30, 60, 100, 150, 200, 300, 450, 700, 1000, 3000, 15000};
		const uint32_t bk = threadIdx.x;
		const uint32_t first_us = bk < 2 ? 0u : (bk > 14 ? 0u : (thr[bk - 2] + 1u) * 1000u);
		first_idx[bk] = (uint16_t)(bk == 0 ? 0u : (bk > 14 ? 0xFFFFu : td_code(first_us) + bk));
	}
	__syncthreads();

	// the batch's work list: first the hot services (rows in use; the heavy ones start first), then the services runs_mark_kernel found
	// in the sorted keys. CTR_NHOT does not move while this kernel runs: rows handed out below count in CTR_NHOT_NEXT.
	const uint32_t nhot = st.hot_rows ? (uint32_t)st.counters[CTR_NHOT] : 0u;
	const uint32_t lt = (1u << lane) - 1u;
	for (uint32_t t = gw; t < nhot + ntouched; t += nwarps) {
		const bool is_hot = t < nhot;
		const uint32_t slot = is_hot ? st.hot_slot[t] : touched[t - nhot];
		const SlotBatch sb = st.slot_batch[slot];
		if (is_hot && sb.minv == 0xFFFFFFFFu) continue;			// a hot service without a sample in this batch (warp-uniform)
		if (lane < 16) { hcnt[wid][lane] = 0; hsum_lo[wid][lane] = 0; hsum_hi[wid][lane] = 0; }
		__syncwarp();
		// one non-empty bin {samples | remainders, usec sum} -> item j of the batch + GY_HISTOGRAM::add_data of its samples
		auto take_bin = [&](uint32_t j, unsigned long long cw, unsigned long long us, uint32_t bin) {
			const unsigned long long cnt = cw & BIN_CNT_MASK, rem = cw >> BIN_CNT_BITS;
			Centroid c; c.mean = __ddiv_rn((double)us, (double)cnt); c.weight = cnt;	// exact integer sum, one rounding
			items[j] = c;
			uint32_t bk = 0;
#pragma unroll
			for (int q = 1; q < 15; ++q) bk += bin >= first_idx[q];
			atomicAdd(&hcnt[wid][bk], (uint32_t)cnt);
			const unsigned long long ms = (us - rem) / 1000ull;		// sum of (usec / 1000) over the bin's samples
			const uint32_t mlo = (uint32_t)ms, mhi = (uint32_t)(ms >> 32);
			const uint32_t old = atomicAdd(&hsum_lo[wid][bk], mlo);
			const uint32_t up = mhi + (old + mlo < old ? 1u : 0u);		// carry out of the low word
			if (up) atomicAdd(&hsum_hi[wid][bk], up);
		};
		uint32_t nitems, nsamples, binmax = 0;		// binmax: samples in the fullest bin (per lane, reduced when needed)
		if (!is_hot) {
			const BatchSeg seg = segs[slot];
			nitems = seg.nruns - seg.run0;				// runs_mark_kernel left the END positions in nruns / nkeys
			nsamples = seg.nkeys - seg.key0;
			for (uint32_t j = lane; j < nitems; j += 32) {
				const RunRec r = pool[seg.run0 + j];
				take_bin(j, r.cw, r.us, run_bin[seg.run0 + j]);
				binmax = max(binmax, (uint32_t)(r.cw & BIN_CNT_MASK));
			}
		}
		else {
			// the service's dense row, in bin order (= value order); the row is left zeroed for the next batch
			unsigned long long *row = st.hot_rows + (size_t)t * HOT_ROW_WORDS;
			uint32_t mine = 0;
			nitems = 0;
			for (uint32_t b0 = 0; b0 < (uint32_t)NBINS; b0 += 32) {
				const uint32_t bin = b0 + lane;
				const uint32_t i0 = hot_word(bin), i1 = i0 + (uint32_t)HOT_ROW_BINS;
				ulonglong2 v = make_ulonglong2(0ull, 0ull);
				if (bin < (uint32_t)NBINS) { v.x = __ldcg(row + i0); v.y = __ldcg(row + i1); }
				const bool ne = (v.x & BIN_CNT_MASK) != 0;
				const uint32_t m = __ballot_sync(0xffffffffu, ne);
				if (ne) {
					take_bin(nitems + __popc(m & lt), v.x, v.y, bin);
					row[i0] = 0ull; row[i1] = 0ull;
					mine += (uint32_t)(v.x & BIN_CNT_MASK);
					binmax = max(binmax, (uint32_t)(v.x & BIN_CNT_MASK));
				}
				nitems += __popc(m);
			}
			nsamples = __reduce_add_sync(0xffffffffu, mine);
		}
		binmax = __reduce_max_sync(0xffffffffu, binmax);
		__syncwarp();
		// nobody else touches this slot's window histogram while the batch is merged (same stream as the flush): plain updates
		if (lane < HIST_MAX_CELL) {
			if (hcnt[wid][lane]) {
				HistCell *c = st.hist_cur + (size_t)slot * HIST_CELLS + lane;
				c->count += hcnt[wid][lane]; c->sum += (long long)(((unsigned long long)hsum_hi[wid][lane] << 32) | hsum_lo[wid][lane]);
			}
		}
		else if (lane == HIST_MAX_CELL) {
			HistCell *c = st.hist_cur + (size_t)slot * HIST_CELLS + HIST_MAX_CELL;
			const long long mx = (long long)(sb.maxv / 1000u);		// max_val_seen_ of add_data (gy_statistics.h:609-611)
			if (mx > c->sum) c->sum = mx;
			// a service that brought hot_min samples in one batch gets a row of dense bins for the batches to come (rows are never
			// taken back: a row belongs to the SLOT, whoever lives in it — routing only, the numbers are the same either way).
			// Not one whose fullest bin holds more than hot_bin_max samples: atomics on one 128-byte line are applied one after the
			// other (measured: ~8 ns each), a few hundred thousand of them on one line would outlast the rest of ingest_kernel —
			// such a service sorts well instead (its keys form long runs).
			uint32_t hot = sb.hot;
			if (!hot && st.hot_rows && nsamples >= st.hot_min && nsamples <= st.hot_max && binmax <= st.hot_bin_max) {
				const unsigned long long h = atomicAdd(st.counters + CTR_NHOT_NEXT, 1ull);
				if (h < (unsigned long long)st.hot_cap) { st.hot_slot[h] = slot; hot = (uint32_t)h + 1u; }
				else atomicAdd(st.counters + CTR_NHOT_NEXT, ~0ull);
			}
			st.slot_batch[slot] = SlotBatch {0xFFFFFFFFu, 0u, 0u, hot};
		}
		__syncwarp();

		TdHead head = st.td_head[slot];
		Centroid *cent = st.td_cent + (size_t)slot * TD_CAP;
		uint32_t nout;
		if (head.n + nitems <= (uint32_t)SMEM_N) nout = warp_merge_compress(work[wid], cent, head.n, items, nitems, cent, st.td);
		else nout = warp_merge_compress(big_scratch[gw], cent, head.n, items, nitems, cent, st.td);
		if (lane == 0) {
			head.n = nout;
			head.total += nsamples;
			if ((double)sb.minv < head.minv) head.minv = (double)sb.minv;
			if ((double)sb.maxv > head.maxv) head.maxv = (double)sb.maxv;
			st.td_head[slot] = head;
		}
		__syncwarp();
	}
}

// ---------------------------------------------------------------------------------------------------
// 5-second window roll: last = cur; all += cur; cur = 0  (one thread per histogram cell)
// ---------------------------------------------------------------------------------------------------
//
// Idle services: the reference deletes a listener that produced no statistics for TIMEOUT_INET_DIAG_SECS
// (300 s, common/gy_socket_stat.h:997) once it is older than twice that (common/gy_socket_stat.cc:3968-3982: tclock != 0,
// tclock + 300 s < now, tstart + 600 s < now) and tells madhava with LISTEN_FLAG_DELETE (:4023-4033). Here the flush records,
// per slot, the first flush that saw it and the last window that held events, and lists the slots that meet the rule.
__global__ void flush_kernel(DevState st, uint32_t nslots, HistCell *__restrict__ ring0, HistCell *__restrict__ ring1, uint32_t tsec, uint32_t idle_secs)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const bool valid = i < (uint64_t)nslots * HIST_CELLS;		// nslots * 16: whole half-warps are valid or not
	const int cell = (int)(i & (HIST_CELLS - 1));
	HistCell c {0, 0};
	unsigned long long cc = 0;
	const uint32_t slot = (uint32_t)(i >> 4);

	if (valid) {
		c = st.hist_cur[i];
		if (cell == HIST_MAX_CELL) cc = st.conn_cur[slot];
	}
	// did the closing window hold any event of this service? (16 lanes = the 15 buckets + the max / conn cell)
	bool auxact = false;				// ACTIVE_CONN_STATS records / API_TRAN errors count as activity of the window too
	if (valid && cell == HIST_MAX_CELL) { const SlotAux a0 = st.slot_aux[slot]; auxact = (a0.act_cur | a0.err_cur) != 0; }
	const uint32_t bal = __ballot_sync(0xffffffffu, valid && (cell == HIST_MAX_CELL ? (cc != 0 || auxact) : c.count != 0));
	const bool active = ((bal >> (threadIdx.x & 16)) & 0xFFFFu) != 0;
	if (!valid) return;

	st.bm_last[i] = st.bm_cur[i]; st.bm_cur[i] = 0;		// CONN_BITMAP::clear every 5 s (gy_socket_stat.h:436)
	st.hist_last[i] = c;
	// rolling levels: the window is added to the current slot of each level (cleared by the host when its epoch changed).
	// A cleared slot's max cell reads 0, which is below any recorded response time or equal to it: harmless for max().
	if (cell == HIST_MAX_CELL) {
		if (c.sum > ring0[i].sum) ring0[i].sum = c.sum;
		if (c.sum > ring1[i].sum) ring1[i].sum = c.sum;
	}
	else {
		ring0[i].count += c.count; ring0[i].sum += c.sum;
		ring1[i].count += c.count; ring1[i].sum += c.sum;
	}
	if (cell == HIST_MAX_CELL) {
		if (c.sum > st.hist_all[i].sum) st.hist_all[i].sum = c.sum;
		st.hist_cur[i].count = 0; st.hist_cur[i].sum = LLONG_MIN;
		st.conn_last[slot] = cc;
		st.conn_all_cnt[slot] += (uint32_t)cc;
		st.conn_all_kb[slot] += cc >> 32;
		st.conn_cur[slot] = 0;
		{
			SlotAux a = st.slot_aux[slot];
			a.act_last = a.act_cur; a.act_cur = 0; a.err_last = a.err_cur; a.err_cur = 0; a.rtt_last = a.rtt_cur; a.rtt_cur = 0;
			st.slot_aux[slot] = a;
		}

		const unsigned long long id = st.slot_id[slot];
		if (id) {
			uint32_t first = st.slot_first_seen[slot], last = st.slot_last_active[slot];
			if (!first) { first = tsec ? tsec : 1u; st.slot_first_seen[slot] = first; }
			if (active) { last = tsec ? tsec : 1u; st.slot_last_active[slot] = last; }
			if (idle_secs && last && (uint64_t)last + idle_secs < tsec && (uint64_t)first + 2ull * idle_secs < tsec) {
				const unsigned long long k = atomicAdd(st.counters + CTR_NEVICT, 1ull);
				st.evict_list[k] = slot;
				st.evict_ids[k] = id;
			}
		}
	}
	else {
		st.hist_all[i].count += c.count;
		st.hist_all[i].sum += c.sum;
		st.hist_cur[i].count = 0; st.hist_cur[i].sum = 0;
	}
}

// ---------------------------------------------------------------------------------------------------
// The state decision of the 5-s reducer, one thread per service slot, right behind flush_kernel: what listener_stats_update does per
// listener around get_curr_state (common/gy_socket_stat.cc:4111-4130 samples of qps_hist_ / active_conn_hist_, :4158-4173 connection
// counts, :4222-4233 level statistics + get_curr_state, :4242-4272 issue_bit_hist_). A slot whose closing window held no event is
// "stale" (:4098-4107): the reference leaves such a listener's state alone, and so does this kernel.
// ---------------------------------------------------------------------------------------------------
struct LevelStat { int64_t p95, p99, p25; uint64_t cnt, sum; double mean; };

__device__ __forceinline__ LevelStat level_stat(const uint64_t *counts, uint64_t sum)
{
	LevelStat r;
	uint64_t c = 0;
	for (int b = 0; b < HIST_MAX_CELL; ++b) c += counts[b];
	r.p95 = resp_bucket_value(hist_pct_bucket(counts, HIST_MAX_CELL, c, 95.0f), c);
	r.p99 = resp_bucket_value(hist_pct_bucket(counts, HIST_MAX_CELL, c, 99.0f), c);
	r.p25 = resp_bucket_value(hist_pct_bucket(counts, HIST_MAX_CELL, c, 25.0f), c);
	r.cnt = c; r.sum = sum;
	r.mean = (double)(long long)sum / (double)(c ? (long long)c : 1ll);		// TIME_HISTOGRAM::get_stats, gy_statistics.h:1358
	return r;
}

__global__ void __launch_bounds__(128) state_kernel(DevState st, uint32_t nslots, uint32_t tsec, uint32_t live0, uint32_t live1)
{
	const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= nslots || !st.slot_id[slot]) return;
	if (st.slot_last_active[slot] != (tsec ? tsec : 1u)) return;			// stale window: state stays

	const size_t base = (size_t)slot * HIST_CELLS;
	uint64_t cnt[HIST_MAX_CELL];
	uint64_t sum;
	gysk_listener_state_in in;
	memset(&in, 0, sizeof(in));

	sum = 0;
	for (int b = 0; b < HIST_MAX_CELL; ++b) { const HistCell c = st.hist_last[base + b]; cnt[b] = c.count; sum += (uint64_t)c.sum; }
	const LevelStat s5 = level_stat(cnt, sum);
	LevelStat lv[NLEVELS];
	for (int l = 0; l < NLEVELS; ++l) {
		const uint32_t live = l ? live1 : live0;
		for (int b = 0; b < HIST_MAX_CELL; ++b) cnt[b] = 0;
		sum = 0;
		for (int k = 0; k < NSLOTS; ++k) {
			if (!((live >> k) & 1u)) continue;
			const HistCell *ring = st.hist_ring + (((size_t)l * NSLOTS + k) * nslots + slot) * HIST_CELLS;
			for (int b = 0; b < HIST_MAX_CELL; ++b) { const HistCell c = ring[b]; cnt[b] += c.count; sum += (uint64_t)c.sum; }
		}
		lv[l] = level_stat(cnt, sum);
	}
	sum = 0;
	for (int b = 0; b < HIST_MAX_CELL; ++b) { const HistCell c = st.hist_all[base + b]; cnt[b] = c.count; sum += (uint64_t)c.sum; }
	const LevelStat sa = level_stat(cnt, sum);

	in.r5p95 = s5.p95; in.r5p99 = s5.p99; in.nqrys_5s = s5.cnt; in.total_resp_msec = s5.sum; in.mean5 = s5.mean;
	in.r300p95 = lv[0].p95; in.r300p99 = lv[0].p99; in.mean300 = lv[0].mean;
	in.r5dp95 = lv[1].p95; in.r5dp99 = lv[1].p99; in.r5dp25 = lv[1].p25; in.tcount_5d = lv[1].cnt; in.mean5d = lv[1].mean;
	in.rallp95 = sa.p95; in.rallp99 = sa.p99; in.meanall = sa.mean;

	SlotState ss = st.slot_state[slot];
	const SlotAux aux = st.slot_aux[slot];
	in.last_qps_count = (int32_t)(s5.cnt / 5);
	if ((uint32_t)aux.act_last) ss.nconn_active = (uint32_t)aux.act_last;		// an ACTIVE_CONN_STATS report arrived in this window

	// GY_HISTOGRAM<int, ...>::add_data of the two per-window samples, then their p95 / p25
	{
		HistCell *q = st.qps_hist + base, *a = st.act_hist + base;
		const int qv = in.last_qps_count, av = (int)ss.nconn_active;
		const int qb = bucket_semi_log_lo(qv), ab = bucket_hash_1_3000(av);
		q[qb].count += 1; q[qb].sum += qv; if ((long long)qv > q[HIST_MAX_CELL].sum) q[HIST_MAX_CELL].sum = qv;
		a[ab].count += 1; a[ab].sum += av; if ((long long)av > a[HIST_MAX_CELL].sum) a[HIST_MAX_CELL].sum = av;
		uint64_t tot = 0;
		for (int b = 0; b < HIST_MAX_CELL; ++b) { cnt[b] = q[b].count; tot += cnt[b]; }
		in.qps_p95 = qps_bucket_value(hist_pct_bucket(cnt, HIST_MAX_CELL, tot, 95.0f), tot);
		in.qps_p25 = qps_bucket_value(hist_pct_bucket(cnt, HIST_MAX_CELL, tot, 25.0f), tot);
		tot = 0;
		for (int b = 0; b < HIST_MAX_CELL; ++b) { cnt[b] = b < 14 ? a[b].count : 0; tot += cnt[b]; }
		in.act_p95 = act_bucket_value(hist_pct_bucket(cnt, 14, tot, 95.0f), tot);
		in.act_p25 = act_bucket_value(hist_pct_bucket(cnt, 14, tot, 25.0f), tot);
	}

	in.nconn = (int32_t)ss.nconn_active;
	in.curr_active_conn = in.nconn;
	for (int b = 0; b < HIST_MAX_CELL; ++b) {
		const int c = __popc(st.bm_last[base + b]);					// CONN_BITMAP::get_conn_breakup
		in.nactive_conn_arr[b] = (uint8_t)c;
		if (in.curr_active_conn < c) in.curr_active_conn = c;
	}
	in.ser_errors = (uint32_t)(aux.err_last >> 32);
	const uint32_t first = st.slot_first_seen[slot];
	const uint32_t age = tsec > first ? tsec - first : 0u;
	in.secs_5d = age + 1u < 432000u ? age + 1u : 432000u;

	classify_listener(in, ss.high_bits, ss.state, ss.issue);
	apply_issue_history(age, in.ser_errors, ss.state, ss.issue, ss.issue_bits);
	st.slot_state[slot] = ss;
}

// one CTA per evicted slot (grid-stride): the id's table entry becomes a tombstone, every per-slot array returns to its
// just-created state and the slot number goes on the free stack for the next unknown id
__global__ void __launch_bounds__(256) evict_kernel(DevState st, uint32_t max_svcs)
{
	const uint32_t nev = (uint32_t)st.counters[CTR_NEVICT];

	for (uint32_t q = blockIdx.x; q < nev; q += gridDim.x) {
		const uint32_t slot = st.evict_list[q];
		const unsigned long long id = st.evict_ids[q];

		if (threadIdx.x == 0) {
			uint32_t pos = table_hash(id) & st.svc_tbl.mask;
			for (uint32_t probe = 0; probe <= st.svc_tbl.mask; ++probe, pos = (pos + 1) & st.svc_tbl.mask) {
				TblEntry *e = &st.svc_tbl.ent[pos];
				if (e->key == id) { e->key = KEY_TOMBSTONE; e->slot1 = 0; break; }
				if (e->key == 0) break;
			}
			st.slot_id[slot] = 0; st.slot_host[slot] = 0; st.slot_first_seen[slot] = 0; st.slot_last_active[slot] = 0;
			st.conn_cur[slot] = 0; st.conn_last[slot] = 0; st.conn_all_cnt[slot] = 0; st.conn_all_kb[slot] = 0;
			st.slot_aux[slot] = SlotAux {0, 0, 0, 0, 0, 0};
			st.slot_state[slot] = SlotState {GYSK_STATE_OK, GYSK_ISSUE_NONE, 0, 0, 0};
			TdHead h; h.total = 0; h.minv = INFINITY; h.maxv = -INFINITY; h.n = 0; h.pad = 0;
			st.td_head[slot] = h;
			const int32_t f = atomicAdd(st.svc_tbl.free_n, 1);
			st.svc_tbl.free_slots[f] = slot;
			atomicAdd(st.counters + CTR_EVICTED_TOTAL, 1ull);
		}
		if (threadIdx.x < HIST_CELLS) {
			const size_t c = (size_t)slot * HIST_CELLS + threadIdx.x;
			const HistCell z {0, threadIdx.x == HIST_MAX_CELL ? LLONG_MIN : 0};
			st.hist_cur[c] = z; st.hist_last[c] = z; st.hist_all[c] = z;
			st.bm_cur[c] = 0; st.bm_last[c] = 0;
			const HistCell zi {0, threadIdx.x == HIST_MAX_CELL ? (long long)INT_MIN : 0};
			st.qps_hist[c] = zi; st.act_hist[c] = zi;
			for (int pl = 0; pl < NLEVELS * NSLOTS; ++pl) st.hist_ring[((size_t)pl * max_svcs + slot) * HIST_CELLS + threadIdx.x] = HistCell {0, 0};
		}
		uint32_t *hw = reinterpret_cast<uint32_t *>(st.hll + ((size_t)slot << st.hll_p));
		for (uint32_t w = threadIdx.x; w < (1u << st.hll_p) / 4u; w += blockDim.x) hw[w] = 0;
		for (uint32_t w = threadIdx.x; w < (uint32_t)TD_CAP; w += blockDim.x) st.td_cent[(size_t)slot * TD_CAP + w] = Centroid {0.0, 0ull};
	}
}

// tombstones only go away by rebuilding: clear the table, re-insert the ids of the live slots (slot numbers stay)
__global__ void rebuild_table_kernel(DevState st, uint32_t max_svcs)
{
	const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= max_svcs) return;
	const unsigned long long id = st.slot_id[slot];
	if (!id) return;
	uint32_t pos = table_hash(id) & st.svc_tbl.mask;
	for (;;) {
		const unsigned long long k = atomicCAS(&st.svc_tbl.ent[pos].key, 0ull, id);
		if (k == 0) { st.svc_tbl.ent[pos].slot1 = slot + 1; return; }
		pos = (pos + 1) & st.svc_tbl.mask;
	}
}

// ---------------------------------------------------------------------------------------------------
// read side: one warp per queried id
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) gather_svcs_kernel(DevState st, const unsigned long long *__restrict__ ids, uint32_t n, uint32_t max_svcs,
		uint32_t live0, uint32_t live1, SvcRaw *__restrict__ out)
{
	__shared__ uint32_t hh[4][64];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t q = blockIdx.x * 4 + wid;

	if (q >= n) return;
	SvcRaw &o = out[q];
	const unsigned long long id = ids[q];
	int slot = -1;
	if (lane == 0 && id) slot = table_lookup(st.svc_tbl, id, false);
	slot = __shfl_sync(0xffffffffu, slot, 0);
	if (lane == 0) { o.id = id; o.found = slot >= 0; o.slot = (uint32_t)slot; }
	if (slot < 0) return;

	if (lane < HIST_CELLS) {
		o.cur[lane] = st.hist_cur[(size_t)slot * HIST_CELLS + lane];
		o.last[lane] = st.hist_last[(size_t)slot * HIST_CELLS + lane];
		o.all[lane] = st.hist_all[(size_t)slot * HIST_CELLS + lane];
		o.bm_cur[lane] = st.bm_cur[(size_t)slot * HIST_CELLS + lane]; o.bm_last[lane] = st.bm_last[(size_t)slot * HIST_CELLS + lane];
		// rolling levels: sum of the slots still inside the level's span
		for (int l = 0; l < NLEVELS; ++l) {
			const uint32_t live = l ? live1 : live0;
			HistCell a {0, 0};
			if (lane == HIST_MAX_CELL) a.sum = LLONG_MIN;
			for (int k = 0; k < NSLOTS; ++k) {
				if (!((live >> k) & 1u)) continue;
				const HistCell x = st.hist_ring[(((size_t)l * NSLOTS + k) * max_svcs + slot) * HIST_CELLS + lane];
				if (lane == HIST_MAX_CELL) a.sum = max(a.sum, x.sum);
				else { a.count += x.count; a.sum += x.sum; }
			}
			o.lvl[l][lane] = a;
		}
	}
	if (lane == 0) {
		o.conn_cur = st.conn_cur[slot]; o.conn_last = st.conn_last[slot];
		o.conn_all_cnt = st.conn_all_cnt[slot]; o.conn_all_kb = st.conn_all_kb[slot];
		o.td = st.td_head[slot];
		o.aux = st.slot_aux[slot];
		o.sst = st.slot_state[slot];
	}
	if (lane < HIST_CELLS) { o.qps[lane] = st.qps_hist[(size_t)slot * HIST_CELLS + lane]; o.act[lane] = st.act_hist[(size_t)slot * HIST_CELLS + lane]; }
	for (int i = lane; i < TD_CAP; i += 32) o.cent[i] = st.td_cent[(size_t)slot * TD_CAP + i];

	hh[wid][lane] = 0; hh[wid][lane + 32] = 0;
	__syncwarp();
	const uint8_t *regs = st.hll + ((size_t)slot << st.hll_p);
	for (uint32_t i = lane; i < (1u << st.hll_p); i += 32) atomicAdd(&hh[wid][regs[i] > 63 ? 63 : regs[i]], 1u);
	__syncwarp();
	o.hll_hist[lane] = hh[wid][lane]; o.hll_hist[lane + 32] = hh[wid][lane + 32];
}

__global__ void gather_tasks_kernel(DevState st, const unsigned long long *__restrict__ ids, uint32_t n, TaskRaw *__restrict__ out)
{
	const uint32_t q = blockIdx.x;
	if (q >= n) return;
	__shared__ int sslot;
	if (threadIdx.x == 0) {
		const unsigned long long id = ids[q];
		sslot = id ? table_lookup(st.task_tbl, id, false) : -1;
		out[q].id = id; out[q].found = sslot >= 0; out[q].slot = (uint32_t)sslot;
	}
	__syncthreads();
	if (sslot < 0) return;
	if (threadIdx.x < 3 * HIST_CELLS) (&out[q].h[0][0])[threadIdx.x] = st.task_hist[(size_t)sslot * 3 * HIST_CELLS + threadIdx.x];
}

__global__ void gather_hll_kernel(DevState st, unsigned long long id, uint8_t *__restrict__ out, int32_t *found)
{
	__shared__ int sslot;
	if (threadIdx.x == 0) { sslot = id ? table_lookup(st.svc_tbl, id, false) : -1; *found = sslot >= 0; }
	__syncthreads();
	if (sslot < 0) return;
	const uint8_t *regs = st.hll + ((size_t)sslot << st.hll_p);
	for (uint32_t i = threadIdx.x; i < (1u << st.hll_p); i += blockDim.x) out[i] = regs[i];
}

__global__ void query_flows_kernel(DevState st, const unsigned long long *__restrict__ keys, uint32_t n, int last_window, gysk_flow_est *__restrict__ out)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const unsigned long long *tbl = last_window ? st.cms_last : st.cms_cur;
	const unsigned long long key = keys[i];
	uint32_t cnt = 0xFFFFFFFFu, kb = 0xFFFFFFFFu;

	for (uint32_t r = 0; r < st.cms_depth; ++r) {
		const unsigned long long c = tbl[((size_t)r << st.cms_log2w) + cms_index(key, r, st.cms_wmask)];
		cnt = min(cnt, (uint32_t)c);
		kb = min(kb, (uint32_t)(c >> 32));
	}
	out[i].flow_key = key; out[i].count = cnt; out[i].kbytes = kb;
}

// ---------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------
static inline uint32_t div_up(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

// cudaFuncSetAttribute is per DEVICE: one process may own an engine on every GPU of the box
static constexpr int MAX_DEVICES = 64;
static int current_device() { int dev = 0; cudaGetDevice(&dev); return dev < 0 || dev >= MAX_DEVICES ? 0 : dev; }
static int sm_count(int dev) { int nsm = 148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev); return nsm; }

int launch_init_state(const DevState &st, uint32_t max_svcs, uint32_t max_tasks, cudaStream_t s)
{
	const uint32_t m = max_svcs > max_tasks ? max_svcs : max_tasks;
	init_state_kernel<<<div_up(m, 256), 256, 0, s>>>(st, max_svcs, max_tasks);
	return 1;
}

int launch_register(const DevState &st, const unsigned long long *d_ids, uint32_t n, int is_task, cudaStream_t s)
{
	if (!n) return 0;
	register_kernel<<<div_up(n, 256), 256, 0, s>>>(st, d_ids, n, is_task);
	return 1;
}

// GYSK_INGEST_VARIANT selects a shape of the warp-autonomous ingest kernel for A/B runs
static int ingest_variant()
{
	static const int v = []{ const char *e = getenv("GYSK_INGEST_VARIANT"); return e ? atoi(e) : 0; }();
	return v;
}

// the radix passes of the RESP keys sort on {slot | bin} = key bits [30, 40 + slot bits): TD_CODE_BITS + slot bits significant
// bits cut into the fewest digits of at most KEY_DIGIT_MAX bits, widths as even as possible (27 bits -> 7 7 7 6). 8-bit passes run
// at 0.39-0.49 ms per 70 M keys, a 9-bit pass (two look-back rows per thread, nine ballots per key) at 0.76 ms (profiles/): four
// of the former beat three of the latter.
static int key_sort_plan(uint32_t max_svcs, SortPlan &P)
{
	int slot_bits = 1;
	while (slot_bits < 24 && (1ull << slot_bits) < max_svcs) slot_bits++;
	const int T = TD_CODE_BITS + slot_bits;
	static const int dmax = []{ const char *e = getenv("GYSK_KEY_DIGIT_MAX"); const int v = e ? atoi(e) : 8; return v == 9 ? 9 : 8; }();
	const int np = (T + dmax - 1) / dmax;
	if (np > OS_MAX_PASSES_VK) return -1;
	int at = KEY_GROUP_SHIFT;
	for (int p = 0; p < np; ++p) { P.bits[p] = T / np + (p < T % np ? 1 : 0); P.shift[p] = at; at += P.bits[p]; }
	P.np = np;
	static const int exp = []{ const char *e = getenv("GYSK_EXP_ABLATE"); return e ? atoi(e) : 0; }();
	P.exp = exp;
	return np;
}

// GYSK_SIDE_DRAIN=1: the connection / process records are queued by ingest_kernel and applied by side_drain_kernel on a second
// stream next to the sort chain. Measured and NOT the default (profiles/r02_side_drain_ab.json): ingest_kernel gets 0.4 ms faster,
// but the block scheduler does not interleave the side kernel with the radix passes (they fill every SM's register file), so its
// 1.5 - 2.5 ms land behind the chain: 8.1 ms per batch against 6.9 ms with the drains inside ingest_kernel.
bool side_drain_enabled()
{
	static const bool on = []{ const char *e = getenv("GYSK_SIDE_DRAIN"); return e && atoi(e) != 0; }();
	return on;
}

template <int WARPS, int MIN_CTAS, int EPT, bool TMA, int DH, bool SIDE>
static void launch_ingest_variant(const DevState &st, const gysk_event *d_ev, uint64_t n, unsigned long long *d_keys, uint32_t *ghist, const SortPlan &plan,
		uint4 *tcpq, uint4 *taskq, int dev, cudaStream_t s)
{
	using Shared = IngestSharedT<WARPS, EPT, TMA, DH, SIDE>;
	static bool attr_set[MAX_DEVICES] = {};
	if (!attr_set[dev]) {
		cudaFuncSetAttribute(ingest_kernel<WARPS, MIN_CTAS, EPT, TMA, DH, SIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Shared));
		attr_set[dev] = true;
	}
	const uint64_t want = (n + (uint64_t)Shared::CHUNK * WARPS - 1) / ((uint64_t)Shared::CHUNK * WARPS);
	const uint64_t full = (uint64_t)sm_count(dev) * MIN_CTAS;
	ingest_kernel<WARPS, MIN_CTAS, EPT, TMA, DH, SIDE><<<(uint32_t)(want < full ? want : full), WARPS * 32, sizeof(Shared), s>>>(st, d_ev, n, d_keys, ghist, plan, tcpq, taskq);
}

int launch_ingest(const DevState &st, const SortTemp &tmp, const gysk_event *d_ev, uint64_t n, uint32_t max_svcs, cudaStream_t s)
{
	if (!n) return 0;
	const int dev = current_device();
	SortPlan plan {};
	if (key_sort_plan(max_svcs, plan) < 0) return -1;
	// the hot rows handed out by the batches so far are in use from this batch on
	if (st.hot_rows) cudaMemcpyAsync(st.counters + CTR_NHOT, st.counters + CTR_NHOT_NEXT, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s);
	// key cursor, digit histograms and tile tickets of this batch's sort
	cudaMemsetAsync(st.counters + CTR_NKEYS, 0, sizeof(unsigned long long), s);
	cudaMemsetAsync(tmp.os_ghist, 0, (OS_MAX_PASSES * RADIX_MAX + OS_MAX_PASSES) * sizeof(uint32_t), s);
	bool wide = false;
	for (int p = 0; p < plan.np; ++p) wide |= plan.bits[p] > 8;
	const bool side = side_drain_enabled();
	if (side) cudaMemsetAsync(st.counters + CTR_NTCPQ, 0, 2 * sizeof(unsigned long long), s);
#define GYSK_LI2(W, C, E, T, D, SD) launch_ingest_variant<W, C, E, T, D, SD>(st, d_ev, n, tmp.keys_a, tmp.os_ghist, plan, tmp.tcpq, tmp.taskq, dev, s)
#define GYSK_LI(W, C, E, T) do { if (wide) GYSK_LI2(W, C, E, T, 512, false); else GYSK_LI2(W, C, E, T, 256, false); } while (0)
	if (side) {			// the shipped shapes; the experiment variants below keep the in-kernel drains
		const int v = ingest_variant();
		if (wide) GYSK_LI2(8, 3, 2, false, 512, true);
		else if (v == 842) GYSK_LI2(8, 4, 2, false, 256, true);
		else GYSK_LI2(8, 3, 2, false, 256, true);
		return 1;
	}
	switch (ingest_variant()) {
	case 842 : GYSK_LI(8, 4, 2, false); break;
	case 852 : GYSK_LI(8, 5, 2, false); break;
	case 834 : GYSK_LI(8, 3, 4, false); break;
	case 844 : GYSK_LI(8, 4, 4, false); break;
	case 832 : GYSK_LI(8, 3, 2, false); break;
	case 482 : GYSK_LI(4, 8, 2, false); break;
	case 1832 : GYSK_LI(8, 3, 2, true); break;	// TMA-staged chunks
	case 1834 : GYSK_LI(8, 3, 4, true); break;
	default : GYSK_LI(8, 3, 2, false); break;
	}
#undef GYSK_LI
#undef GYSK_LI2
	return 1;
}

// the batch's queued connection / process records -> count-min, HLL, exact cells, process histograms (side stream)
int launch_side_drain(const DevState &st, const SortTemp &tmp, uint64_t n_events, cudaStream_t s)
{
	if (!n_events || !side_drain_enabled()) return 0;
	static const int per_sm = []{ const char *e = getenv("GYSK_SIDE_CTAS"); const int v = e ? atoi(e) : 1; return v >= 1 && v <= 8 ? v : 1; }();
	const uint64_t want = (n_events + SD_WARPS * 32 - 1) / (SD_WARPS * 32);
	const uint64_t full = (uint64_t)sm_count(current_device()) * per_sm;
	side_drain_kernel<<<(uint32_t)(want < full ? want : full), SD_WARPS * 32, 0, s>>>(st, tmp.tcpq, tmp.taskq);
	return 1;
}

// plain-key digit plan: the significant bits [lo1, hi1) then [lo2, hi2) (lo2 >= hi1; hi2 <= lo2 for a single range) cut into
// digits in order, a digit may straddle the gap. Digits are 8 bits wide unless 9-bit digits save a whole pass.
static int build_digit_specs(int lo1, int hi1, int lo2, int hi2, DigitSpecs &P)
{
	int p1 = lo1, p2 = lo2;			// next unsorted bit of each range

	P.np = 0;
	if (hi2 < lo2) hi2 = lo2;
	const int T = (hi1 - lo1) + (hi2 - lo2);
	const int p8 = (T + 7) / 8, p9 = (T + 8) / 9;
	int wide = p9 < p8 ? T - 8 * p9 : 0;	// number of 9-bit passes (the first ones)
	while ((p1 < hi1 || p2 < hi2) && P.np < OS_MAX_PASSES) {
		DigitSpec D {0, 0, 0, 0};
		int need = wide > 0 ? 9 : 8;
		if (wide > 0) --wide;
		if (p1 < hi1) { D.s1 = p1; D.b1 = hi1 - p1 < need ? hi1 - p1 : need; p1 += D.b1; need -= D.b1; }
		if (need && p1 >= hi1 && p2 < hi2) {
			const int take = hi2 - p2 < need ? hi2 - p2 : need;
			if (D.b1) { D.s2 = p2; D.b2 = take; } else { D.s1 = p2; D.b1 = take; }
			p2 += take;
		}
		P.d[P.np++] = D;
	}
	return (p1 < hi1 || p2 < hi2) ? -1 : 0;		// more than 64 significant bits cannot happen
}

// host-side view of the pass plan (C ABI: gysk_sort_plan): per pass {shift1, bits1, shift2, bits2}
int radix_sort_plan(int lo1, int hi1, int lo2, int hi2, int out[][4], int cap)
{
	DigitSpecs P;
	if (build_digit_specs(lo1, hi1, lo2, hi2, P)) return -1;
	for (int p = 0; p < P.np && p < cap; ++p) { out[p][0] = P.d[p].s1; out[p][1] = P.d[p].b1; out[p][2] = P.d[p].s2; out[p][3] = P.d[p].b2; }
	return P.np;
}

static void os_set_attrs(int dev)
{
	static bool attr_set[MAX_DEVICES] = {};
	if (attr_set[dev]) return;
	cudaFuncSetAttribute(os_pass_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OneSweepSharedT<6>));
	cudaFuncSetAttribute(os_pass_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OneSweepSharedT<7>));
	cudaFuncSetAttribute(os_pass_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OneSweepSharedT<8>));
	cudaFuncSetAttribute(os_pass_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OneSweepSharedT<9>));
	cudaFuncSetAttribute(os_hist_kernel<8, 257>, cudaFuncAttributeMaxDynamicSharedMemorySize, OS_MAX_PASSES * 8 * 257 * (int)sizeof(uint32_t));
	cudaFuncSetAttribute(os_hist_kernel<4, 513>, cudaFuncAttributeMaxDynamicSharedMemorySize, OS_MAX_PASSES * 4 * 513 * (int)sizeof(uint32_t));
	attr_set[dev] = true;
}

static uint32_t next_epoch(const SortTemp &tmp, uint32_t max_tiles, cudaStream_t s)
{
	uint32_t e = ++*tmp.epoch;
	if (e == 0) {		// 2^32 passes later: stale words could pass for current ones, start over
		cudaMemsetAsync(tmp.tile_status, 0, (size_t)max_tiles * RADIX_MAX * sizeof(unsigned long long), s);
		e = ++*tmp.epoch;
	}
	return e;
}

static const int g_rank_mode = []{ const char *e = getenv("GYSK_OS_RANK"); return e ? atoi(e) : 0; }();

// one radix pass with the kernel instantiation of the digit's width: a 7-bit digit has half the per-digit work (per-warp counters to
// clear and prefix, status words to publish and look back through, ballots per key) of an 8-bit one
static void launch_os_pass(int bits, uint32_t ntiles, const unsigned long long *in, unsigned long long *out, const unsigned long long *d_n, const DigitSpec &D,
		const uint32_t *ghist, unsigned long long *status, uint32_t *ticket, uint32_t epoch, cudaStream_t s)
{
	static const bool narrow = []{ const char *e = getenv("GYSK_OS_NARROW"); return !e || atoi(e) != 0; }();
	// GYSK_OS_PERSIST=0: one CTA per possible tile (A/B runs)
	static const bool persist = []{ const char *e = getenv("GYSK_OS_PERSIST"); return !e || atoi(e) != 0; }();
	if (persist) ntiles = std::min<uint32_t>(ntiles, (uint32_t)sm_count(current_device()) * 4u);		// __launch_bounds__(OS_THREADS, 4)
	if (bits > 8) os_pass_kernel<9><<<ntiles, OS_THREADS, sizeof(OneSweepSharedT<9>), s>>>(in, out, d_n, D, ghist, status, ticket, epoch, g_rank_mode);
	else if (bits == 8 || !narrow) os_pass_kernel<8><<<ntiles, OS_THREADS, sizeof(OneSweepSharedT<8>), s>>>(in, out, d_n, D, ghist, status, ticket, epoch, g_rank_mode);
	else if (bits == 7) os_pass_kernel<7><<<ntiles, OS_THREADS, sizeof(OneSweepSharedT<7>), s>>>(in, out, d_n, D, ghist, status, ticket, epoch, g_rank_mode);
	else os_pass_kernel<6><<<ntiles, OS_THREADS, sizeof(OneSweepSharedT<6>), s>>>(in, out, d_n, D, ghist, status, ticket, epoch, g_rank_mode);
}

// stable LSD radix sort of the *d_n keys in keys_a on their own bits (plain mode, the top-N sorts): n_max >= *d_n sizes the grids
int launch_radix_sort(const SortTemp &tmp, const unsigned long long *d_n, uint64_t n_max, int lo1, int hi1, int lo2, int hi2, int *which, cudaStream_t s)
{
	int launches = 0;
	unsigned long long *bufs[2] = { tmp.keys_a, tmp.keys_b };
	int w = 0;
	DigitSpecs P;

	*which = 0;
	if (!n_max) return 0;
	if (build_digit_specs(lo1, hi1, lo2, hi2, P) || n_max >= (1ull << 30)) return -1;	// status words carry 30-bit counts
	bool any9 = false;
	for (int p = 0; p < P.np; ++p) any9 |= P.d[p].b1 + P.d[p].b2 > 8;
	const int copies = any9 ? 4 : 8, stride = (any9 ? 512 : 256) + 1;
	const int dev = current_device();
	os_set_attrs(dev);
	const uint32_t ntiles = div_up(n_max, SORT_TILE);
	uint32_t *ghist = tmp.os_ghist, *tickets = tmp.os_ghist + OS_MAX_PASSES * RADIX_MAX;

	cudaMemsetAsync(ghist, 0, (OS_MAX_PASSES * RADIX_MAX + OS_MAX_PASSES) * sizeof(uint32_t), s);
	const uint32_t hgrid = std::min<uint32_t>(div_up(n_max, 512 * 2 * 4), (uint32_t)sm_count(dev) * 3);
	if (any9) os_hist_kernel<4, 513><<<hgrid, 512, (size_t)P.np * copies * stride * sizeof(uint32_t), s>>>(bufs[w], d_n, P, ghist);
	else os_hist_kernel<8, 257><<<hgrid, 512, (size_t)P.np * copies * stride * sizeof(uint32_t), s>>>(bufs[w], d_n, P, ghist);
	launches++;
	for (int p = 0; p < P.np; ++p) {
		const bool nine = P.d[p].b1 + P.d[p].b2 > 8;
		const uint32_t epoch = next_epoch(tmp, tmp.max_tiles, s);
		(void)nine;
		launch_os_pass(P.d[p].b1 + P.d[p].b2, ntiles, bufs[w], bufs[w ^ 1], d_n, P.d[p], ghist + p * RADIX_MAX, tmp.tile_status, tickets + p, epoch, s);
		launches++;
		w ^= 1;
	}
	*which = w;
	return launches;
}

// after the ingest kernel of a batch: sort its RESP keys by {slot, bin}, reduce them to one record per run and fold every touched
// service's runs into its window histogram and its digest. Nothing here needs a number from the device on the host: the key
// count lives in st.counters[CTR_NKEYS], the digit histograms in tmp.os_ghist (both written by ingest_kernel); grids are sized by
// n_events, the largest possible key count, and surplus CTAs leave at once.
int launch_batch_merge(const DevState &st, const SortTemp &tmp, uint64_t n_events, uint32_t max_svcs, cudaStream_t s)
{
	if (!n_events) return 0;
	int launches = 0;
	unsigned long long *d_nkeys = st.counters + CTR_NKEYS, *d_ntouched = st.counters + CTR_NTOUCHED;
	const int dev = current_device();
	const int nsm = sm_count(dev);
	os_set_attrs(dev);

	SortPlan plan {};
	if (key_sort_plan(max_svcs, plan) < 0) return -1;
	const uint32_t ntiles = div_up(n_events, SORT_TILE);
	unsigned long long *bufs[2] = { tmp.keys_a, tmp.keys_b };
	uint32_t *ghist = tmp.os_ghist, *tickets = tmp.os_ghist + OS_MAX_PASSES * RADIX_MAX;
	int w = 0;
	for (int p = 0; p < plan.np; ++p) {
		const DigitSpec D { plan.shift[p], plan.bits[p], 0, 0 };
		const uint32_t epoch = next_epoch(tmp, tmp.max_tiles, s);
		launch_os_pass(plan.bits[p], ntiles, bufs[w], bufs[w ^ 1], d_nkeys, D, ghist + p * RADIX_MAX, tmp.tile_status, tickets + p, epoch, s);
		launches++;
		w ^= 1;
	}
	const unsigned long long *src = bufs[w];

	cudaMemsetAsync(d_ntouched, 0, sizeof(unsigned long long), s);
	const uint32_t epoch = next_epoch(tmp, tmp.max_tiles, s);
	// GYSK_RM_THREADS=512: tiles of 4096 keys (half the look-back steps) — A/B runs
	static const int rm_threads = []{ const char *e = getenv("GYSK_RM_THREADS"); return e && atoi(e) == 512 ? 512 : 256; }();
#define RM_ARGS src, d_nkeys, tmp.tile_status, epoch, reinterpret_cast<RunRec *>(tmp.pool), tmp.run_bin, tmp.chunk_run, \
		reinterpret_cast<BatchSeg *>(tmp.segs), tmp.touched, d_ntouched, st.counters + CTR_NRUNS
	if (rm_threads == 256) runs_mark_kernel<256><<<div_up(n_events, 256 * RM_V), 256, 0, s>>>(RM_ARGS);
	else runs_mark_kernel<512><<<div_up(n_events, 512 * RM_V), 512, 0, s>>>(RM_ARGS);
#undef RM_ARGS
	runs_sum_kernel<<<nsm * 8, 256, 0, s>>>(src, d_nkeys, tmp.chunk_run, reinterpret_cast<RunRec *>(tmp.pool));
	// GYSK_MERGE_SMEM_N=512 selects the larger work area (A/B runs)
	static const int merge_smem_n = []{ const char *e = getenv("GYSK_MERGE_SMEM_N"); return e && atoi(e) == 512 ? 512 : 384; }();
	const int merge_ctas = std::min(nsm, TD_MERGE_MAX_SMS) * (merge_smem_n == 512 ? 5 : TD_MERGE_CTAS_PER_SM);
	if (merge_smem_n == 512) bins_merge_kernel<512><<<merge_ctas, TD_WARPS * 32, 0, s>>>(st, tmp.touched, d_ntouched,
			reinterpret_cast<const RunRec *>(tmp.pool), tmp.run_bin, reinterpret_cast<const BatchSeg *>(tmp.segs), tmp.items_scratch, tmp.big_scratch);
	else bins_merge_kernel<384><<<merge_ctas, TD_WARPS * 32, 0, s>>>(st, tmp.touched, d_ntouched,
			reinterpret_cast<const RunRec *>(tmp.pool), tmp.run_bin, reinterpret_cast<const BatchSeg *>(tmp.segs), tmp.items_scratch, tmp.big_scratch);
	return launches + 3;
}

// ---------------------------------------------------------------------------------------------------
// top-N services of the last closed window (BOUNDED_PRIO_QUEUE uses of partha_listener_state, gy_mconnhdlr.cc:11262-11304:
// top listeners by qps / active conns / network): score every service, radix-sort (score, slot) keys, read the tail
// ---------------------------------------------------------------------------------------------------
__global__ void topn_score_kernel(DevState st, uint32_t nslots, int metric, int host_filter, unsigned long long *__restrict__ keys,
		unsigned long long *d_n)
{
	const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot == 0) *d_n = nslots;
	if (slot >= nslots) return;
	unsigned long long score = 0;

	if (host_filter < 0 || st.slot_host[slot] == (uint32_t)host_filter) {
		if (metric == GYSK_TOPN_QPS) {
			for (int b = 0; b < HIST_MAX_CELL; ++b) score += st.hist_last[(size_t)slot * HIST_CELLS + b].count;
		}
		else if (metric == GYSK_TOPN_CONNS) score = (uint32_t)st.conn_last[slot];
		else if (metric == GYSK_TOPN_ISSUE) {
			// ptopissue of partha_listener_state (gy_mconnhdlr.cc:11262-11271): listeners with curr_state_ > STATE_OK, worse state first
			// (LISTEN_TOPN::is_comp_issue, gy_msocket.h:745; its tie-break, tasks_delay_usec_, is not on this path)
			const SlotState ss = st.slot_state[slot];
			score = ss.state > GYSK_STATE_OK && ss.state <= GYSK_STATE_DOWN ? ss.state : 0;
		}
		else score = st.conn_last[slot] >> 32;
	}
	if (score > 0xFFFFFFFFull) score = 0xFFFFFFFFull;
	keys[slot] = (score << 32) | slot;
}

__global__ void topn_pick_kernel(DevState st, const unsigned long long *__restrict__ sorted, uint32_t nslots, uint32_t want, gysk_topn_entry *__restrict__ out)
{
	const uint32_t i = threadIdx.x;
	if (i >= want) return;
	gysk_topn_entry o; o.glob_id = 0; o.score = 0; o.host_idx = 0; o.pad = 0;
	if (i < nslots) {
		const unsigned long long k = sorted[nslots - 1 - i];		// descending
		const uint32_t slot = (uint32_t)k;
		o.glob_id = st.slot_id[slot]; o.score = k >> 32; o.host_idx = st.slot_host[slot];
	}
	out[i] = o;
}

int launch_topn(const DevState &st, const SortTemp &tmp, uint32_t nslots, int metric, int host_filter, uint32_t want, gysk_topn_entry *d_out, cudaStream_t s)
{
	if (!nslots) return 0;
	unsigned long long *d_n = st.counters + CTR_NKEYS;
	int which = 0, launches = 2;
	topn_score_kernel<<<div_up(nslots, 256), 256, 0, s>>>(st, nslots, metric, host_filter, tmp.keys_a, d_n);
	const int sorted = launch_radix_sort(tmp, d_n, nslots, 32, 64, 64, 64, &which, s);
	if (sorted < 0) return sorted;
	launches += sorted;
	topn_pick_kernel<<<1, 64, 0, s>>>(st, which ? tmp.keys_b : tmp.keys_a, nslots, want, d_out);
	return launches;
}

// per-task window of the three MTASK_HIST histograms: totals now minus totals at the previous flush (nothing on the ingest path)
__global__ void task_flush_kernel(DevState st, uint32_t max_tasks)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;		// (task, histogram)
	if (i >= max_tasks * 3u) return;
	const HistCell *h = st.task_hist + (size_t)i * HIST_CELLS;
	HistCell tot {0, 0};
	for (int b = 0; b < HIST_MAX_CELL; ++b) { tot.count += h[b].count; tot.sum += h[b].sum; }
	const HistCell prev = st.task_prev[i];
	st.task_last[i] = HistCell {tot.count - prev.count, tot.sum - prev.sum};
	st.task_prev[i] = tot;
}

int launch_task_flush(const DevState &st, uint32_t max_tasks, cudaStream_t s)
{
	task_flush_kernel<<<div_up((uint64_t)max_tasks * 3, 256), 256, 0, s>>>(st, max_tasks);
	return 1;
}

// top-N aggregated processes of the last closed window by cpu / cpu delay / blkio delay: the atask_top_cpu_ / _cpu_delay_ /
// _io_delay_ queues of partha_aggr_task_state (server/gy_mconnhdlr.cc:10020-10065; entries with a zero metric never enter)
__global__ void topn_task_score_kernel(DevState st, uint32_t ntasks, int metric, unsigned long long *__restrict__ keys, unsigned long long *d_n)
{
	const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot == 0) *d_n = ntasks;
	if (slot >= ntasks) return;
	unsigned long long score = st.task_slot_id[slot] ? (unsigned long long)st.task_last[(size_t)slot * 3 + metric].sum : 0ull;
	if ((long long)score < 0) score = 0;
	if (score > 0xFFFFFFFFull) score = 0xFFFFFFFFull;
	keys[slot] = (score << 32) | slot;
}

__global__ void topn_task_pick_kernel(DevState st, const unsigned long long *__restrict__ sorted, uint32_t ntasks, uint32_t want, gysk_topn_entry *__restrict__ out)
{
	const uint32_t i = threadIdx.x;
	if (i >= want) return;
	gysk_topn_entry o; o.glob_id = 0; o.score = 0; o.host_idx = 0; o.pad = 0;
	if (i < ntasks) {
		const unsigned long long k = sorted[ntasks - 1 - i];		// descending
		const uint32_t slot = (uint32_t)k;
		o.glob_id = st.task_slot_id[slot]; o.score = k >> 32; o.host_idx = st.task_slot_host[slot];
	}
	out[i] = o;
}

int launch_topn_tasks(const DevState &st, const SortTemp &tmp, uint32_t ntasks, int metric, uint32_t want, gysk_topn_entry *d_out, cudaStream_t s)
{
	if (!ntasks) return 0;
	int which = 0, launches = 2;
	unsigned long long *d_n = st.counters + CTR_NKEYS;
	topn_task_score_kernel<<<div_up(ntasks, 256), 256, 0, s>>>(st, ntasks, metric, tmp.keys_a, d_n);
	const int sorted = launch_radix_sort(tmp, d_n, ntasks, 32, 64, 64, 64, &which, s);
	if (sorted < 0) return sorted;
	launches += sorted;
	topn_task_pick_kernel<<<1, 64, 0, s>>>(st, which ? tmp.keys_b : tmp.keys_a, ntasks, want, d_out);
	return launches;
}

int launch_flush(const DevState &st, uint32_t nslots, HistCell *ring_plane0, HistCell *ring_plane1, uint32_t tsec, uint32_t idle_secs,
		uint32_t live_mask0, uint32_t live_mask1, cudaStream_t s)
{
	if (!nslots) return 0;
	cudaMemsetAsync(st.counters + CTR_NEVICT, 0, sizeof(unsigned long long), s);
	flush_kernel<<<div_up((uint64_t)nslots * HIST_CELLS, 256), 256, 0, s>>>(st, nslots, ring_plane0, ring_plane1, tsec, idle_secs);
	state_kernel<<<div_up(nslots, 128), 128, 0, s>>>(st, nslots, tsec, live_mask0, live_mask1);
	if (!idle_secs) return 2;
	evict_kernel<<<296, 256, 0, s>>>(st, nslots);		// grid-stride over the (device-side) eviction list
	return 3;
}

int launch_rebuild_table(const DevState &st, uint32_t max_svcs, cudaStream_t s)
{
	cudaMemsetAsync(st.svc_tbl.ent, 0, ((size_t)st.svc_tbl.mask + 1) * sizeof(TblEntry), s);
	rebuild_table_kernel<<<div_up(max_svcs, 256), 256, 0, s>>>(st, max_svcs);
	return 1;
}

int launch_gather_svcs(const DevState &st, const unsigned long long *d_ids, uint32_t n, uint32_t max_svcs, uint32_t live_mask0, uint32_t live_mask1,
		SvcRaw *d_out, cudaStream_t s)
{
	if (!n) return 0;
	gather_svcs_kernel<<<div_up(n, 4), 128, 0, s>>>(st, d_ids, n, max_svcs, live_mask0, live_mask1, d_out);
	return 1;
}

int launch_gather_tasks(const DevState &st, const unsigned long long *d_ids, uint32_t n, TaskRaw *d_out, cudaStream_t s)
{
	if (!n) return 0;
	gather_tasks_kernel<<<n, 64, 0, s>>>(st, d_ids, n, d_out);
	return 1;
}

int launch_gather_hll(const DevState &st, unsigned long long id, uint8_t *d_out, int32_t *d_found, cudaStream_t s)
{
	gather_hll_kernel<<<1, 256, 0, s>>>(st, id, d_out, d_found);
	return 1;
}

int launch_query_flows(const DevState &st, const unsigned long long *d_keys, uint32_t n, int last_window, gysk_flow_est *d_out, cudaStream_t s)
{
	if (!n) return 0;
	query_flows_kernel<<<div_up(n, 256), 256, 0, s>>>(st, d_keys, n, last_window, d_out);
	return 1;
}

} // namespace gysk
