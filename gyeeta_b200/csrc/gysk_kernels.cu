// gysk_kernels.cu — hand-written sm_100a kernels of the streaming-sketch engine.
//
//   ingest_kernel        one pass over a batch of 32-byte events: id -> slot, then per event type
//                          RESP : GY_HISTOGRAM::add_data (RESP_TIME_HASH)            common/gy_statistics.h:596-623, :1698
//                                 + emit (slot, usec) sort key for the t-digest chain
//                          TCP  : count-min cell adds, HLL register max, per-service exact cell
//                          TASK : MAGGR_TASK::set_local_task_state (3 histograms)     server/gy_msocket.h:1009-1018
//   rs_* / scan_*        stable LSD radix sort of the (slot, usec) keys (8-bit digits)
//   td_segments/update   batched merging t-digest (K_1 scale, delta = 100)           DESIGN.md §t-digest
//   flush_kernel         5-s window roll                                               common/gy_socket_stat.cc:3898
//   gather_* / query_*   read side
#include "gysk_kernels.cuh"

#include <cfloat>
#include <climits>
#include <cmath>
#include <algorithm>
#include <type_traits>
#include <cstdlib>
#include <cstring>

namespace gysk {

static constexpr unsigned long long KEY_SENTINEL = ~0ull;
static constexpr unsigned long long VALUE_MASK = (1ull << VALUE_BITS) - 1;
__device__ __forceinline__ uint32_t key_slot(unsigned long long k) { return (uint32_t)(k >> KEY_SLOT_SHIFT); }
__device__ __forceinline__ uint32_t key_usec(unsigned long long k) { return (uint32_t)(k >> KEY_VALUE_SHIFT) & (uint32_t)VALUE_MASK; }

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
//   warp : lanes updating the same cell are grouped with match.any, reduced with redux and represented by one leader;
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
		atomicMax(&st.task_hist[(cell & ~CELL_TASK) | 15u].sum, (long long)vmax);
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

__device__ __forceinline__ void hll_update(uint8_t *regs, uint32_t idx, uint32_t rank)
{
	uint32_t *wp = reinterpret_cast<uint32_t *>(regs) + (idx >> 2);
	const uint32_t sh = (idx & 3u) * 8u;
	uint32_t w = __ldca(wp);				// stale is harmless: the CAS re-validates

	while (((w >> sh) & 0xFFu) < rank) {
		const uint32_t nw = (w & ~(0xFFu << sh)) | (rank << sh);
		const uint32_t old = atomicCAS(wp, w, nw);
		if (old == w) break;
		w = old;
	}
}

// The kernel runs as a persistent tile pipeline. A CTA takes a tile of INGEST_TILE events and
//   phase 1 (one event per thread and round, fully converged): 2 x 128-bit load, shard filter, id -> slot lookup. A RESP
//           sample becomes its sort key {slot, usec, client port & 31} in the tile's key queue — its histogram cell, CONN_BITMAP
//           bit and t-digest share are all produced later from the SORTED keys, where equal cells are contiguous runs
//           (td_sums_kernel). TCP / TASK events leave their decoded record in shared memory and their index in the queue of
//           their kind (ballot + one shared-memory atomic per warp and kind);
//   phase 2 (converged per kind): every warp strides over one queue at a time, so the count-min rows (one (event, row) pair
//           per lane), the HLL updates and the three task histograms (one (event, histogram) pair per lane) each run with all
//           lanes doing the same thing instead of serialising 70/20/10-divergent branches;
//   phase 3: the tile's RESP keys go out as one contiguous, coalesced run (one global cursor bump per tile).

struct IngestRec { uint32_t slot; uint32_t value; unsigned long long flow_key; };

// ---- TMA (bulk async copy) staging of the next event tile into shared memory, completion on an mbarrier ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");		// make the init visible to the async proxy
}

__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, unsigned long long *bar)
{
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");		// earlier generic-proxy reads of dst are done (after bar.sync)
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			:: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
	asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}"
			:: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

template <int INGEST_THREADS, bool STAGE, int INGEST_EPT = 4>
struct IngestSharedT
{
	static constexpr int INGEST_TILE = INGEST_THREADS * INGEST_EPT;
	using HotTable = HotTableT<(INGEST_TILE >= 1024 ? 10 : 9)>;
	alignas(128) uint4	evbuf[STAGE ? INGEST_TILE * 2 : 1];	// next tile of 32-byte events, filled by cp.async.bulk
	unsigned long long	mbar;
	HotTable	hot;
	// per-tile state, double-buffered by tile parity: phase 1 of tile t+1 fills one set while stragglers still drain the other
	unsigned long long kq[2][INGEST_TILE];		// RESP sort keys of the tile, compacted
	IngestRec	rec[2][INGEST_TILE];		// decoded TCP events from the front, TASK events from the back (together <= tile)
	// packed {n_resp | n_tcp | n_task}, QBITS bits each; three sets in rotation so that a set is cleared a full tile before its
	// next use (after the barrier of tile t: the set of tile t+2). Tiles of up to 512 events fit 3 x 10 bits in ONE 32-bit word
	// (native shared-memory add; the 64-bit add is a compare-and-swap loop)
	static constexpr int QBITS = INGEST_TILE <= 512 ? 10 : 21;
	using QWord = typename std::conditional<INGEST_TILE <= 512, uint32_t, unsigned long long>::type;
	QWord		qn[3];
	unsigned long long key_base[2];
	uint32_t	key_seq[2];			// tile sequence number + 1 once key_base of that parity is valid
	uint32_t	max_value;			// largest RESP msec seen by this CTA (sizes the radix sort)
};

// warp-aggregated append: returns this lane's position in the queue (valid where pred)
__device__ __forceinline__ uint32_t queue_reserve(bool pred, uint32_t *qn)
{
	const uint32_t m = __ballot_sync(0xffffffffu, pred);
	if (!m) return 0;
	const int lane = threadIdx.x & 31;
	uint32_t base = 0;
	if (lane == __ffs(m) - 1) base = atomicAdd(qn, (uint32_t)__popc(m));
	base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
	return base + __popc(m & ((1u << lane) - 1u));
}

template <int INGEST_THREADS, int MIN_CTAS, bool STAGE, int INGEST_EPT, bool PIPE>
__global__ void __launch_bounds__(INGEST_THREADS, MIN_CTAS) ingest_kernel(DevState st, const gysk_event *__restrict__ ev, uint64_t n,
		unsigned long long *__restrict__ keys, int prefetch_next)
{
	using IngestShared = IngestSharedT<INGEST_THREADS, STAGE, INGEST_EPT>;
	using HotTable = typename IngestShared::HotTable;
	constexpr int INGEST_TILE = IngestShared::INGEST_TILE;
	constexpr int QBITS = IngestShared::QBITS;
	constexpr uint32_t QMASK = (1u << QBITS) - 1u;
	using QWord = typename IngestShared::QWord;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	IngestShared &S = *reinterpret_cast<IngestShared *>(smem_raw);
	uint32_t c_in = 0, c_foreign = 0;		// per thread: < 2^32 events per launch
	unsigned long long t_resp = 0, t_tcp = 0, t_task = 0;	// thread 0: per-tile queue lengths summed
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint64_t ntiles = (n + INGEST_TILE - 1) / INGEST_TILE;
	uint32_t max_us = 0;

	for (int i = threadIdx.x; i < HotTable::N; i += INGEST_THREADS) { S.hot.tag[i] = 0; S.hot.count[i] = 0; S.hot.sum[i] = 0; S.hot.vmax[i] = INT_MIN; }
	if (threadIdx.x < 3) S.qn[threadIdx.x] = 0;
	if (threadIdx.x == 0) { S.max_value = 0; S.key_seq[0] = 0; S.key_seq[1] = 0; }
	if (STAGE && threadIdx.x == 0) mbar_init(&S.mbar, 1);
	__syncthreads();
	// prologue: the first tile of this CTA starts flying into shared memory
	if (STAGE && threadIdx.x == 0 && blockIdx.x < ntiles) {
		const uint64_t b0 = (uint64_t)blockIdx.x * INGEST_TILE;
		const uint64_t cnt = n - b0 < (uint64_t)INGEST_TILE ? n - b0 : (uint64_t)INGEST_TILE;
		tma_load_1d(S.evbuf, ev + b0, (uint32_t)cnt * 32u, &S.mbar);
	}

	uint4 ra[INGEST_EPT], rb[INGEST_EPT];
	auto load_tile = [&](uint64_t tb) {
#pragma unroll
		for (int k = 0; k < INGEST_EPT; ++k) {
			const uint64_t i = tb + (uint64_t)k * INGEST_THREADS + threadIdx.x;
			if (i < n) {
				if (STAGE) {
					ra[k] = S.evbuf[2 * (k * INGEST_THREADS + threadIdx.x)];
					rb[k] = S.evbuf[2 * (k * INGEST_THREADS + threadIdx.x) + 1];
				}
				else {
					ra[k] = __ldcs(reinterpret_cast<const uint4 *>(ev + i));		// streamed once: evict-first, keep L2 for
					rb[k] = __ldcs(reinterpret_cast<const uint4 *>(ev + i) + 1);	// the id table / histogram / count-min lines
				}
			}
			else { ra[k] = make_uint4(0, 0, 0, 0); rb[k] = make_uint4(0, 0, 0, 0xFFFFu); }	// type 0xFFFF: padding, not counted
		}
	};
	if (PIPE && !STAGE) load_tile((uint64_t)blockIdx.x * INGEST_TILE);

	int par = 0, qi = 0;
	uint32_t seq = 1;
	for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, par ^= 1, ++seq, qi = qi == 2 ? 0 : qi + 1) {
		const uint64_t tbase = tile * INGEST_TILE;
		QWord *qn64 = &S.qn[qi];
		unsigned long long *kq = S.kq[par];
		IngestRec *rec = S.rec[par];
		if (STAGE) mbar_wait(&S.mbar, (uint32_t)par);		// this tile's events have landed (one completion per tile)

		// ---------------- phase 1: decode + lookup + enqueue ----------------
		// The tile's events are already on their way (or here): their loads were issued before the previous tile's barrier
		// (PIPE), so the only memory latency left on this phase's critical path is the id-table probe.
		if (!PIPE || STAGE) load_tile(tbase);
		// this CTA's next tile starts moving from HBM to L2 now (no registers held): its loads one tile later find L2
		if (!STAGE && prefetch_next) {
#pragma unroll
			for (int k = 0; k < INGEST_EPT; ++k) {
				const uint64_t i = tbase + (uint64_t)gridDim.x * INGEST_TILE + (uint64_t)k * INGEST_THREADS + threadIdx.x;
				if (i < n) {
					if (prefetch_next == 2) asm volatile("prefetch.global.L2::evict_last [%0];" :: "l"(ev + i));
					else asm volatile("prefetch.global.L2 [%0];" :: "l"(ev + i));
				}
			}
		}
		// decode; put the first id-table probe of all EPT events in flight before any of them is resolved
		uint4 praw[INGEST_EPT];
		uint32_t ppos[INGEST_EPT];
		uint32_t kind[INGEST_EPT];		// 0 none, GYSK_EV_RESP, 1 tcp (stored as GYSK_EV_ACCEPT), GYSK_EV_TASK
#pragma unroll
		for (int k = 0; k < INGEST_EPT; ++k) {
			const unsigned long long svc = ((unsigned long long)ra[k].y << 32) | ra[k].x;
			const uint32_t value = rb[k].x, host_idx = rb[k].y;
			const uint32_t type = rb[k].w & 0xFFFFu;
			const bool pad = tbase + (uint64_t)k * INGEST_THREADS + threadIdx.x >= n;
			const bool is_resp = type == GYSK_EV_RESP, is_task = type == GYSK_EV_TASK;
			const bool is_tcp = type >= GYSK_EV_CONNECT && type <= GYSK_EV_CLOSE_SER;
			bool mine = !pad;

			kind[k] = 0; ppos[k] = 0; praw[k] = make_uint4(0, 0, 0, 0);
			if (mine && st.world > 1 && (host_idx % st.world) != st.rank) { c_foreign++; mine = false; }
			if (mine) {
				c_in++;
				// usec -> msec as SVC_INFO_CAP::upd_stats_on_req (gy_proto_parser.cc:2678); validity rule of
				// handle_ipv4_resp_event (gy_socket_stat.cc:1519-1524): drop beyond 1 000 000 msec
				if (svc + 1ull > 1ull && (is_tcp || is_task || (is_resp && value < 1000001000u))) {	// id not 0 / ~0 (tombstone); msec <= 1 000 000
					kind[k] = is_resp ? (uint32_t)GYSK_EV_RESP : (is_task ? (uint32_t)GYSK_EV_TASK : (uint32_t)GYSK_EV_ACCEPT);
					praw[k] = table_probe_first(is_task ? st.task_tbl : st.svc_tbl, svc, ppos[k]);
				}
			}
		}
#pragma unroll
		for (int k = 0; k < INGEST_EPT; ++k) {
			const bool is_resp = kind[k] == GYSK_EV_RESP, is_task = kind[k] == GYSK_EV_TASK, is_tcp = kind[k] == GYSK_EV_ACCEPT;
			int slot = -1;
			if (kind[k]) {
				// one id lookup for all three event kinds (services and tasks live in separate tables)
				slot = table_resolve(is_task ? st.task_tbl : st.svc_tbl, ((unsigned long long)ra[k].y << 32) | ra[k].x, st.auto_register, rb[k].y, ppos[k], praw[k]);
			}
			const bool ok = slot >= 0;
			if (ok && is_resp) max_us = max(max_us, rb[k].x);
			// one shared-memory atomic per warp reserves queue space for all three kinds: {resp : 21 | tcp : 21 | task : 21}
			const uint32_t m_resp = __ballot_sync(0xffffffffu, ok && is_resp), m_tcp = __ballot_sync(0xffffffffu, ok && is_tcp),
					m_task = __ballot_sync(0xffffffffu, ok && is_task);
			QWord qbase = 0;
			if (lane == 0 && (m_resp | m_tcp | m_task))
				qbase = atomicAdd(qn64, (QWord)__popc(m_resp) | ((QWord)__popc(m_tcp) << QBITS) | ((QWord)__popc(m_task) << (2 * QBITS)));
			qbase = __shfl_sync(0xffffffffu, qbase, 0);
			const uint32_t lt = (1u << lane) - 1u;
			const uint32_t q_resp = ((uint32_t)qbase & QMASK) + __popc(m_resp & lt), q_tcp = ((uint32_t)(qbase >> QBITS) & QMASK) + __popc(m_tcp & lt),
					q_task = (uint32_t)(qbase >> (2 * QBITS)) + __popc(m_task & lt);
			if (ok) {
				if (is_resp) {
					// {slot, usec, client port & 31 (CONN_BITMAP index, common/gy_socket_stat.h:403-410)}
					kq[q_resp] = ((unsigned long long)(uint32_t)slot << KEY_SLOT_SHIFT) | ((unsigned long long)rb[k].x << KEY_VALUE_SHIFT) | (ra[k].z & 0x1Fu);
				}
				else {
					IngestRec r; r.slot = (uint32_t)slot; r.value = rb[k].x; r.flow_key = ((unsigned long long)ra[k].w << 32) | ra[k].z;
					rec[is_tcp ? q_tcp : (uint32_t)INGEST_TILE - 1u - q_task] = r;
				}
			}
		}
		if (PIPE && !STAGE) load_tile(tbase + (uint64_t)gridDim.x * INGEST_TILE);	// next tile of this CTA (all padding past the end)
		__syncthreads();			// the only block barrier of the tile: queues of this parity are complete
		const QWord qv = *qn64;
		const uint32_t n_resp = (uint32_t)qv & QMASK, n_tcp = (uint32_t)(qv >> QBITS) & QMASK, n_task = (uint32_t)(qv >> (2 * QBITS));
		// thread 0 bumps the global key cursor now; its round trip to L2 hides behind the TCP and TASK phases
		if (threadIdx.x == 0) {
			// the set of tile t+2 (== tile t-1): every warp has read its counts (it passed this barrier), and nobody appends
			// to it before the next barrier
			S.qn[qi == 0 ? 2 : qi - 1] = 0;
			t_resp += n_resp; t_tcp += n_tcp; t_task += n_task;
			S.key_base[par] = n_resp ? atomicAdd(st.counters + CTR_NKEYS, (unsigned long long)n_resp) : 0ull;
			__threadfence_block();
			*((volatile uint32_t *)&S.key_seq[par]) = seq;
			// every thread has consumed evbuf (barrier above): stage the next tile of this CTA while phase 2 runs
			if (STAGE && tile + gridDim.x < ntiles) {
				const uint64_t b1 = (tile + gridDim.x) * INGEST_TILE;
				const uint64_t cnt = n - b1 < (uint64_t)INGEST_TILE ? n - b1 : (uint64_t)INGEST_TILE;
				tma_load_1d(S.evbuf, ev + b1, (uint32_t)cnt * 32u, &S.mbar);
			}
		}

		// ---------------- phase 2a: TCP — count-min rows, one (event, row) pair per lane ----------------
		{
			const uint32_t npairs = n_tcp * st.cms_depth;
			const bool d4 = st.cms_depth == 4;		// the default depth: no integer division on the pair index
			for (uint32_t p = threadIdx.x; p < npairs; p += INGEST_THREADS) {
				const uint32_t e = d4 ? p >> 2 : p / st.cms_depth, row = d4 ? p & 3u : p - e * st.cms_depth;
				const IngestRec r = rec[e];
				red_add_u64(st.cms_cur + ((size_t)row << st.cms_log2w) + cms_index(r.flow_key, row, st.cms_wmask), cms_increment(r.value));
			}
			// HLL register + the service's exact {count, kbytes} cell
			for (uint32_t base = wid * 32; base < n_tcp; base += INGEST_THREADS) {
				const uint32_t q = base + lane;
				const bool act = q < n_tcp;
				uint32_t cell = 0; int kb = 0;
				if (act) {
					const IngestRec r = rec[q];
					uint32_t idx, rank;
					hll_idx_rank(r.flow_key, st.hll_p, idx, rank);
					hll_update(st.hll + ((size_t)r.slot << st.hll_p), idx, rank);
					cell = r.slot;
					kb = (int)(r.value >> 10);
				}
				cell_add(st, S.hot, act, cell, kb);
			}
		}
		// ---------------- phase 2b: TASK — MAGGR_TASK::set_local_task_state, one (event, histogram) pair per lane ----------------
		{
			const uint32_t ntrip = n_task * 3u;
			for (uint32_t base = wid * 32; base < ntrip; base += INGEST_THREADS) {
				const uint32_t p = base + lane;
				const bool act = p < ntrip;
				uint32_t cell = 0; int d = 0;
				if (act) {
					const uint32_t e = p / 3u, h = p - e * 3u;
					const IngestRec r = rec[(uint32_t)INGEST_TILE - 1u - e];
					// GY_HISTOGRAM<int, ...>::add_data(int): the three values narrow to int (server/gy_msocket.h:1014-1016)
					d = h == 0 ? (int)r.value : (h == 1 ? (int)(uint32_t)r.flow_key : (int)(uint32_t)(r.flow_key >> 32));
					const uint32_t b = h == 0 ? (uint32_t)bucket_hash_1_3000(d) : (uint32_t)bucket_duration(d);
					cell = CELL_TASK | (r.slot * 3u * HIST_CELLS + h * HIST_CELLS + b);
				}
				cell_add(st, S.hot, act, cell, d);
			}
		}
		// ---------------- phase 3: the tile's RESP keys leave as one coalesced run ----------------
		if (n_resp) {
			while (*((volatile uint32_t *)&S.key_seq[par]) != seq) { }		// thread 0's cursor bump has landed (normally long ago)
			unsigned long long *dst = keys + *((volatile unsigned long long *)&S.key_base[par]);
			for (uint32_t q = threadIdx.x; q < n_resp; q += INGEST_THREADS) __stcs(dst + q, kq[q]);
		}
		// no barrier here: the next tile fills the other parity; this parity is reused only after the next tile's barrier
	}

	max_us = __reduce_max_sync(0xffffffffu, max_us);
	if (lane == 0 && max_us) atomicMax(&S.max_value, max_us / 1000u);	// msec is enough: bits(usec) <= bits(msec) + 10
	__syncthreads();
	if (threadIdx.x == 0 && S.max_value) atomicMax(st.counters + CTR_MAXVAL, (unsigned long long)S.max_value);
	// retire: one RED group per privatised cell
	for (int i = threadIdx.x; i < HotTable::N; i += INGEST_THREADS) {
		if (S.hot.tag[i] && S.hot.count[i]) cell_add_global(st, S.hot.tag[i] - 1, S.hot.count[i], S.hot.sum[i], S.hot.vmax[i]);
	}

	// statsmap-style counters (gy_mconnhdlr.cc:4708-4715): warp-reduce, one atomic per warp and counter
	c_in = __reduce_add_sync(0xffffffffu, c_in);
	c_foreign = __reduce_add_sync(0xffffffffu, c_foreign);
	if (lane == 0) {
		if (c_in) atomicAdd(st.counters + CTR_IN, (unsigned long long)c_in);
		if (c_foreign) atomicAdd(st.counters + CTR_FOREIGN, (unsigned long long)c_foreign);
		// dropped = taken in but not queued (svc_id 0, bad type or value, table full, unknown id): derived after all adds landed
		if (c_in) atomicAdd(st.counters + CTR_DROPPED, (unsigned long long)c_in);
	}
	if (threadIdx.x == 0) {
		if (t_resp) atomicAdd(st.counters + CTR_RESP, t_resp);
		if (t_tcp) atomicAdd(st.counters + CTR_TCP, t_tcp);
		if (t_task) atomicAdd(st.counters + CTR_TASK, t_task);
		const unsigned long long q = t_resp + t_tcp + t_task;
		if (q) atomicAdd(st.counters + CTR_DROPPED, 0ull - q);	// two's complement: dropped += in - queued
	}
}

// ---------------------------------------------------------------------------------------------------
// stable LSD radix sort, 8- or 9-bit digits, tile = SORT_TILE keys per CTA of 256 threads
// ---------------------------------------------------------------------------------------------------
static constexpr int RADIX_MAX_BITS = 9;
static constexpr int RADIX_MAX = 1 << RADIX_MAX_BITS;

// one radix pass sorts on a digit made of up to two bit fields of the key, so that the unused bits between the usec field
// and the slot field of a key never cost a pass: digit = ((k >> s1) & m1) | (((k >> s2) & m2) << b1)
struct DigitSpec { int s1, b1, s2, b2; };
__device__ __forceinline__ uint32_t key_digit(unsigned long long k, const DigitSpec &D)
{
	return ((uint32_t)(k >> D.s1) & ((1u << D.b1) - 1u)) | (((uint32_t)(k >> D.s2) & ((1u << D.b2) - 1u)) << D.b1);
}

// ---------------------------------------------------------------------------------------------------
// one-sweep radix pass: 16 B of HBM traffic per key and pass (read once, write once)
//
//   os_hist_kernel   one read of the keys fills the GLOBAL digit histograms of every pass (digits are fixed before the first
//                    pass, and a stable pass does not change how many keys carry a digit value);
//   os_pass_kernel   a CTA takes the next tile (ticket from an atomic counter, so every predecessor tile is already running),
//                    ranks its keys per digit, publishes the tile's digit counts and obtains the number of keys with the same
//                    digit in all earlier tiles by decoupled look-back over the status words of its predecessors
//                    (status word = 2-bit state | 30-bit count: 1 = this tile's count, 2 = inclusive prefix up to this tile),
//                    reorders the tile by digit in shared memory and writes every digit's run to its final place.
// Stability: tiles are ordered by ticket = tile index, ranks inside a tile follow the input order (warp, round, lane).
// Digits are 8 bits wide; when the significant bits do not fit ceil(bits / 9) + ... passes of 8 (e.g. 41 bits), some passes take
// 9 bits (512 digits, two per thread in the per-digit steps) instead of the sort paying a whole extra pass.
// ---------------------------------------------------------------------------------------------------
static constexpr int OS_THREADS = 256;			// thread t owns digits t, t + 256 in the per-digit steps
static constexpr int OS_WARPS = OS_THREADS / 32;
static constexpr int OS_KPT = SORT_TILE / OS_THREADS;	// 16 keys per thread
static constexpr int OS_MAX_PASSES = 8;
static constexpr uint32_t OS_FLAG_AGG = 1u << 30, OS_FLAG_PREFIX = 2u << 30, OS_COUNT_MASK = (1u << 30) - 1u;

struct DigitSpecs { DigitSpec d[OS_MAX_PASSES]; int np; };

// lane-privatised histogram copies (lane & (copies - 1)), skewed by one bank each: 8 copies of 257 words per pass for 8-bit
// digits, 4 copies of 513 words when a pass has 9 bits
template <int copies, int stride>
__global__ void __launch_bounds__(512) os_hist_kernel(const unsigned long long *__restrict__ keys, uint64_t n, DigitSpecs P,
		uint32_t *__restrict__ ghist /* [np][RADIX_MAX] */)
{
	extern __shared__ __align__(16) unsigned char osh_smem[];
	uint32_t *h = reinterpret_cast<uint32_t *>(osh_smem);		// [np][copies][stride]
	const int copy = threadIdx.x & (copies - 1);
	constexpr int pstride = copies * stride;

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
		uint32_t n, DigitSpec D, const uint32_t *__restrict__ ghist /* [RADIX] of this pass */, uint32_t *__restrict__ status /* [ntiles][RADIX] */,
		uint32_t *__restrict__ ticket, int rank_mode /* 0 auto, 1 match.any, 2 ballots */)
{
	constexpr int RADIX = 1 << RBITS;
	constexpr int DPT = RADIX / OS_THREADS;		// digits per thread in the per-digit steps
	extern __shared__ __align__(16) unsigned char os_smem[];
	OneSweepSharedT<RBITS> &S = *reinterpret_cast<OneSweepSharedT<RBITS> *>(os_smem);
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t lt_mask = (1u << lane) - 1u;

	if (threadIdx.x == 0) S.tile = atomicAdd(ticket, 1u);
	for (int i = threadIdx.x; i < OS_WARPS * RADIX; i += OS_THREADS) (&S.whist[0][0])[i] = 0;
	__syncthreads();
	const uint32_t tile = S.tile;
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
#pragma unroll
	for (int r = 0; r < OS_KPT; ++r) {
		const bool valid = wbase + (uint32_t)r * 32 + lane < n;
		const uint32_t d = valid ? key_digit(k[r], D) : ((uint32_t)RADIX + lane);
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

	if constexpr (DPT == 1) {
		// thread d: prefix over the warps, the tile's count of digit d -> published at once, so successors can look back through it
		const uint32_t d = threadIdx.x;
		uint32_t dtotal = 0;
#pragma unroll
		for (int w = 0; w < OS_WARPS; ++w) { const uint32_t t = S.whist[w][d]; S.whist[w][d] = dtotal; dtotal += t; }
		st_volatile_u32(status + (size_t)tile * RADIX + d, (tile == 0 ? OS_FLAG_PREFIX : OS_FLAG_AGG) | dtotal);

		// {global count of digit d, tile count of digit d} -> exclusive scans over the digits in one go
		const unsigned long long sc = os_block_exclusive_scan(((unsigned long long)ghist[d] << 16) | dtotal, S.scan[0], nullptr);
		const uint32_t gexcl = (uint32_t)(sc >> 16), dstart = (uint32_t)(sc & 0xFFFFu);

		// decoupled look-back: keys with digit d in the tiles before this one
		uint32_t excl = 0;
		if (tile > 0) {
			uint32_t p = tile - 1;
			for (;;) {
				const uint32_t v = ld_volatile_u32(status + (size_t)p * RADIX + d);
				if (!(v >> 30)) continue;				// predecessor has its ticket, so it is running: its count will come
				excl += v & OS_COUNT_MASK;
				if (v & OS_FLAG_PREFIX) break;
				--p;
			}
			st_volatile_u32(status + (size_t)tile * RADIX + d, OS_FLAG_PREFIX | (excl + dtotal));
		}
		S.dstart[d] = dstart;
		S.goff[d] = gexcl + excl - dstart;
	}
	else {
		// thread t owns digits t and t + 256: same steps, the scan runs in digit order with a carry between the two halves
		uint32_t dtotal[DPT];
#pragma unroll
		for (int j = 0; j < DPT; ++j) {
			const uint32_t d = threadIdx.x + j * OS_THREADS;
			uint32_t run = 0;
#pragma unroll
			for (int w = 0; w < OS_WARPS; ++w) { const uint32_t t = S.whist[w][d]; S.whist[w][d] = run; run += t; }
			dtotal[j] = run;
			st_volatile_u32(status + (size_t)tile * RADIX + d, (tile == 0 ? OS_FLAG_PREFIX : OS_FLAG_AGG) | run);
		}
		unsigned long long carry = 0;
#pragma unroll
		for (int j = 0; j < DPT; ++j) {
			const uint32_t d = threadIdx.x + j * OS_THREADS;
			unsigned long long tot = 0;
			const unsigned long long sc = carry + os_block_exclusive_scan(((unsigned long long)ghist[d] << 16) | dtotal[j], S.scan[j], &tot);
			carry += tot;
			const uint32_t gexcl = (uint32_t)(sc >> 16), dstart = (uint32_t)(sc & 0xFFFFu);
			uint32_t excl = 0;
			if (tile > 0) {
				uint32_t p = tile - 1;
				for (;;) {
					const uint32_t v = ld_volatile_u32(status + (size_t)p * RADIX + d);
					if (!(v >> 30)) continue;
					excl += v & OS_COUNT_MASK;
					if (v & OS_FLAG_PREFIX) break;
					--p;
				}
				st_volatile_u32(status + (size_t)tile * RADIX + d, OS_FLAG_PREFIX | (excl + dtotal[j]));
			}
			S.dstart[d] = dstart;
			S.goff[d] = gexcl + excl - dstart;
		}
	}
	__syncthreads();

	// reorder the tile by digit in shared memory
#pragma unroll
	for (int r = 0; r < OS_KPT; ++r) {
		if (wbase + (uint32_t)r * 32 + lane < n) {
			const uint32_t dd = key_digit(k[r], D);
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
}

// ---------------------------------------------------------------------------------------------------
// batched merging t-digest
// ---------------------------------------------------------------------------------------------------
static constexpr int TSEG_V = 4;			// consecutive keys per thread

__global__ void __launch_bounds__(256) td_segments_kernel(const unsigned long long *__restrict__ keys, const unsigned long long *__restrict__ d_n,
		uint32_t *__restrict__ seg_start, uint32_t *__restrict__ seg_end, uint32_t *__restrict__ touched, unsigned long long *ntouched)
{
	const uint64_t n = *d_n;
	const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * TSEG_V;
	const int lane = threadIdx.x & 31;
	uint32_t sl[TSEG_V + 2];			// slots of key i0-1, i0 .. i0+3, i0+4 (0xFFFFFFFF outside the array)

#pragma unroll
	for (int t = 0; t < TSEG_V + 2; ++t) sl[t] = 0xFFFFFFFFu;
	if (i0 + TSEG_V <= n) {
		const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(keys + i0), c = *reinterpret_cast<const ulonglong2 *>(keys + i0 + 2);
		sl[1] = key_slot(a.x); sl[2] = key_slot(a.y);
		sl[3] = key_slot(c.x); sl[4] = key_slot(c.y);
	}
	else {
#pragma unroll
		for (int t = 0; t < TSEG_V; ++t) if (i0 + t < n) sl[1 + t] = key_slot(keys[i0 + t]);
	}
	// neighbours: from the adjacent lanes, the warp's edge lanes load them
	const uint32_t up = __shfl_up_sync(0xffffffffu, sl[TSEG_V], 1), down = __shfl_down_sync(0xffffffffu, sl[1], 1);
	sl[0] = lane ? up : ((i0 && i0 - 1 < n) ? key_slot(keys[i0 - 1]) : 0xFFFFFFFFu);
	sl[TSEG_V + 1] = lane < 31 ? down : (i0 + TSEG_V < n ? key_slot(keys[i0 + TSEG_V]) : 0xFFFFFFFFu);

	uint32_t nstart = 0;
#pragma unroll
	for (int t = 1; t <= TSEG_V; ++t) nstart += (i0 + t - 1 < n && sl[t] != sl[t - 1]) ? 1u : 0u;

	// one cursor bump per warp for all the runs that start in it (the cold tail has a new service almost every sample)
	uint32_t incl = nstart;
#pragma unroll
	for (int off = 1; off < 32; off <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += v; }
	const uint32_t wtotal = __shfl_sync(0xffffffffu, incl, 31);
	unsigned long long base = 0;
	if (wtotal && lane == 0) base = atomicAdd(ntouched, (unsigned long long)wtotal);
	base = __shfl_sync(0xffffffffu, base, 0) + (incl - nstart);

#pragma unroll
	for (int t = 1; t <= TSEG_V; ++t) {
		const uint64_t i = i0 + t - 1;
		if (i >= n) break;
		if (sl[t] != sl[t - 1]) { seg_start[sl[t]] = (uint32_t)i; touched[base++] = sl[t]; }
		if (sl[t] != sl[t + 1]) seg_end[sl[t]] = (uint32_t)(i + 1);
	}
}

static constexpr int TD_WARPS = 4;
static constexpr int PLAN_STRIDE = TD_CAP + 1;

// (1) plan: one thread per touched service runs the greedy chain over n unit-weight samples. Cluster j of the run is
// [bounds[j], bounds[j+1]) with bounds[j+1] = max(bounds[j] + 1, floor(n q(k(bounds[j]/n) + 1))): the boundaries depend on n
// only, so they can be fixed before any sample is summed. Also clears the cluster-sum row of the service.
__global__ void td_plan_kernel(TdParams P, const uint32_t *__restrict__ seg_start, const uint32_t *__restrict__ seg_end,
		const uint32_t *__restrict__ touched, const unsigned long long *__restrict__ ntouched_p, uint32_t *__restrict__ plan_bounds,
		uint32_t *__restrict__ plan_n, unsigned long long *__restrict__ newsum)
{
	const uint32_t ntouched = (uint32_t)*ntouched_p;

	for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < ntouched; t += gridDim.x * blockDim.x) {
		const uint32_t slot = touched[t];
		const uint32_t n = seg_end[slot] - seg_start[slot];
		uint32_t *bounds = plan_bounds + (size_t)slot * PLAN_STRIDE;
		unsigned long long *sums = newsum + (size_t)slot * TD_CAP;
		uint32_t nnew = 0, s = 0;

		while (s < n) {
			const double wl = td_wlimit(s, n, P);
			unsigned long long ee = (unsigned long long)floor(wl);
			if (ee > n) ee = n;
			if (ee < (unsigned long long)s + 1) ee = s + 1;
			if (nnew == TD_CAP - 1) ee = n;			// the last slot absorbs whatever is left
			bounds[nnew] = s;
			sums[nnew] = 0;
			nnew++;
			s = (uint32_t)ee;
		}
		bounds[nnew] = n;
		plan_n[slot] = nnew | (nnew == n ? 0x80000000u : 0u);	// flag: every cluster is a single sample, cluster id == rank
	}
}

// (2) sums: one thread per sorted sample. cluster id = position of the sample's rank in the service's bounds; histogram bucket
// = RESP_TIME_HASH of its msec value. Both are monotone in the rank, so runs of equal (service, cluster, bucket) are contiguous
// in the sorted order: a lane combines its own consecutive samples, the warp groups the runs with match.any (one group for the
// whole warp inside a hot service) and the group leader issues
//   one 64-bit RED with the exact usec sum of the run into the cluster sum (t-digest), and
//   GY_HISTOGRAM::add_data for the whole run (common/gy_statistics.h:596-623): count and msec sum of the bucket cell, plus the
//   run's CONN_BITMAP bits (TCP_LISTENER::CONN_BITMAP::add_response, common/gy_socket_stat.h:403-410, transposed: one mask over
//   (client port & 31) per bucket).
// max_val_seen_ is the service's last sorted sample: td_merge_kernel records it. Skew-immune: a hot service's samples are spread
// over as many warps as it has samples / 128, and no cell sees more than one RED per warp.
__device__ __forceinline__ uint32_t td_cluster_of(const uint32_t *__restrict__ bounds, uint32_t nn, uint32_t r, uint32_t &lo_out, uint32_t &hi_out)
{
	uint32_t lo = 0, hi = nn - 1;			// largest j with bounds[j] <= r
	while (lo < hi) { const uint32_t mid = (lo + hi + 1) >> 1; if (bounds[mid] <= r) lo = mid; else hi = mid - 1; }
	lo_out = bounds[lo]; hi_out = bounds[lo + 1];
	return lo;
}

static constexpr int TDS_V = 4;			// consecutive sorted samples per lane

__global__ void __launch_bounds__(256) td_sums_kernel(DevState st, const unsigned long long *__restrict__ keys, const unsigned long long *__restrict__ d_n,
		const uint32_t *__restrict__ seg_start, const uint32_t *__restrict__ plan_bounds, const uint32_t *__restrict__ plan_n,
		unsigned long long *__restrict__ newsum)
{
	const uint64_t n = *d_n;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * TDS_V;
	const int lane = threadIdx.x & 31;

	for (uint64_t base = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x - lane) * TDS_V; base < n; base += stride) {
		const uint64_t i0 = base + (uint64_t)lane * TDS_V;
		unsigned long long kk[TDS_V];
		if (i0 + TDS_V <= n) {
			const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(keys + i0), c = *reinterpret_cast<const ulonglong2 *>(keys + i0 + 2);
			kk[0] = a.x; kk[1] = a.y; kk[2] = c.x; kk[3] = c.y;
		}
		else {
#pragma unroll
			for (int t = 0; t < TDS_V; ++t) kk[t] = i0 + t < n ? keys[i0 + t] : KEY_SENTINEL;
		}

		// the whole warp (128 consecutive samples) usually sits inside one cluster of one hot service: lane 0 looks its first
		// sample up, every sample first checks that range
		uint32_t slot0 = 0xFFFFFFFFu, j0 = 0, lo0 = 1, hi0 = 0;
		if (lane == 0 && kk[0] != KEY_SENTINEL) {
			slot0 = key_slot(kk[0]);
			const uint32_t nn = plan_n[slot0], r = (uint32_t)(i0 - seg_start[slot0]);
			if (nn & 0x80000000u) { j0 = r; lo0 = r; hi0 = r + 1; }
			else j0 = td_cluster_of(plan_bounds + (size_t)slot0 * PLAN_STRIDE, nn, r, lo0, hi0);
		}
		slot0 = __shfl_sync(0xffffffffu, slot0, 0); j0 = __shfl_sync(0xffffffffu, j0, 0);
		lo0 = __shfl_sync(0xffffffffu, lo0, 0); hi0 = __shfl_sync(0xffffffffu, hi0, 0);

		// own samples -> runs of equal (service, cluster, bucket); sorted input keeps them contiguous. Sample t carries the totals
		// of its run so far; only the last sample of a run (its tail) is emitted. Everything is indexed statically (registers).
		unsigned long long pg[TDS_V];		// (slot * TD_CAP + cluster) << 4 | bucket
		unsigned long long ps[TDS_V];		// usec sum
		uint32_t pc[TDS_V], pm[TDS_V], pb[TDS_V];	// samples, msec sum, CONN_BITMAP bits
		bool tail[TDS_V];
		uint32_t cslot = 0xFFFFFFFFu, cstart = 0, cnn = 0, clo = 1, chi = 0, cj = 0;	// cached lookup of the previous sample
#pragma unroll
		for (int t = 0; t < TDS_V; ++t) {
			const bool valid = kk[t] != KEY_SENTINEL;
			tail[t] = valid;
			pg[t] = 0xFFFFFFFFFFFFFF00ull + lane; ps[t] = 0; pc[t] = 0; pm[t] = 0; pb[t] = 0;
			if (!valid) continue;
			const uint32_t slot = key_slot(kk[t]), v = key_usec(kk[t]), bit = 1u << ((uint32_t)kk[t] & 0x1Fu);
			if (slot != cslot) { cslot = slot; cstart = seg_start[slot]; cnn = plan_n[slot]; clo = 1; chi = 0; }
			const uint32_t r = (uint32_t)(i0 + t - cstart);
			uint32_t j;
			if (cnn & 0x80000000u) j = r;
			else if (slot == slot0 && r >= lo0 && r < hi0) j = j0;
			else if (r >= clo && r < chi) j = cj;
			else { cj = td_cluster_of(plan_bounds + (size_t)slot * PLAN_STRIDE, cnn, r, clo, chi); j = cj; }
			const uint32_t ms = v / 1000u;			// usec -> msec as SVC_INFO_CAP::upd_stats_on_req (gy_proto_parser.cc:2678)
			pg[t] = ((unsigned long long)(slot * (uint32_t)TD_CAP + j) << 4) | (uint32_t)bucket_resp_time((long long)ms);
			ps[t] = v; pc[t] = 1; pm[t] = ms; pb[t] = bit;
			if (t > 0 && pg[t] == pg[t - 1]) {		// continues the previous sample's run (an invalid predecessor never matches)
				ps[t] += ps[t - 1]; pc[t] += pc[t - 1]; pm[t] += pm[t - 1]; pb[t] |= pb[t - 1];
				tail[t - 1] = false;
			}
		}

		// emit the run tails: round t handles every lane's run ending at its sample t; inside a hot service the only tail of a
		// lane is its last sample and the whole warp forms one group
#pragma unroll
		for (int t = 0; t < TDS_V; ++t) {
			const bool act = tail[t];
			if (!__any_sync(0xffffffffu, act)) continue;
			const unsigned long long gid = act ? pg[t] : (0xFFFFFFFFFFFFFF00ull + lane);
			const unsigned long long v = act ? ps[t] : 0ull;		// < 2^32
			uint32_t cnt = act ? pc[t] : 0u, msum = act ? pm[t] : 0u, bits = act ? pb[t] : 0u;	// msum <= 4e6 per lane
			const uint32_t m = __match_any_sync(0xffffffffu, gid);
			unsigned long long gsum = v;
			if (m == 0xffffffffu) {
				gsum = (unsigned long long)__reduce_add_sync(0xffffffffu, (uint32_t)v & 0xFFFFu) +
						((unsigned long long)__reduce_add_sync(0xffffffffu, (uint32_t)(v >> 16)) << 16);
				cnt = __reduce_add_sync(0xffffffffu, cnt);
				msum = __reduce_add_sync(0xffffffffu, msum);
				bits = __reduce_or_sync(0xffffffffu, bits);
			}
			else {
				const uint32_t maxcnt = __reduce_max_sync(0xffffffffu, (uint32_t)__popc(m));
				uint32_t rest = m & ~(1u << lane);
				const uint32_t cnt0 = cnt, msum0 = msum, bits0 = bits;
				for (uint32_t u = 1; u < maxcnt; ++u) {
					const int src = rest ? (__ffs(rest) - 1) : lane;
					const unsigned long long ov = __shfl_sync(0xffffffffu, v, src);
					const uint32_t oc = __shfl_sync(0xffffffffu, cnt0, src), om = __shfl_sync(0xffffffffu, msum0, src), ob = __shfl_sync(0xffffffffu, bits0, src);
					if (rest) { gsum += ov; cnt += oc; msum += om; bits |= ob; rest &= rest - 1; }
				}
			}
			if (act && (m & ((1u << lane) - 1u)) == 0) {
				const uint32_t cj2 = (uint32_t)(gid >> 4);			// slot * TD_CAP + cluster
				const uint32_t cell = (cj2 / (uint32_t)TD_CAP) * HIST_CELLS + ((uint32_t)gid & 15u);
				red_add_u64(newsum + cj2, gsum);
				red_add_u64(&st.hist_cur[cell].count, cnt);
				red_add_u64((unsigned long long *)&st.hist_cur[cell].sum, msum);
				atomicOr(st.bm_cur + cell, bits);
			}
		}
	}
}

// (3) merge: one warp per touched service turns (sums, bounds) into the new clusters, merges them with the old centroids
// (old first on ties) and runs the greedy pass again
__global__ void __launch_bounds__(TD_WARPS * 32) td_merge_kernel(DevState st, const unsigned long long *__restrict__ keys,
		const uint32_t *__restrict__ seg_start, const uint32_t *__restrict__ seg_end, const uint32_t *__restrict__ touched,
		const unsigned long long *__restrict__ ntouched_p, const uint32_t *__restrict__ plan_bounds, const uint32_t *__restrict__ plan_n,
		const unsigned long long *__restrict__ newsum)
{
	__shared__ TdScratch scratch[TD_WARPS];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	TdScratch &S = scratch[wid];
	const uint32_t ntouched = (uint32_t)*ntouched_p;
	const uint32_t nwarps = gridDim.x * TD_WARPS;
	for (uint32_t t = blockIdx.x * TD_WARPS + wid; t < ntouched; t += nwarps) {
		const uint32_t slot = touched[t];
		const uint32_t s0 = seg_start[slot];
		const uint32_t n = seg_end[slot] - s0;
		const uint32_t nnew = plan_n[slot] & 0x7FFFFFFFu;
		const uint32_t *bounds = plan_bounds + (size_t)slot * PLAN_STRIDE;
		const unsigned long long *sums = newsum + (size_t)slot * TD_CAP;

		for (uint32_t j = lane; j < nnew; j += 32) {
			const uint32_t w = bounds[j + 1] - bounds[j];
			S.newc[j].mean = __ddiv_rn((double)sums[j], (double)w);		// cluster sums are exact integers
			S.newc[j].weight = w;
		}
		const uint32_t us_max = key_usec(keys[s0 + n - 1]);
		const double bmin = (double)key_usec(keys[s0]), bmax = (double)us_max;
		// max_val_seen_ of GY_HISTOGRAM::add_data (gy_statistics.h:609-611): the batch maximum is the last sorted sample
		if (lane == 0) atomicMax(&st.hist_cur[(size_t)slot * HIST_CELLS + HIST_MAX_CELL].sum, (long long)(us_max / 1000u));
		__syncwarp();

		TdHead head = st.td_head[slot];
		Centroid *cent = st.td_cent + (size_t)slot * TD_CAP;
		const uint32_t nout = warp_merge_compress(S, cent, head.n, S.newc, nnew, cent, st.td);
		if (lane == 0) {
			head.n = nout;
			head.total += n;
			if (bmin < head.minv) head.minv = bmin;
			if (bmax > head.maxv) head.maxv = bmax;
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
	const uint32_t bal = __ballot_sync(0xffffffffu, valid && (cell == HIST_MAX_CELL ? cc != 0 : c.count != 0));
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

// one CTA per evicted slot (grid-stride): the id's table entry becomes a tombstone, every per-slot array returns to its
// just-created state and the slot number goes on the free stack for the next unknown id
__global__ void __launch_bounds__(256) evict_kernel(DevState st, uint32_t max_svcs)
{
	const uint32_t nev = (uint32_t)st.counters[CTR_NEVICT];

	for (uint32_t q = blockIdx.x; q < nev; q += gridDim.x) {
		const uint32_t slot = st.evict_list[q];
		const unsigned long long id = st.evict_ids[q];

		if (threadIdx.x == 0) {
			uint32_t pos = uint64_hash(id) & st.svc_tbl.mask;
			for (uint32_t probe = 0; probe <= st.svc_tbl.mask; ++probe, pos = (pos + 1) & st.svc_tbl.mask) {
				TblEntry *e = &st.svc_tbl.ent[pos];
				if (e->key == id) { e->key = KEY_TOMBSTONE; e->slot1 = 0; break; }
				if (e->key == 0) break;
			}
			st.slot_id[slot] = 0; st.slot_host[slot] = 0; st.slot_first_seen[slot] = 0; st.slot_last_active[slot] = 0;
			st.conn_cur[slot] = 0; st.conn_last[slot] = 0; st.conn_all_cnt[slot] = 0; st.conn_all_kb[slot] = 0;
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
	uint32_t pos = uint64_hash(id) & st.svc_tbl.mask;
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
	}
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

// GYSK_INGEST_VARIANT selects a CTA shape of the tile pipeline for A/B runs: 256 (default, measured best) / 128 / 2563 (TMA-staged
// tiles) / 2562 / 2568 (2 / 8 events per thread)
static int ingest_variant()
{
	static const int v = []{ const char *e = getenv("GYSK_INGEST_VARIANT"); return e ? atoi(e) : 2562; }();
	return v;
}

template <int THREADS, int MIN_CTAS, bool STAGE, int EPT = 4, bool PIPE = false>
static void launch_ingest_variant(const DevState &st, const gysk_event *d_ev, uint64_t n, unsigned long long *d_keys, int nsm, cudaStream_t s)
{
	using Shared = IngestSharedT<THREADS, STAGE, EPT>;
	static bool attr_set = false;
	if (!attr_set) { cudaFuncSetAttribute(ingest_kernel<THREADS, MIN_CTAS, STAGE, EPT, PIPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Shared)); attr_set = true; }
	const uint64_t want = (n + Shared::INGEST_TILE - 1) / Shared::INGEST_TILE;
	const uint32_t grid = (uint32_t)(want < (uint64_t)nsm * MIN_CTAS ? want : (uint64_t)nsm * MIN_CTAS);
	// A/B switch (measured: L2 prefetch of the CTA's next tile costs more issue slots than the latency it hides: 2.58 -> 2.78 ms)
	static const int prefetch_next = []{ const char *e = getenv("GYSK_INGEST_PREFETCH"); return e ? atoi(e) : 0; }();
	ingest_kernel<THREADS, MIN_CTAS, STAGE, EPT, PIPE><<<grid, THREADS, sizeof(Shared), s>>>(st, d_ev, n, d_keys, prefetch_next);
}

int launch_ingest(const DevState &st, const gysk_event *d_ev, uint64_t n, unsigned long long *d_keys, cudaStream_t s)
{
	if (!n) return 0;
	int dev = 0, nsm = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
	// variants for A/B runs (GYSK_INGEST_VARIANT): 256 = 256 thr x 4 CTAs/SM, events by direct streaming loads (measured best so
	// far); 2563 = same shape, next tile staged by TMA bulk copy (3 CTAs/SM: +32 KB smem); 128 = 128 thr x 8 CTAs/SM
	static const int variant = ingest_variant();
	cudaMemsetAsync(st.counters + CTR_NKEYS, 0, 2 * sizeof(unsigned long long), s);	// key cursor + max RESP msec of this batch
	if (variant == 2563) launch_ingest_variant<256, 3, true>(st, d_ev, n, d_keys, nsm, s);
	else if (variant == 128) launch_ingest_variant<128, 8, false>(st, d_ev, n, d_keys, nsm, s);
	else if (variant == 2562) launch_ingest_variant<256, 5, false, 2>(st, d_ev, n, d_keys, nsm, s);	// 2 events per thread: fewer live registers, 5 CTAs/SM
	else if (variant == 2568) launch_ingest_variant<256, 3, false, 8>(st, d_ev, n, d_keys, nsm, s);	// 8 events per thread: fewer barriers per event
	else if (variant == 2662) launch_ingest_variant<256, 6, false, 2>(st, d_ev, n, d_keys, nsm, s);
	else if (variant == 25620) launch_ingest_variant<256, 5, false, 2, true>(st, d_ev, n, d_keys, nsm, s);	// next tile's loads issued before the barrier (measured: 2.74 vs 2.61 ms)
	else if (variant == 256) launch_ingest_variant<256, 4, false>(st, d_ev, n, d_keys, nsm, s);
	else launch_ingest_variant<256, 5, false, 2>(st, d_ev, n, d_keys, nsm, s);
	return 1;
}

// stable LSD radix sort of the n keys in bufs[0] on the significant key bits [lo1, hi1) then [lo2, hi2) (lo2 >= hi1; pass
// hi2 <= lo2 for a single range): the significant bits are cut into digits in order, a digit may straddle the gap. Digits are
// 8 bits wide unless 9-bit digits save a whole pass (41-45 significant bits: 5 passes instead of 6). Result in bufs[*which].
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

int launch_radix_sort(const SortTemp &tmp, uint64_t n_keys, int lo1, int hi1, int lo2, int hi2, int *which, cudaStream_t s)
{
	int launches = 0;
	unsigned long long *bufs[2] = { tmp.keys_a, tmp.keys_b };
	int w = 0;
	DigitSpecs P;

	*which = 0;
	if (!n_keys) return 0;
	if (build_digit_specs(lo1, hi1, lo2, hi2, P) || n_keys >= (1ull << 30)) return -1;	// status words carry 30-bit counts
	bool any9 = false;
	for (int p = 0; p < P.np; ++p) any9 |= P.d[p].b1 + P.d[p].b2 > 8;
	const int copies = any9 ? 4 : 8, stride = (any9 ? 512 : 256) + 1;

	// one-sweep passes: global digit histograms of all passes from one read, then 16 B per key and pass
	static bool attr_set = false;
	if (!attr_set) {
		cudaFuncSetAttribute(os_pass_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OneSweepSharedT<8>));
		cudaFuncSetAttribute(os_pass_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OneSweepSharedT<9>));
		cudaFuncSetAttribute(os_hist_kernel<8, 257>, cudaFuncAttributeMaxDynamicSharedMemorySize, OS_MAX_PASSES * 8 * 257 * (int)sizeof(uint32_t));
		cudaFuncSetAttribute(os_hist_kernel<4, 513>, cudaFuncAttributeMaxDynamicSharedMemorySize, OS_MAX_PASSES * 4 * 513 * (int)sizeof(uint32_t));
		attr_set = true;
	}
	static const int rank_mode = []{ const char *e = getenv("GYSK_OS_RANK"); return e ? atoi(e) : 0; }();
	const uint32_t n = (uint32_t)n_keys;
	const uint32_t ntiles = div_up(n, SORT_TILE);
	uint32_t *ghist = tmp.os_ghist, *tickets = tmp.os_ghist + OS_MAX_PASSES * RADIX_MAX;
	int dev = 0, nsm = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);

	cudaMemsetAsync(ghist, 0, (OS_MAX_PASSES * RADIX_MAX + OS_MAX_PASSES) * sizeof(uint32_t), s);
	const uint32_t hgrid = std::min<uint32_t>(div_up(n, 512 * 2 * 4), (uint32_t)nsm * 3);
	if (any9) os_hist_kernel<4, 513><<<hgrid, 512, (size_t)P.np * copies * stride * sizeof(uint32_t), s>>>(bufs[w], n, P, ghist);
	else os_hist_kernel<8, 257><<<hgrid, 512, (size_t)P.np * copies * stride * sizeof(uint32_t), s>>>(bufs[w], n, P, ghist);
	launches++;
	for (int p = 0; p < P.np; ++p) {
		const bool nine = P.d[p].b1 + P.d[p].b2 > 8;
		cudaMemsetAsync(tmp.tile_status, 0, (size_t)ntiles * (nine ? 512 : 256) * sizeof(uint32_t), s);
		if (nine) os_pass_kernel<9><<<ntiles, OS_THREADS, sizeof(OneSweepSharedT<9>), s>>>(bufs[w], bufs[w ^ 1], n, P.d[p], ghist + p * RADIX_MAX, tmp.tile_status, tickets + p, rank_mode);
		else os_pass_kernel<8><<<ntiles, OS_THREADS, sizeof(OneSweepSharedT<8>), s>>>(bufs[w], bufs[w ^ 1], n, P.d[p], ghist + p * RADIX_MAX, tmp.tile_status, tickets + p, rank_mode);
		launches++;
		w ^= 1;
	}
	*which = w;
	return launches;
}

// sort the (slot, usec) keys produced by ingest, then fold every touched service's new samples into its digest
int launch_tdigest_update(const DevState &st, const SortTemp &tmp, uint64_t n, uint32_t nslots, int value_bits, cudaStream_t s)
{
	if (!n) return 0;		// n = number of RESP keys of this batch (read back by the host), nslots = services registered so far
	int launches = 0;
	const uint32_t ntiles = div_up(n, SORT_TILE);
	unsigned long long *d_nkeys = st.counters + CTR_NKEYS, *d_ntouched = st.counters + CTR_NTOUCHED;

	// radix passes cover only the bits that can differ: usec bits of the batch maximum, then the slot bits in use
	uint32_t slot_bits = 1;
	while (slot_bits < 32 && (1ull << slot_bits) < nslots) slot_bits++;

	const unsigned long long *src = tmp.keys_a;
	unsigned long long *bufs[2] = { tmp.keys_a, tmp.keys_b };
	int which = 0;

	cudaMemsetAsync(d_ntouched, 0, sizeof(unsigned long long), s);

	// with the warp-autonomous ingest the keys sit at their events' positions with sentinels in between: the first pass reads
	// n_events slots and compacts, the later passes run over the n keys
	const int sorted = launch_radix_sort(tmp, n, KEY_VALUE_SHIFT, KEY_VALUE_SHIFT + value_bits, KEY_SLOT_SHIFT, KEY_SLOT_SHIFT + (int)slot_bits, &which, s);
	if (sorted < 0) return sorted;
	launches += sorted;
	src = bufs[which];

	td_segments_kernel<<<div_up(n, 256 * TSEG_V), 256, 0, s>>>(src, d_nkeys, tmp.seg_start, tmp.seg_end, tmp.touched, d_ntouched);
	int dev = 0, nsm = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
	td_plan_kernel<<<nsm * 2, 128, 0, s>>>(st.td, tmp.seg_start, tmp.seg_end, tmp.touched, d_ntouched, tmp.plan_bounds, tmp.plan_n, tmp.newsum);
	td_sums_kernel<<<nsm * 8, 256, 0, s>>>(st, src, d_nkeys, tmp.seg_start, tmp.plan_bounds, tmp.plan_n, tmp.newsum);
	td_merge_kernel<<<nsm * 7, TD_WARPS * 32, 0, s>>>(st, src, tmp.seg_start, tmp.seg_end, tmp.touched, d_ntouched, tmp.plan_bounds, tmp.plan_n, tmp.newsum);
	return launches + 4;
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
	const int sorted = launch_radix_sort(tmp, nslots, 32, 64, 64, 64, &which, s);
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
__global__ void topn_task_score_kernel(DevState st, uint32_t ntasks, int metric, unsigned long long *__restrict__ keys)
{
	const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
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
	topn_task_score_kernel<<<div_up(ntasks, 256), 256, 0, s>>>(st, ntasks, metric, tmp.keys_a);
	const int sorted = launch_radix_sort(tmp, ntasks, 32, 64, 64, 64, &which, s);
	if (sorted < 0) return sorted;
	launches += sorted;
	topn_task_pick_kernel<<<1, 64, 0, s>>>(st, which ? tmp.keys_b : tmp.keys_a, ntasks, want, d_out);
	return launches;
}

int launch_flush(const DevState &st, uint32_t nslots, HistCell *ring_plane0, HistCell *ring_plane1, uint32_t tsec, uint32_t idle_secs, cudaStream_t s)
{
	if (!nslots) return 0;
	cudaMemsetAsync(st.counters + CTR_NEVICT, 0, sizeof(unsigned long long), s);
	flush_kernel<<<div_up((uint64_t)nslots * HIST_CELLS, 256), 256, 0, s>>>(st, nslots, ring_plane0, ring_plane1, tsec, idle_secs);
	if (!idle_secs) return 1;
	evict_kernel<<<296, 256, 0, s>>>(st, nslots);		// grid-stride over the (device-side) eviction list
	return 2;
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
