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
#include <cstdlib>

namespace gysk {

static constexpr unsigned long long KEY_SENTINEL = ~0ull;
static constexpr unsigned long long VALUE_MASK = (1ull << VALUE_BITS) - 1;

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

	if (i < n && ids[i]) table_lookup(is_task ? st.task_tbl : st.svc_tbl, ids[i], true);
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
	uint32_t		bits[N];		// CONN_BITMAP bits of a RESP cell
};
static constexpr uint32_t CELL_TASK = 1u << 30;			// cell ids: svc hist = slot*16 + bucket, conn = slot*16 + 15,
								//           task = CELL_TASK | (tslot*48 + hist*16 + bucket)

__device__ __forceinline__ bool cell_is_conn(uint32_t cell) { return !(cell & CELL_TASK) && (cell & 15u) == (uint32_t)HIST_MAX_CELL; }

// one RED per field, nothing is read back: {count, sum} (+ max_val_seen_ for histogram cells)
__device__ __forceinline__ void cell_add_global(const DevState &st, uint32_t cell, uint32_t cnt, unsigned long long sum, int vmax, uint32_t bits)
{
	if (cell & CELL_TASK) {
		HistCell *c = st.task_hist + (cell & ~CELL_TASK);
		red_add_u64(&c->count, cnt); red_add_u64((unsigned long long *)&c->sum, sum);
		atomicMax(&st.task_hist[(cell & ~CELL_TASK) | 15u].sum, (long long)vmax);
	}
	else if ((cell & 15u) == (uint32_t)HIST_MAX_CELL) {
		red_add_u64(st.conn_cur + (cell >> 4), (unsigned long long)cnt + (sum << 32));	// packed {count, kbytes}
	}
	else {
		HistCell *c = st.hist_cur + cell;
		red_add_u64(&c->count, cnt); red_add_u64((unsigned long long *)&c->sum, sum);
		atomicMax(&st.hist_cur[cell | 15u].sum, (long long)vmax);
		// TCP_LISTENER::CONN_BITMAP::add_response (common/gy_socket_stat.h:403-410), transposed: the cell index is the same
		if (bits) atomicOr(st.bm_cur + cell, bits);
	}
}

// all 32 lanes call this; lanes with active == false only take part in the collectives.
// Group sums use a shuffle loop bounded by the largest group of the warp (typically 1-4): redux with per-lane masks would
// make the compiler iterate over every distinct group. No global load anywhere: the updates are fire-and-forget REDs.
template <typename HotTable>
__device__ __forceinline__ void cell_add(const DevState &st, HotTable &hot, bool active, uint32_t cell, int data, uint32_t bits = 0)
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
		const uint32_t obits = __shfl_sync(0xffffffffu, bits, src);
		if (rest) { sum += other; gmax = max(gmax, other); bits |= obits; rest &= rest - 1; }
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
	if (hit) {
		atomicAdd(&hot.count[h], cnt); atomicAdd(&hot.sum[h], (unsigned long long)sum); atomicMax(&hot.vmax[h], gmax);
		if (bits) atomicOr(&hot.bits[h], bits);
	}
	else cell_add_global(st, cell, cnt, (unsigned long long)sum, gmax, bits);
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
//   phase 1 (one event per thread and round, fully converged): 2 x 128-bit load, shard filter, id -> slot lookup, then the
//           decoded record {slot, value, flow_key} goes to shared memory and its index to the queue of its kind
//           (RESP / TCP / TASK; ballot + one shared-memory atomic per warp and kind);
//   phase 2 (converged per kind): every warp strides over one queue at a time, so the RESP histogram code, the count-min
//           rows (one (event, row) pair per lane), the HLL updates and the three task histograms (one (event, histogram)
//           pair per lane) each run with all lanes doing the same thing instead of serialising 70/20/10-divergent branches.
// RESP sort keys are written compacted (one global cursor bump per tile), so the radix sort never sees a non-RESP slot.

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
	using HotTable = HotTableT<(INGEST_THREADS >= 256 ? 10 : 9)>;
	alignas(128) uint4	evbuf[STAGE ? INGEST_TILE * 2 : 1];	// next tile of 32-byte events, filled by cp.async.bulk
	unsigned long long	mbar;
	HotTable	hot;
	IngestRec	rec[INGEST_TILE];
	uint16_t	q_resp[INGEST_TILE], q_tcp[INGEST_TILE], q_task[INGEST_TILE];
	uint32_t	qn[2][4];		// per tile parity: n_resp, n_tcp, n_task (double-buffered: no barrier to reset them)
	unsigned long long key_base;
	uint32_t	max_value;		// largest RESP usec seen by this CTA (sizes the radix sort)
};

__device__ __forceinline__ void queue_push(bool pred, uint16_t *q, uint32_t *qn, uint16_t item)
{
	const uint32_t m = __ballot_sync(0xffffffffu, pred);
	if (!m) return;
	const int lane = threadIdx.x & 31;
	uint32_t base = 0;
	if (lane == __ffs(m) - 1) base = atomicAdd(qn, (uint32_t)__popc(m));
	base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
	if (pred) q[base + __popc(m & ((1u << lane) - 1u))] = item;
}

template <int INGEST_THREADS, int MIN_CTAS, bool STAGE, int INGEST_EPT>
__global__ void __launch_bounds__(INGEST_THREADS, MIN_CTAS) ingest_kernel(DevState st, const gysk_event *__restrict__ ev, uint64_t n,
		unsigned long long *__restrict__ keys)
{
	using IngestShared = IngestSharedT<INGEST_THREADS, STAGE, INGEST_EPT>;
	using HotTable = typename IngestShared::HotTable;
	constexpr int INGEST_TILE = IngestShared::INGEST_TILE;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	IngestShared &S = *reinterpret_cast<IngestShared *>(smem_raw);
	unsigned long long c_in = 0, c_drop = 0, c_resp = 0, c_tcp = 0, c_task = 0, c_foreign = 0;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint64_t ntiles = (n + INGEST_TILE - 1) / INGEST_TILE;

	for (int i = threadIdx.x; i < HotTable::N; i += INGEST_THREADS) { S.hot.tag[i] = 0; S.hot.count[i] = 0; S.hot.sum[i] = 0; S.hot.vmax[i] = INT_MIN; S.hot.bits[i] = 0; }
	if (threadIdx.x < 8) (&S.qn[0][0])[threadIdx.x] = 0;
	if (threadIdx.x == 0) S.max_value = 0;
	if (STAGE && threadIdx.x == 0) mbar_init(&S.mbar, 1);
	__syncthreads();
	// prologue: the first tile of this CTA starts flying into shared memory
	if (STAGE && threadIdx.x == 0 && blockIdx.x < ntiles) {
		const uint64_t b0 = (uint64_t)blockIdx.x * INGEST_TILE;
		const uint64_t cnt = n - b0 < (uint64_t)INGEST_TILE ? n - b0 : (uint64_t)INGEST_TILE;
		tma_load_1d(S.evbuf, ev + b0, (uint32_t)cnt * 32u, &S.mbar);
	}

	int par = 0;
	for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, par ^= 1) {
		const uint64_t tbase = tile * INGEST_TILE;
		uint32_t *qn = S.qn[par];
		if (STAGE) mbar_wait(&S.mbar, (uint32_t)par);		// this tile's events have landed (one completion per tile)

		// ---------------- phase 1: decode + lookup + enqueue ----------------
		uint4 ra[INGEST_EPT], rb[INGEST_EPT];
#pragma unroll
		for (int k = 0; k < INGEST_EPT; ++k) {
			const uint64_t i = tbase + (uint64_t)k * INGEST_THREADS + threadIdx.x;
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
		// decode; put the first id-table probe of all EPT events in flight before any of them is resolved
		unsigned long long svc[INGEST_EPT];
		uint4 praw[INGEST_EPT];
		uint32_t ppos[INGEST_EPT];
		uint32_t kind[INGEST_EPT];		// 0 none, GYSK_EV_RESP, 1 tcp (stored as GYSK_EV_ACCEPT), GYSK_EV_TASK
#pragma unroll
		for (int k = 0; k < INGEST_EPT; ++k) {
			svc[k] = ((unsigned long long)ra[k].y << 32) | ra[k].x;
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
				if (svc[k] != 0 && (is_tcp || is_task || (is_resp && value / 1000u <= 1000000u))) {
					kind[k] = is_resp ? (uint32_t)GYSK_EV_RESP : (is_task ? (uint32_t)GYSK_EV_TASK : (uint32_t)GYSK_EV_ACCEPT);
					praw[k] = table_probe_first(is_task ? st.task_tbl : st.svc_tbl, svc[k], ppos[k]);
				}
				else c_drop++;
			}
		}
#pragma unroll
		for (int k = 0; k < INGEST_EPT; ++k) {
			const bool is_resp = kind[k] == GYSK_EV_RESP, is_task = kind[k] == GYSK_EV_TASK, is_tcp = kind[k] == GYSK_EV_ACCEPT;
			int slot = -1;
			if (kind[k]) {
				// one id lookup for all three event kinds (services and tasks live in separate tables)
				slot = table_resolve(is_task ? st.task_tbl : st.svc_tbl, svc[k], st.auto_register, rb[k].y, ppos[k], praw[k]);
				if (slot < 0) c_drop++;
				else if (is_resp) c_resp++;
				else if (is_tcp) c_tcp++;
				else c_task++;
			}
			const uint16_t pos = (uint16_t)(k * INGEST_THREADS + threadIdx.x);
			if (slot >= 0) {
				IngestRec r; r.slot = (uint32_t)slot; r.value = rb[k].x; r.flow_key = ((unsigned long long)ra[k].w << 32) | ra[k].z;
				S.rec[pos] = r;
			}
			queue_push(slot >= 0 && is_resp, S.q_resp, &qn[0], pos);
			queue_push(slot >= 0 && is_tcp, S.q_tcp, &qn[1], pos);
			queue_push(slot >= 0 && is_task, S.q_task, &qn[2], pos);
		}
		__syncthreads();
		const uint32_t n_resp = qn[0], n_tcp = qn[1], n_task = qn[2];
		// thread 0 bumps the global key cursor now; its round trip to L2 hides behind the TCP and TASK phases
		if (threadIdx.x == 0) {
			S.qn[par ^ 1][0] = 0; S.qn[par ^ 1][1] = 0; S.qn[par ^ 1][2] = 0;
			S.key_base = n_resp ? atomicAdd(st.counters + CTR_NKEYS, (unsigned long long)n_resp) : 0ull;
			// every thread has consumed evbuf (barrier above): stage the next tile of this CTA while phase 2 runs
			if (STAGE && tile + gridDim.x < ntiles) {
				const uint64_t b1 = (tile + gridDim.x) * INGEST_TILE;
				const uint64_t cnt = n - b1 < (uint64_t)INGEST_TILE ? n - b1 : (uint64_t)INGEST_TILE;
				tma_load_1d(S.evbuf, ev + b1, (uint32_t)cnt * 32u, &S.mbar);
			}
		}

		// ---------------- phase 2b: TCP — count-min rows, one (event, row) pair per lane ----------------
		{
			const uint32_t npairs = n_tcp * st.cms_depth;
			for (uint32_t p = threadIdx.x; p < npairs; p += INGEST_THREADS) {
				const uint32_t e = p / st.cms_depth, row = p - e * st.cms_depth;
				const IngestRec r = S.rec[S.q_tcp[e]];
				red_add_u64(st.cms_cur + ((size_t)row << st.cms_log2w) + cms_index(r.flow_key, row, st.cms_wmask), cms_increment(r.value));
			}
			// HLL register + the service's exact {count, kbytes} cell
			for (uint32_t base = wid * 32; base < n_tcp; base += INGEST_THREADS) {
				const uint32_t q = base + lane;
				const bool act = q < n_tcp;
				uint32_t cell = 0; int kb = 0;
				if (act) {
					const IngestRec r = S.rec[S.q_tcp[q]];
					uint32_t idx, rank;
					hll_idx_rank(r.flow_key, st.hll_p, idx, rank);
					hll_update(st.hll + ((size_t)r.slot << st.hll_p), idx, rank);
					cell = r.slot * HIST_CELLS + HIST_MAX_CELL;
					kb = (int)(r.value >> 10);
				}
				cell_add(st, S.hot, act, cell, kb);
			}
		}
		// ---------------- phase 2c: TASK — MAGGR_TASK::set_local_task_state, one (event, histogram) pair per lane ----------------
		{
			const uint32_t ntrip = n_task * 3u;
			for (uint32_t base = wid * 32; base < ntrip; base += INGEST_THREADS) {
				const uint32_t p = base + lane;
				const bool act = p < ntrip;
				uint32_t cell = 0; int d = 0;
				if (act) {
					const uint32_t e = p / 3u, h = p - e * 3u;
					const IngestRec r = S.rec[S.q_task[e]];
					// GY_HISTOGRAM<int, ...>::add_data(int): the three values narrow to int (server/gy_msocket.h:1014-1016)
					d = h == 0 ? (int)r.value : (h == 1 ? (int)(uint32_t)r.flow_key : (int)(uint32_t)(r.flow_key >> 32));
					const uint32_t b = h == 0 ? (uint32_t)bucket_hash_1_3000(d) : (uint32_t)bucket_duration(d);
					cell = CELL_TASK | (r.slot * 3u * HIST_CELLS + h * HIST_CELLS + b);
				}
				cell_add(st, S.hot, act, cell, d);
			}
		}
		__syncthreads();			// key_base is visible
		// ---------------- phase 2d: RESP — GY_HISTOGRAM<int64_t, RESP_TIME_HASH>::add_data, gy_statistics.h:596-623 ----------------
		{
			const unsigned long long kb = S.key_base;
			for (uint32_t base = wid * 32; base < n_resp; base += INGEST_THREADS) {
				const uint32_t q = base + lane;
				const bool act = q < n_resp;
				uint32_t cell = 0, bit = 0; int ms = 0;
				if (act) {
					const IngestRec r = S.rec[S.q_resp[q]];
					bit = 1u << ((uint32_t)r.flow_key & 0x1Fu);		// CONN_BITMAP: client port & 0x1F
					ms = (int)(r.value / 1000u);
					cell = r.slot * HIST_CELLS + (uint32_t)bucket_resp_time((long long)ms);
					__stcs(keys + kb + q, ((unsigned long long)r.slot << VALUE_BITS) | r.value);
				}
				const uint32_t wmax = __reduce_max_sync(0xffffffffu, act ? (uint32_t)ms : 0u);	// msec is enough: bits(usec) <= bits(msec) + 10
				if (lane == 0 && wmax > S.max_value) atomicMax(&S.max_value, wmax);
				cell_add(st, S.hot, act, cell, ms, bit);
			}
		}
		__syncthreads();			// rec / queues may be overwritten by the next tile
	}

	if (threadIdx.x == 0 && S.max_value) atomicMax(st.counters + CTR_MAXVAL, (unsigned long long)S.max_value);
	// retire: one RED group per privatised cell
	for (int i = threadIdx.x; i < HotTable::N; i += INGEST_THREADS) {
		if (S.hot.tag[i] && S.hot.count[i]) cell_add_global(st, S.hot.tag[i] - 1, S.hot.count[i], S.hot.sum[i], S.hot.vmax[i], S.hot.bits[i]);
	}

	// statsmap-style counters (gy_mconnhdlr.cc:4708-4715): warp-reduce, one atomic per warp and counter
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) {
		c_in += __shfl_down_sync(0xffffffffu, c_in, off);
		c_drop += __shfl_down_sync(0xffffffffu, c_drop, off);
		c_resp += __shfl_down_sync(0xffffffffu, c_resp, off);
		c_tcp += __shfl_down_sync(0xffffffffu, c_tcp, off);
		c_task += __shfl_down_sync(0xffffffffu, c_task, off);
		c_foreign += __shfl_down_sync(0xffffffffu, c_foreign, off);
	}
	if (lane == 0) {
		if (c_in) atomicAdd(st.counters + CTR_IN, c_in);
		if (c_drop) atomicAdd(st.counters + CTR_DROPPED, c_drop);
		if (c_resp) atomicAdd(st.counters + CTR_RESP, c_resp);
		if (c_tcp) atomicAdd(st.counters + CTR_TCP, c_tcp);
		if (c_task) atomicAdd(st.counters + CTR_TASK, c_task);
		if (c_foreign) atomicAdd(st.counters + CTR_FOREIGN, c_foreign);
	}
}

// ---------------------------------------------------------------------------------------------------
// stable LSD radix sort, 8-bit digits, tile = SORT_TILE keys per CTA of 256 threads
// ---------------------------------------------------------------------------------------------------
static constexpr int RS_THREADS = 512;
static constexpr int RS_WARPS = RS_THREADS / 32;
static constexpr int RS_ROUNDS = SORT_TILE / RS_THREADS;	// 16 keys per thread
static constexpr int RADIX = 256;

// one radix pass sorts on an 8-bit digit made of up to two bit fields of the key, so that the unused bits between the usec field
// and the slot field of a key never cost a pass: digit = ((k >> s1) & m1) | (((k >> s2) & m2) << b1)
struct DigitSpec { int s1, b1, s2, b2; };
__device__ __forceinline__ uint32_t key_digit(unsigned long long k, const DigitSpec &D)
{
	return ((uint32_t)(k >> D.s1) & ((1u << D.b1) - 1u)) | (((uint32_t)(k >> D.s2) & ((1u << D.b2) - 1u)) << D.b1);
}

// per-tile digit histogram -> tile_hist[digit * ntiles + tile]
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const unsigned long long *__restrict__ keys, uint64_t n_host,
		const unsigned long long *__restrict__ d_n, DigitSpec D, uint32_t *__restrict__ tile_hist, uint32_t ntiles)
{
	__shared__ uint32_t hist[RADIX];
	const uint64_t n = d_n ? *d_n : n_host;
	const uint32_t tile = blockIdx.x;

	if (threadIdx.x < RADIX) hist[threadIdx.x] = 0;
	__syncthreads();

	const uint64_t base = (uint64_t)tile * SORT_TILE;
#pragma unroll
	for (int r = 0; r < RS_ROUNDS; ++r) {
		const uint64_t i = base + (uint64_t)r * RS_THREADS + threadIdx.x;
		if (i < n) {
			const unsigned long long k = keys[i];
			if (k != KEY_SENTINEL) atomicAdd(&hist[key_digit(k, D)], 1u);
		}
	}
	__syncthreads();
	if (threadIdx.x < RADIX) tile_hist[(size_t)threadIdx.x * ntiles + tile] = hist[threadIdx.x];
}

// exclusive scan of a u32 array of length len: reduce / scan block sums / apply
static constexpr int SCAN_THREADS = 256;
static constexpr int SCAN_ITEMS = 8;
static constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total_out, uint32_t *smem /* >= 32 */)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	uint32_t incl = v;

#pragma unroll
	for (int off = 1; off < 32; off <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
		if (lane >= off) incl += t;
	}
	if (lane == 31) smem[wid] = incl;
	__syncthreads();
	if (wid == 0) {
		const int nw = blockDim.x >> 5;
		uint32_t w = lane < nw ? smem[lane] : 0, wi = w;
#pragma unroll
		for (int off = 1; off < 32; off <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, wi, off);
			if (lane >= off) wi += t;
		}
		smem[lane] = wi - w;			// exclusive warp offsets
		if (lane == nw - 1 && total_out) *total_out = wi;
	}
	__syncthreads();
	const uint32_t res = smem[wid] + incl - v;
	__syncthreads();
	return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint32_t *__restrict__ in, uint32_t len, uint32_t *__restrict__ block_sums)
{
	__shared__ uint32_t smem[32];
	__shared__ uint32_t total;
	const uint32_t base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
	uint32_t s = 0;

#pragma unroll
	for (int j = 0; j < SCAN_ITEMS; ++j) if (base + j < len) s += in[base + j];
	block_exclusive_scan(s, &total, smem);
	if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single CTA: exclusive scan of block_sums[nblocks] in place; grand total -> *total_out (u64 counter)
__global__ void __launch_bounds__(1024) scan_blocksums_kernel(uint32_t *block_sums, uint32_t nblocks, unsigned long long *total_out)
{
	__shared__ uint32_t smem[32];
	__shared__ uint32_t chunk_total;
	uint32_t carry = 0;

	for (uint32_t base = 0; base < nblocks; base += blockDim.x) {
		const uint32_t i = base + threadIdx.x;
		const uint32_t v = i < nblocks ? block_sums[i] : 0;
		const uint32_t ex = block_exclusive_scan(v, &chunk_total, smem);
		if (i < nblocks) block_sums[i] = carry + ex;
		carry += chunk_total;
		__syncthreads();
	}
	if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(uint32_t *__restrict__ data, uint32_t len, const uint32_t *__restrict__ block_sums)
{
	__shared__ uint32_t smem[32];
	const uint32_t base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
	uint32_t v[SCAN_ITEMS], s = 0;

#pragma unroll
	for (int j = 0; j < SCAN_ITEMS; ++j) { v[j] = base + j < len ? data[base + j] : 0; s += v[j]; }
	uint32_t ex = block_exclusive_scan(s, nullptr, smem) + block_sums[blockIdx.x];
#pragma unroll
	for (int j = 0; j < SCAN_ITEMS; ++j) { if (base + j < len) data[base + j] = ex; ex += v[j]; }
}

// scatter one tile to its stable positions. Thread t of warp w holds keys w*256 + r*32 + lane (r = 0..7), i.e. ascending input
// order is (warp, round, lane); ranks come from match.any groups so equal digits keep that order. The tile is first reordered by
// digit in shared memory, then written out by consecutive threads: every digit's keys leave as one contiguous run (full 32-byte
// sectors) instead of 4096 scattered 8-byte stores.
struct ScatterShared
{
	unsigned long long	keys[SORT_TILE];		// tile reordered by digit
	uint32_t		whist[RS_WARPS][RADIX];		// per-warp digit counts, then local start of (warp, digit)
	uint32_t		dstart[RADIX];			// tile-local start of each digit
	uint32_t		goff[RADIX];			// global position of the tile's first key of each digit
	uint32_t		wsum[RS_WARPS];
	uint32_t		nvalid;				// keys of the tile that are not sentinels
};

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const unsigned long long *__restrict__ in, unsigned long long *__restrict__ out,
		uint64_t n_host, const unsigned long long *__restrict__ d_n, DigitSpec D, const uint32_t *__restrict__ tile_offs, uint32_t ntiles)
{
	extern __shared__ __align__(16) unsigned char rs_smem[];
	ScatterShared &S = *reinterpret_cast<ScatterShared *>(rs_smem);
	const uint64_t n = d_n ? *d_n : n_host;
	const uint32_t tile = blockIdx.x;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t lt_mask = (1u << lane) - 1u;

	if ((uint64_t)tile * SORT_TILE >= n) return;

	for (int i = threadIdx.x; i < RS_WARPS * RADIX; i += RS_THREADS) (&S.whist[0][0])[i] = 0;
	__syncthreads();

	unsigned long long k[RS_ROUNDS];
	uint32_t rank[RS_ROUNDS];
	const uint64_t wbase = (uint64_t)tile * SORT_TILE + (uint64_t)wid * (RS_ROUNDS * 32);

#pragma unroll
	for (int r = 0; r < RS_ROUNDS; ++r) {
		const uint64_t i = wbase + (uint64_t)r * 32 + lane;
		k[r] = i < n ? in[i] : KEY_SENTINEL;
	}

	// phase A: rank of every key among the keys of its digit inside this warp's chunk (rounds in order, lanes in order),
	// leaving the per-warp digit counts in whist. One match.any per round; the group leader bumps the warp counter and
	// hands the previous value to its group.
#pragma unroll
	for (int r = 0; r < RS_ROUNDS; ++r) {
		const bool valid = k[r] != KEY_SENTINEL;
		const uint32_t d = valid ? key_digit(k[r], D) : (0x100u + lane);
		const uint32_t m = __match_any_sync(0xffffffffu, d);
		const int leader = __ffs(m) - 1;
		uint32_t old = 0;
		if (valid && lane == leader) { old = S.whist[wid][d]; S.whist[wid][d] = old + __popc(m); }
		__syncwarp();
		old = __shfl_sync(0xffffffffu, old, leader);
		rank[r] = old + __popc(m & lt_mask);
	}
	__syncthreads();

	// per digit: exclusive prefix over warps, total, and (block scan over the 256 digit totals) the tile-local digit start
	uint32_t dtotal = 0;
	if (threadIdx.x < RADIX) {
		const uint32_t d = threadIdx.x;
		uint32_t run = 0;
#pragma unroll
		for (int w = 0; w < RS_WARPS; ++w) { const uint32_t t = S.whist[w][d]; S.whist[w][d] = run; run += t; }
		dtotal = run;
		S.goff[d] = tile_offs[(size_t)d * ntiles + tile];
	}
	{
		uint32_t incl = dtotal;
#pragma unroll
		for (int off = 1; off < 32; off <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += t; }
		if (lane == 31) S.wsum[wid] = incl;
		__syncthreads();
		if (threadIdx.x < RADIX) {
			uint32_t woff = 0;
			for (int w = 0; w < wid; ++w) woff += S.wsum[w];		// wid < 8 here
			S.dstart[threadIdx.x] = woff + incl - dtotal;
			if (threadIdx.x == RADIX - 1) S.nvalid = woff + incl;
		}
	}
	__syncthreads();

	// phase B1: reorder the tile by digit in shared memory
#pragma unroll
	for (int r = 0; r < RS_ROUNDS; ++r) {
		if (k[r] != KEY_SENTINEL) {
			const uint32_t d = key_digit(k[r], D);
			S.keys[S.dstart[d] + S.whist[wid][d] + rank[r]] = k[r];
		}
	}
	__syncthreads();

	// phase B2: consecutive threads write consecutive keys; a digit's run goes to goff[d] onwards
	const uint32_t nvalid = S.nvalid;
	for (uint32_t i = threadIdx.x; i < nvalid; i += RS_THREADS) {
		const unsigned long long key = S.keys[i];
		const uint32_t d = key_digit(key, D);
		out[S.goff[d] + (i - S.dstart[d])] = key;
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
		sl[1] = (uint32_t)(a.x >> VALUE_BITS); sl[2] = (uint32_t)(a.y >> VALUE_BITS);
		sl[3] = (uint32_t)(c.x >> VALUE_BITS); sl[4] = (uint32_t)(c.y >> VALUE_BITS);
	}
	else {
#pragma unroll
		for (int t = 0; t < TSEG_V; ++t) if (i0 + t < n) sl[1 + t] = (uint32_t)(keys[i0 + t] >> VALUE_BITS);
	}
	// neighbours: from the adjacent lanes, the warp's edge lanes load them
	const uint32_t up = __shfl_up_sync(0xffffffffu, sl[TSEG_V], 1), down = __shfl_down_sync(0xffffffffu, sl[1], 1);
	sl[0] = lane ? up : ((i0 && i0 - 1 < n) ? (uint32_t)(keys[i0 - 1] >> VALUE_BITS) : 0xFFFFFFFFu);
	sl[TSEG_V + 1] = lane < 31 ? down : (i0 + TSEG_V < n ? (uint32_t)(keys[i0 + TSEG_V] >> VALUE_BITS) : 0xFFFFFFFFu);

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

// (2) sums: one thread per sorted sample. cluster id = position of the sample's rank in the service's bounds; runs of equal
// (service, cluster) are contiguous in the sorted order, so a warp reduces them with match.any groups and issues one
// 64-bit RED per group. Skew-immune: a hot service's samples are spread over as many warps as it has samples / 32.
__device__ __forceinline__ uint32_t td_cluster_of(const uint32_t *__restrict__ bounds, uint32_t nn, uint32_t r, uint32_t &lo_out, uint32_t &hi_out)
{
	uint32_t lo = 0, hi = nn - 1;			// largest j with bounds[j] <= r
	while (lo < hi) { const uint32_t mid = (lo + hi + 1) >> 1; if (bounds[mid] <= r) lo = mid; else hi = mid - 1; }
	lo_out = bounds[lo]; hi_out = bounds[lo + 1];
	return lo;
}

static constexpr int TDS_V = 4;			// consecutive sorted samples per lane

__global__ void __launch_bounds__(256) td_sums_kernel(const unsigned long long *__restrict__ keys, const unsigned long long *__restrict__ d_n,
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
			slot0 = (uint32_t)(kk[0] >> VALUE_BITS);
			const uint32_t nn = plan_n[slot0], r = (uint32_t)(i0 - seg_start[slot0]);
			if (nn & 0x80000000u) { j0 = r; lo0 = r; hi0 = r + 1; }
			else j0 = td_cluster_of(plan_bounds + (size_t)slot0 * PLAN_STRIDE, nn, r, lo0, hi0);
		}
		slot0 = __shfl_sync(0xffffffffu, slot0, 0); j0 = __shfl_sync(0xffffffffu, j0, 0);
		lo0 = __shfl_sync(0xffffffffu, lo0, 0); hi0 = __shfl_sync(0xffffffffu, hi0, 0);

		// own samples -> runs of equal (service, cluster); sorted input keeps them contiguous
		uint32_t pg[TDS_V];
		unsigned long long ps[TDS_V];
		int np = 0;
		uint32_t cslot = 0xFFFFFFFFu, cstart = 0, cnn = 0, clo = 1, chi = 0, cj = 0;	// cached lookup of the previous sample
#pragma unroll
		for (int t = 0; t < TDS_V; ++t) {
			if (kk[t] == KEY_SENTINEL) continue;
			const uint32_t slot = (uint32_t)(kk[t] >> VALUE_BITS), v = (uint32_t)(kk[t] & VALUE_MASK);
			if (slot != cslot) { cslot = slot; cstart = seg_start[slot]; cnn = plan_n[slot]; clo = 1; chi = 0; }
			const uint32_t r = (uint32_t)(i0 + t - cstart);
			uint32_t j;
			if (cnn & 0x80000000u) j = r;
			else if (slot == slot0 && r >= lo0 && r < hi0) j = j0;
			else if (r >= clo && r < chi) j = cj;
			else { cj = td_cluster_of(plan_bounds + (size_t)slot * PLAN_STRIDE, cnn, r, clo, chi); j = cj; }
			const uint32_t gid = slot * (uint32_t)TD_CAP + j;
			if (np && pg[np - 1] == gid) ps[np - 1] += v;
			else { pg[np] = gid; ps[np] = v; np++; }
		}

		// emit the runs: round t handles every lane's t-th run; usually one round, one group for the whole warp
		const int maxp = (int)__reduce_max_sync(0xffffffffu, (uint32_t)np);
		for (int t = 0; t < maxp; ++t) {
			const bool act = t < np;
			const uint32_t gid = act ? pg[t] : (0xFFFFFF00u + lane);
			const unsigned long long v = act ? ps[t] : 0ull;		// < 2^33
			const uint32_t m = __match_any_sync(0xffffffffu, gid);
			unsigned long long gsum = v;
			if (m == 0xffffffffu) {
				gsum = (unsigned long long)__reduce_add_sync(0xffffffffu, (uint32_t)v & 0xFFFFu) +
						((unsigned long long)__reduce_add_sync(0xffffffffu, (uint32_t)(v >> 16)) << 16);
			}
			else {
				const uint32_t maxcnt = __reduce_max_sync(0xffffffffu, (uint32_t)__popc(m));
				uint32_t rest = m & ~(1u << lane);
				for (uint32_t u = 1; u < maxcnt; ++u) {
					const int src = rest ? (__ffs(rest) - 1) : lane;
					const unsigned long long other = __shfl_sync(0xffffffffu, v, src);
					if (rest) { gsum += other; rest &= rest - 1; }
				}
			}
			if (act && (m & ((1u << lane) - 1u)) == 0) red_add_u64(newsum + gid, gsum);	// gid == slot * TD_CAP + j
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
		const double bmin = (double)(keys[s0] & VALUE_MASK), bmax = (double)(keys[s0 + n - 1] & VALUE_MASK);
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
__global__ void flush_kernel(DevState st, uint32_t nslots, HistCell *__restrict__ ring0, HistCell *__restrict__ ring1)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;

	if (i >= (uint64_t)nslots * HIST_CELLS) return;
	const int cell = (int)(i & (HIST_CELLS - 1));
	const HistCell c = st.hist_cur[i];

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
		const uint32_t slot = (uint32_t)(i >> 4);
		const unsigned long long cc = st.conn_cur[slot];
		st.conn_last[slot] = cc;
		st.conn_all_cnt[slot] += (uint32_t)cc;
		st.conn_all_kb[slot] += cc >> 32;
		st.conn_cur[slot] = 0;
	}
	else {
		st.hist_all[i].count += c.count;
		st.hist_all[i].sum += c.sum;
		st.hist_cur[i].count = 0; st.hist_cur[i].sum = 0;
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

// ---------------------------------------------------------------------------------------------------
// warp-autonomous variant of the tile pipeline: every warp owns a 128-event tile, its records and its three queues, so
// the two phases need only __syncwarp — no block barrier, no global key cursor. A warp writes its RESP sort keys to the 128
// key slots that belong to its events (keys first, sentinels behind); the first radix pass compacts the sentinels away.
// The hot-cell table stays shared by the CTA (shared-memory atomics).
// ---------------------------------------------------------------------------------------------------
static constexpr int WI_WARPS = 8;
static constexpr int WI_EPT = 4;
static constexpr int WI_TILE = 32 * WI_EPT;		// events per warp tile

struct WarpIngestShared
{
	using HotTable = HotTableT<10>;
	HotTable	hot;
	IngestRec	rec[WI_WARPS][WI_TILE];
	uint8_t		q_resp[WI_WARPS][WI_TILE], q_tcp[WI_WARPS][WI_TILE], q_task[WI_WARPS][WI_TILE];
};

template <int MIN_CTAS>
__global__ void __launch_bounds__(WI_WARPS * 32, MIN_CTAS) ingest_warp_kernel(DevState st, const gysk_event *__restrict__ ev, uint64_t n,
		unsigned long long *__restrict__ keys)
{
	__shared__ WarpIngestShared S;
	unsigned long long c_in = 0, c_drop = 0, c_resp = 0, c_tcp = 0, c_task = 0, c_foreign = 0;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t lt_mask = (1u << lane) - 1u;
	const uint64_t ntiles = (n + WI_TILE - 1) / WI_TILE;
	uint32_t max_ms = 0;

	for (int i = threadIdx.x; i < WarpIngestShared::HotTable::N; i += WI_WARPS * 32) {
		S.hot.tag[i] = 0; S.hot.count[i] = 0; S.hot.sum[i] = 0; S.hot.vmax[i] = INT_MIN; S.hot.bits[i] = 0;
	}
	__syncthreads();

	IngestRec *rec = S.rec[wid];
	uint8_t *q_resp = S.q_resp[wid], *q_tcp = S.q_tcp[wid], *q_task = S.q_task[wid];

	for (uint64_t tile = (uint64_t)blockIdx.x * WI_WARPS + wid; tile < ntiles; tile += (uint64_t)gridDim.x * WI_WARPS) {
		const uint64_t tbase = tile * WI_TILE;
		uint32_t n_resp = 0, n_tcp = 0, n_task = 0;		// warp-uniform queue lengths

		// ---------------- phase 1: decode + lookup + enqueue (all lanes converged) ----------------
		uint4 ra[WI_EPT], rb[WI_EPT];
#pragma unroll
		for (int k = 0; k < WI_EPT; ++k) {
			const uint64_t i = tbase + (uint64_t)k * 32 + lane;
			if (i < n) {
				ra[k] = __ldcs(reinterpret_cast<const uint4 *>(ev + i));		// streamed once: evict-first
				rb[k] = __ldcs(reinterpret_cast<const uint4 *>(ev + i) + 1);
			}
			else { ra[k] = make_uint4(0, 0, 0, 0); rb[k] = make_uint4(0, 0, 0, 0xFFFFu); }
		}
#pragma unroll
		for (int k = 0; k < WI_EPT; ++k) {
			const unsigned long long svc_id = ((unsigned long long)ra[k].y << 32) | ra[k].x;
			const unsigned long long flow_key = ((unsigned long long)ra[k].w << 32) | ra[k].z;
			const uint32_t value = rb[k].x, host_idx = rb[k].y;
			const uint32_t type = rb[k].w & 0xFFFFu;
			const bool pad = tbase + (uint64_t)k * 32 + lane >= n;
			const bool is_resp = type == GYSK_EV_RESP, is_task = type == GYSK_EV_TASK;
			const bool is_tcp = type >= GYSK_EV_CONNECT && type <= GYSK_EV_CLOSE_SER;
			// usec -> msec as SVC_INFO_CAP::upd_stats_on_req (gy_proto_parser.cc:2678); validity rule of
			// handle_ipv4_resp_event (gy_socket_stat.cc:1519-1524): drop beyond 1 000 000 msec
			const uint32_t ms = value / 1000u;
			int slot = -1;
			bool mine = !pad;

			if (mine && st.world > 1 && (host_idx % st.world) != st.rank) { c_foreign++; mine = false; }
			if (mine) {
				c_in++;
				if (svc_id != 0 && (is_tcp || is_task || (is_resp && ms <= 1000000u)))
					slot = table_lookup(is_task ? st.task_tbl : st.svc_tbl, svc_id, st.auto_register, host_idx);
				if (slot < 0) c_drop++;
				else if (is_resp) c_resp++;
				else if (is_tcp) c_tcp++;
				else c_task++;
			}
			const uint8_t pos = (uint8_t)(k * 32 + lane);
			if (slot >= 0) { IngestRec r; r.slot = (uint32_t)slot; r.value = value; r.flow_key = flow_key; rec[pos] = r; }
			const uint32_t m_resp = __ballot_sync(0xffffffffu, slot >= 0 && is_resp);
			const uint32_t m_tcp = __ballot_sync(0xffffffffu, slot >= 0 && is_tcp);
			const uint32_t m_task = __ballot_sync(0xffffffffu, slot >= 0 && is_task);
			if (slot >= 0) {
				if (is_resp) q_resp[n_resp + __popc(m_resp & lt_mask)] = pos;
				else if (is_tcp) q_tcp[n_tcp + __popc(m_tcp & lt_mask)] = pos;
				else q_task[n_task + __popc(m_task & lt_mask)] = pos;
			}
			n_resp += __popc(m_resp); n_tcp += __popc(m_tcp); n_task += __popc(m_task);
		}
		__syncwarp();

		// ---------------- phase 2a: TCP — count-min rows, one (event, row) pair per lane ----------------
		{
			const uint32_t npairs = n_tcp * st.cms_depth;
			for (uint32_t p = lane; p < npairs; p += 32) {
				const uint32_t e = p / st.cms_depth, row = p - e * st.cms_depth;
				const IngestRec r = rec[q_tcp[e]];
				red_add_u64(st.cms_cur + ((size_t)row << st.cms_log2w) + cms_index(r.flow_key, row, st.cms_wmask), cms_increment(r.value));
			}
			for (uint32_t base = 0; base < n_tcp; base += 32) {
				const uint32_t q = base + lane;
				const bool act = q < n_tcp;
				uint32_t cell = 0; int kb = 0;
				if (act) {
					const IngestRec r = rec[q_tcp[q]];
					uint32_t idx, rank;
					hll_idx_rank(r.flow_key, st.hll_p, idx, rank);
					hll_update(st.hll + ((size_t)r.slot << st.hll_p), idx, rank);
					cell = r.slot * HIST_CELLS + HIST_MAX_CELL;
					kb = (int)(r.value >> 10);
				}
				cell_add(st, S.hot, act, cell, kb);
			}
		}
		// ---------------- phase 2b: TASK — one (event, histogram) pair per lane ----------------
		{
			const uint32_t ntrip = n_task * 3u;
			for (uint32_t base = 0; base < ntrip; base += 32) {
				const uint32_t p = base + lane;
				const bool act = p < ntrip;
				uint32_t cell = 0; int d = 0;
				if (act) {
					const uint32_t e = p / 3u, h = p - e * 3u;
					const IngestRec r = rec[q_task[e]];
					d = h == 0 ? (int)r.value : (h == 1 ? (int)(uint32_t)r.flow_key : (int)(uint32_t)(r.flow_key >> 32));
					const uint32_t b = h == 0 ? (uint32_t)bucket_hash_1_3000(d) : (uint32_t)bucket_duration(d);
					cell = CELL_TASK | (r.slot * 3u * HIST_CELLS + h * HIST_CELLS + b);
				}
				cell_add(st, S.hot, act, cell, d);
			}
		}
		// ---------------- phase 2c: RESP — histogram cell + CONN_BITMAP bit + sort key ----------------
		for (uint32_t base = 0; base < WI_TILE; base += 32) {
			const uint32_t q = base + lane;
			const bool act = q < n_resp;
			uint32_t cell = 0, bit = 0; int ms = 0;
			unsigned long long key = KEY_SENTINEL;
			if (act) {
				const IngestRec r = rec[q_resp[q]];
				bit = 1u << ((uint32_t)r.flow_key & 0x1Fu);
				ms = (int)(r.value / 1000u);
				cell = r.slot * HIST_CELLS + (uint32_t)bucket_resp_time((long long)ms);
				key = ((unsigned long long)r.slot << VALUE_BITS) | r.value;
				max_ms = max(max_ms, (uint32_t)ms);
			}
			if (tbase + q < n) __stcs(keys + tbase + q, key);		// the tile's key slots: RESP keys, then sentinels
			if (base < n_resp) cell_add(st, S.hot, act, cell, ms, bit);
		}
		__syncwarp();		// rec / queues are rewritten by the next tile
	}

	__syncthreads();
	for (int i = threadIdx.x; i < WarpIngestShared::HotTable::N; i += WI_WARPS * 32) {
		if (S.hot.tag[i] && S.hot.count[i]) cell_add_global(st, S.hot.tag[i] - 1, S.hot.count[i], S.hot.sum[i], S.hot.vmax[i], S.hot.bits[i]);
	}

	max_ms = __reduce_max_sync(0xffffffffu, max_ms);
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) {
		c_in += __shfl_down_sync(0xffffffffu, c_in, off);
		c_drop += __shfl_down_sync(0xffffffffu, c_drop, off);
		c_resp += __shfl_down_sync(0xffffffffu, c_resp, off);
		c_tcp += __shfl_down_sync(0xffffffffu, c_tcp, off);
		c_task += __shfl_down_sync(0xffffffffu, c_task, off);
		c_foreign += __shfl_down_sync(0xffffffffu, c_foreign, off);
	}
	if (lane == 0) {
		if (c_in) atomicAdd(st.counters + CTR_IN, c_in);
		if (c_drop) atomicAdd(st.counters + CTR_DROPPED, c_drop);
		if (c_resp) { atomicAdd(st.counters + CTR_RESP, c_resp); atomicAdd(st.counters + CTR_NKEYS, c_resp); }
		if (c_tcp) atomicAdd(st.counters + CTR_TCP, c_tcp);
		if (c_task) atomicAdd(st.counters + CTR_TASK, c_task);
		if (c_foreign) atomicAdd(st.counters + CTR_FOREIGN, c_foreign);
		if (max_ms) atomicMax(st.counters + CTR_MAXVAL, (unsigned long long)max_ms);
	}
}

// 256 (default, measured best) / 128 / 2563 (TMA-staged) = CTA tile pipeline shapes; 4 / 40 = warp-autonomous tiles (keys left in
// event order with sentinels) with 4 / 5 CTAs per SM
static int ingest_variant()
{
	static const int v = []{ const char *e = getenv("GYSK_INGEST_VARIANT"); return e ? atoi(e) : 256; }();
	return v;
}

bool ingest_keys_compact() { return ingest_variant() != 40 && ingest_variant() != 4; }

template <int THREADS, int MIN_CTAS, bool STAGE, int EPT = 4>
static void launch_ingest_variant(const DevState &st, const gysk_event *d_ev, uint64_t n, unsigned long long *d_keys, int nsm, cudaStream_t s)
{
	using Shared = IngestSharedT<THREADS, STAGE, EPT>;
	static bool attr_set = false;
	if (!attr_set) { cudaFuncSetAttribute(ingest_kernel<THREADS, MIN_CTAS, STAGE, EPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Shared)); attr_set = true; }
	const uint64_t want = (n + Shared::INGEST_TILE - 1) / Shared::INGEST_TILE;
	const uint32_t grid = (uint32_t)(want < (uint64_t)nsm * MIN_CTAS ? want : (uint64_t)nsm * MIN_CTAS);
	ingest_kernel<THREADS, MIN_CTAS, STAGE, EPT><<<grid, THREADS, sizeof(Shared), s>>>(st, d_ev, n, d_keys);
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
	else if (variant == 256) launch_ingest_variant<256, 4, false>(st, d_ev, n, d_keys, nsm, s);
	else if (variant == 2562) launch_ingest_variant<256, 5, false, 2>(st, d_ev, n, d_keys, nsm, s);	// 2 events per thread: fewer live registers, 5 CTAs/SM
	else if (variant == 2568) launch_ingest_variant<256, 3, false, 8>(st, d_ev, n, d_keys, nsm, s);	// 8 events per thread: fewer barriers per event
	else {
		const uint64_t want = (n + (uint64_t)WI_TILE * WI_WARPS - 1) / ((uint64_t)WI_TILE * WI_WARPS);
		const int per_sm = variant == 4 ? 4 : 5;			// 4: 64 registers, no spills; 5: 48 registers, small spills
		const uint32_t grid = (uint32_t)(want < (uint64_t)nsm * per_sm ? want : (uint64_t)nsm * per_sm);
		if (per_sm == 4) ingest_warp_kernel<4><<<grid, WI_WARPS * 32, 0, s>>>(st, d_ev, n, d_keys);
		else ingest_warp_kernel<5><<<grid, WI_WARPS * 32, 0, s>>>(st, d_ev, n, d_keys);
	}
	return 1;
}

static int launch_exclusive_scan(uint32_t *d_data, uint32_t len, uint32_t *d_block_sums, unsigned long long *d_total, cudaStream_t s)
{
	const uint32_t nblocks = div_up(len, SCAN_CHUNK);
	scan_reduce_kernel<<<nblocks, SCAN_THREADS, 0, s>>>(d_data, len, d_block_sums);
	scan_blocksums_kernel<<<1, 1024, 0, s>>>(d_block_sums, nblocks, d_total);
	scan_apply_kernel<<<nblocks, SCAN_THREADS, 0, s>>>(d_data, len, d_block_sums);
	return 3;
}

// stable LSD radix sort of bufs[start] (n_upper >= *d_n keys) on the significant key bits [lo1, hi1) then [lo2, hi2) (lo2 >= hi1;
// pass hi2 <= lo2 for a single range): the significant bits are cut into 8-bit digits in order, a digit may straddle the gap.
// Result in bufs[*which].
int launch_radix_sort_from(const SortTemp &tmp, int start, uint64_t n_first, uint64_t n_upper, const unsigned long long *d_n, int lo1, int hi1, int lo2, int hi2,
		int *which, cudaStream_t s)
{
	int launches = 0;
	static bool attr_set = false;
	if (!attr_set) { cudaFuncSetAttribute(rs_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterShared)); attr_set = true; }
	bool first = true;
	unsigned long long *bufs[2] = { tmp.keys_a, tmp.keys_b };
	int w = start;
	int p1 = lo1, p2 = lo2;			// next unsorted bit of each range

	if (hi2 < lo2) hi2 = lo2;
	while (p1 < hi1 || p2 < hi2) {
		DigitSpec D {0, 0, 0, 0};
		int need = 8;
		if (p1 < hi1) { D.s1 = p1; D.b1 = hi1 - p1 < need ? hi1 - p1 : need; p1 += D.b1; need -= D.b1; }
		if (need && p1 >= hi1 && p2 < hi2) {
			const int take = hi2 - p2 < need ? hi2 - p2 : need;
			if (D.b1) { D.s2 = p2; D.b2 = take; } else { D.s1 = p2; D.b1 = take; }
			p2 += take;
		}
		// the first pass may run over n_first >= n_upper slots holding sentinels (skipped, so its output is compact)
		const uint64_t nn = first ? n_first : n_upper;
		const unsigned long long *dn = (first && n_first != n_upper) ? nullptr : d_n;
		const uint32_t ntiles = div_up(nn, SORT_TILE);
		rs_hist_kernel<<<ntiles, RS_THREADS, 0, s>>>(bufs[w], nn, dn, D, tmp.tile_hist, ntiles);
		launches += 1 + launch_exclusive_scan(tmp.tile_hist, RADIX * ntiles, tmp.scan_tmp, nullptr, s);
		rs_scatter_kernel<<<ntiles, RS_THREADS, sizeof(ScatterShared), s>>>(bufs[w], bufs[w ^ 1], nn, dn, D, tmp.tile_hist, ntiles);
		launches++;
		w ^= 1; first = false;
	}
	*which = w;
	return launches;
}

int launch_radix_sort(const SortTemp &tmp, uint64_t n_upper, const unsigned long long *d_n, int bit_lo, int bit_hi, int *which, cudaStream_t s)
{
	return launch_radix_sort_from(tmp, 0, n_upper, n_upper, d_n, bit_lo, bit_hi, bit_hi, bit_hi, which, s);
}

// sort the (slot, usec) keys produced by ingest, then fold every touched service's new samples into its digest
int launch_tdigest_update(const DevState &st, const SortTemp &tmp, uint64_t n_events, uint64_t n, uint32_t nslots, int value_bits, cudaStream_t s)
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
	launches += launch_radix_sort_from(tmp, 0, ingest_keys_compact() ? n : n_events, n, d_nkeys, 0, value_bits, VALUE_BITS, VALUE_BITS + (int)slot_bits, &which, s);
	src = bufs[which];

	td_segments_kernel<<<div_up(n, 256 * TSEG_V), 256, 0, s>>>(src, d_nkeys, tmp.seg_start, tmp.seg_end, tmp.touched, d_ntouched);
	int dev = 0, nsm = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
	td_plan_kernel<<<nsm * 2, 128, 0, s>>>(st.td, tmp.seg_start, tmp.seg_end, tmp.touched, d_ntouched, tmp.plan_bounds, tmp.plan_n, tmp.newsum);
	td_sums_kernel<<<nsm * 8, 256, 0, s>>>(src, d_nkeys, tmp.seg_start, tmp.plan_bounds, tmp.plan_n, tmp.newsum);
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
	launches += launch_radix_sort(tmp, nslots, d_n, 32, 64, &which, s);
	topn_pick_kernel<<<1, 64, 0, s>>>(st, which ? tmp.keys_b : tmp.keys_a, nslots, want, d_out);
	return launches;
}

int launch_flush(const DevState &st, uint32_t nslots, HistCell *ring_plane0, HistCell *ring_plane1, cudaStream_t s)
{
	if (!nslots) return 0;
	flush_kernel<<<div_up((uint64_t)nslots * HIST_CELLS, 256), 256, 0, s>>>(st, nslots, ring_plane0, ring_plane1);
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
