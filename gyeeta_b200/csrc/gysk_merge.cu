// gysk_merge.cu — the multi-GPU merge step.
//
// Ingest is sharded by host (host_idx % world), so every per-service sketch lives wholly on one GPU and the ingest path
// has no exchange. What needs one exchange per query window are the answers for LOGICAL services spanning hosts (the
// reference's cross-host key is the svc-mesh cluster id, common/gy_comm_proto.h:2479-2506) and the global flow sketch:
//
//   gysk_set_logical_map   glob_id -> logical id (same list on every rank => same dense logical index everywhere)
//   gysk_merge_prepare     fold this GPU's member services into per-logical arrays laid out in ONE arena:
//                            [u64 SUM region : global CMS cur/last | histogram last/all | conn cells]
//                            [i64 MAX region : max_val_seen_ last/all]   [u8 MAX region : HLL registers]
//                          and a fixed t-digest slab (not element-wise mergeable)
//   (caller)               all-reduce each region once, all-gather the slab        — NCCL via torch.distributed
//   gysk_merge_finish      rank-ascending merge + compress of the gathered digests
//   gysk_query_logical     same summary fields as gysk_query_svcs, for logical ids
//
// It is the additive roll-up of MS_CLUSTER_STATE::STATE_ONE::add_stats (common/gy_comm_proto.h:3199-3214) /
// SHCONN_HANDLER::aggregate_cluster_state (server/gy_shconnhdlr.cc:4583) and of GY_HISTOGRAM::update_from_serialized
// (common/gy_statistics.h:625-650): integer sums are order independent => bit-exact at any GPU count.
#include "gysk_engine.h"

#include <climits>
#include <dlfcn.h>
#include <nccl.h>		// types only: the library is dlopen()ed, libgysketch.so carries no link-time dependency on it

using namespace gysk;

namespace gysk {

struct SlabEntry { TdHead head; Centroid cent[TD_CAP]; };

// The map keeps every {glob_id, logical} pair it was given; which of them live on this GPU, and in which slot, is looked up at
// every merge: a service that registers after gysk_set_logical_map takes part from its first window on, one that was evicted (or whose
// slot now belongs to another id) drops out. An id without a slot maps to the engine's null slot (index max_svcs: always in its
// just-created state, the identity of every fold), which the fold kernels skip.
__global__ void resolve_members_kernel(DevState st, const unsigned long long *__restrict__ member_ids, uint32_t n, uint32_t *__restrict__ members,
		uint32_t null_slot)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int slot = member_ids[i] ? table_lookup(st.svc_tbl, member_ids[i], false) : -1;
	members[i] = slot >= 0 ? (uint32_t)slot : null_slot;
}

// one thread per (logical, cell)
__global__ void fold_hist_kernel(DevState st, const uint32_t *__restrict__ offs, const uint32_t *__restrict__ members, uint32_t nl, uint32_t null_slot,
		HistCell *__restrict__ l_last, HistCell *__restrict__ l_all, unsigned long long *__restrict__ l_conn, long long *__restrict__ l_hmax)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (uint64_t)nl * HIST_CELLS) return;
	const uint32_t l = (uint32_t)(i >> 4);
	const int cell = (int)(i & 15);
	const uint32_t b = offs[l], e = offs[l + 1];

	if (cell < HIST_MAX_CELL) {
		HistCell a {0, 0}, c {0, 0};
		for (uint32_t m = b; m < e; ++m) {
			if (members[m] == null_slot) continue;
			const HistCell x = st.hist_last[(size_t)members[m] * HIST_CELLS + cell], y = st.hist_all[(size_t)members[m] * HIST_CELLS + cell];
			a.count += x.count; a.sum += x.sum; c.count += y.count; c.sum += y.sum;
		}
		l_last[i] = a; l_all[i] = c;
	}
	else {
		long long ml = LLONG_MIN, ma = LLONG_MIN;
		unsigned long long lc = 0, lk = 0, ac = 0, ak = 0;
		for (uint32_t m = b; m < e; ++m) {
			const uint32_t s = members[m];
			if (s == null_slot) continue;
			ml = max(ml, st.hist_last[(size_t)s * HIST_CELLS + HIST_MAX_CELL].sum);
			ma = max(ma, st.hist_all[(size_t)s * HIST_CELLS + HIST_MAX_CELL].sum);
			const unsigned long long cl = st.conn_last[s];
			lc += (uint32_t)cl; lk += cl >> 32; ac += st.conn_all_cnt[s]; ak += st.conn_all_kb[s];
		}
		l_last[i] = HistCell {0, 0}; l_all[i] = HistCell {0, 0};
		l_hmax[2 * l] = ml; l_hmax[2 * l + 1] = ma;
		l_conn[4 * l] = lc; l_conn[4 * l + 1] = lk; l_conn[4 * l + 2] = ac; l_conn[4 * l + 3] = ak;
	}
}

// one thread per (logical, 4 registers): per-byte max over the member services
__global__ void fold_hll_kernel(DevState st, const uint32_t *__restrict__ offs, const uint32_t *__restrict__ members, uint32_t nl, uint32_t null_slot,
		uint32_t *__restrict__ l_hll)
{
	const uint32_t words = 1u << (st.hll_p - 2);
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (uint64_t)nl * words) return;
	const uint32_t l = (uint32_t)(i / words), w = (uint32_t)(i % words);
	uint32_t acc = 0;

	for (uint32_t m = offs[l]; m < offs[l + 1]; ++m) {
		if (members[m] == null_slot) continue;
		acc = __vmaxu4(acc, reinterpret_cast<const uint32_t *>(st.hll + ((size_t)members[m] << st.hll_p))[w]);
	}
	l_hll[i] = acc;
}

static constexpr int MG_WARPS = 2;		// 2 x 17.8 KB of scratch: static shared memory

// one warp per logical service: fold member digests one after the other (member order = slot order of the map call)
__global__ void __launch_bounds__(MG_WARPS * 32) fold_td_kernel(DevState st, const uint32_t *__restrict__ offs, const uint32_t *__restrict__ members,
		uint32_t nl, uint32_t null_slot, SlabEntry *__restrict__ slab)
{
	__shared__ TdScratch scratch[MG_WARPS];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	TdScratch &S = scratch[wid];

	for (uint32_t l = blockIdx.x * MG_WARPS + wid; l < nl; l += gridDim.x * MG_WARPS) {
		uint32_t nacc = 0;
		unsigned long long total = 0;
		double mn = INFINITY, mx = -INFINITY;

		for (uint32_t m = offs[l]; m < offs[l + 1]; ++m) {
			const uint32_t s = members[m];
			if (s == null_slot) continue;
			const TdHead h = st.td_head[s];
			if (!h.n) continue;
			nacc = warp_merge_compress(S, S.newc, nacc, st.td_cent + (size_t)s * TD_CAP, h.n, S.newc, st.td);
			total += h.total; mn = fmin(mn, h.minv); mx = fmax(mx, h.maxv);
		}
		for (uint32_t c = lane; c < TD_CAP; c += 32) slab[l].cent[c] = c < nacc ? S.newc[c] : Centroid {0.0, 0};
		if (lane == 0) { TdHead h; h.total = total; h.minv = mn; h.maxv = mx; h.n = nacc; h.pad = 0; slab[l].head = h; }
		__syncwarp();
	}
}

// one warp per logical service over the all-gathered slabs [world][nl]
__global__ void __launch_bounds__(MG_WARPS * 32) finish_td_kernel(const SlabEntry *__restrict__ gathered, uint32_t world, uint32_t nl,
		SlabEntry *__restrict__ out, TdParams P)
{
	__shared__ TdScratch scratch[MG_WARPS];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	TdScratch &S = scratch[wid];

	for (uint32_t l = blockIdx.x * MG_WARPS + wid; l < nl; l += gridDim.x * MG_WARPS) {
		uint32_t nacc = 0;
		unsigned long long total = 0;
		double mn = INFINITY, mx = -INFINITY;

		for (uint32_t r = 0; r < world; ++r) {			// fixed rank-ascending order => deterministic result
			const SlabEntry &e = gathered[(size_t)r * nl + l];
			if (!e.head.n) continue;
			nacc = warp_merge_compress(S, S.newc, nacc, e.cent, e.head.n, S.newc, P);
			total += e.head.total; mn = fmin(mn, e.head.minv); mx = fmax(mx, e.head.maxv);
		}
		for (uint32_t c = lane; c < TD_CAP; c += 32) out[l].cent[c] = c < nacc ? S.newc[c] : Centroid {0.0, 0};
		if (lane == 0) { TdHead h; h.total = total; h.minv = mn; h.maxv = mx; h.n = nacc; h.pad = 0; out[l].head = h; }
		__syncwarp();
	}
}

// read side: logical arrays -> SvcRaw (so the host summary code is shared with gysk_query_svcs)
__global__ void __launch_bounds__(128) gather_logical_kernel(const int32_t *__restrict__ lidx, uint32_t n, uint32_t hll_p,
		const HistCell *__restrict__ l_last, const HistCell *__restrict__ l_all, const unsigned long long *__restrict__ l_conn,
		const long long *__restrict__ l_hmax, const uint8_t *__restrict__ l_hll, const SlabEntry *__restrict__ slab, SvcRaw *__restrict__ out)
{
	__shared__ uint32_t hh[4][64];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t q = blockIdx.x * 4 + wid;

	if (q >= n) return;
	SvcRaw &o = out[q];
	const int32_t l = lidx[q];
	if (lane == 0) { o.id = 0; o.found = l >= 0; o.slot = (uint32_t)l; }
	if (l < 0) return;

	if (lane < HIST_CELLS) {
		HistCell a = l_last[(size_t)l * HIST_CELLS + lane], b = l_all[(size_t)l * HIST_CELLS + lane];
		if (lane == HIST_MAX_CELL) { a.sum = l_hmax[2 * l]; b.sum = l_hmax[2 * l + 1]; }
		o.last[lane] = a; o.all[lane] = b; o.cur[lane] = HistCell {0, 0};
		o.lvl[0][lane] = HistCell {0, 0}; o.lvl[1][lane] = HistCell {0, 0};
		o.bm_cur[lane] = 0; o.bm_last[lane] = 0;
	}
	if (lane == 0) {
		o.conn_cur = 0;
		o.conn_last = (l_conn[4 * l] & 0xFFFFFFFFull) | (l_conn[4 * l + 1] << 32);
		o.conn_all_cnt = l_conn[4 * l + 2]; o.conn_all_kb = l_conn[4 * l + 3];
		o.td = slab[l].head;
	}
	for (int i = lane; i < TD_CAP; i += 32) o.cent[i] = slab[l].cent[i];

	hh[wid][lane] = 0; hh[wid][lane + 32] = 0;
	__syncwarp();
	const uint8_t *regs = l_hll + ((size_t)l << hll_p);
	for (uint32_t i = lane; i < (1u << hll_p); i += 32) atomicAdd(&hh[wid][regs[i] > 63 ? 63 : regs[i]], 1u);
	__syncwarp();
	o.hll_hist[lane] = hh[wid][lane]; o.hll_hist[lane + 32] = hh[wid][lane + 32];
}

} // namespace gysk

namespace {

inline uint32_t div_up(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }
inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

} // namespace

// ---- NCCL inside the library: the whole merge step of one query window as ONE call -------------------
//
// A C++ madhava has no torch.distributed: it calls gysk_merge_global(engine, comm) per engine (one thread per GPU, or inside its
// own ncclGroupStart/End when one thread drives all eight). libnccl.so.2 is resolved at first use with dlopen — a process that
// already carries NCCL (torch) shares that copy — so single-GPU deployments and the CPU-side ABI tests need no NCCL at all.
namespace {

struct NcclApi
{
	void *h {nullptr};
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) {nullptr};
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) {nullptr};
	ncclResult_t (*CommDestroy)(ncclComm_t) {nullptr};
	ncclResult_t (*CommCount)(const ncclComm_t, int *) {nullptr};
	ncclResult_t (*CommUserRank)(const ncclComm_t, int *) {nullptr};
	ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) {nullptr};
	ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) {nullptr};
	ncclResult_t (*GroupStart)() {nullptr};
	ncclResult_t (*GroupEnd)() {nullptr};
	const char *(*GetErrorString)(ncclResult_t) {nullptr};
	std::string err;
};

NcclApi *nccl_api()
{
	static NcclApi api;
	static std::once_flag once;
	std::call_once(once, [] {
		for (const char *name : {"libnccl.so.2", "libnccl.so"}) { api.h = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (api.h) break; }
		if (!api.h) { api.err = std::string("dlopen(libnccl.so.2): ") + (dlerror() ? dlerror() : "not found"); return; }
		auto sym = [&](const char *n) { void *p = dlsym(api.h, n); if (!p && api.err.empty()) api.err = std::string("libnccl lacks ") + n; return p; };
		api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
		api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
		api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
		api.CommCount = (decltype(api.CommCount))sym("ncclCommCount");
		api.CommUserRank = (decltype(api.CommUserRank))sym("ncclCommUserRank");
		api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
		api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
		api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
		api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
		api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
	});
	return api.err.empty() ? &api : nullptr;
}

int nccl_fail(gysk_engine *e, const char *what, ncclResult_t r)
{
	NcclApi *a = nccl_api();
	char buf[256];
	snprintf(buf, sizeof(buf), "%s: %s", what, a && a->GetErrorString ? a->GetErrorString(r) : "nccl error");
	return fail(e, GYSK_ERR_CUDA, buf);		// sticky, like a CUDA error
}

} // namespace

namespace gysk {
void merge_release(gysk_engine *e)
{
	if (e->mg.comm && e->mg.comm_owned) { NcclApi *a = nccl_api(); if (a) a->CommDestroy((ncclComm_t)e->mg.comm); }
	e->mg.comm = nullptr;
}
} // namespace gysk


extern "C" {

int gysk_set_logical_map(gysk_engine *e, const uint64_t *glob_ids, const uint64_t *logical_ids, uint32_t n)
{
	CHECK_ENGINE(e);
	if ((!glob_ids || !logical_ids) && n) return GYSK_ERR_INVAL;
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	int rc = sync_locked(e);
	if (rc) return rc;

	MergeState &mg = e->mg;

	// dense logical index in order of first appearance: identical on every rank when the same list is passed
	mg.index.clear(); mg.logical_ids.clear();
	std::vector<uint32_t> lidx(n);
	for (uint32_t i = 0; i < n; ++i) {
		auto it = mg.index.find(logical_ids[i]);
		if (it == mg.index.end()) {
			it = mg.index.emplace(logical_ids[i], (uint32_t)mg.logical_ids.size()).first;
			mg.logical_ids.push_back(logical_ids[i]);
		}
		lidx[i] = it->second;
	}
	const uint32_t nl = (uint32_t)mg.logical_ids.size();

	// CSR logical -> every {glob_id} mapped to it, in map order; the slots are looked up at merge time (resolve_members_kernel)
	std::vector<uint32_t> offs(nl + 1, 0), members(n, e->cfg.max_svcs);
	std::vector<uint64_t> member_ids(n);
	for (uint32_t i = 0; i < n; ++i) offs[lidx[i] + 1]++;
	for (uint32_t l = 0; l < nl; ++l) offs[l + 1] += offs[l];
	{
		std::vector<uint32_t> cur(offs.begin(), offs.end() - 1);
		for (uint32_t i = 0; i < n; ++i) member_ids[cur[lidx[i]]++] = glob_ids[i];
	}

	// (re)allocate the arena
	auto dfree = [&](void *p) { if (p) { cudaFree(p); e->dallocs.erase(std::remove(e->dallocs.begin(), e->dallocs.end(), p), e->dallocs.end()); } };
	dfree(mg.d_offsets); dfree(mg.d_members); dfree(mg.d_member_ids); dfree(mg.arena); dfree(mg.slab); dfree(mg.final_slab);
	{
		std::vector<uint64_t> ids_keep(std::move(mg.logical_ids));
		std::unordered_map<uint64_t, uint32_t> idx_keep(std::move(mg.index));
		mg = MergeState {};
		mg.logical_ids = std::move(ids_keep); mg.index = std::move(idx_keep);
	}
	mg.nlogical = nl;

	const size_t ncms = (size_t)e->cfg.cms_depth << e->cfg.cms_log2_width;
	const size_t b_cms = ncms * 8, b_hist = (size_t)nl * HIST_CELLS * sizeof(HistCell), b_conn = (size_t)nl * 4 * 8;
	size_t off = 0;
	mg.off_sum = off;
	const size_t o_cms_cur = off; off += align256(b_cms);
	const size_t o_cms_last = off; off += align256(b_cms);
	const size_t o_hl = off; off += align256(b_hist);
	const size_t o_ha = off; off += align256(b_hist);
	const size_t o_conn = off; off += align256(b_conn);
	mg.bytes_sum = off - mg.off_sum;
	mg.off_maxi64 = off; const size_t o_hmax = off; off += align256((size_t)nl * 2 * 8); mg.bytes_maxi64 = off - mg.off_maxi64;
	mg.off_maxu8 = off; const size_t o_hll = off; off += align256((size_t)nl << e->cfg.hll_p); mg.bytes_maxu8 = off - mg.off_maxu8;
	mg.arena_bytes = off;

	if ((rc = dalloc(e, &mg.arena, mg.arena_bytes))) return rc;
	mg.g_cms_cur = reinterpret_cast<unsigned long long *>(mg.arena + o_cms_cur);
	mg.g_cms_last = reinterpret_cast<unsigned long long *>(mg.arena + o_cms_last);
	mg.l_hist_last = reinterpret_cast<HistCell *>(mg.arena + o_hl);
	mg.l_hist_all = reinterpret_cast<HistCell *>(mg.arena + o_ha);
	mg.l_conn = reinterpret_cast<unsigned long long *>(mg.arena + o_conn);
	mg.l_hmax = reinterpret_cast<long long *>(mg.arena + o_hmax);
	mg.l_hll = mg.arena + o_hll;
	mg.slab_bytes = (size_t)(nl ? nl : 1) * sizeof(SlabEntry);
	if ((rc = dalloc(e, &mg.slab, mg.slab_bytes))) return rc;
	if ((rc = dalloc(e, &mg.final_slab, mg.slab_bytes))) return rc;
	if ((rc = dalloc(e, &mg.d_offsets, (size_t)nl + 1))) return rc;
	if ((rc = dalloc(e, &mg.d_members, members.size() + 1))) return rc;
	if ((rc = dalloc(e, &mg.d_member_ids, member_ids.size() + 1))) return rc;
	mg.nmembers = (uint32_t)members.size();
	if (!member_ids.empty()) CU(e, cudaMemcpyAsync(mg.d_member_ids, member_ids.data(), member_ids.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, e->stream));
	CU(e, cudaMemcpyAsync(mg.d_offsets, offs.data(), ((size_t)nl + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
	if (!members.empty()) CU(e, cudaMemcpyAsync(mg.d_members, members.data(), members.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
	CU(e, cudaStreamSynchronize(e->stream));
	return post_launch(e, "set_logical_map");
}

int gysk_merge_prepare(gysk_engine *e)
{
	CHECK_ENGINE(e);
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	MergeState &mg = e->mg;
	if (!mg.arena) return fail(e, GYSK_ERR_INVAL, "gysk_merge_prepare: call gysk_set_logical_map first");
	int rc = submit_stage(e);
	if (rc) return rc;

	const size_t b_cms = ((size_t)e->cfg.cms_depth << e->cfg.cms_log2_width) * 8;
	const uint32_t nl = mg.nlogical;

	CU(e, cudaMemcpyAsync(mg.g_cms_cur, e->st.cms_cur, b_cms, cudaMemcpyDeviceToDevice, e->stream));
	CU(e, cudaMemcpyAsync(mg.g_cms_last, e->st.cms_last, b_cms, cudaMemcpyDeviceToDevice, e->stream));
	if (nl) {
		const uint32_t null_slot = e->cfg.max_svcs;
		if (mg.nmembers) {
			resolve_members_kernel<<<div_up(mg.nmembers, 256), 256, 0, e->stream>>>(e->st, mg.d_member_ids, mg.nmembers, mg.d_members, null_slot);
			e->kernel_launches++;
		}
		fold_hist_kernel<<<div_up((uint64_t)nl * HIST_CELLS, 256), 256, 0, e->stream>>>(e->st, mg.d_offsets, mg.d_members, nl, null_slot,
				mg.l_hist_last, mg.l_hist_all, mg.l_conn, mg.l_hmax);
		fold_hll_kernel<<<div_up((uint64_t)nl << (e->cfg.hll_p - 2), 256), 256, 0, e->stream>>>(e->st, mg.d_offsets, mg.d_members, nl, null_slot,
				reinterpret_cast<uint32_t *>(mg.l_hll));
		fold_td_kernel<<<std::min<uint32_t>(div_up(nl, MG_WARPS), 148 * 8), MG_WARPS * 32, 0, e->stream>>>(e->st, mg.d_offsets, mg.d_members, nl, null_slot,
				reinterpret_cast<SlabEntry *>(mg.slab));
		e->kernel_launches += 3;
	}
	// no host sync: the caller enqueues the collectives on gysk_stream(e) (stream order) or calls gysk_sync() first
	mg.prepared = true; mg.finished = false;
	return post_launch(e, "merge_prepare");
}

int gysk_merge_buffers(gysk_engine *e, gysk_buffer_desc *out, uint32_t cap, uint32_t *n)
{
	CHECK_ENGINE(e);
	if (!out || !n) return GYSK_ERR_INVAL;
	GYSK_ENTER(e);
	MergeState &mg = e->mg;
	if (!mg.arena) return fail(e, GYSK_ERR_INVAL, "gysk_merge_buffers: call gysk_set_logical_map first");
	if (cap < 3) return GYSK_ERR_NOSPC;
	out[0] = gysk_buffer_desc {"sum_u64: cms_cur|cms_last|hist_last|hist_all|conn", mg.arena + mg.off_sum, mg.bytes_sum, GYSK_RED_SUM_U64, 0};
	out[1] = gysk_buffer_desc {"max_i64: hist max_val_seen", mg.arena + mg.off_maxi64, mg.bytes_maxi64, GYSK_RED_MAX_I64, 0};
	out[2] = gysk_buffer_desc {"max_u8: hll registers", mg.arena + mg.off_maxu8, mg.bytes_maxu8, GYSK_RED_MAX_U8, 0};
	*n = 3;
	return GYSK_OK;
}

int gysk_merge_tdigest_slab(gysk_engine *e, void **dptr, uint64_t *nbytes)
{
	CHECK_ENGINE(e);
	if (!dptr || !nbytes) return GYSK_ERR_INVAL;
	GYSK_ENTER(e);
	if (!e->mg.slab) return fail(e, GYSK_ERR_INVAL, "gysk_merge_tdigest_slab: call gysk_set_logical_map first");
	*dptr = e->mg.slab; *nbytes = (uint64_t)e->mg.nlogical * sizeof(SlabEntry);
	return GYSK_OK;
}

int gysk_merge_finish(gysk_engine *e, const void *d_gathered, uint32_t world)
{
	CHECK_ENGINE(e);
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	MergeState &mg = e->mg;
	if (!mg.prepared) return fail(e, GYSK_ERR_INVAL, "gysk_merge_finish: call gysk_merge_prepare first");
	if (!world) world = 1;
	const SlabEntry *src = d_gathered ? static_cast<const SlabEntry *>(d_gathered) : reinterpret_cast<const SlabEntry *>(mg.slab);
	if (!d_gathered) world = 1;
	if (mg.nlogical) {
		finish_td_kernel<<<std::min<uint32_t>(div_up(mg.nlogical, MG_WARPS), 148 * 8), MG_WARPS * 32, 0, e->stream>>>(src, world, mg.nlogical,
				reinterpret_cast<SlabEntry *>(mg.final_slab), e->st.td);
		e->kernel_launches++;
	}
	mg.finished = true;			// stream-ordered; the query calls synchronise
	return post_launch(e, "merge_finish");
}

int gysk_query_logical(gysk_engine *e, const uint64_t *logical_ids, uint32_t n, gysk_svc_summary *out)
{
	CHECK_ENGINE(e);
	if ((!logical_ids || !out) && n) return GYSK_ERR_INVAL;
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	MergeState &mg = e->mg;
	if (!mg.finished) return fail(e, GYSK_ERR_INVAL, "gysk_query_logical: no finished merge");

	int32_t *h_l = reinterpret_cast<int32_t *>(e->h_qids), *d_l = reinterpret_cast<int32_t *>(e->d_qids);
	for (uint32_t off = 0; off < n; off += QCHUNK) {
		const uint32_t m = std::min(QCHUNK, n - off);
		for (uint32_t i = 0; i < m; ++i) {
			auto it = mg.index.find(logical_ids[off + i]);
			h_l[i] = it == mg.index.end() ? -1 : (int32_t)it->second;
		}
		CU(e, cudaMemcpyAsync(d_l, h_l, (size_t)m * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
		gather_logical_kernel<<<div_up(m, 4), 128, 0, e->stream>>>(d_l, m, e->cfg.hll_p, mg.l_hist_last, mg.l_hist_all, mg.l_conn, mg.l_hmax,
				mg.l_hll, reinterpret_cast<const SlabEntry *>(mg.final_slab), e->d_svcraw);
		e->kernel_launches++;
		CU(e, cudaMemcpyAsync(e->h_svcraw, e->d_svcraw, (size_t)m * sizeof(SvcRaw), cudaMemcpyDeviceToHost, e->stream));
		CU(e, cudaStreamSynchronize(e->stream));
		for (uint32_t i = 0; i < m; ++i) summarize_raw(e, e->h_svcraw[i], logical_ids[off + i], out[off + i]);
	}
	return post_launch(e, "query_logical");
}

// global count-min point query on the merged table
int gysk_query_flows_global(gysk_engine *e, const uint64_t *keys, uint32_t n, int last_window, gysk_flow_est *out)
{
	CHECK_ENGINE(e);
	if ((!keys || !out) && n) return GYSK_ERR_INVAL;
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	MergeState &mg = e->mg;
	if (!mg.prepared) return fail(e, GYSK_ERR_INVAL, "gysk_query_flows_global: no merge");
	DevState st = e->st;
	st.cms_cur = mg.g_cms_cur; st.cms_last = mg.g_cms_last;
	for (uint32_t off = 0; off < n; off += QCHUNK) {
		const uint32_t m = std::min(QCHUNK, n - off);
		memcpy(e->h_qids, keys + off, (size_t)m * sizeof(uint64_t));
		CU(e, cudaMemcpyAsync(e->d_qids, e->h_qids, (size_t)m * sizeof(uint64_t), cudaMemcpyHostToDevice, e->stream));
		e->kernel_launches += launch_query_flows(st, e->d_qids, m, last_window, e->d_flowout, e->stream);
		CU(e, cudaMemcpyAsync(e->h_flowout, e->d_flowout, (size_t)m * sizeof(gysk_flow_est), cudaMemcpyDeviceToHost, e->stream));
		CU(e, cudaStreamSynchronize(e->stream));
		memcpy(out + off, e->h_flowout, (size_t)m * sizeof(gysk_flow_est));
	}
	return post_launch(e, "query_flows_global");
}

#define NC(e, call) do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) return nccl_fail((e), #call, r__); } while (0)

int gysk_nccl_unique_id(uint8_t out[GYSK_NCCL_UNIQUE_ID_BYTES])
{
	NcclApi *a = nccl_api();
	if (!out) return GYSK_ERR_INVAL;
	if (!a) return GYSK_ERR_NOTSUP;
	static_assert(sizeof(ncclUniqueId) == GYSK_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
	ncclUniqueId id;
	if (a->GetUniqueId(&id) != ncclSuccess) return GYSK_ERR_CUDA;
	memcpy(out, &id, sizeof(id));
	return GYSK_OK;
}

int gysk_nccl_comm_init(gysk_engine *e, const uint8_t uid[GYSK_NCCL_UNIQUE_ID_BYTES], uint32_t nranks, uint32_t rank)
{
	CHECK_ENGINE(e);
	if (!uid || !nranks || rank >= nranks) return GYSK_ERR_INVAL;
	NcclApi *a = nccl_api();
	if (!a) return fail(e, GYSK_ERR_NOTSUP, "libnccl.so.2 could not be loaded");
	GYSK_ENTER(e);
	CU(e, cudaSetDevice(e->dev));
	if (e->mg.comm) { a->CommDestroy((ncclComm_t)e->mg.comm); e->mg.comm = nullptr; }
	ncclUniqueId id;
	memcpy(&id, uid, sizeof(id));
	ncclComm_t c = nullptr;
	NC(e, a->CommInitRank(&c, (int)nranks, id, (int)rank));
	e->mg.comm = c; e->mg.comm_world = nranks; e->mg.comm_owned = true;
	return GYSK_OK;
}

// prepare (fold) -> one grouped NCCL launch: all-reduce per reduction kind + all-gather of the t-digest slabs -> finish.
// comm == NULL uses the communicator of gysk_nccl_comm_init. Everything is enqueued on the engine's stream; nothing blocks.
int gysk_merge_global(gysk_engine *e, void *comm)
{
	CHECK_ENGINE(e);
	NcclApi *a = nccl_api();
	if (!a) return fail(e, GYSK_ERR_NOTSUP, "libnccl.so.2 could not be loaded");
	ncclComm_t c = comm ? (ncclComm_t)comm : (ncclComm_t)e->mg.comm;
	if (!c) return fail(e, GYSK_ERR_INVAL, "gysk_merge_global: no communicator (pass one or call gysk_nccl_comm_init)");
	int rc = gysk_merge_prepare(e);
	if (rc) return rc;
	int world = 0;
	{
		GYSK_ENTER(e);
		CU(e, cudaSetDevice(e->dev));
		MergeState &mg = e->mg;
		NC(e, a->CommCount(c, &world));
		if (world < 1) return fail(e, GYSK_ERR_INVAL, "gysk_merge_global: empty communicator");
		if (mg.gathered_world != (uint32_t)world) {
			if (mg.gathered) { cudaFree(mg.gathered); e->dallocs.erase(std::remove(e->dallocs.begin(), e->dallocs.end(), (void *)mg.gathered), e->dallocs.end()); mg.gathered = nullptr; }
			if ((rc = dalloc(e, &mg.gathered, mg.slab_bytes * (size_t)world, false))) return rc;
			mg.gathered_world = (uint32_t)world;
		}
		const size_t slab = (size_t)mg.nlogical * sizeof(SlabEntry);
		NC(e, a->GroupStart());
		NC(e, a->AllReduce(mg.arena + mg.off_sum, mg.arena + mg.off_sum, mg.bytes_sum / 8, ncclUint64, ncclSum, c, e->stream));
		NC(e, a->AllReduce(mg.arena + mg.off_maxi64, mg.arena + mg.off_maxi64, mg.bytes_maxi64 / 8, ncclInt64, ncclMax, c, e->stream));
		NC(e, a->AllReduce(mg.arena + mg.off_maxu8, mg.arena + mg.off_maxu8, mg.bytes_maxu8, ncclUint8, ncclMax, c, e->stream));
		if (slab) NC(e, a->AllGather(mg.slab, mg.gathered, slab, ncclUint8, c, e->stream));
		NC(e, a->GroupEnd());
		e->merges++;
	}
	return gysk_merge_finish(e, e->mg.nlogical ? e->mg.gathered : nullptr, (uint32_t)world);
}

} // extern "C"
