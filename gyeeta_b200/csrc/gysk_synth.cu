// gysk_synth.cu — on-device synthetic event source for the sustained-stream run (BASELINE.json configs[4], SURVEY.md §8d "Config 5":
// "events generated on-device from a counter-based RNG (Philox) so host PCIe is not the limiter"). Builds libgysynth.so, a
// bench / test utility: NOT part of libgysketch.so and not on the product path. Same distributions as bench.py::gen_events_gpu
// (70 / 20 / 10 RESP / TCP / TASK, Zipf services, log-normal values), plus service churn so that idle eviction has work to do.
//
// Event i of a fill is a pure function of (seed, rank, counter_base + i): Philox4x32-10, three calls per event.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

#include "../../include/gysketch.h"

extern "C" {

typedef struct gysyn_params
{
	uint64_t		seed;
	uint32_t		rank, world;
	uint32_t		nsvc, ntask;			// ids / cdf entries
	const uint64_t		*d_svc_ids, *d_task_ids;	// device arrays
	const double		*d_cdf_svc, *d_cdf_task;	// inclusive Zipf CDFs, device arrays
	uint32_t		nhosts, nclients;
	uint32_t		tsec;				// stamped into every event (informational for the engine)
	// churn: a service of Zipf rank s >= tail_start belongs to group s % churn_groups and is alive only while
	// (window / churn_epoch) % churn_groups == its group; a draw of a silent service is folded onto rank s % tail_start
	uint32_t		tail_start, churn_groups, churn_epoch, window;
	float			resp_mu, resp_sigma;		// ln usec
} gysyn_params;

int gysyn_fill(void *d_out, uint64_t n, uint64_t counter_base, const gysyn_params *p, void *stream);

}

namespace {

struct U4 { uint32_t x, y, z, w; };

__device__ __forceinline__ U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1)
{
#pragma unroll
	for (int r = 0; r < 10; ++r) {
		const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
		const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
		c = U4 { hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0 };
		k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
	}
	return c;
}

__device__ __forceinline__ uint32_t cdf_rank(const double *__restrict__ cdf, uint32_t n, double u)
{
	uint32_t lo = 0, hi = n;			// first index with cdf[i] >= u  (torch.searchsorted, left)
	while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(cdf + mid) < u) lo = mid + 1; else hi = mid; }
	return lo < n ? lo : n - 1;
}

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

__device__ __forceinline__ float lognormal(float mu, float sigma, float z, float cap)
{
	return fminf(__expf(fmaf(sigma, z, mu)), cap);
}

__global__ void __launch_bounds__(256) synth_kernel(gysk_event *__restrict__ out, uint64_t n, uint64_t base, gysyn_params P)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint64_t ctr = base + i;
	const uint32_t k0 = (uint32_t)P.seed, k1 = (uint32_t)(P.seed >> 32);
	const U4 a = philox4x32_10(U4 { (uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, P.rank }, k0, k1);
	const U4 b = philox4x32_10(U4 { (uint32_t)ctr, (uint32_t)(ctr >> 32), 1u, P.rank }, k0, k1);

	const double u_svc = (double)((((uint64_t)a.x << 32) | a.y) >> 11) * (1.0 / 9007199254740992.0);
	uint32_t srank = cdf_rank(P.d_cdf_svc, P.nsvc, u_svc);
	if (P.churn_groups && srank >= P.tail_start && (P.window / P.churn_epoch) % P.churn_groups != srank % P.churn_groups) srank %= P.tail_start;
	const float kind = (float)a.z * (1.0f / 4294967296.0f);
	const bool is_resp = kind < 0.70f, is_task = kind >= 0.90f && P.ntask;

	// Box-Muller pair
	const float u1 = ((float)(b.x >> 8) + 1.0f) * (1.0f / 16777216.0f), u2 = (float)(b.y >> 8) * (1.0f / 16777216.0f);
	const float rad = sqrtf(-2.0f * __logf(u1));
	float sn, cs;
	sincospif(2.0f * u2, &sn, &cs);
	const float z0 = rad * cs, z1 = rad * sn;

	gysk_event e;
	e.host_idx = (srank % (P.nhosts / (P.world ? P.world : 1))) * (P.world ? P.world : 1) + P.rank;
	e.tsec = P.tsec; e.flags = 0;
	if (is_task) {
		const U4 c = philox4x32_10(U4 { (uint32_t)ctr, (uint32_t)(ctr >> 32), 2u, P.rank }, k0, k1);
		const double u_t = (double)((((uint64_t)c.x << 32) | c.y) >> 11) * (1.0 / 9007199254740992.0);
		e.svc_id = P.d_task_ids[cdf_rank(P.d_cdf_task, P.ntask, u_t)];
		const uint32_t cpu_delay = (uint32_t)lognormal(3.4011974f /* ln 30 */, 2.0f, z0, 1.0e5f);
		const uint32_t blkio = (uint32_t)lognormal(1.6094379f /* ln 5 */, 2.5f, z1, 1.0e5f);
		e.flow_key = (uint64_t)cpu_delay | ((uint64_t)blkio << 32);
		e.value = (uint32_t)((float)c.z * (400.0f / 4294967296.0f));
		e.type = GYSK_EV_TASK;
	}
	else {
		e.svc_id = P.d_svc_ids[srank];
		const uint64_t cli = (uint64_t)(b.z % (P.nclients / 8u)) * 8u + (srank % 8u);
		e.flow_key = mix64(cli + (1ull << 48));
		if (is_resp) {
			e.value = (uint32_t)lognormal(P.resp_mu, P.resp_sigma, z0, 9.0e8f);
			e.type = GYSK_EV_RESP;
		}
		else {
			e.value = (uint32_t)lognormal(8.3177662f /* ln 4096 */, 2.0f, z0, 4.0e9f);
			const float tu = (float)a.w * (1.0f / 4294967296.0f);
			e.type = tu < 0.45f ? GYSK_EV_ACCEPT : (tu < 0.90f ? GYSK_EV_CLOSE_SER : GYSK_EV_CONNECT);
		}
	}
	uint4 *d = reinterpret_cast<uint4 *>(out + i);
	d[0] = make_uint4((uint32_t)e.svc_id, (uint32_t)(e.svc_id >> 32), (uint32_t)e.flow_key, (uint32_t)(e.flow_key >> 32));
	d[1] = make_uint4(e.value, e.host_idx, e.tsec, (uint32_t)e.type | ((uint32_t)e.flags << 16));
}

} // namespace

extern "C" int gysyn_fill(void *d_out, uint64_t n, uint64_t counter_base, const gysyn_params *p, void *stream)
{
	if (!d_out || !p || !p->d_svc_ids || !p->d_cdf_svc || !p->nsvc || !p->nhosts || p->nclients < 8) return -1;
	if ((p->churn_groups && (!p->churn_epoch || !p->tail_start)) || (p->ntask && (!p->d_task_ids || !p->d_cdf_task))) return -1;
	if (!n) return 0;
	const uint64_t blocks = (n + 255) / 256;
	if (blocks > 0x7FFFFFFFull) return -1;
	synth_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(static_cast<gysk_event *>(d_out), n, counter_base, *p);
	return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ---- producer threads of bench.py's e2e_wire leg -----------------------------------------------------------------------------
// N native threads (madhava's L2 handle_l2_misc threads, server/gy_mconnhdlr.cc:5128), each handing its own prebuilt rounds of
// {TCP_CONN_NOTIFY message, AGGR_TASK_STATE_NOTIFY message, raw tcp_ipv4_resp_event_t array} to the engine through the C ABI's
// function pointers (this utility does not link libgysketch.so). Returns the seconds from the common start to the last thread's
// return; the caller adds the closing query + sync.
extern "C" {
typedef int (*gysyn_ingest_msg_fn)(void *e, const uint8_t *host_id, uint32_t host_idx, void *msg, uint32_t msglen);
typedef int (*gysyn_ingest_raw_fn)(void *e, const uint8_t *host_id, uint32_t host_idx, uint32_t kind, const void *events, uint32_t n);
typedef struct gysyn_wire_round { void *msg1; void *msg2; const void *raw; uint32_t len1, len2, nraw, raw_kind; } gysyn_wire_round;

double gysyn_wire_run(void *engine, gysyn_ingest_msg_fn fmsg, gysyn_ingest_raw_fn fraw, const uint8_t *host_id, const gysyn_wire_round *rounds,
		uint32_t nthreads, uint32_t nrounds, uint32_t iters, int *nerr)
{
	std::atomic<int> ready {0}, errors {0};
	std::atomic<bool> go {false};
	std::vector<std::thread> thr;
	for (uint32_t t = 0; t < nthreads; ++t) {
		thr.emplace_back([&, t]() {
			const gysyn_wire_round *mine = rounds + (size_t)t * nrounds;
			ready.fetch_add(1);
			while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
			for (uint32_t i = 0; i < iters; ++i) {
				const gysyn_wire_round &r = mine[i % nrounds];
				int rc = fmsg(engine, host_id, t, r.msg1, r.len1);
				rc |= fmsg(engine, host_id, t, r.msg2, r.len2);
				rc |= fraw(engine, host_id, t, r.raw_kind, r.raw, r.nraw);
				if (rc) { errors.fetch_add(1); return; }
			}
		});
	}
	while (ready.load() < (int)nthreads) std::this_thread::yield();
	const auto t0 = std::chrono::steady_clock::now();
	go.store(true, std::memory_order_release);
	for (auto &x : thr) x.join();
	const auto t1 = std::chrono::steady_clock::now();
	if (nerr) *nerr = errors.load();
	return std::chrono::duration<double>(t1 - t0).count();
}
}
