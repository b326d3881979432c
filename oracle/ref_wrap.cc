// TEST INFRASTRUCTURE ONLY (see oracle/README.md). Never linked into the product.
//
// C-ABI wrapper around the REFERENCE's own GY_HISTOGRAM / bucket-hash / jhash code.
// The reference sources are compiled from where they lie under /root/reference:
// oracle/Makefile generates oracle/_ref/gy_hist_trim.h at build time from
// common/gy_statistics.h lines 455-980 (HIST_SERIAL .. GY_HISTOGRAM_DATA) and
// 1565-2063 (bucket hash classes), wrapped in namespace gyeeta — the folly / liburcu
// dependent parts of that header cannot be built here. Nothing from
// the reference is copied into the repository; oracle/_ref/ is git-ignored.
//
// Everything exported here exists so that tests can pin oracle/gysk_oracle.c (the CPU
// restatement) against outputs of the reference itself, and so bench.py --impl reference
// can time the reference's own add_data loop ("kind": "reference").

#include "gy_common_inc.h"
#include "gy_hist_trim.h"   // generated into oracle/_ref by the Makefile

#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

using namespace gyeeta;

namespace {

// class ids shared with oracle/gysk_oracle.h (GYO_CLS_*)
enum {
	CLS_RESP_TIME = 0, CLS_SEMI_LOG, CLS_SEMI_LOG_LO, CLS_DURATION, CLS_HASH_10_5000,
	CLS_HASH_5_250, CLS_HASH_1_3000, CLS_PERCENT, CLS_FD_I8_9_26_5, CLS_FD_INT_M15_M3_4, CLS_MAX
};

using FD_I8  = FIXED_DIFF_HASH<int8_t, 9, 26, 5>;       // test/test_histogram.cc:17
using FD_N4  = FIXED_DIFF_HASH<int, -15, -3, 4>;        // test/test_histogram.cc:92

struct ser_t { uint64_t count; int64_t sum; };
static_assert(sizeof(ser_t) == sizeof(HIST_SERIAL), "HIST_SERIAL layout");

template <typename T, typename H>
int run_hist(const int64_t *vals, size_t n, const float *pcts, size_t npct, ser_t *out_stats,
		uint64_t *out_total, int64_t *out_max, int64_t *out_pct, int64_t *out_bucket_ids, float *out_avg)
{
	GY_HISTOGRAM<T, H>	hist(1);

	for (size_t i = 0; i < n; ++i) {
		size_t b = hist.add_data((T)vals[i], 1);
		if (out_bucket_ids) out_bucket_ids[i] = (int64_t)b;
	}

	size_t		total;
	T		maxv;
	uint64_t	ec, sc;

	HIST_SERIAL	ser[GY_HISTOGRAM<T, H>::maxbuckets_];

	hist.get_serialized(ser, total, maxv, ec, sc);
	if (out_stats) std::memcpy(out_stats, ser, sizeof(ser));
	if (out_total) *out_total = total;
	if (out_max) *out_max = (int64_t)maxv;

	if (npct) {
		std::vector<HIST_DATA>	hd(npct);
		for (size_t i = 0; i < npct; ++i) hd[i].percentile = pcts[i];
		float		avg = 0;
		hist.get_percentiles(hd.data(), npct, total, maxv, &avg);
		for (size_t i = 0; i < npct; ++i) out_pct[i] = hd[i].data_value;
		if (out_avg) *out_avg = avg;
	}

	return (int)GY_HISTOGRAM<T, H>::maxbuckets_;
}

template <typename T, typename H>
int pct_from_serial(const ser_t *stats, uint64_t total, int64_t maxv, const float *pcts, size_t npct, int64_t *out_pct, float *out_avg)
{
	GY_HISTOGRAM<T, H>	hist(1);

	hist.update_from_serialized(reinterpret_cast<const HIST_SERIAL *>(stats), total, (T)maxv);

	std::vector<HIST_DATA>	hd(npct);
	for (size_t i = 0; i < npct; ++i) hd[i].percentile = pcts[i];
	size_t		t;
	T		m;
	float		avg = 0;
	hist.get_percentiles(hd.data(), npct, t, m, &avg);
	for (size_t i = 0; i < npct; ++i) out_pct[i] = hd[i].data_value;
	if (out_avg) *out_avg = avg;
	return (int)GY_HISTOGRAM<T, H>::maxbuckets_;
}

#define DISPATCH(FN, ...) \
	switch (cls) { \
	case CLS_RESP_TIME	: return t_is_int ? FN<int, RESP_TIME_HASH>(__VA_ARGS__) : FN<int64_t, RESP_TIME_HASH>(__VA_ARGS__); \
	case CLS_SEMI_LOG	: return t_is_int ? FN<int, SEMI_LOG_HASH>(__VA_ARGS__) : FN<int64_t, SEMI_LOG_HASH>(__VA_ARGS__); \
	case CLS_SEMI_LOG_LO	: return t_is_int ? FN<int, SEMI_LOG_HASH_LO>(__VA_ARGS__) : FN<int64_t, SEMI_LOG_HASH_LO>(__VA_ARGS__); \
	case CLS_DURATION	: return t_is_int ? FN<int, DURATION_HASH>(__VA_ARGS__) : FN<int64_t, DURATION_HASH>(__VA_ARGS__); \
	case CLS_HASH_10_5000	: return t_is_int ? FN<int, HASH_10_5000>(__VA_ARGS__) : FN<int64_t, HASH_10_5000>(__VA_ARGS__); \
	case CLS_HASH_5_250	: return t_is_int ? FN<int, HASH_5_250>(__VA_ARGS__) : FN<int64_t, HASH_5_250>(__VA_ARGS__); \
	case CLS_HASH_1_3000	: return t_is_int ? FN<int, HASH_1_3000>(__VA_ARGS__) : FN<int64_t, HASH_1_3000>(__VA_ARGS__); \
	case CLS_PERCENT	: return t_is_int ? FN<int, PERCENT_HASH>(__VA_ARGS__) : FN<int64_t, PERCENT_HASH>(__VA_ARGS__); \
	case CLS_FD_I8_9_26_5	: return FN<int8_t, FD_I8>(__VA_ARGS__); \
	case CLS_FD_INT_M15_M3_4: return FN<int, FD_N4>(__VA_ARGS__); \
	default			: return -1; \
	}

} // namespace

extern "C" {

// Runs a fresh GY_HISTOGRAM<T, cls> over vals[0..n): returns nbuckets, fills serialized stats,
// total, max, requested percentiles (get_percentiles), per-sample bucket ids and the float avg.
int gyref_hist_run(int cls, int t_is_int, const int64_t *vals, size_t n, const float *pcts, size_t npct,
		void *out_stats, uint64_t *out_total, int64_t *out_max, int64_t *out_pct, int64_t *out_bucket_ids, float *out_avg)
{
	DISPATCH(run_hist, vals, n, pcts, npct, (ser_t *)out_stats, out_total, out_max, out_pct, out_bucket_ids, out_avg)
}

// update_from_serialized() into an empty histogram followed by get_percentiles(): proves a GPU
// export is accepted by the reference's own deserialiser and yields the reference's percentiles.
int gyref_hist_pct_from_serial(int cls, int t_is_int, const void *stats, uint64_t total, int64_t maxv,
		const float *pcts, size_t npct, int64_t *out_pct, float *out_avg)
{
	DISPATCH(pct_from_serial, (const ser_t *)stats, total, maxv, pcts, npct, out_pct, out_avg)
}

uint32_t gyref_uint64_hash(uint64_t k)				{ return get_uint64_hash(k); }
uint32_t gyref_jhash_2words(uint32_t a, uint32_t b, uint32_t iv)	{ return jhash_2words(a, b, iv); }
uint32_t gyref_jhash2(const uint32_t *k, uint32_t len, uint32_t iv)	{ return jhash2(k, len, iv); }
uint32_t gyref_jhash(const void *k, uint32_t len, uint32_t iv)		{ return jhash(k, len, iv); }

size_t gyref_sizeof_hist_resp(void)				{ return sizeof(GY_HISTOGRAM<int64_t, RESP_TIME_HASH>); }

// CPU baseline ("kind": "reference"): the reference's own GY_HISTOGRAM::add_data over (slot, value) samples that the caller has
// PRE-SHARDED: thread t owns samples [offs[t], offs[t+1]) (its slots are private to it: slot % nthreads == t, mirrors
// l1_thr_num % maxthr, gy_mconnhdlr.cc:16252) and a private array of histograms. Nothing but add_data in the timed region.
double gyref_bench_resp_hist(const uint32_t *slots, const int64_t *vals_ms, const uint64_t *offs, uint32_t nslots, int nthreads, uint64_t *out_total)
{
	using H = GY_HISTOGRAM<int64_t, RESP_TIME_HASH>;

	if (nthreads < 1) nthreads = 1;

	std::vector<std::vector<H>>	tbl(nthreads);
	for (int t = 0; t < nthreads; ++t) tbl[t].assign(nslots / nthreads + 1, H(1));

	struct timespec		t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);

	auto work = [&](int t) {
		auto & mine = tbl[t];
		for (uint64_t i = offs[t]; i < offs[t + 1]; ++i) mine[slots[i] / nthreads].add_data(vals_ms[i], 1);
	};

	if (nthreads == 1) work(0);
	else {
		std::vector<std::thread> thr;
		for (int t = 0; t < nthreads; ++t) thr.emplace_back(work, t);
		for (auto & th : thr) th.join();
	}

	clock_gettime(CLOCK_MONOTONIC, &t1);

	uint64_t tot = 0;
	for (auto & v : tbl) for (auto & h : v) tot += h.get_total_count();
	if (out_total) *out_total = tot;

	return (t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec) * 1e-9;
}

} // extern "C"
