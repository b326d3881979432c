// gysk_tdigest.cuh — warp-level building blocks of the batched merging t-digest (shared by the ingest-side update
// kernel and the multi-GPU merge kernels). Definitions: DESIGN.md §2; CPU statement: oracle/gysk_oracle.c.
#pragma once

#include "gysk_device.cuh"

namespace gysk {

// K_1 scale function of the merging t-digest (Dunning), k spanning [-delta/2, delta/2]: k(q) = delta/pi asin(2q - 1).
// td_q_next(q0) = q(k(q0) + 1), the upper end of the unit-k interval starting at q0, without inverse trig:
// sin(asin(2 q0 - 1) + pi/delta) = (2 q0 - 1) cos(pi/delta) + 2 sqrt(q0 (1 - q0)) sin(pi/delta). Only IEEE + - * / sqrt with
// explicit round-to-nearest (no FMA contraction), in the same order as oracle/gysk_oracle.c::td_q_next => identical bits.
struct TdRung { double C, S, qclamp; };		// cos(pi/delta'), sin(pi/delta'), (1 + C)/2 — computed once on the host
// A greedy pass over items that cannot be split needs up to ~1.3 delta clusters; a pass that would exceed TD_CAP is repeated on
// the next, coarser rung (delta' = delta x {1, .92, .85, .78, .72, .66}); only the last rung lets the last slot absorb the rest.
static constexpr int TD_LADDER = 6;
struct TdParams { TdRung r[TD_LADDER]; };

__device__ __forceinline__ double td_q_next(double q0, const TdRung &P)
{
	if (q0 >= P.qclamp) return 1.0;
	const double t = __dsub_rn(__dmul_rn(2.0, q0), 1.0);
	const double r = __dsqrt_rn(__dmul_rn(q0, __dsub_rn(1.0, q0)));
	const double a = __dmul_rn(t, P.C);
	const double b = __dmul_rn(__dmul_rn(2.0, r), P.S);
	return __dmul_rn(__dadd_rn(__dadd_rn(a, b), 1.0), 0.5);
}

__device__ __forceinline__ double td_wlimit(unsigned long long wsofar, unsigned long long W, const TdRung &P)
{
	const double q0 = wsofar ? __ddiv_rn((double)wsofar, (double)W) : 0.0;
	return __dmul_rn((double)W, td_q_next(q0, P));
}

template <int NMAX_>
struct TdWorkT				// per warp; NMAX = 2 x TD_CAP: 9.7 KB
{
	static constexpr int NMAX = NMAX_;			// capacity of the merged list, a multiple of 32
	double			mean[NMAX];			// merged list: means ...
	unsigned long long	pref[NMAX + 1];			// ... and exclusive weight prefix: weight of item i = pref[i+1] - pref[i]
	uint16_t		bounds[TD_CAP + 2];
	uint16_t		nxt[NMAX];
};
using TdWork = TdWorkT<2 * TD_CAP>;			// shared memory: two lists of up to TD_CAP centroids
using TdWorkBig = TdWorkT<1120>;			// global scratch: TD_CAP old centroids + up to NBINS (848) items of a batch
struct TdScratch : TdWork		// + an accumulator list for the folds of the merge step: 14.3 KB
{
	Centroid		newc[TD_CAP];
};

// Stable merge by mean of two mean-sorted centroid lists (`a` first on ties), then the greedy K_1 pass; one warp.
// Both inputs are fully consumed into S.mean / S.pref before `out` is written, so `out` may alias `a` or `b`.
// a and b may live in shared or global memory. Returns the number of centroids written to out (<= TD_CAP).
template <typename Work>
__device__ __forceinline__ uint32_t warp_merge_compress(Work &S, const Centroid *a, uint32_t na, const Centroid *b, uint32_t nb,
		Centroid *out, const TdParams &P)
{
	const int lane = threadIdx.x & 31;
	const uint32_t nm = na + nb;
	constexpr int IPL = Work::NMAX / 32;		// merged items per lane

	for (uint32_t j = lane; j < na; j += 32) {
		const Centroid c = a[j];
		uint32_t lo = 0, hi = nb;			// # of b with mean < c.mean
		while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (b[mid].mean < c.mean) lo = mid + 1; else hi = mid; }
		S.mean[j + lo] = c.mean; S.pref[j + lo + 1] = c.weight;
	}
	for (uint32_t j = lane; j < nb; j += 32) {
		const Centroid c = b[j];
		uint32_t lo = 0, hi = na;			// # of a with mean <= c.mean
		while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a[mid].mean <= c.mean) lo = mid + 1; else hi = mid; }
		S.mean[j + lo] = c.mean; S.pref[j + lo + 1] = c.weight;
	}
	__syncwarp();

	// in-place weight prefix: pref[i+1] holds w_i on entry and sum(w_0..w_i) on exit; lane owns IPL consecutive items
	{
		unsigned long long tot = 0;
		for (int j = 0; j < IPL; ++j) { const uint32_t i = lane * IPL + j; if (i < nm) tot += S.pref[i + 1]; }
		unsigned long long incl = tot;
#pragma unroll
		for (int off = 1; off < 32; off <<= 1) {
			const unsigned long long tt = __shfl_up_sync(0xffffffffu, incl, off);
			if (lane >= off) incl += tt;
		}
		unsigned long long run = incl - tot;
		for (int j = 0; j < IPL; ++j) { const uint32_t i = lane * IPL + j; if (i < nm) { run += S.pref[i + 1]; S.pref[i + 1] = run; } }
		if (lane == 0) S.pref[0] = 0;
	}
	__syncwarp();

	// greedy chain over the merged list: a cluster that starts at item i (after weight P = pref[i]) takes items while the
	// running total stays <= W q(k(P/W) + 1), and at least one item. The successor nxt[i] of EVERY possible start is
	// evaluated in parallel (the expensive double sqrt/div part); the chain itself is then a pointer walk by lane 0.
	uint32_t nout = 0;
	if (nm) {
		const unsigned long long W = S.pref[nm];
		for (int k = 0; k < TD_LADDER; ++k) {
			const TdRung R = P.r[k];
			for (uint32_t i = lane; i < nm; i += 32) {
				const double wl = td_wlimit(S.pref[i], W, R);
				uint32_t e = i + 1;				// largest e in [i+1, nm] with pref[e] <= wl
				while (e < nm && (double)S.pref[e + 1] <= wl) ++e;
				S.nxt[i] = (uint16_t)e;
			}
			__syncwarp();
			if (lane == 0) {
				const bool final = k == TD_LADDER - 1;
				uint32_t cs = 0;
				nout = 0;
				while (cs < nm) {
					if (nout == TD_CAP) { nout = TD_CAP + 1; break; }		// would need more than TD_CAP clusters
					uint32_t e = S.nxt[cs];
					if (final && nout == TD_CAP - 1) e = nm;			// the last slot absorbs whatever is left
					S.bounds[nout++] = (uint16_t)cs;
					cs = e;
				}
				if (nout <= TD_CAP) S.bounds[nout] = (uint16_t)nm;
			}
			nout = __shfl_sync(0xffffffffu, nout, 0);
			__syncwarp();
			if (nout <= TD_CAP) break;
		}
	}

	for (uint32_t c = lane; c < nout; c += 32) {
		double csum = 0.0;
		const uint32_t lo = S.bounds[c], hi = S.bounds[c + 1];
		for (uint32_t i = lo; i < hi; ++i) {
			csum = __dadd_rn(csum, __dmul_rn(S.mean[i], (double)(S.pref[i + 1] - S.pref[i])));
		}
		const unsigned long long cw = S.pref[hi] - S.pref[lo];
		Centroid o; o.mean = __ddiv_rn(csum, (double)cw); o.weight = cw;
		out[c] = o;
	}
	__syncwarp();
	return nout;
}

} // namespace gysk
