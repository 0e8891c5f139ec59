// bsgpu_device.cuh -- device-side building blocks shared by the kernels in bsgpu_kernels.cu.
//
// Everything here computes what the reference computes per site (citations to /root/reference), arranged for
// one-site-per-thread execution on sm_100a.  The translation unit is compiled with -fmad=false so that no a*b+c
// in this file is contracted: the reference binary is built without FMA (src/Makefile:44, x86-64 baseline) and
// structural ties between genotype likelihoods only survive if the operation sequence is the same.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "bsgpu_math.cuh"

namespace bsgpu {

constexpr int kMaxQual = 43;
constexpr int kFltQual = 63;
constexpr double kLn10 = 2.30258509299404568402;

// Tables and constants computed once on the host with the C library (so they are bit-identical to what the
// reference computes, src/genotype_model.c:10-21, src/stats_utils.c:14-21) and uploaded at bsgpu_init.
struct Tables {                   // copied into shared memory once per CTA
	double qp[kMaxQual + 1][4];   // k, ln k, ln(1/2 + k), ln(1 + k)
	MathTables math;              // reduction tables of fast_log / fast_exp (bsgpu_math.cuh)
};
struct DevConst {
	Tables tab;
	double lfact[256];            // log-factorials for the Fisher test (het sites only: read through L1, not staged)
	double l, t;                  // 1 - under_conv, over_conv
	double lrb, lrb1;             // ln(ref_bias), ln((1 + ref_bias) / 2)
	int min_qual;
	int pad_;
	// gather kernel: packed base byte -> increment of the narrow accumulators {bases A,C | bases G,T}; each half-word is
	// 5-bit count | 11-bit quality; zero for a byte that does not count (q < min_qual, q = 63)
	uint32_t pile_lut[256][2];
};

// What a thread holds for its site before the model runs: the reference's pileup record in registers.
struct SiteCounts {
	uint32_t cnt[2][8];
	uint32_t n;
	float qsum[8];
	float mapq2;
};

// ---------------------------------------------------------------------------------------------------------------
// Exact division helpers.  a / b is computed as q = a*r, q' = fma(fma(-q, b, a), r, q) with r = RN(1/b) (Markstein):
// correctly rounded, i.e. bit-identical to the IEEE quotient the reference computes, at 3 FP64 issues instead of the
// ~35 of the generic division sequence (checked against a/b on 3e8 random operands).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double div_by(double a, double b, double r) {
	const double q = a * r;
	return fma(fma(-q, b, a), r, q);
}
constexpr double kInvLn10 = 1.0 / kLn10;
// Guard band of the genotype call: relative distance of the two best log-likelihoods below which the call is reported as
// "may differ from the reference" (the parity bar on gt_prob[] itself is 1e-9 relative).  Entries of the guard list are
// kind << 56 | site id; kinds: 1 the two best genotypes are closer than the band or EQUAL (the reference adds the same terms
// in another order for some genotype pairs, so a tie here need not be a tie there), 2 QUAL / GQ within its band of an
// integer, 3 FS within its band.
constexpr double kTieBand = 1.0e-9;
constexpr int kGuardCap = 65536;           // flagged sites listed per context between two reads (all of them are counted)
constexpr int kGuardList = 16;             // d_counters[16 ..): the list; [8] = its length; [4] near ties, [5] exact ties, [6] QUAL, [7] FS

// ---------------------------------------------------------------------------------------------------------------
// Closed-form ML conversion fraction.  src/genotype_model.c:23-42.  Z is garbage (NaN) when a + b == 0; the caller
// only uses it for classes that have counts, exactly like the reference.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void conv_ml(double a, double b, double ka, double kb, double l, double t, double *Z) {
	const double lpt = l + t, lmt = l - t;
	const double d0 = (a + b) * lmt;
	const double d = d0 != 0.0 ? d0 : 1.0;      // empty pair: result unused; keeps the reciprocal off its special-case path
	const double r = __drcp_rn(d);
	const double two_m = 2.0 - lpt;
	double num[3];
	num[0] = a * (lpt + 2.0 * kb) - b * (two_m + 2.0 * ka);
	num[1] = a * (2.0 + lpt + 4.0 * kb) - b * (two_m + 4.0 * ka);
	num[2] = a * (lpt + 4.0 * kb) - b * (two_m + 4.0 * ka);
#pragma unroll
	for (int i = 0; i < 3; i++) {
		double s = div_by(num[i], d, r);
		s = s < -1.0 ? -1.0 : (s > 1.0 ? 1.0 : s);
		Z[i] = 0.5 * (lmt * s + 2.0 - lpt);
	}
}

// number of copies of base b in genotype g (AA AC AG AT CC CG CT GG GT TT)
__host__ __device__ constexpr int gt_hits(int g, int b) {
	constexpr int a0[10] = {0, 0, 0, 0, 1, 1, 1, 2, 2, 3};
	constexpr int a1[10] = {0, 1, 2, 3, 1, 2, 3, 2, 3, 3};
	return (a0[g] == b) + (a1[g] == b);
}

// exclusive prefix sum of a small per-lane count over the warp; *total = warp sum
__device__ __forceinline__ int warp_offsets(int c, int lane, int *total) {
	int incl = c;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
	*total = __shfl_sync(0xffffffffu, incl, 31);
	return incl - c;
}

// ---------------------------------------------------------------------------------------------------------------
// 10 genotype log-likelihoods + reference prior + normalisation.  src/genotype_model.c:44-246
//
// Warp-cooperative: all 32 lanes of a warp call this together, one site per lane.  Each ll[g] receives the prior and
// then one addend per non-empty class, classes in order 0..7, exactly as the reference does.  The transcendental
// calls are what the per-site cost is made of, and which of them a site needs depends on which classes it has counts
// in (3 logs for a typical A/T site, 6 for C/G, up to 12; 2-4 of the 10 exps are not vanishing).  Evaluating them
// under per-lane branches makes every warp pay for the union.  Instead the lanes pool their arguments in a per-warp
// shared-memory list (`wbuf`, 32 x 12 doubles), the warp evaluates the list 32 at a time fully converged, and each
// lane reads its results back.  Values and order of every addition are unchanged.
//   * exp(x) with x < -45 is replaced by 0: such a term (< 3e-20) cannot change a double sum that contains the
//     exp(0) = 1 of the best genotype, so the result is the same double.
// Returns max_gt | tie << 8 (tie: 0, 1 = runner-up inside the guard band, 2 = equal to the best) and writes log10 posteriors
// to prob[10].
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int genotype_model(const uint32_t cnt[8], const int qual[8], int rf,
		const DevConst *__restrict__ dc, const Tables *__restrict__ tb, double prob[10],
		double *__restrict__ wbuf, int lane, bool covered) {
	const double (*__restrict__ qp)[4] = tb->qp;
	const MathTables *__restrict__ mt = &tb->math;
	double ll[10];
	const double l = dc->l, t = dc->t;
#pragma unroll
	for (int g = 0; g < 10; g++) {
		double v = 0.0;
#pragma unroll
		for (int b = 0; b < 4; b++) {
			if (gt_hits(g, b) == 2) v = (rf == b + 1) ? dc->lrb : v;
			else if (gt_hits(g, b) == 1) v = (rf == b + 1) ? dc->lrb1 : v;
		}
		ll[g] = v;
	}
	// classes 0-3 (:109-164).  Branch-free: an empty class has n = 0 and qual 0, so its three products are signed zeros
	// and the additions leave ll untouched.
#pragma unroll
	for (int j = 0; j < 4; j++) {
		const double n = (double)cnt[j];
		const double *q = qp[qual[j]];
		const double v0 = n * q[1], v1 = n * q[2], v2 = n * q[3];
#pragma unroll
		for (int g = 0; g < 10; g++) ll[g] += gt_hits(g, j) == 2 ? v2 : (gt_hits(g, j) == 1 ? v1 : v0);
	}
	// :165-171
	double Z[6];
	const double n4 = (double)cnt[4], n5 = (double)cnt[5], n6 = (double)cnt[6], n7 = (double)cnt[7];
	const double *q4 = qp[qual[4]], *q5 = qp[qual[5]], *q6 = qp[qual[6]], *q7 = qp[qual[7]];
	const double k4 = q4[0], k5 = q5[0], k6 = q6[0], k7 = q7[0];
	conv_ml(n5, n7, k5, k7, l, t, Z);
	conv_ml(n6, n4, k6, k4, l, t, Z + 3);

	// ---- pooled log() of the informative-class arguments (:173-230)
	const bool c4 = cnt[4] != 0, c5 = cnt[5] != 0, c6 = cnt[6] != 0, c7 = cnt[7] != 0;
	const int nlog = 3 * ((int)c4 + (int)c5 + (int)c6 + (int)c7);
	int total;
	const int o4 = warp_offsets(nlog, lane, &total);
	const int o5 = o4 + (c4 ? 3 : 0), o6 = o5 + (c5 ? 3 : 0), o7 = o6 + (c6 ? 3 : 0);
	if (c4) { wbuf[o4] = 1.0 - 0.5 * Z[4] + k4; wbuf[o4 + 1] = 1.0 - Z[3] + k4; wbuf[o4 + 2] = 0.5 * (1.0 - Z[5]) + k4; }
	if (c5) { wbuf[o5] = Z[0] + k5; wbuf[o5 + 1] = 0.5 * Z[2] + k5; wbuf[o5 + 2] = 0.5 * Z[1] + k5; }
	if (c6) { wbuf[o6] = Z[3] + k6; wbuf[o6 + 1] = 0.5 * Z[5] + k6; wbuf[o6 + 2] = 0.5 * Z[4] + k6; }
	if (c7) { wbuf[o7] = 1.0 - Z[0] + k7; wbuf[o7 + 1] = 1.0 - 0.5 * Z[1] + k7; wbuf[o7 + 2] = 0.5 * (1.0 - Z[2]) + k7; }
	__syncwarp();
	// two independent evaluations per trip for instruction-level parallelism
	for (int i = lane; i < total; i += 64) {
		const bool two = i + 32 < total;
		const double a = wbuf[i], b = two ? wbuf[i + 32] : 1.0;
		const double la = fast_log(a, mt), lb = fast_log(b, mt);
		wbuf[i] = la;
		if (two) wbuf[i + 32] = lb;
	}
	__syncwarp();
	// Add the four informative classes, in class order.  Branch-free: a class without counts has n = 0 and reads 0 for
	// its logs, so every addend is a signed zero and ll is unchanged.
	{   // informative A
		const double a0 = c4 ? wbuf[o4] : 0.0, a1 = c4 ? wbuf[o4 + 1] : 0.0, a2 = c4 ? wbuf[o4 + 2] : 0.0;
		const double kk = n4 * q4[1], half = n4 * q4[2], one = n4 * q4[3];
		const double ag = a0 * n4, gg = a1 * n4, mix = a2 * n4;
		ll[0] += one; ll[2] += ag; ll[7] += gg; ll[5] += mix; ll[8] += mix;
		ll[1] += half; ll[3] += half; ll[4] += kk; ll[6] += kk; ll[9] += kk;
	}
	{   // informative C
		const double a0 = c5 ? wbuf[o5] : 0.0, a1 = c5 ? wbuf[o5 + 1] : 0.0, a2 = c5 ? wbuf[o5 + 2] : 0.0;
		const double kk = n5 * q5[1];
		const double cc = a0 * n5, mix = a1 * n5, ct = a2 * n5;
		ll[4] += cc; ll[1] += mix; ll[5] += mix; ll[6] += ct;
		ll[0] += kk; ll[2] += kk; ll[3] += kk; ll[7] += kk; ll[8] += kk; ll[9] += kk;
	}
	{   // informative G
		const double a0 = c6 ? wbuf[o6] : 0.0, a1 = c6 ? wbuf[o6 + 1] : 0.0, a2 = c6 ? wbuf[o6 + 2] : 0.0;
		const double kk = n6 * q6[1];
		const double gg = a0 * n6, mix = a1 * n6, ag = a2 * n6;
		ll[7] += gg; ll[5] += mix; ll[8] += mix; ll[2] += ag;
		ll[0] += kk; ll[1] += kk; ll[3] += kk; ll[4] += kk; ll[6] += kk; ll[9] += kk;
	}
	{   // informative T
		const double a0 = c7 ? wbuf[o7] : 0.0, a1 = c7 ? wbuf[o7 + 1] : 0.0, a2 = c7 ? wbuf[o7 + 2] : 0.0;
		const double kk = n7 * q7[1], half = n7 * q7[2], one = n7 * q7[3];
		const double cc = a0 * n7, ct = a1 * n7, mix = a2 * n7;
		ll[9] += one; ll[4] += cc; ll[6] += ct; ll[1] += mix; ll[5] += mix;
		ll[3] += half; ll[8] += half; ll[0] += kk; ll[2] += kk; ll[7] += kk;
	}
	__syncwarp();
	// first strict maximum (:231-239)
	double top = ll[0];
	int best = 0;
#pragma unroll
	for (int g = 1; g < 10; g++) if (ll[g] > top) { top = ll[g]; best = g; }
	// ---- exp() of the differences (:240-242), all ten evaluated straight: the table-driven exp is ~18 instructions, which is
	// less than what pooling the 2-4 non-vanishing ones across the warp costs in mask / offset / shared-memory traffic.
	// The same differences feed the guard band: the transcendental terms of ll[] are within 1.5 ulp of libm's, so the order
	// of two genotypes whose likelihoods differ by less than kTieBand (relative) is not guaranteed to be the reference's.
	// Such sites are reported (bsgpu_stats.near_tie_sites / exact_tie_sites, bsgpu_guard_read), not hidden.
	double sum = 0.0;
#pragma unroll
	for (int g = 0; g < 10; g++) {
		const double x = ll[g] - top;
		const double e = fast_exp(x < -45.0 ? -45.0 : x, mt);
		sum += x < -45.0 ? 0.0 : (x == 0.0 ? 1.0 : e);
	}
	// A runner-up inside the band (x >= -band) contributes e^x >= 1 - band to the sum, so only sites with sum >= 2 - 2 band (the
	// call holds less than half of the posterior mass: rare at sequencing depth) are looked at genotype by genotype.
	int tie = 0;
#ifndef BSGPU_NO_GUARD
	// (an uncovered site has ten equal likelihoods and no call: without `covered` its lane would drag most warps through the
	// look, 3 % empty sites put one into 62 % of the warps; band <= 1e-9 |top| is far below the 0.005 of the first test)
	if (covered && sum >= 1.99) {
		const double band = -kTieBand * (fabs(top) > 1.0 ? fabs(top) : 1.0);
		int ntop = 0, near = 0;
#pragma unroll
		for (int g = 0; g < 10; g++) {
			const double x = ll[g] - top;
			ntop += x == 0.0;
			near |= x != 0.0 && x >= band;
		}
		tie = ntop > 1 ? 2 : near;
	}
#endif
	sum = fast_log(sum, mt);          // sum is in [1, 10]
#pragma unroll
	for (int g = 0; g < 10; g++) prob[g] = div_by(ll[g] - top - sum, kLn10, kInvLn10);
	return best | tie << 8;          // the tie flag rides in the same register as the call (the model runs at the register limit)
}

// ---------------------------------------------------------------------------------------------------------------
// Two-sided Fisher exact test.  src/stats_utils.c:25-91, lfact2 include/bs_call.h:335
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double lfact(const double *__restrict__ tab, int x) {
	return x < 256 ? tab[x] : lgamma((double)(x + 1));
}

__device__ __forceinline__ double table_prob(const double *__restrict__ tab, double knst, int c0, int c1, int c2, int c3) {
	return exp(knst - lfact(tab, c0) - lfact(tab, c1) - lfact(tab, c2) - lfact(tab, c3));
}

__device__ __forceinline__ void tail_walk(int dn0, int dn1, int up0, int up1, int steps, double &l, double &p) {
	for (int i = 0; i < steps; i++) {
		l *= (double)((dn0 - i) * (dn1 - i)) / (double)((up0 + i + 1) * (up1 + i + 1));
		p += l;
	}
}

static __device__ __noinline__ double fisher_exact(const double *__restrict__ tab, int c0, int c1, int c2, int c3) {
	const int row0 = c0 + c1, row1 = c2 + c3, col0 = c0 + c2, col1 = c1 + c3;
	const int n = row0 + row1;
	if (n == 0) return 1.0;
	const double delta = (double)c0 - (double)(row0 * col0) / (double)n;
	const double knst = lfact(tab, col0) + lfact(tab, col1) + lfact(tab, row0) + lfact(tab, row1) - lfact(tab, n);
	double l = table_prob(tab, knst, c0, c1, c2, c3);
	double p = l;
	const int lead = min(c0, c3), cntr = min(c1, c2);
	if (delta > 0.0) {
		tail_walk(c1, c2, c0, c3, cntr, l, p);
		const int k = (int)ceil(2.0 * delta);
		if (k <= lead) {
			c0 -= k; c3 -= k; c1 += k; c2 += k;
			l = table_prob(tab, knst, c0, c1, c2, c3);
			p += l;
			tail_walk(c0, c3, c1, c2, lead - k, l, p);
		}
	} else {
		tail_walk(c0, c3, c1, c2, lead, l, p);
		int k = (int)ceil(-2.0 * delta);
		if (!k) k = 1;
		if (k <= cntr) {
			c0 += k; c3 += k; c1 -= k; c2 -= k;
			l = table_prob(tab, knst, c0, c1, c2, c3);
			p += l;
			tail_walk(c1, c2, c0, c3, cntr - k, l, p);
		}
	}
	return p;
}

// allele x strand table of the called het genotype (src/call_genotypes.c:62-104); bit j of a set = class j.
// The GT case keeps the reference's counts[0][6] in the ori-1 cell (line 98).
__device__ __forceinline__ int class_sum(const uint32_t c[8], uint32_t set) {
	int s = 0;
#pragma unroll
	for (int j = 0; j < 8; j++) s += (set >> j & 1) ? (int)c[j] : 0;
	return s;
}

__device__ __forceinline__ double strand_bias(const SiteCounts &s, int max_gt, const double *__restrict__ lfact_tab) {
	uint32_t a1, a2;
	switch (max_gt) {
	case 1: a1 = 0x11; a2 = 0xa2; break;   // AC
	case 2: a1 = 0x01; a2 = 0x44; break;   // AG
	case 3: a1 = 0x11; a2 = 0x88; break;   // AT
	case 5: a1 = 0xa2; a2 = 0x54; break;   // CG
	case 6: a1 = 0x22; a2 = 0x08; break;   // CT
	case 8: a1 = 0x54; a2 = 0x88; break;   // GT
	default: return 0.0;
	}
	int f0 = class_sum(s.cnt[0], a1), f1 = class_sum(s.cnt[0], a2);
	int f2 = class_sum(s.cnt[1], a1), f3 = class_sum(s.cnt[1], a2);
	if (max_gt == 8) f2 = (int)(s.cnt[1][2] + s.cnt[1][4] + s.cnt[0][6]);
	double z = fisher_exact(lfact_tab, f0, f1, f2, f3);
	if (z < 1.0e-20) z = 1.0e-20;
	return div_by(log(z), kLn10, kInvLn10);
}

// ---------------------------------------------------------------------------------------------------------------
// Per-site body of call_thread (src/call_genotypes.c:43-115): summarise, model, strand bias; writes the 200-byte
// gt_meth image as 25 eight-byte words into `rec` (a row of the CTA's staging tile in shared memory).
// Returns false for a site with no counted base (record zeroed, the caller sets skip).  Byte 1 of the record's last word
// (padding of gt_meth) carries the guard flag of the call -- 0, 1 = the two best genotypes are closer than the guard band
// (the call may differ from the reference's), 2 = they are exactly equal -- for the caller to read and clear.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool call_site(const SiteCounts &s, int rf, const DevConst *__restrict__ dc,
		const Tables *__restrict__ tb, uint64_t *rec, double *__restrict__ wbuf, int lane) {
	uint32_t tot[8];
	int qual[8];
	float tq = 0.0f;
#pragma unroll
	for (int j = 0; j < 8; j++) {
		tot[j] = s.cnt[0][j] + s.cnt[1][j];
		// float divide, double add, narrowed to float, floorf (:50).  Branch-free: an empty class computes 1/1 instead of
		// 0/0 (a zero or NaN operand sends the FP32 division into its special-case subroutine) and is discarded.
		const float nn = tot[j] ? (float)tot[j] : 1.0f;
		const float qs = tot[j] ? s.qsum[j] : 1.0f;
		const int qj = (int)floorf((float)(0.5 + (double)(qs / nn)));
		qual[j] = tot[j] ? qj : 0;
		tq += tot[j] ? s.qsum[j] : 0.0f;
	}
	// the whole warp runs the model together (a lane without counts contributes nothing to the pooled lists)
	double prob[10];
	const int best_tie = genotype_model(tot, qual, rf, dc, tb, prob, wbuf, lane, s.n != 0);
	const int best = best_tie & 0xff;
	__syncwarp();                      // wbuf may alias this warp's output rows: everyone is done with it
	if (!s.n) {
#pragma unroll
		for (int i = 0; i < 25; i++) rec[i] = 0;
		return false;
	}
	// a lane without counts computes 1/1 (discarded) instead of 0/0, which would drag its whole warp through the FP32
	// division's special-case subroutine
	const float fn = s.n ? (float)s.n : 1.0f;
	const int aq = (int)floorf((float)(0.5 + (double)((s.n ? tq : 1.0f) / fn)));
	const int mq = (int)(0.5 + sqrt((double)((s.n ? s.mapq2 : 1.0f) / fn)));
	const double fs = strand_bias(s, best, dc->lfact);
#pragma unroll
	for (int j = 0; j < 8; j++) rec[j] = tot[j];
#pragma unroll
	for (int j = 0; j < 4; j++) rec[8 + j] = (uint64_t)(uint32_t)qual[2 * j] | ((uint64_t)(uint32_t)qual[2 * j + 1] << 32);
#pragma unroll
	for (int g = 0; g < 10; g++) rec[12 + g] = (uint64_t)__double_as_longlong(prob[g]);
	rec[22] = (uint64_t)__double_as_longlong(fs);
	rec[23] = (uint64_t)(uint32_t)mq | ((uint64_t)(uint32_t)aq << 32);
	rec[24] = (uint64_t)(uint32_t)best_tie;
	return true;
}

}  // namespace bsgpu
