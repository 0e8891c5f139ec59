// bsgpu_device.cuh -- device-side building blocks shared by the kernels in bsgpu_kernels.cu.
//
// Everything here computes what the reference computes per site (citations to /root/reference), arranged for
// one-site-per-thread execution on sm_100a.  The translation unit is compiled with -fmad=false so that no a*b+c
// in this file is contracted: the reference binary is built without FMA (src/Makefile:44, x86-64 baseline) and
// structural ties between genotype likelihoods only survive if the operation sequence is the same.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bsgpu {

constexpr int kMaxQual = 43;
constexpr int kFltQual = 63;
constexpr double kLn10 = 2.30258509299404568402;

// Tables and constants computed once on the host with the C library (so they are bit-identical to what the
// reference computes, src/genotype_model.c:10-21, src/stats_utils.c:14-21) and uploaded at bsgpu_init.
struct DevConst {
	double qp[kMaxQual + 1][4];   // k, ln k, ln(1/2 + k), ln(1 + k)
	double lfact[256];
	double l, t;                  // 1 - under_conv, over_conv
	double lrb, lrb1;             // ln(ref_bias), ln((1 + ref_bias) / 2)
	int min_qual;
	int pad_;
};

// What a thread holds for its site before the model runs: the reference's pileup record in registers.
struct SiteCounts {
	uint32_t cnt[2][8];
	uint32_t n;
	float qsum[8];
	float mapq2;
};

// ---------------------------------------------------------------------------------------------------------------
// Closed-form ML conversion fraction.  src/genotype_model.c:23-42
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void conv_ml(double a, double b, double ka, double kb, double l, double t, double *Z) {
	const double lpt = l + t, lmt = l - t;
	const double d = (a + b) * lmt;
	const double two_m = 2.0 - lpt;
	double num[3];
	num[0] = a * (lpt + 2.0 * kb) - b * (two_m + 2.0 * ka);
	num[1] = a * (2.0 + lpt + 4.0 * kb) - b * (two_m + 4.0 * ka);
	num[2] = a * (lpt + 4.0 * kb) - b * (two_m + 4.0 * ka);
#pragma unroll
	for (int i = 0; i < 3; i++) {
		double s = num[i] / d;
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

// ---------------------------------------------------------------------------------------------------------------
// 10 genotype log-likelihoods + reference prior + normalisation.  src/genotype_model.c:44-246
// Each ll[g] receives the prior and then one addend per non-empty class, classes in order 0..7, exactly as the
// reference does; returns max_gt and writes log10 posteriors to prob[10].
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int genotype_model(const uint32_t cnt[8], const int qual[8], int rf,
		const DevConst *__restrict__ dc, const double (*__restrict__ qp)[4], double prob[10]) {
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
	// classes 0-3 (:109-164)
#pragma unroll
	for (int j = 0; j < 4; j++) {
		if (cnt[j]) {
			const double n = (double)cnt[j];
			const double *q = qp[qual[j]];
			const double v0 = n * q[1], v1 = n * q[2], v2 = n * q[3];
#pragma unroll
			for (int g = 0; g < 10; g++) ll[g] += gt_hits(g, j) == 2 ? v2 : (gt_hits(g, j) == 1 ? v1 : v0);
		}
	}
	// :165-171
	double Z[6];
	const double n4 = (double)cnt[4], n5 = (double)cnt[5], n6 = (double)cnt[6], n7 = (double)cnt[7];
	const double k4 = qp[qual[4]][0], k5 = qp[qual[5]][0], k6 = qp[qual[6]][0], k7 = qp[qual[7]][0];
	if (cnt[5] | cnt[7]) conv_ml(n5, n7, k5, k7, l, t, Z);
	if (cnt[4] | cnt[6]) conv_ml(n6, n4, k6, k4, l, t, Z + 3);
	if (cnt[4]) {   // informative A (:173-187)
		const double *q = qp[qual[4]];
		const double kk = n4 * q[1], half = n4 * q[2], one = n4 * q[3];
		const double ag = log(1.0 - 0.5 * Z[4] + k4) * n4;
		const double gg = log(1.0 - Z[3] + k4) * n4;
		const double mix = log(0.5 * (1.0 - Z[5]) + k4) * n4;
		ll[0] += one; ll[2] += ag; ll[7] += gg; ll[5] += mix; ll[8] += mix;
		ll[1] += half; ll[3] += half; ll[4] += kk; ll[6] += kk; ll[9] += kk;
	}
	if (cnt[5]) {   // informative C (:188-201)
		const double kk = n5 * qp[qual[5]][1];
		const double cc = log(Z[0] + k5) * n5;
		const double mix = log(0.5 * Z[2] + k5) * n5;
		const double ct = log(0.5 * Z[1] + k5) * n5;
		ll[4] += cc; ll[1] += mix; ll[5] += mix; ll[6] += ct;
		ll[0] += kk; ll[2] += kk; ll[3] += kk; ll[7] += kk; ll[8] += kk; ll[9] += kk;
	}
	if (cnt[6]) {   // informative G (:202-215)
		const double kk = n6 * qp[qual[6]][1];
		const double gg = log(Z[3] + k6) * n6;
		const double mix = log(0.5 * Z[5] + k6) * n6;
		const double ag = log(0.5 * Z[4] + k6) * n6;
		ll[7] += gg; ll[5] += mix; ll[8] += mix; ll[2] += ag;
		ll[0] += kk; ll[1] += kk; ll[3] += kk; ll[4] += kk; ll[6] += kk; ll[9] += kk;
	}
	if (cnt[7]) {   // informative T (:216-230)
		const double *q = qp[qual[7]];
		const double kk = n7 * q[1], half = n7 * q[2], one = n7 * q[3];
		const double cc = log(1.0 - Z[0] + k7) * n7;
		const double ct = log(1.0 - 0.5 * Z[1] + k7) * n7;
		const double mix = log(0.5 * (1.0 - Z[2]) + k7) * n7;
		ll[9] += one; ll[4] += cc; ll[6] += ct; ll[1] += mix; ll[5] += mix;
		ll[3] += half; ll[8] += half; ll[0] += kk; ll[2] += kk; ll[7] += kk;
	}
	// first strict maximum (:231-239)
	double top = ll[0];
	int best = 0;
#pragma unroll
	for (int g = 1; g < 10; g++) if (ll[g] > top) { top = ll[g]; best = g; }
	double sum = 0.0;
#pragma unroll
	for (int g = 0; g < 10; g++) sum += exp(ll[g] - top);
	sum = log(sum);
#pragma unroll
	for (int g = 0; g < 10; g++) prob[g] = (ll[g] - top - sum) / kLn10;
	return best;
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
	return log(z) / kLn10;
}

// ---------------------------------------------------------------------------------------------------------------
// Per-site body of call_thread (src/call_genotypes.c:43-115): summarise, model, strand bias; writes the 200-byte
// gt_meth image as 25 eight-byte words into `rec` (a row of the CTA's staging tile in shared memory).
// Returns false for a site with no counted base (record zeroed, the caller sets skip).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool call_site(const SiteCounts &s, int rf, const DevConst *__restrict__ dc,
		const double (*__restrict__ qp)[4], const double *__restrict__ lfact_tab, uint64_t *rec) {
	if (!s.n) {
#pragma unroll
		for (int i = 0; i < 25; i++) rec[i] = 0;
		return false;
	}
	uint32_t tot[8];
	int qual[8];
	float tq = 0.0f;
#pragma unroll
	for (int j = 0; j < 8; j++) {
		tot[j] = s.cnt[0][j] + s.cnt[1][j];
		const float nn = (float)tot[j];
		if (nn > 0) {
			tq += s.qsum[j];
			// float divide, double add, narrowed to float, floorf  (:50)
			qual[j] = (int)floorf((float)(0.5 + (double)(s.qsum[j] / nn)));
		} else qual[j] = 0;
	}
	const int aq = (int)floorf((float)(0.5 + (double)(tq / (float)s.n)));
	const int mq = (int)(0.5 + sqrt((double)(s.mapq2 / (float)s.n)));
	double prob[10];
	const int best = genotype_model(tot, qual, rf, dc, qp, prob);
	const double fs = strand_bias(s, best, lfact_tab);
#pragma unroll
	for (int j = 0; j < 8; j++) rec[j] = tot[j];
#pragma unroll
	for (int j = 0; j < 4; j++) rec[8 + j] = (uint64_t)(uint32_t)qual[2 * j] | ((uint64_t)(uint32_t)qual[2 * j + 1] << 32);
#pragma unroll
	for (int g = 0; g < 10; g++) rec[12 + g] = (uint64_t)__double_as_longlong(prob[g]);
	rec[22] = (uint64_t)__double_as_longlong(fs);
	rec[23] = (uint64_t)(uint32_t)mq | ((uint64_t)(uint32_t)aq << 32);
	rec[24] = (uint64_t)best;
	return true;
}

}  // namespace bsgpu
