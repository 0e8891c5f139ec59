/*
 * bs_oracle.c -- TEST INFRASTRUCTURE ONLY (see bs_oracle.h for the rules and for how parity is pinned).
 *
 * Restates, function by function, what the bs_call 2.1.7 hot path computes.  Citations are to files under
 * /root/reference.  The arithmetic (operand order, float/double mixing, integer widths) follows the reference
 * because results must be bit-identical; the code structure is this repo's own (table-driven model, explicit
 * out-of-place normalisation) and is built with the reference's flags (-O3, no -march, no -ffast-math) so that
 * the compiler never contracts a*b+c into an FMA.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "bs_oracle.h"

#define BSO_MAX_QUAL 43
#define BSO_FLT_QUAL 63
#define BSO_LN10 (2.30258509299404568402)
#define BSO_LFACT_N 256

void bso_default_params(bso_params *p) {
	/* include/bs_call.h:14-18, 26 */
	memset(p, 0, sizeof(*p));
	p->under_conv = 0.01;
	p->over_conv = 0.05;
	p->ref_bias = 2.0;
	p->min_qual = 20;
}

/* ------------------------------------------------------------------------------------------------
 * Quality -> error-probability table.  src/genotype_model.c:10-21
 * ------------------------------------------------------------------------------------------------ */
typedef struct { double e, k, ln_k, ln_k_half, ln_k_one; } qp_t;
static qp_t qp_tab[BSO_MAX_QUAL + 1];
static double lfact_tab[BSO_LFACT_N];
static pthread_once_t tab_once = PTHREAD_ONCE_INIT;

static void build_tables(void) {
	for (int q = 0; q <= BSO_MAX_QUAL; q++) {
		qp_t *t = qp_tab + q;
		double e = exp(-.1 * (double)q * BSO_LN10);
		t->e = e > .5 ? .5 : e;
		t->k = t->e / (3.0 - 4.0 * t->e);
		t->ln_k = log(t->k);
		t->ln_k_half = log(0.5 + t->k);
		t->ln_k_one = log(1.0 + t->k);
	}
	/* src/stats_utils.c:14-21 */
	double acc = 0.0;
	lfact_tab[0] = lfact_tab[1] = 0.0;
	for (int i = 2; i < BSO_LFACT_N; i++) {
		acc += log((double)i);
		lfact_tab[i] = acc;
	}
}

void bso_qprob_table(double *out) {
	pthread_once(&tab_once, build_tables);
	memcpy(out, qp_tab, sizeof(qp_tab));
}

void bso_lfact_table(double *out) {
	pthread_once(&tab_once, build_tables);
	memcpy(out, lfact_tab, sizeof(lfact_tab));
}

/* ------------------------------------------------------------------------------------------------
 * Closed-form ML conversion fraction.  src/genotype_model.c:23-42
 * a = count of the "C-like" observation, b = count of the "T-like" one, ka/kb their error terms.
 * ------------------------------------------------------------------------------------------------ */
static void conv_ml(double a, double b, double ka, double kb, double l, double t, double *Z) {
	const double lpt = l + t, lmt = l - t;
	const double d = (a + b) * lmt;
	/* numerators for (w,p) = (1,1), (1,1/2), (1/2,1) */
	const double num[3] = {
		a * (lpt + 2.0 * kb) - b * (2.0 - lpt + 2.0 * ka),
		a * (2.0 + lpt + 4.0 * kb) - b * (2.0 - lpt + 4.0 * ka),
		a * (lpt + 4.0 * kb) - b * (2.0 - lpt + 4.0 * ka)
	};
	for (int i = 0; i < 3; i++) {
		double s = num[i] / d;
		if (s < -1.0) s = -1.0;
		else if (s > 1.0) s = 1.0;
		Z[i] = 0.5 * (lmt * s + 2.0 - lpt);
	}
}

/* genotype g -> its two alleles (A0 C1 G2 T3); order AA AC AG AT CC CG CT GG GT TT (src/init_param.c:16) */
static const int8_t gt_al[10][2] = {
	{0,0},{0,1},{0,2},{0,3},{1,1},{1,2},{1,3},{2,2},{2,3},{3,3}
};

/* ------------------------------------------------------------------------------------------------
 * 10-genotype log-likelihoods with reference prior.  src/genotype_model.c:44-246
 * Every ll[g] receives: prior, then one addend per non-empty class in class order 0..7 -- the same
 * sequence of additions the reference performs, so sums are bit-identical.
 * ------------------------------------------------------------------------------------------------ */
void bso_calc_gt_prob(bso_gt_meth *gt, const bso_params *p, int rf) {
	pthread_once(&tab_once, build_tables);
	qp_t qp[8];
	double n[8], ll[10];
	for (int j = 0; j < 8; j++) {
		qp[j] = qp_tab[gt->qual[j]];
		n[j] = (double)gt->counts[j];
	}
	const double l = 1.0 - p->under_conv, t = p->over_conv;
	for (int g = 0; g < 10; g++) ll[g] = 0.0;
	if (rf >= 1 && rf <= 4) {                       /* :87-108 */
		const double lrb = log(p->ref_bias), lrb1 = log(0.5 * (1.0 + p->ref_bias));
		const int b = rf - 1;
		for (int g = 0; g < 10; g++) {
			const int hits = (gt_al[g][0] == b) + (gt_al[g][1] == b);
			if (hits == 2) ll[g] = lrb;
			else if (hits == 1) ll[g] = lrb1;
		}
	}
	/* classes 0-3: non-informative A,C,G,T   :109-164 */
	for (int j = 0; j < 4; j++) {
		if (!n[j]) continue;
		const double v2 = n[j] * qp[j].ln_k_one, v1 = n[j] * qp[j].ln_k_half, v0 = n[j] * qp[j].ln_k;
		for (int g = 0; g < 10; g++) {
			const int hits = (gt_al[g][0] == j) + (gt_al[g][1] == j);
			ll[g] += hits == 2 ? v2 : (hits == 1 ? v1 : v0);
		}
	}
	/* :165-171  Z[0..2] from informative C (5) / T (7); Z[3..5] from informative G (6) / A (4) */
	double Z[6] = { -1.0, -1.0, -1.0, -1.0, -1.0, -1.0 };
	if (n[5] + n[7] > 0.0) conv_ml(n[5], n[7], qp[5].k, qp[7].k, l, t, Z);
	if (n[4] + n[6] > 0.0) conv_ml(n[6], n[4], qp[6].k, qp[4].k, l, t, Z + 3);
	double term[10];
	if (n[4]) {                                     /* informative A  :173-187 */
		const double k = qp[4].k;
		const double one = n[4] * qp[4].ln_k_one, half = n[4] * qp[4].ln_k_half, kk = n[4] * qp[4].ln_k;
		const double mix = log(0.5 * (1.0 - Z[5]) + k) * n[4];
		term[0] = one;
		term[2] = log(1.0 - 0.5 * Z[4] + k) * n[4];
		term[7] = log(1.0 - Z[3] + k) * n[4];
		term[5] = term[8] = mix;
		term[1] = term[3] = half;
		term[4] = term[6] = term[9] = kk;
		for (int g = 0; g < 10; g++) ll[g] += term[g];
	}
	if (n[5]) {                                     /* informative C  :188-201 */
		const double k = qp[5].k;
		const double kk = n[5] * qp[5].ln_k;
		const double mix = log(0.5 * Z[2] + k) * n[5];
		term[4] = log(Z[0] + k) * n[5];
		term[1] = term[5] = mix;
		term[6] = log(0.5 * Z[1] + k) * n[5];
		term[0] = term[2] = term[3] = term[7] = term[8] = term[9] = kk;
		for (int g = 0; g < 10; g++) ll[g] += term[g];
	}
	if (n[6]) {                                     /* informative G  :202-215 */
		const double k = qp[6].k;
		const double kk = n[6] * qp[6].ln_k;
		const double mix = log(0.5 * Z[5] + k) * n[6];
		term[7] = log(Z[3] + k) * n[6];
		term[5] = term[8] = mix;
		term[2] = log(0.5 * Z[4] + k) * n[6];
		term[0] = term[1] = term[3] = term[4] = term[6] = term[9] = kk;
		for (int g = 0; g < 10; g++) ll[g] += term[g];
	}
	if (n[7]) {                                     /* informative T  :216-230 */
		const double k = qp[7].k;
		const double one = n[7] * qp[7].ln_k_one, half = n[7] * qp[7].ln_k_half, kk = n[7] * qp[7].ln_k;
		const double mix = log(0.5 * (1.0 - Z[2]) + k) * n[7];
		term[9] = one;
		term[4] = log(1.0 - Z[0] + k) * n[7];
		term[6] = log(1.0 - 0.5 * Z[1] + k) * n[7];
		term[1] = term[5] = mix;
		term[3] = term[8] = half;
		term[0] = term[2] = term[7] = kk;
		for (int g = 0; g < 10; g++) ll[g] += term[g];
	}
	/* first strict maximum, then log-sum-exp normalisation to log10 posteriors  :231-245 */
	int best = 0;
	for (int g = 1; g < 10; g++) if (ll[g] > ll[best]) best = g;
	const double top = ll[best];
	gt->max_gt = (uint8_t)best;
	double sum = 0.0;
	for (int g = 0; g < 10; g++) sum += exp(ll[g] - top);
	sum = log(sum);
	for (int g = 0; g < 10; g++) gt->gt_prob[g] = (ll[g] - top - sum) / BSO_LN10;
}

/* ------------------------------------------------------------------------------------------------
 * Two-sided Fisher exact test on a 2x2 table.  src/stats_utils.c:25-91; lfact2 include/bs_call.h:335
 * ------------------------------------------------------------------------------------------------ */
static inline double lfact(int x) { return x < BSO_LFACT_N ? lfact_tab[x] : lgamma((double)(x + 1)); }

static inline double table_prob(double knst, const int c[4]) {
	return exp(knst - lfact(c[0]) - lfact(c[1]) - lfact(c[2]) - lfact(c[3]));
}

/* walk `steps` tables further away from independence: `dn` is the diagonal being decreased, `up` the one
 * being increased; *l is the running table probability and each step's value is added to *p */
static inline void tail_walk(int dn0, int dn1, int up0, int up1, int steps, double *l, double *p) {
	for (int i = 0; i < steps; i++) {
		*l *= (double)((dn0 - i) * (dn1 - i)) / (double)((up0 + i + 1) * (up1 + i + 1));
		*p += *l;
	}
}

double bso_fisher(const int c_in[4]) {
	pthread_once(&tab_once, build_tables);
	int c[4] = { c_in[0], c_in[1], c_in[2], c_in[3] };
	const int row0 = c[0] + c[1], row1 = c[2] + c[3], col0 = c[0] + c[2], col1 = c[1] + c[3];
	const int n = row0 + row1;
	if (n == 0) return 1.0;
	const double delta = (double)c[0] - (double)(row0 * col0) / (double)n;
	const double knst = lfact(col0) + lfact(col1) + lfact(row0) + lfact(row1) - lfact(n);
	double l = table_prob(knst, c);
	double p = l;
	const int lead = c[0] < c[3] ? c[0] : c[3];       /* min of leading diagonal */
	const int cntr = c[1] < c[2] ? c[1] : c[2];       /* min of counter diagonal */
	if (delta > 0.0) {
		tail_walk(c[1], c[2], c[0], c[3], cntr, &l, &p);
		int k = (int)ceil(2.0 * delta);
		if (k <= lead) {
			c[0] -= k; c[3] -= k; c[1] += k; c[2] += k;
			l = table_prob(knst, c);
			p += l;
			tail_walk(c[0], c[3], c[1], c[2], lead - k, &l, &p);
		}
	} else {
		tail_walk(c[0], c[3], c[1], c[2], lead, &l, &p);
		int k = (int)ceil(-2.0 * delta);
		if (!k) k = 1;
		if (k <= cntr) {
			c[0] += k; c[3] += k; c[1] -= k; c[2] -= k;
			l = table_prob(knst, c);
			p += l;
			tail_walk(c[1], c[2], c[0], c[3], cntr - k, &l, &p);
		}
	}
	return p;
}

/* ------------------------------------------------------------------------------------------------
 * Per-site body of call_thread.  src/call_genotypes.c:43-115
 * ------------------------------------------------------------------------------------------------ */
void bso_summarise(const bso_pileup *tp, bso_gt_meth *tg) {
	/* :45-59.  `0.5 + float` is evaluated in double, then narrowed to float for floorf */
	float tot = 0.0f;
	for (int j = 0; j < 8; j++) {
		const uint32_t cj = tp->counts[0][j] + tp->counts[1][j];
		const float nn = (float)cj;
		if (nn > 0) {
			tot += tp->quality[j];
			tg->qual[j] = (int)floorf(0.5 + tp->quality[j] / nn);
		} else tg->qual[j] = 0;
		tg->counts[j] = cj;
	}
	tg->aq = (int)floorf(0.5 + tot / (float)tp->n);
	tg->mq = (int)(0.5 + sqrt(tp->mapq2 / (float)tp->n));
}

/* allele x strand table for the het genotypes, :62-104.  Class sets per allele; the GT case reproduces the
 * reference's use of counts[0][6] in the ori-1 cell (line 98). Returns 0 when max_gt is not a het. */
int bso_strand_table(const bso_pileup *tp, int max_gt, int ftab[4]) {
	static const uint8_t het_sets[10][2] = {
		/* bit j set = class j belongs to the allele */
		{0, 0},
		{0x11, 0xa2},   /* AC: {0,4} | {1,5,7} */
		{0x01, 0x44},   /* AG: {0}   | {2,6}   */
		{0x11, 0x88},   /* AT: {0,4} | {3,7}   */
		{0, 0},
		{0xa2, 0x54},   /* CG: {1,5,7} | {2,4,6} */
		{0x22, 0x08},   /* CT: {1,5} | {3} */
		{0, 0},
		{0x54, 0x88},   /* GT: {2,4,6} | {3,7} */
		{0, 0}
	};
	if (max_gt < 0 || max_gt > 9 || !het_sets[max_gt][0]) return 0;
	for (int o = 0; o < 2; o++) for (int a = 0; a < 2; a++) {
		int s = 0;
		for (int j = 0; j < 8; j++) if (het_sets[max_gt][a] >> j & 1) s += (int)tp->counts[o][j];
		ftab[2 * o + a] = s;
	}
	if (max_gt == 8) ftab[2] = (int)(tp->counts[1][2] + tp->counts[1][4] + tp->counts[0][6]);
	return 1;
}

void bso_call_site(const bso_pileup *tp, int rf, const bso_params *p, bso_gt_meth *out, uint8_t *skip) {
	memset(out, 0, sizeof(*out));
	if (!tp->n) { *skip = 1; return; }
	bso_summarise(tp, out);
	bso_calc_gt_prob(out, p, rf);
	double fs = 0.0;
	int ftab[4];
	if (bso_strand_table(tp, out->max_gt, ftab)) {
		double z = bso_fisher(ftab);
		if (z < 1.0e-20) z = 1.0e-20;
		fs = log(z) / BSO_LN10;
	}
	out->fisher_strand = fs;
	*skip = 0;
}

typedef struct {
	const bso_pileup *tp; const uint8_t *ref; size_t n; const bso_params *p;
	bso_gt_meth *out; uint8_t *skip; size_t first, step;
} site_job;

static void *site_worker(void *arg) {
	site_job *j = arg;
	/* thread i takes sites i, i+step, ... like the reference (src/call_genotypes.c:259-270) */
	for (size_t i = j->first; i < j->n; i += j->step) bso_call_site(j->tp + i, j->ref[i], j->p, j->out + i, j->skip + i);
	return NULL;
}

void bso_call_sites(const bso_pileup *tp, const uint8_t *ref, size_t n, const bso_params *p,
		bso_gt_meth *out, uint8_t *skip, int nthreads) {
	pthread_once(&tab_once, build_tables);
	if (nthreads < 1) nthreads = 1;
	site_job *jobs = malloc(sizeof(site_job) * nthreads);
	pthread_t *thr = malloc(sizeof(pthread_t) * nthreads);
	for (int i = 0; i < nthreads; i++) {
		jobs[i] = (site_job){ tp, ref, n, p, out, skip, (size_t)i, (size_t)nthreads };
		if (i) pthread_create(thr + i, NULL, site_worker, jobs + i);
	}
	site_worker(jobs);
	for (int i = 1; i < nthreads; i++) pthread_join(thr[i], NULL);
	free(jobs);
	free(thr);
}

/* ------------------------------------------------------------------------------------------------
 * Pileup over normalised templates.  src/call_genotypes.c:180-226, class table :17-19
 * ------------------------------------------------------------------------------------------------ */
static inline int counted_anywhere(uint8_t b) { const uint8_t q = b >> 2; return q > 0 && q != BSO_FLT_QUAL; }

void bso_pileup_block(const bso_template *t, size_t n, const uint8_t *bases, uint32_t x, uint32_t y,
		const bso_params *p, bso_pileup *out) {
	/* class = base, +4 when the base is methylation-informative on this bisulfite strand */
	static const uint8_t cls[3][4] = { {0,1,2,3}, {0,5,2,7}, {4,1,6,3} };
	const uint32_t sz = y - x + 1;
	memset(out, 0, sizeof(bso_pileup) * sz);
	for (size_t it = 0; it < n; it++, t++) {
		int ori = t->orientation;
		for (int k = 0; k < 2; k++) {
			if (!t->present[k] || !t->read_len[k]) continue;
			const uint8_t *sp = bases + t->read_off[k];
			const uint32_t rl = t->read_len[k];
			uint32_t first = 0, last = rl;
			while (first < rl && !counted_anywhere(sp[first])) first++;
			if (first == rl) continue;                   /* nothing usable: no strand flip either */
			while (!counted_anywhere(sp[last - 1])) last--;
			const float mq2 = (float)(t->mapq[k] * t->mapq[k]);
			uint32_t pos = (k ? t->reverse_position : t->forward_position) + first;
			for (uint32_t j = first; j < last && pos <= y; j++, pos++) {
				const uint8_t q = sp[j] >> 2;
				if (q >= p->min_qual && q != BSO_FLT_QUAL) {
					bso_pileup *s = out + (pos - x);
					const int c = cls[t->bs_strand][sp[j] & 3];
					s->n++;
					s->quality[c] += (float)q;
					s->mapq2 += mq2;
					s->counts[ori][c]++;
				}
			}
			ori ^= 1;
		}
	}
}

/* ------------------------------------------------------------------------------------------------
 * Template normalisation.
 *   trim_read          src/read_utils.c:13-26
 *   trim_soft_clips    src/al_utils.c:122-162
 *   handle_overlap     src/al_utils.c:164-318
 *   indel normalise    src/process_template.c:66-110
 * Works on a private mutable copy of one template.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
	uint32_t pos[2];              /* [0] forward_position, [1] reverse_position */
	uint32_t span[2];
	uint8_t *rd[2];
	uint32_t len[2];
	int present[2];
	bso_misms *mm[2];
	uint32_t nmm[2];
	uint32_t tl[2], tr[2];        /* trim_left / trim_right of src/process_template.c:42-45 */
	int *orig[2];                 /* original read position of every byte (profile only) */
} tmpl_w;

/* ---- --report-file side channels gathered on this path (bs_stats, include/bs_call.h:124-146) ---- */
#define PROF_ALLOC 70000
static struct {
	int on;
	uint64_t (*mem)[4];           /* stats->meth_profile->memory */
	size_t used;
	uint64_t base_filter[5], filter_cts[15], filter_bases[15];
} prof;

void bso_profile_enable(int on) {
	if (on && !prof.mem) prof.mem = calloc(PROF_ALLOC, sizeof(*prof.mem));
	prof.on = on;
}
void bso_profile_reset(void) {
	if (prof.mem) memset(prof.mem, 0, PROF_ALLOC * sizeof(*prof.mem));
	prof.used = 0;
	memset(prof.base_filter, 0, sizeof(prof.base_filter));
	memset(prof.filter_cts, 0, sizeof(prof.filter_cts));
	memset(prof.filter_bases, 0, sizeof(prof.filter_bases));
}
int bso_profile_is_on(void) { return prof.on; }
void bso_profile_tally(int reason_cts, uint64_t cts, int reason_bases, uint64_t bases) {
	if (!prof.on) return;
	prof.filter_cts[reason_cts] += cts;
	prof.filter_bases[reason_bases] += bases;
}
void bso_profile_read(bso_profile *out) {
	memset(out, 0, sizeof(*out));
	out->used = (uint32_t)prof.used;
	for (size_t i = 0; i < prof.used && i < BSO_PROFILE_MAX; i++) memcpy(out->conv_cts[i], prof.mem[i], sizeof(prof.mem[i]));
	memcpy(out->base_filter, prof.base_filter, sizeof(prof.base_filter));
	memcpy(out->filter_cts, prof.filter_cts, sizeof(prof.filter_cts));
	memcpy(out->filter_bases, prof.filter_bases, sizeof(prof.filter_bases));
}

/* meth_profile() for one normalised template (src/meth_profile.c:48-76).  refcodes[0] is position x. */
static void profile_template(const tmpl_w *w, int bs_strand, int max_pos, const uint8_t *refcodes, uint32_t x) {
	/* [prev][cur] -> 4: a C not followed by G or N (looked up one step ahead), 8: a G not preceded by C or N */
	static const uint8_t ctx_tab[5][5] = {
		{0, 0, 0, 0, 0}, {0, 0, 0, 8, 0}, {0, 4, 4, 0, 4}, {0, 0, 0, 8, 0}, {0, 0, 0, 8, 0} };
	/* flt_tab rows (src/init_param.c:57-69): per strand, per base, for MIN_QUAL <= q < FLT_QUAL; low two bits pick the counter */
	static const uint8_t row[3][4] = { {11, 6, 10, 7}, {11, 4, 10, 5}, {9, 6, 8, 7} };
	if ((size_t)max_pos + 1 > prof.used) {
		/* gt_vector_reserve(.., zero_mem = true) clears from the old `used` to the end of the allocation
		 * (gt/src/gt_vector.c:34-37), then used is raised */
		memset(prof.mem + prof.used, 0, (PROF_ALLOC - prof.used) * sizeof(*prof.mem));
		prof.used = (size_t)max_pos + 1;
	}
	uint64_t (*mc)[4] = prof.mem + 1;
	for (int k = 0; k < 2; k++) {
		if (!w->present[k] || !w->len[k]) continue;
		const uint32_t pos = w->pos[k];
		const uint8_t *rf = refcodes + (pos - x);
		uint32_t prev = 0, cur = 0;
		if (pos > x) { prev = rf[-1]; cur = *rf++; }
		uint32_t mask = prev < 5 && cur < 5 ? ctx_tab[prev][cur] : 0;
		for (uint32_t j = 0; j < w->len[k]; j++) {
			const uint8_t b = w->rd[k][j];
			const uint32_t q = b >> 2;
			const uint32_t xx = (q >= 20 && q < BSO_FLT_QUAL) ? row[bs_strand][b & 3] : 0;      /* MIN_QUAL, include/bs_call.h:28 */
			const uint32_t carried = (xx & mask) >> 1;
			prev = cur; cur = *rf++;
			mask = prev < 5 && cur < 5 ? ctx_tab[prev][cur] : 0;
			mc[w->orig[k][j]][xx & 3] += (((xx & mask) | carried) >> 2) & 1;
		}
	}
}

static void mark_trimmed(uint8_t *sp, uint32_t rl, uint32_t left, uint32_t right) {
	for (uint32_t i = 0; i < left && i < rl; i++) sp[i] = (sp[i] & 3) | (BSO_FLT_QUAL << 2);
	/* the base bits of a right-trimmed byte come from the mirrored left index (reference quirk, :22) */
	for (uint32_t i = 0; i < right && i < rl; i++) sp[rl - i - 1] = (sp[i] & 3) | (BSO_FLT_QUAL << 2);
}

static void cut_left(tmpl_w *w, int k, uint32_t l) {
	if (!l) return;
	if (l >= w->len[k]) { w->len[k] = 0; return; }
	w->len[k] -= l;
	memmove(w->rd[k], w->rd[k] + l, w->len[k]);
}

static void cut_right(tmpl_w *w, int k, uint32_t l) {
	if (!l) return;
	if (l >= w->len[k]) w->len[k] = 0;
	else w->len[k] -= l;
}

static int strip_soft_clips(tmpl_w *w) {
	for (int k = 0; k < 2; k++) {
		if (!w->present[k] || !w->len[k]) continue;
		const uint32_t rl = w->len[k], n0 = w->nmm[k];
		uint32_t kept = 0, shift = 0, nclip = 0;
		for (uint32_t z = 0; z < n0; z++) {
			bso_misms m = w->mm[k][z];
			if (m.type == BSO_SOFT) {
				if (z && z != n0 - 1) return -1;
				nclip++;
				if (!m.position) {
					if (m.size >= rl) return -1;
					shift = m.size;
					cut_left(w, k, shift);
					w->tl[k] = shift;
					if (prof.on) prof.base_filter[2] += shift;
				} else {
					if (m.position + m.size != rl) return -1;
					cut_right(w, k, m.size);
					w->tr[k] = m.size;
					if (prof.on) prof.base_filter[2] += m.size;
				}
			} else {
				if (nclip) m.position -= shift;
				w->mm[k][kept++] = m;
			}
		}
		w->nmm[k] = kept;
	}
	return 0;
}

static uint32_t mean_untrimmed_qual(const uint8_t *sp, uint32_t rl) {
	uint32_t tot = 0, n = 0;
	for (uint32_t i = 0; i < rl; i++) {
		const uint8_t q = sp[i] >> 2;
		if (q != BSO_FLT_QUAL) { tot += q; n++; }
	}
	return n ? tot / n : 0;
}

/* drop the first z entries of the event list of mate k */
static void drop_events(tmpl_w *w, int k, uint32_t z) {
	if (z) memmove(w->mm[k], w->mm[k] + z, sizeof(bso_misms) * (w->nmm[k] - z));
	w->nmm[k] -= z;
}

/* returns -1 when nothing is shared, else (mate that was cut) | (cut at its right end) << 1 */
static int overlap_cut(tmpl_w *w) {
	if (!(w->present[0] && w->len[0] && w->present[1] && w->len[1])) return -1;
	const int rev = !(w->pos[0] <= w->pos[1]);
	const int32_t overlap = rev ? (int32_t)(w->span[1] + w->pos[1] - w->pos[0]) : (int32_t)(w->span[0] - w->pos[1] + w->pos[0]);
	if (!(w->pos[0] + w->span[0] >= w->pos[1])) return -1;
	/* which mate loses the shared part: the one with the shorter reference span, else the lower mean quality,
	 * else mate 0  (:185-202) */
	int tr;
	if (w->span[0] > w->span[1]) tr = 1;
	else if (w->span[0] < w->span[1]) tr = 0;
	else tr = mean_untrimmed_qual(w->rd[0], w->len[0]) <= mean_untrimmed_qual(w->rd[1], w->len[1]) ? 0 : 1;
	const int at_right = (rev == tr);             /* cut the right end of the left-hand mate, else the left end of the right-hand one */
	if (!at_right) w->pos[tr] += (uint32_t)overlap; /* :204-207 */
	const uint32_t rl = w->len[tr];
	uint32_t nmm = w->nmm[tr];
	bso_misms *mm = w->mm[tr];
	if (!nmm) {
		if (at_right) cut_right(w, tr, (uint32_t)overlap);
		else cut_left(w, tr, (uint32_t)overlap);
		return tr | at_right << 1;
	}
	int done = 0;
	int64_t adj = 0;
	if (at_right) {                                /* :220-245 */
		const uint32_t keep = w->span[tr] - (uint32_t)overlap;
		for (uint32_t z = 0; z < nmm; z++) {
			bso_misms *m = mm + z;
			if ((int64_t)m->position + adj >= (int64_t)keep) {
				const int64_t trim = (int64_t)(uint32_t)(rl - keep) + adj;
				cut_right(w, tr, (uint32_t)trim);
				w->nmm[tr] = z;
				done = 1;
				break;
			}
			if (m->type == BSO_INS) {
				if ((int64_t)m->position + adj + (int64_t)m->size >= (int64_t)keep) {
					const int64_t trim = (int64_t)(uint32_t)(rl - m->position);
					m->size = (uint32_t)((int64_t)keep - ((int64_t)m->position + adj));
					cut_right(w, tr, (uint32_t)trim);
					w->nmm[tr] = z + 1;
					done = 1;
					break;
				}
				adj += m->size;
			} else if (m->type == BSO_DEL) adj -= m->size;
		}
		if (!done) cut_right(w, tr, (uint32_t)overlap);
	} else {                                       /* :246-303 */
		const uint32_t cut = (uint32_t)overlap;
		for (uint32_t z = 0; z < nmm; z++) {
			bso_misms *m = mm + z;
			if ((int64_t)m->position + adj >= (int64_t)cut) {
				const uint32_t trim = (uint32_t)((int64_t)overlap - adj);
				cut_left(w, tr, trim);
				for (uint32_t z1 = z; z1 < nmm; z1++) mm[z1].position -= trim;
				drop_events(w, tr, z);
				done = 1;
				break;
			}
			if (m->type == BSO_INS) {
				if ((int64_t)m->position + adj + (int64_t)m->size >= (int64_t)cut) {
					m->size = (uint32_t)((int64_t)m->position + (int64_t)m->size + adj - (int64_t)cut);
					const uint32_t trim = m->position;
					cut_left(w, tr, trim);
					const uint32_t z2 = m->size ? z : z + 1;
					for (uint32_t z1 = z2; z1 < nmm; z1++) mm[z1].position -= trim;
					drop_events(w, tr, z2);
					done = 1;
					break;
				}
				adj += m->size;
			} else if (m->type == BSO_DEL) adj -= m->size;
		}
		if (!done) {
			cut_left(w, tr, (uint32_t)((int64_t)overlap - adj));
			w->nmm[tr] = 0;
		}
	}
	return tr | at_right << 1;
}

/* handle_overlap including its bookkeeping (src/al_utils.c:304-314): what was cut is added to the base_overlap tally
 * and to trim_right (cut at the right end) or trim_left of the mate that lost it */
static void resolve_overlap(tmpl_w *w) {
	const uint32_t before[2] = { w->len[0], w->len[1] };
	const int c = overlap_cut(w);
	if (c < 0) return;
	const int tr = c & 1;
	if (prof.on) prof.base_filter[3] += (before[0] - w->len[0]) + (before[1] - w->len[1]);
	if (c >> 1) w->tr[tr] += before[tr] - w->len[tr];
	else w->tl[tr] += before[tr] - w->len[tr];
}

/* rewrite mate k into reference coordinates: zero-fill INS (reference bases missing from the read), drop DEL */
static uint32_t to_ref_coords(tmpl_w *w, int k) {
	uint8_t *sp = w->rd[k];
	uint32_t used = w->len[k];
	uint32_t adj = 0;
	for (uint32_t z = 0; z < w->nmm[k]; z++) {
		const bso_misms *m = w->mm[k] + z;
		const uint32_t at = m->position + adj;
		int *og = w->orig[k];
		if (m->type == BSO_INS) {
			memmove(sp + at + m->size, sp + at, used - at);
			memset(sp + at, 0, m->size);
			if (og) {
				memmove(og + at + m->size, og + at, sizeof(int) * (used - at));
				for (uint32_t i = 0; i < m->size; i++) og[at + i] = -1;
			}
			adj += m->size;
			used += m->size;
		} else if (m->type == BSO_DEL) {
			memmove(sp + at, sp + at + m->size, used - at - m->size);
			if (og) memmove(og + at, og + at + m->size, sizeof(int) * (used - at - m->size));
			adj -= m->size;
			used -= m->size;
		}
	}
	return used;
}

/* the block's reference window for the profile: set by bso_process_block around its call of bso_normalise_block */
static const uint8_t *prof_ref;
static uint32_t prof_x;

int bso_normalise_block(const bso_template *t, size_t n, const uint8_t *bases, const bso_misms *mm,
		const bso_params *p, bso_template *out_t, uint8_t *out_bases, size_t out_cap, size_t *out_used) {
	size_t off = 0;
	int ret = 0;
	for (size_t i = 0; i < n; i++, t++) {
		tmpl_w w;
		memset(&w, 0, sizeof(w));
		w.pos[0] = t->forward_position;
		w.pos[1] = t->reverse_position;
		for (int k = 0; k < 2; k++) {
			w.span[k] = t->reference_span[k];
			w.present[k] = t->present[k];
			w.len[k] = t->present[k] ? t->read_len[k] : 0;
			w.nmm[k] = t->mm_n[k];
			uint32_t grow = 0;
			for (uint32_t z = 0; z < w.nmm[k]; z++) if (mm[t->mm_off[k] + z].type == BSO_INS) grow += mm[t->mm_off[k] + z].size;
			w.rd[k] = malloc((size_t)w.len[k] + grow + 8);
			if (w.len[k]) memcpy(w.rd[k], bases + t->read_off[k], w.len[k]);
			w.mm[k] = malloc(sizeof(bso_misms) * (w.nmm[k] + 1));
			if (w.nmm[k]) memcpy(w.mm[k], mm + t->mm_off[k], sizeof(bso_misms) * w.nmm[k]);
		}
		/* -L/-R apply to read 1 / read 2 in BAM orientation; mate slot [0] holds R1 iff the template is FORWARD
		 * (src/process_template.c:36-41) */
		const int msk = t->orientation == 0 ? 0 : 1;
		for (int r = 0; r < 2; r++) {
			const int k = r ^ msk;
			if ((p->left_trim[r] || p->right_trim[r]) && w.present[k] && w.len[k])
				mark_trimmed(w.rd[k], w.len[k], p->left_trim[r], p->right_trim[r]);
		}
		if (strip_soft_clips(&w) < 0) ret = -1;
		else {
			resolve_overlap(&w);
			int max_pos = 0;
			for (int k = 0; k < 2; k++) {
				if (!w.present[k]) continue;
				if (prof.on) {
					/* tallies over the bytes that are left (src/process_template.c:52-62) and the map back to positions in
					 * the original read: slot 0 counts up from trim_left, slot 1 down from rdl + trim_right - 1 (:80-91) */
					const int rdl = (int)w.len[k];
					for (int j = 0; j < rdl; j++) {
						const uint8_t q = w.rd[k][j] >> 2;
						if (q == BSO_FLT_QUAL) prof.base_filter[1]++;
						else if (q < p->min_qual) prof.base_filter[4]++;
						else prof.base_filter[0]++;
					}
					prof.filter_cts[0]++;
					prof.filter_bases[0] += (uint64_t)rdl;
					uint32_t grow = 0;
					for (uint32_t z = 0; z < w.nmm[k]; z++) if (w.mm[k][z].type == BSO_INS) grow += w.mm[k][z].size;
					w.orig[k] = malloc(sizeof(int) * ((size_t)rdl + grow + 8));
					int mpos;
					if (k) {
						const int top = rdl + (int)w.tr[k] - 1;
						for (int j = 0; j < rdl; j++) w.orig[k][j] = top - j;
						mpos = top;
					} else {
						const int first = (int)w.tl[k];
						for (int j = 0; j < rdl; j++) w.orig[k][j] = first + j;
						mpos = first + rdl;
					}
					if (mpos > max_pos) max_pos = mpos;
				}
				w.len[k] = to_ref_coords(&w, k);
			}
			if (prof.on && prof_ref) profile_template(&w, t->bs_strand, max_pos, prof_ref, prof_x);
			for (int k = 0; k < 2; k++) { free(w.orig[k]); w.orig[k] = NULL; }
		}
		bso_template *o = out_t + i;
		memset(o, 0, sizeof(*o));
		o->forward_position = w.pos[0];
		o->reverse_position = w.pos[1];
		o->orientation = t->orientation;
		o->bs_strand = t->bs_strand;
		for (int k = 0; k < 2; k++) {
			o->reference_span[k] = w.span[k];
			o->mapq[k] = t->mapq[k];
			o->present[k] = (uint8_t)w.present[k];
			o->read_off[k] = (uint32_t)off;
			o->read_len[k] = w.len[k];
			if (w.len[k]) {
				if (off + w.len[k] > out_cap) ret = ret ? ret : -3;
				else memcpy(out_bases + off, w.rd[k], w.len[k]);
				off += w.len[k];
			}
			free(w.rd[k]);
			free(w.mm[k]);
		}
		if (ret) break;
	}
	if (out_used) *out_used = off;
	return ret;
}

int bso_process_block(const bso_template *t, size_t n, const uint8_t *bases, const bso_misms *mm,
		const uint8_t *refcodes, uint32_t y, const bso_params *p,
		uint32_t *x_out, bso_pileup *pile_out, bso_gt_vcf *vcf_out) {
	if (!n) return -1;
	/* window: two positions before the first template's start (src/process_template.c:24-28) */
	uint32_t x = t[0].forward_position ? t[0].forward_position : t[0].reverse_position;
	x = x > 2 ? x - 2 : 1;
	if (y < x) return -1;
	const uint32_t sz = y - x + 1;
	size_t cap = 0;
	for (size_t i = 0; i < n; i++) for (int k = 0; k < 2; k++) {
		cap += t[i].read_len[k];
		for (uint32_t z = 0; z < t[i].mm_n[k]; z++) if (mm[t[i].mm_off[k] + z].type == BSO_INS) cap += mm[t[i].mm_off[k] + z].size;
	}
	bso_template *nt = malloc(sizeof(bso_template) * n);
	uint8_t *nb = malloc(cap + 16);
	size_t used = 0;
	prof_ref = refcodes; prof_x = x;              /* with the profile on, refcodes holds one more code (y + 1) */
	int ret = bso_normalise_block(t, n, bases, mm, p, nt, nb, cap + 16, &used);
	prof_ref = NULL;
	/* call_genotypes_ML asserts that no template begins before the window -- the smaller of its two positions, whether or not
	 * a read lies there (src/call_genotypes.c:182-186; the reference's build keeps its asserts).  It happens with -k and -d
	 * together: a lone mate is then kept with the position its absent partner claimed, which may lie before the block
	 * (src/get_template_vector.c:247-268).  The reference aborts; the restatement refuses the block. */
	for (size_t i = 0; !ret && i < n; i++) {
		uint32_t x1 = nt[i].forward_position;
		if (x1 == 0) x1 = nt[i].reverse_position;
		else if (nt[i].reverse_position > 0 && nt[i].reverse_position < x1) x1 = nt[i].reverse_position;
		if (x1 < x) ret = -7;
	}
	if (!ret) {
		bso_pileup *pl = pile_out ? pile_out : malloc(sizeof(bso_pileup) * sz);
		bso_pileup_block(nt, n, nb, x, y, p, pl);
		if (vcf_out) {
			for (uint32_t i = 0; i < sz; i++) {
				memset(vcf_out + i, 0, sizeof(bso_gt_vcf));
				bso_call_site(pl + i, refcodes[i], p, &vcf_out[i].gtm, &vcf_out[i].skip);
				vcf_out[i].ready = 1;
			}
		}
		if (!pile_out) free(pl);
		if (x_out) *x_out = x;
	}
	free(nt);
	free(nb);
	return ret;
}

/* ------------------------------------------------------------------------------------------------
 * Host twin of the device generator k_synth_sites (bs_call_b200/csrc/bsgpu_kernels.cu): the same counter-based
 * draws in the same order, so record i is the same on both sides (config 2 of BASELINE.json; distribution in
 * SURVEY.md section 8d).  Used by the CPU baseline / reference arm of bench.py, which must not launch kernels.
 * ------------------------------------------------------------------------------------------------ */
static inline uint64_t mix64(uint64_t z) {
	z += 0x9e3779b97f4a7c15ull;
	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
	return z ^ (z >> 31);
}
typedef struct { uint64_t s; } rng_t;
static inline uint64_t rng_next(rng_t *r) { r->s += 0x9e3779b97f4a7c15ull; uint64_t z = r->s; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
static inline float rng_unif(rng_t *r) { return (float)(rng_next(r) >> 40) * (1.0f / 16777216.0f); }

static void synth_one(uint64_t seed, uint64_t idx, float mean_depth, bso_pileup *out, uint8_t *ref) {
	rng_t r = { mix64(seed ^ mix64(idx)) };
	float u = rng_unif(&r);
	const int rb = u < 0.295f ? 0 : (u < 0.5f ? 1 : (u < 0.705f ? 2 : 3));
	int a0 = rb, a1 = rb;
	u = rng_unif(&r);
	if (u < 0.001f) a1 = (rb + 1 + (int)(rng_next(&r) % 3)) & 3;
	else if (u < 0.0015f) a0 = a1 = (rb + 1 + (int)(rng_next(&r) % 3)) & 3;
	const float meth = rng_unif(&r) < 0.02f ? 0.7f : 0.01f;
	int depth = 0;
	if (rng_unif(&r) >= 0.03f) {
		double p = exp(-(double)mean_depth), c = p;
		const double uu = (double)(rng_next(&r) >> 11) * (1.0 / 9007199254740992.0);
		while (uu > c && depth < 1000) { depth++; p *= (double)mean_depth / depth; c += p; }
	}
	uint32_t qs[8] = {0};
	memset(out, 0, sizeof(*out));
	for (int k = 0; k < depth; k++) {
		const uint64_t bits = rng_next(&r);
		int b = (bits & 1) ? a1 : a0;
		const int st = 1 + (int)((bits >> 1) & 1), ori = (int)((bits >> 2) & 1);
		const float uc = (float)((bits >> 8) & 0xffffff) * (1.0f / 16777216.0f);
		const int converts = uc >= meth && uc < meth + (1.0f - meth) * 0.99f;
		if (st == 1 && b == 1 && converts) b = 3;
		if (st == 2 && b == 2 && converts) b = 0;
		if (((bits >> 32) & 0x3ff) < 3) b = (b + 1 + (int)((bits >> 42) % 3)) & 3;
		const int q = 20 + (int)((bits >> 48) % 24);
		const int cl = st == 1 ? (b == 1 ? 5 : (b == 3 ? 7 : b)) : (b == 0 ? 4 : (b == 2 ? 6 : b));
		out->counts[ori][cl]++;
		qs[cl] += q;
	}
	for (int j = 0; j < 8; j++) out->quality[j] = (float)qs[j];
	out->n = (uint32_t)depth;
	out->mapq2 = 3600.0f * (float)depth;
	*ref = (uint8_t)(rb + 1);
}

typedef struct { uint64_t seed, first; size_t n, t0, step; float mean; bso_pileup *p; uint8_t *ref; } synth_job;
static void *synth_worker(void *arg) {
	synth_job *j = arg;
	for (size_t i = j->t0; i < j->n; i += j->step) synth_one(j->seed, j->first + i, j->mean, j->p + i, j->ref + i);
	return NULL;
}

void bso_synth_sites(uint64_t seed, uint64_t first, size_t n, double mean_depth, bso_pileup *out, uint8_t *ref, int nthreads) {
	if (nthreads < 1) nthreads = 1;
	synth_job *jobs = malloc(sizeof(synth_job) * nthreads);
	pthread_t *thr = malloc(sizeof(pthread_t) * nthreads);
	for (int i = 0; i < nthreads; i++) {
		jobs[i] = (synth_job){ seed, first, n, (size_t)i, (size_t)nthreads, (float)mean_depth, out, ref };
		if (i) pthread_create(thr + i, NULL, synth_worker, jobs + i);
	}
	synth_worker(jobs);
	for (int i = 1; i < nthreads; i++) pthread_join(thr[i], NULL);
	free(jobs);
	free(thr);
}
