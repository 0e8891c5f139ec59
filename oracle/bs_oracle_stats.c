/* bs_oracle_stats.c -- TEST INFRASTRUCTURE (CPU oracle).  Restatement of the statistics the reference's writer gathers for
 * --report-file while it prints a block (src/print_vcf.c:382-526, inside _print_vcf_entry()), as a function of the block:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu legs may call it.  Pinned against the compiled print_vcf.c by
 * tests/test_oracle_vs_reference.py (bsref_print_block_ann with the harness's bs_stats live).
 *
 * What is counted, per site the writer visits (called, depth > 0, position beyond the last one printed, :112-131):
 *   always             cov[dp].all, cov[dp].gc_pcent[gc of the site's 100-base bin]                           (:385-398)
 *   if the site is not skipped (hom-ref A/T without -A, outside the region / contig, :139, 154-157):
 *     snps or multi / qual[variant_sites] / cov[dp].var   for EVERY such site: the test is `alt[0] != '.'` on a pointer that
 *                                         the record builder has advanced to the string's terminator (:178-182, 400), so it
 *                                         always holds; `alt[1] == ','` then looks one byte PAST that terminator.  In the
 *                                         reference object as gcc builds it that byte is a comma exactly behind the empty ALT
 *                                         string (homozygous reference calls) and a letter behind all others -- so a
 *                                         homozygous reference site counts as "multi" and every other site as "snp".  Undefined
 *                                         behaviour in the source; pinned here to what the compiled print_vcf.c does.
 *     qd / fs / mq histograms by (value, het call), filter_counts[het][flt & 31], qual[all_sites][phred]      (:423-427)
 *     dbSNP_sites / dbSNP_var (known position; "var" = the always-true test above)                             (:428-443)
 *     CpG pairs, qual[CpG_*_sites], cov[dp].CpG, cov[d_inf].CpG_inf and the methylation posterior of the strand (:444-515)
 *     mut_counts / dbSNP_mut_counts by mut_type[call][reference base]                                          (:516-523)
 * "passed" = no filter bit at all (mac1, bit 7, included). */
#include <math.h>
#include <string.h>
#include "bs_oracle.h"

#define LN10 2.30258509299404568402      /* LOG10, include/bs_call.h:36 */

enum { mut_AC = 0, mut_AG, mut_AT, mut_CA, mut_CG, mut_CT, mut_GA, mut_GC, mut_GT, mut_TA, mut_TC, mut_TG, mut_no };      /* include/bs_call.h:46 */
static const int mut_type[10][5] = {                                                       /* src/print_vcf.c:47-58 */
	{mut_no, mut_no, mut_CA, mut_GA, mut_TA}, {mut_no, mut_AC, mut_CA, mut_no, mut_no}, {mut_no, mut_AG, mut_no, mut_GA, mut_no},
	{mut_no, mut_AT, mut_no, mut_no, mut_TA}, {mut_no, mut_AC, mut_no, mut_GC, mut_TC}, {mut_no, mut_no, mut_CG, mut_GC, mut_no},
	{mut_no, mut_no, mut_CT, mut_no, mut_TC}, {mut_no, mut_AG, mut_CG, mut_no, mut_TG}, {mut_no, mut_no, mut_no, mut_GT, mut_TG},
	{mut_no, mut_AT, mut_CT, mut_GT, mut_no},
};
static const int is_het[10] = { 0, 1, 1, 1, 0, 1, 1, 0, 1, 0 };      /* defs.gt_het, src/init_param.c:54 */

static int site_call(const bso_gt_vcf *v) {                        /* :579-588 */
	if (v->skip) return 0;
	int gt = 0;
	double z = v->gtm.gt_prob[0];
	for (int i = 1; i < 10; i++) if (v->gtm.gt_prob[i] > z) { z = v->gtm.gt_prob[i]; gt = i; }
	return gt + 1;
}

static int db_flags(const bso_dbsnp *db, uint32_t pos) {
	if (!db || !db->n) return 0;
	uint32_t lo = 0, hi = db->n;
	while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (db->pos[mid] < pos) lo = mid + 1; else hi = mid; }
	return lo < db->n && db->pos[lo] == pos ? db->flags[lo] : 0;
}

static double lfact(uint32_t x) {                                  /* lfact2(), include/bs_call.h:335 over src/stats_utils.c:14-21 */
	static double store[256];
	static int ready = 0;
	if (!ready) {
		double l = 0.0;
		store[0] = store[1] = 0.0;
		for (int i = 2; i < 256; i++) { l += log((double)i); store[i] = l; }
		ready = 1;
	}
	return x < 256 ? store[x] : lgamma((double)(x + 1));
}

static void hist(uint64_t (*v)[2], int n, uint64_t *overflow, int ct, int var) {      /* add_flt_counts, :22-27 */
	if (ct >= 0 && ct < n) v[ct][var ? 1 : 0]++;
	else if (overflow) (*overflow)++;
}

static bso_cov_stats *cov_of(bso_site_stats *st, uint32_t c) {
	static bso_cov_stats sink;
	if (c < BSO_STATS_COV_MAX) return st->cov + c;
	st->cov_overflow++;
	return &sink;
}

void bso_stats_block(const bso_gt_vcf *vcf, uint32_t sz, const uint8_t *refcodes, uint32_t x, uint32_t ctg_end, int all_positions,
		uint32_t reg_start, uint32_t reg_stop, const bso_dbsnp *db, const uint8_t *gc, int nbins, uint32_t start_pos,
		bso_site_stats *st, bso_stats_state *state) {
	static double logp[100];
	if (logp[0] == 0.0) for (int i = 0; i < 100; i++) logp[i] = log(0.01 * (double)(i + 1));      /* src/init_param.c:56 */
	for (uint32_t i = 0; i < sz; i++) {
		int g[5];
		for (int k = 0; k < 5; k++) { const int64_t j = (int64_t)i + k - 2; g[k] = j < 0 ? 0 : site_call(vcf + (j < (int64_t)sz ? j : (int64_t)sz - 1)); }
		if (!g[2]) continue;
		const bso_gt_meth *gtm = &vcf[i].gtm;
		const uint64_t *counts = gtm->counts;
		uint32_t dp1 = 0, d_inf = 0;
		for (int k = 0; k < 4; k++) dp1 += (uint32_t)counts[k];
		for (int k = 4; k < 8; k++) d_inf += (uint32_t)counts[k];
		const uint32_t dp = dp1 + d_inf, pos = x + i;
		if (!dp) continue;
		/* reference context as the writer's window holds it (see bs_oracle_writer.c) */
		uint8_t rc[5];
		{
			const int64_t last = (int64_t)sz - 1;
			const int64_t wstart = (int64_t)i + 2 <= last ? (int64_t)i - 2 : last - 4;
			int wiped = 0;
			for (int64_t j = wstart < 0 ? 0 : wstart; j < (int64_t)i - 2; j++) if (refcodes[j] == 0) wiped = 1;
			for (int k = 0; k < 5; k++) {
				const int64_t j = (int64_t)i + k - 2;
				uint8_t c = j >= 0 && !wiped ? refcodes[j] : 0;
				if (c == 0 && j >= 0) wiped = 1;
				rc[k] = c;
			}
		}
		const int rfix = rc[2], gt = g[2] - 1;
		const int rs_found = db_flags(db, pos);
		int skip = !all_positions && !(rs_found & 2) && ((gt == 0 && rfix == 1) || (gt == 9 && rfix == 4));
		const double z1 = exp(gtm->gt_prob[gt] * LN10);
		int phred;
		if (z1 >= 1.0) phred = 255;
		else { phred = (int)(-10.0 * log(1.0 - z1) / LN10); if (phred > 255) phred = 255; }
		const int fs = (int)(-gtm->fisher_strand * 10.0 + 0.5);
		const uint32_t qd = dp1 > 0 ? (uint32_t)phred / dp1 : (uint32_t)phred;
		if (!skip) {
			if (reg_start || reg_stop) skip = pos < reg_start || pos > reg_stop;
			else skip = pos > ctg_end;
		}
		uint32_t flt = 0;
		if (!skip) {
			if (phred < 20) flt |= 1;
			if (qd < 2) flt |= 2;
			if (fs > 60) flt |= 4;
			if (gtm->mq < 40) flt |= 8;
			if (!flt) {
				const uint64_t *c = counts;
				int mac1 = 0;
				switch (gt) {
				case 1: mac1 = c[1] + c[5] + c[7] <= 1 || c[0] + c[4] <= 1; break;
				case 2: mac1 = c[2] + c[6] <= 1 || c[0] <= 1; break;
				case 3: mac1 = c[3] + c[7] <= 1 || c[0] + c[4] <= 1; break;
				case 5: mac1 = c[2] + c[6] + c[4] <= 1 || c[1] + c[5] + c[7] <= 1; break;
				case 6: mac1 = c[3] <= 1 || c[1] + c[5] <= 1; break;
				case 8: mac1 = c[3] + c[7] <= 1 || c[2] + c[6] + c[4] <= 1; break;
				}
				if (mac1) flt |= 128;
			}
		}
		/* ---- the statistics (:382-526) */
		bso_cov_stats *gcov = cov_of(st, dp);
		gcov->all++;
		{
			const int bn = (int)((pos - start_pos) / 100);      /* unsigned subtraction, then int, as in the reference */
			if (gc && bn >= 0 && bn < nbins) { const int v = gc[bn]; if (v <= 100) gcov->gc_pcent[v]++; }
		}
		if (skip) continue;
		const int het = is_het[gt];
		/* "variant site": always; "multi" for a homozygous reference call (see the header of this file) */
		{
			static const int hom_allele[10] = { 1, 0, 0, 0, 2, 0, 0, 3, 0, 4 };
			uint64_t *ctr = hom_allele[gt] && hom_allele[gt] == rfix ? st->multi : st->snps;
			ctr[0]++;
			if (!flt) ctr[1]++;
		}
		st->qual[1][phred]++;
		gcov->var++;
		hist(st->qd_stats, 256, NULL, (int)qd, het);
		hist(st->fs_stats, BSO_STATS_FS_MAX, &st->fs_overflow, fs, het);
		hist(st->mq_stats, 256, NULL, gtm->mq, het);
		st->filter_counts[het][flt & 31]++;
		st->qual[0][phred]++;
		if (rs_found) {
			st->dbSNP_sites[0]++; st->dbSNP_var[0]++;
			if (!flt) { st->dbSNP_sites[1]++; st->dbSNP_var[1]++; }
		}
		if ((g[2] == 5 && g[3] == 8) || (g[2] == 8 && g[1] == 5)) {      /* cpg == "CG" (:229-232) */
			int ref_cpg = 0, cpg_ok = 0;
			uint32_t a = 0, b = 0;
			if (gt == 4) {                         /* cs_str "+" */
				state->prev_cpg_x = pos;
				state->prev_cpg_flt = flt != 0;
				ref_cpg = rc[2] == 2 && rc[3] == 3;
				a = (uint32_t)counts[5]; b = (uint32_t)counts[7];
				cpg_ok = 1;
			} else if (gt == 7) {                  /* cs_str "-" */
				ref_cpg = rc[1] == 2 && rc[2] == 3;
				if (pos - state->prev_cpg_x == 1) {
					uint64_t *ctr = ref_cpg ? st->CpG_ref : st->CpG_nonref;
					ctr[0]++;
					if (!(state->prev_cpg_flt || flt)) ctr[1]++;
				}
				a = (uint32_t)counts[6]; b = (uint32_t)counts[4];
				cpg_ok = 1;
			}
			if (cpg_ok) {
				st->qual[ref_cpg ? 2 : 3][phred]++;
				gcov->CpG[ref_cpg ? 0 : 1]++;
				cov_of(st, d_inf)->CpG_inf[ref_cpg ? 0 : 1]++;
				if (a + b) {
					double meth[101];
					const double konst = lfact(a + b + 1) - lfact(a) - lfact(b);
					double sum = 0.0;
					if (a) meth[0] = 0.0; else sum = meth[0] = exp(konst);
					if (b) meth[100] = 0.0; else sum = (meth[100] = exp(konst));
					const double da = (double)a, dbl = (double)b;
					for (int k = 1; k < 100; k++) sum += (meth[k] = exp(konst + logp[k - 1] * da + logp[99 - k] * dbl));
					double (*dst)[101] = ref_cpg ? st->CpG_ref_meth : st->CpG_nonref_meth;
					for (int k = 0; k < 101; k++) {
						const double z = meth[k] / sum;
						dst[0][k] += z;
						if (!flt) dst[1][k] += z;
					}
				}
			}
		}
		const int mut = mut_type[gt][rfix];
		if (mut != mut_no) {
			st->mut_counts[mut][0]++;
			if (!flt) st->mut_counts[mut][1]++;
			if (rs_found) {
				st->dbSNP_mut_counts[mut][0]++;
				if (!flt) st->dbSNP_mut_counts[mut][1]++;
			}
		}
	}
}
