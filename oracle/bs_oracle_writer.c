/*
 * bs_oracle_writer.c -- TEST INFRASTRUCTURE ONLY (see bs_oracle.h).
 *
 * CPU restatement of what the reference's print thread derives from a block of gt_vcf records and serialises as BCF
 * (src/print_vcf.c: print_vcf_entry 548-594, flush_vcf_entries 535-546, _print_vcf_entry 32-381; driven per block by
 * src/process.c:89-104).  The reference walks the block with a five-site sliding window; here every site is a pure
 * function of the block (its own record, the genotype calls and reference codes two sites either side), which is the
 * formulation the device kernel uses.  Pinned against the reference's compiled print_vcf.c by
 * tests/test_oracle_vs_reference.py (oracle/ref_harness.c:bsref_print_block).
 *
 * The byte encoding is BCF2 (VCF/BCF specification v4.3, section 6.3) as htslib's bcf_enc_* / bcf_write produce it;
 * htslib itself is an external dependency of bs_call, absent from /root/reference and from this image.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "bs_oracle.h"

#define LN10 2.30258509299404568402          /* LOG10, include/bs_call.h:36 */

enum { T_INT8 = 1, T_INT16 = 2, T_INT32 = 3, T_FLOAT = 5, T_CHAR = 7 };

typedef struct { uint8_t *p; size_t n; } wbuf;
static void put(wbuf *w, int c) { w->p[w->n++] = (uint8_t)c; }
static void put_mem(wbuf *w, const void *m, size_t l) { memcpy(w->p + w->n, m, l); w->n += l; }

/* type byte: length << 4 | type, lengths >= 15 spill into a typed integer */
static void enc_size(wbuf *w, int size, int type) {
	if (size < 15) { put(w, size << 4 | type); return; }
	put(w, 15 << 4 | type);
	if (size < 128) { put(w, 1 << 4 | T_INT8); put(w, size); }
	else if (size < 32768) { int16_t v = (int16_t)size; put(w, 1 << 4 | T_INT16); put_mem(w, &v, 2); }
	else { int32_t v = size; put(w, 1 << 4 | T_INT32); put_mem(w, &v, 4); }
}
/* one integer in the narrowest type; the lowest eight values of int8 / int16 are reserved */
static void enc_int1(wbuf *w, int32_t x) {
	if (x <= 127 && x >= -120) { enc_size(w, 1, T_INT8); put(w, x); }
	else if (x <= 32767 && x >= -32760) { int16_t v = (int16_t)x; enc_size(w, 1, T_INT16); put_mem(w, &v, 2); }
	else { enc_size(w, 1, T_INT32); put_mem(w, &x, 4); }
}
/* a vector of integers (no missing values occur here): one type for all */
static void enc_vint(wbuf *w, int n, const int32_t *a) {
	if (n == 1) { enc_int1(w, a[0]); return; }
	int32_t mx = INT32_MIN + 1, mn = INT32_MAX;
	for (int i = 0; i < n; i++) { if (a[i] > mx) mx = a[i]; if (a[i] < mn) mn = a[i]; }
	if (mx <= 127 && mn >= -120) { enc_size(w, n, T_INT8); for (int i = 0; i < n; i++) put(w, a[i]); }
	else if (mx <= 32767 && mn >= -32760) { enc_size(w, n, T_INT16); for (int i = 0; i < n; i++) { int16_t v = (int16_t)a[i]; put_mem(w, &v, 2); } }
	else { enc_size(w, n, T_INT32); put_mem(w, a, (size_t)n * 4); }
}
static void enc_chars(wbuf *w, const char *s, int l) { enc_size(w, l, T_CHAR); put_mem(w, s, (size_t)l); }

/* genotype index 0..9 = AA AC AG AT CC CG CT GG GT TT; reference code 0..4 = N A C G T */
static const uint8_t allele_of[10][2] = { {1, 1}, {1, 2}, {1, 3}, {1, 4}, {2, 2}, {2, 3}, {2, 4}, {3, 3}, {3, 4}, {4, 4} };
static const int has_c[10] = { 0, 1, 0, 0, 1, 1, 1, 0, 0, 0 }, has_g[10] = { 0, 0, 1, 0, 0, 1, 0, 1, 1, 0 };      /* :104-105 */
static const int is_het[10] = { 0, 1, 1, 1, 0, 1, 1, 0, 1, 0 };

static int gl_index(int a, int b) { return a < b ? a * (9 - a) / 2 + b - 5 : b * (9 - b) / 2 + a - 5; }      /* alleles 1..4 -> 0..9 */

/* the call the writer makes for a site (:579-588): first maximum of gt_prob, 0 for a skipped site */
static int site_call(const bso_gt_vcf *v) {
	if (v->skip) return 0;
	int gt = 0;
	double z = v->gtm.gt_prob[0];
	for (int i = 1; i < 10; i++) if (v->gtm.gt_prob[i] > z) { z = v->gtm.gt_prob[i]; gt = i; }
	return gt + 1;
}

/* dbSNP hit of position pos (what dbSNP_lookup_name() gives the writer, src/dbSNP.c:305-346): flags 0 / 1 / 3 and the ID bytes */
static int db_lookup(const bso_dbsnp *db, uint32_t pos, const uint8_t **name, uint32_t *len) {
	if (!db || !db->n) return 0;
	uint32_t lo = 0, hi = db->n;
	while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (db->pos[mid] < pos) lo = mid + 1; else hi = mid; }
	if (lo == db->n || db->pos[lo] != pos) return 0;
	*name = db->names + db->name_off[lo];
	*len = db->name_off[lo + 1] - db->name_off[lo];
	return db->flags[lo];
}

size_t bso_print_site(const bso_gt_vcf *vcf, uint32_t sz, const uint8_t *refcodes, uint32_t x, uint32_t i, int rid, uint32_t ctg_end,
		const int *ids, int all_positions, uint8_t *out) {
	return bso_print_site_ann(vcf, sz, refcodes, x, i, rid, ctg_end, ids, all_positions, 0, 0, NULL, out);
}

/* reg_start / reg_stop: ctg->curr_reg (0, 0: none -- then the contig end clips, src/print_vcf.c:154-157); db: the contig's
 * dbSNP entries (NULL: no -D index) */
size_t bso_print_site_ann(const bso_gt_vcf *vcf, uint32_t sz, const uint8_t *refcodes, uint32_t x, uint32_t i, int rid, uint32_t ctg_end,
		const int *ids, int all_positions, uint32_t reg_start, uint32_t reg_stop, const bso_dbsnp *db, uint8_t *out) {
	int g[5];
	/* calls of the five sites around this one; before the block: none; beyond its end: flush_vcf_entries shifts the
	 * window without clearing the slot it vacates (:538), so the last site's call is seen again */
	for (int k = 0; k < 5; k++) { const int64_t j = (int64_t)i + k - 2; g[k] = j < 0 ? 0 : site_call(vcf + (j < (int64_t)sz ? j : (int64_t)sz - 1)); }
	if (!g[2]) return 0;
	const bso_gt_meth *gtm = &vcf[i].gtm;
	uint32_t dp1 = 0, dinf = 0;
	for (int k = 0; k < 4; k++) dp1 += (uint32_t)gtm->counts[k];
	for (int k = 4; k < 8; k++) dinf += (uint32_t)gtm->counts[k];
	if (!(dp1 + dinf)) return 0;
	/* Reference context: the five codes around the site, except that the window is filled by strncpy() from a string in
	 * which N is the terminator (:572-578): once an N has been copied everything after it in the seven-code window reads
	 * as N.  The window normally starts at site - 2; for the last two sites of the block it is the one left over from the
	 * block's last site (flush_vcf_entries only shifts it, :539-543), so an N up to two codes further left wipes them too;
	 * codes before the block start are N. */
	uint8_t rc[5];
	{
		const int64_t last = (int64_t)sz - 1;
		const int64_t wstart = (int64_t)i + 2 <= last ? (int64_t)i - 2 : last - 4;      /* first code of the window in use */
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
	const int hom_ref_at = (gt == 0 && rfix == 1) || (gt == 9 && rfix == 4);      /* gt_flag, :91-102 */
	const uint8_t *rs = NULL;
	uint32_t rs_len = 0;
	const int rs_found = db_lookup(db, x + i, &rs, &rs_len);                       /* :133 */
	if (!all_positions && !(rs_found & 2) && hom_ref_at) return 0;                 /* :139: an "always output" dbSNP site is kept */
	if (reg_start || reg_stop) { if (x + i < reg_start || x + i > reg_stop) return 0; }      /* :154-157 */
	else if (x + i > ctg_end) return 0;
	/* phred-scaled probability that the call is wrong (:140-148), quality by depth, strand bias, filters (:151-153, 184-187) */
	const double z1 = exp(gtm->gt_prob[gt] * LN10);
	int phred;
	if (z1 >= 1.0) phred = 255;
	else { phred = (int)(-10.0 * log(1.0 - z1) / LN10); if (phred > 255) phred = 255; }
	const int fs = (int)(-gtm->fisher_strand * 10.0 + 0.5);
	const uint32_t qd = dp1 > 0 ? (uint32_t)phred / dp1 : (uint32_t)phred;
	uint32_t flt = 0;
	if (phred < 20) flt |= 1;
	if (qd < 2) flt |= 2;
	if (fs > 60) flt |= 4;
	if (gtm->mq < 40) flt |= 8;
	int fid = ids[0];
	if (!flt) {
		const uint64_t *c = gtm->counts;
		int mac1 = 0;
		switch (gt) {                                                                  /* :190-210 */
		case 1: mac1 = c[1] + c[5] + c[7] <= 1 || c[0] + c[4] <= 1; break;
		case 2: mac1 = c[2] + c[6] <= 1 || c[0] <= 1; break;
		case 3: mac1 = c[3] + c[7] <= 1 || c[0] + c[4] <= 1; break;
		case 5: mac1 = c[2] + c[6] + c[4] <= 1 || c[1] + c[5] + c[7] <= 1; break;
		case 6: mac1 = c[3] <= 1 || c[1] + c[5] <= 1; break;
		case 8: mac1 = c[3] + c[7] <= 1 || c[2] + c[6] + c[4] <= 1; break;
		}
		if (mac1) { flt |= 128; fid = ids[2]; }
	} else fid = ids[1];
	static const char base_char[] = "NACGT", iupac[] = "NAMRWCSYGKT";
	char ref_ctx[5], call_ctx[5];
	for (int k = 0; k < 5; k++) { ref_ctx[k] = base_char[rc[k]]; call_ctx[k] = iupac[g[k]]; }
	/* ALT: the alleles of the call that are not the reference base, in base order (ref_alt / all_idx, :34-45, 62-73) */
	int alts[2] = { 0, 0 }, n_alt = 0;
	for (int k = 0; k < 2; k++) { const int a = allele_of[gt][k]; if (a != rfix && (!n_alt || alts[0] != a)) alts[n_alt++] = a; }

	uint8_t sh[384], in[256];
	wbuf S = { sh, 0 }, I = { in, 0 };
	if (!rs_found) enc_size(&S, 0, T_CHAR);                  /* ID (:163-167) */
	else enc_chars(&S, (const char *)rs, (int)rs_len);
	enc_chars(&S, ref_ctx + 2, 1);                           /* REF */
	for (int k = 0; k < n_alt; k++) enc_chars(&S, base_char + alts[k], 1);
	enc_int1(&S, fid);                                       /* FILTER */
	enc_int1(&S, ids[3]);                                    /* INFO CX */
	enc_chars(&S, ref_ctx, 5);

	/* GT as the reference's gt_int table has it (:75-86): 2,2 hom-ref; 4,4 hom-alt; 2,4 het with the reference allele; and
	 * 4,8 -- not 4,6 -- for a het of two ALT alleles */
	int n_fmt = 11;
	{
		const int a0 = allele_of[gt][0], a1 = allele_of[gt][1];
		int32_t v[2];
		if (a0 == a1) v[0] = v[1] = a0 == rfix ? 2 : 4;
		else if (a0 == rfix || a1 == rfix) { v[0] = 2; v[1] = 4; }
		else { v[0] = 4; v[1] = 8; }
		enc_int1(&I, ids[4]);
		enc_vint(&I, 2, v);
	}
	{                                                        /* FT (:277-301) */
		static const char *const name[4] = { "q20", "qd2", "fs60", "mq40" };
		char fb[24];
		int l = 0;
		if (flt & 15) {
			/* every name goes out WITH its terminating NUL (the copy loop at :289 steps over it), ';' between names */
			for (int b = 0; b < 4; b++) if (flt & (1u << b)) { if (l) fb[l++] = ';'; const int nl = (int)strlen(name[b]) + 1; memcpy(fb + l, name[b], (size_t)nl); l += nl; }
		} else { memcpy(fb, "PASS", 4); l = 4; }
		enc_int1(&I, ids[5]);
		enc_chars(&I, fb, l);
	}
	enc_int1(&I, ids[8]); enc_int1(&I, (int32_t)dp1);        /* DP */
	enc_int1(&I, ids[9]); enc_int1(&I, gtm->mq);             /* MQ */
	enc_int1(&I, ids[7]); enc_int1(&I, phred);               /* GQ */
	enc_int1(&I, ids[10]); enc_int1(&I, (int32_t)qd);        /* QD */
	{                                                        /* GL (:317-345): RR, then per ALT allele R/A (if the reference is known) and A/A */
		float gl[6];
		int n = 0;
		double z = rfix ? gtm->gt_prob[gl_index(rfix, rfix)] : -99.999;
		if (z < -99.999) z = -99.999;
		gl[n++] = (float)z;
		for (int k = 0; k < 2 && alts[k] > 0; k++) {
			if (rfix) { z = gtm->gt_prob[gl_index(rfix, alts[k])]; if (z < -99.999) z = -99.999; gl[n++] = (float)z; }
			z = gtm->gt_prob[gl_index(alts[k], alts[k])];
			if (z < -99.999) z = -99.999;
			gl[n++] = (float)z;
		}
		enc_int1(&I, ids[6]);
		enc_size(&I, n, T_FLOAT);
		put_mem(&I, gl, (size_t)n * 4);
	}
	{                                                        /* MC8, AMQ (:347-358) */
		int32_t v[8];
		for (int k = 0; k < 8; k++) v[k] = (int32_t)gtm->counts[k];
		enc_int1(&I, ids[11]);
		enc_vint(&I, 8, v);
		int n = 0;
		for (int k = 0; k < 8; k++) if (gtm->counts[k] > 0) v[n++] = gtm->qual[k];
		if (n) { enc_int1(&I, ids[12]); enc_vint(&I, n, v); n_fmt++; }
	}
	{                                                        /* CS: strand(s) on which the call has a cytosine (:60-61) */
		const char *cs = has_c[gt] ? (has_g[gt] ? "+-" : "+") : has_g[gt] ? "-" : "NA";
		enc_int1(&I, ids[13]); enc_chars(&I, cs, (int)strlen(cs));
	}
	{                                                        /* CG: CpG status from the calls either side (:229-270) */
		const int c0 = g[2], nx = g[3], pv = g[1];
		char cg;
		if ((c0 == 5 && nx == 8) || (c0 == 8 && pv == 5)) cg = 'C';          /* "CG": one character is written (:366-367) */
		else if (c0 == 5) cg = nx ? (has_g[nx - 1] ? 'H' : 'N') : '?';
		else if (c0 == 8) cg = pv ? (has_c[pv - 1] ? 'H' : 'N') : '?';
		else if (has_c[c0 - 1]) cg = nx ? (has_g[nx - 1] ? 'H' : 'N') : '?';
		else if (has_g[c0 - 1]) cg = pv ? (has_c[pv - 1] ? 'H' : 'N') : '.';
		else cg = '.';
		enc_int1(&I, ids[14]);
		enc_chars(&I, &cg, 1);
	}
	enc_int1(&I, ids[3]); enc_chars(&I, call_ctx, 5);        /* CX */
	if (is_het[gt]) { enc_int1(&I, ids[15]); enc_int1(&I, fs); n_fmt++; }      /* FS */

	/* the record as bcf_write lays it out */
	uint32_t h[8];
	const float qual = (float)phred;
	h[0] = (uint32_t)S.n + 24; h[1] = (uint32_t)I.n; h[2] = (uint32_t)rid; h[3] = x + i - 1; h[4] = 1;
	memcpy(h + 5, &qual, 4);
	h[6] = (uint32_t)(1 + n_alt) << 16 | 1u;
	h[7] = (uint32_t)n_fmt << 24 | 1u;
	memcpy(out, h, 32);
	memcpy(out + 32, sh, S.n);
	memcpy(out + 32 + S.n, in, I.n);
	return 32 + S.n + I.n;
}

int bso_print_block(const bso_gt_vcf *vcf, uint32_t sz, const uint8_t *refcodes, uint32_t x, int rid, uint32_t ctg_end,
		const int *vcf_ids, int all_positions, uint8_t *out, size_t cap, size_t *nbytes, size_t *nrec) {
	return bso_print_block_ann(vcf, sz, refcodes, x, rid, ctg_end, vcf_ids, all_positions, 0, 0, NULL, out, cap, nbytes, nrec);
}

int bso_print_block_ann(const bso_gt_vcf *vcf, uint32_t sz, const uint8_t *refcodes, uint32_t x, int rid, uint32_t ctg_end,
		const int *vcf_ids, int all_positions, uint32_t reg_start, uint32_t reg_stop, const bso_dbsnp *db,
		uint8_t *out, size_t cap, size_t *nbytes, size_t *nrec) {
	size_t at = 0, n = 0;
	for (uint32_t i = 0; i < sz; i++) {
		if (at + BSO_BCF_MAX_RECORD > cap) return -3;
		const size_t l = bso_print_site_ann(vcf, sz, refcodes, x, i, rid, ctg_end, vcf_ids, all_positions, reg_start, reg_stop, db, out + at);
		if (l) { at += l; n++; }
	}
	*nbytes = at; *nrec = n;
	return 0;
}

/* Walks two BCF record streams side by side (test infrastructure for full-size diffs: Python cannot walk millions of
 * records).  Counts records, records whose 32 leading bytes (lengths, CHROM, POS, rlen, QUAL, n_allele | n_info,
 * n_fmt | n_sample) agree, records that agree byte for byte, and checks that POS never decreases within a CHROM in `a`.
 * out[0] records in a, out[1] records in b, out[2] fixed part equal, out[3] identical, out[4] order violations in a,
 * out[5] index of the first record that differs in its fixed part (or -1).  Returns 0, or -1 on a malformed stream. */
int bso_bcf_diff(const uint8_t *a, size_t na, const uint8_t *b, size_t nb, long long *out) {
	size_t ia = 0, ib = 0;
	long long ra = 0, rb = 0, fixed = 0, same = 0, bad_order = 0, first_diff = -1;
	int32_t last_chrom = -1, last_pos = -1;
	while (ia < na || ib < nb) {
		size_t la = 0, lb = 0;
		if (ia < na) {
			if (ia + 32 > na) return -1;
			uint32_t s, i;
			memcpy(&s, a + ia, 4); memcpy(&i, a + ia + 4, 4);
			la = 8 + (size_t)s + i;
			if (la < 32 || ia + la > na) return -1;
			int32_t chrom, pos;
			memcpy(&chrom, a + ia + 8, 4); memcpy(&pos, a + ia + 12, 4);
			if (chrom == last_chrom && pos <= last_pos) bad_order++;
			last_chrom = chrom; last_pos = pos;
			ra++;
		}
		if (ib < nb) {
			if (ib + 32 > nb) return -1;
			uint32_t s, i;
			memcpy(&s, b + ib, 4); memcpy(&i, b + ib + 4, 4);
			lb = 8 + (size_t)s + i;
			if (lb < 32 || ib + lb > nb) return -1;
			rb++;
		}
		if (la && lb) {
			if (!memcmp(a + ia, b + ib, 32)) fixed++;
			else if (first_diff < 0) first_diff = ra - 1;
			if (la == lb && !memcmp(a + ia, b + ib, la)) same++;
		}
		ia += la; ib += lb;
	}
	out[0] = ra; out[1] = rb; out[2] = fixed; out[3] = same; out[4] = bad_order; out[5] = first_diff;
	return 0;
}
