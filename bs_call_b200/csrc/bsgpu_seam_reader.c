/*
 * bsgpu_seam_reader.c -- seams C and D: a link-compatible replacement for the reference's src/get_template_vector.c,
 * src/process_template.c AND src/call_genotypes.c (include/bs_call.h:355-360):
 *
 *     gt_status read_input(htsFile *sam_input, gt_vector *align_list, sr_param *param);
 *     gt_status process_template_vector(...), void call_genotypes_ML(...)      -- not reached any more
 *     void init_calc_threads(sr_param *param);  void join_calc_threads(sr_param *param);
 *
 * Build bs_call with this file in place of those three (and link libbsgpu.so).  main(), option parsing, htslib input,
 * the reference FASTA, the VCF/BCF header and file, and (seam C) the print thread with its writer stay the reference's.
 *
 * read_input here does what the reference's does at its top -- fetch one alignment record after the other with
 * sam_read1() / sam_itr_next() over the regions (src/get_template_vector.c:66-110, src/input_sam.c:228-229) -- and puts
 * every record of a wanted contig, as the BAM record it is, into a streaming session of the device library
 * (bsgpu_bam_reserve / _commit, include/bsgpu.h): record decode and filters, mate pairing, duplicate removal, block
 * cutting, normalisation, pileup and model all happen behind that.  A second thread takes the results:
 *   seam C (default): gt_vcf[] of every block, published to the reference's print thread by the protocol of
 *       src/call_genotypes.c:228-258 -- work->vcf points INTO the session's page-locked result, nothing is copied;
 *   seam D (BSGPU_SEAM_RECORDS=1): the BCF records of every block (print_vcf_entry's work done on the device), handed to
 *       htslib's bcf_write() one by one -- which writes BCF as it is and formats VCF text for -O v / z.  -D: the contig's dbSNP
 *       entries go to the device writer when the contig starts (contig_annotation); --report-file: the site statistics of
 *       src/print_vcf.c:382-526 are gathered on the device and folded into bs_stats at join_calc_threads.
 * Contig sequences: the session wants the codes 0..4 of a whole contig; they are taken from the reference's own
 * get_sequence_string() when the first record of a contig arrives and kept to the end of the run (one byte per position).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>
#include <sys/stat.h>
#include <htslib/sam.h>
#include <htslib/vcf.h>
#ifdef BSGPU_SEAM_BULK_IO
#include <htslib/bgzf.h>
#endif

#include "gem_tools.h"
#include "bs_call.h"
#include "bsgpu.h"

static bsgpu_ctx *g_ctx;
static int g_profile;
static int g_site_stats;         /* seam D with --report-file: the statistics of print_vcf.c:382-526 are gathered on the device too */
static int g_site_ctgs;

static void die(const char *what) {
	gt_fatal_error_msg("bsgpu: %s: %s\n", what, bsgpu_last_error());
}

static void timed_wait(pthread_cond_t *c, pthread_mutex_t *m) {
	struct timespec ts;
	clock_gettime(CLOCK_REALTIME, &ts);
	ts.tv_sec += 5;
	pthread_cond_timedwait(c, m, &ts);
}

static bsgpu_params g_params;

/* BSGPU_SEAM_TIMING=1: wall-clock stamps of the seam's phases on stderr (start-up cost against streaming time) */
static int g_timing = -1;
static double g_t0;
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; }
static void stamp(const char *what) {
	if (g_timing < 0) { const char *e = getenv("BSGPU_SEAM_TIMING"); g_timing = e != NULL && atoi(e) != 0; g_t0 = now_s(); }
	if (g_timing) fprintf(stderr, "bsgpu seam: %8.3f s  %s\n", now_s() - g_t0, what);
}

static void params_of(const sr_param * const param, bsgpu_params * const p) {
	bsgpu_default_params(p);
	p->under_conv = param->under_conv;
	p->over_conv = param->over_conv;
	p->ref_bias = param->ref_bias;
	p->min_qual = param->min_qual;
	for (int i = 0; i < 2; i++) { p->left_trim[i] = param->left_trim[i]; p->right_trim[i] = param->right_trim[i]; }
	const char *dev = getenv("BSGPU_DEVICE");
	p->device = dev ? atoi(dev) : 0;
}

void init_calc_threads(sr_param * const param) {
	work_t * const work = &param->work;
	params_of(param, &g_params);
	stamp("init_calc_threads");
	if (bsgpu_init(&g_params, &g_ctx) != BSGPU_OK) die("bsgpu_init");
	stamp("bsgpu_init done");
	g_profile = 0;
	work->calc_end = false;
	work->n_calc_threads = 0;
	work->calc_threads_complete = 0;
	work->calc_threads = NULL;
}

static void fold_profile(bs_stats * const stats) {
	bsgpu_profile *pr = malloc(sizeof(bsgpu_profile));
	if (pr == NULL || bsgpu_profile_read(g_ctx, pr, 1) != BSGPU_OK) die("bsgpu_profile_read");
	if (pr->used > gt_vector_get_used(stats->meth_profile)) {
		const uint64_t old = gt_vector_get_used(stats->meth_profile);
		gt_vector_reserve(stats->meth_profile, pr->used, false);
		memset(gt_vector_get_mem(stats->meth_profile, meth_cts) + old, 0, (pr->used - old) * sizeof(meth_cts));
		gt_vector_set_used(stats->meth_profile, pr->used);
	}
	meth_cts *mc = gt_vector_get_mem(stats->meth_profile, meth_cts);
	for (uint32_t i = 0; i < pr->used; i++) for (int k = 0; k < 4; k++) mc[i].conv_cts[k] += pr->conv_cts[i][k];
	for (int k = 0; k < 5; k++) stats->base_filter[k] += pr->base_filter[k];
	for (int k = 0; k < 15; k++) { stats->filter_cts[k] += pr->filter_cts[k]; stats->filter_bases[k] += pr->filter_bases[k]; }
	free(pr);
}

/* what _print_vcf_entry() would have added to bs_stats and the contigs' ctg_stats for the sites the device wrote (seam D) */
static void add_vec(gt_vector * const v, const uint64_t (*src)[2], const int n) {
	int top = -1;
	for (int i = 0; i < n; i++) if (src[i][0] | src[i][1]) top = i;
	if (top < 0) return;
	if ((uint64_t)top >= v->elements_allocated) gt_vector_reserve(v, top + 1, true);
	if ((uint64_t)top >= v->used) v->used = top + 1;
	for (int i = 0; i <= top; i++) { fstats_cts * const c = gt_vector_get_elm(v, i, fstats_cts); c->cts[0] += src[i][0]; c->cts[1] += src[i][1]; }
}

/* Which written sites does the linked reference writer file under "multi" instead of "snps"?  Its test (src/print_vcf.c:400-401)
 * looks at the byte BEHIND the terminator of the site's ALT string -- the record builder has walked `alt` to the end by then
 * (:178-182) -- i.e. at whatever the linker put behind that literal in the program's merged string pool: undefined behaviour,
 * different from one link of the same sources to the next (in oracle/_ref/libbsref.so a comma follows the empty string of a
 * homozygous reference call, in oracle/_ref/bs_call nothing of the kind does).  The device reports the two groups separately
 * (bsgpu_site_stats.multi = homozygous reference calls, .snps = every other written site); this probe reads the same byte of
 * the same pool -- identical literals of all objects of a link are merged into one copy -- and the fold follows it. */
static __attribute__((noinline)) int comma_behind(const char *lit) {
	const volatile char *p = lit;
	__asm__ volatile("" : "+r"(p));
	while (*p) p++;
	return p[1] == ',';
}
static int homref_is_multi(void) {
	const char *e = getenv("BSGPU_STATS_HOMREF_MULTI");        /* override, for a writer object that lives in another link */
	return e != NULL ? atoi(e) != 0 : comma_behind("");
}
static int others_are_multi(void) {
	static const char * const alts[10] = {"A", "AC", "AG", "AT", "C", "CG", "CT", "G", "GT", "T"};
	int n = 0;
	for (int i = 0; i < 10; i++) n += comma_behind(alts[i]);
	return n == 10;
}
int bsgpu_seam_variant_rule(void) { return homref_is_multi() | others_are_multi() << 1; }       /* for tests: bit 0 hom-ref -> multi, bit 1 others -> multi */

static void fold_site_stats(sr_param * const param) {
	work_t * const work = &param->work;
	bs_stats * const stats = work->stats;
	const int hm = homref_is_multi(), om = others_are_multi();
	bsgpu_site_stats *ss = malloc(sizeof(bsgpu_site_stats));
	bsgpu_ctg_site_stats *cs = calloc((size_t)(g_site_ctgs > 0 ? g_site_ctgs : 1), sizeof(bsgpu_ctg_site_stats));
	if (ss == NULL || cs == NULL || bsgpu_site_stats_read(g_ctx, ss, cs, g_site_ctgs, 1) != BSGPU_OK) die("bsgpu_site_stats_read");
	for (int k = 0; k < 2; k++) {
		*(hm ? &stats->multi[k] : &stats->snps[k]) += ss->multi[k]; *(om ? &stats->multi[k] : &stats->snps[k]) += ss->snps[k];
		stats->dbSNP_sites[k] += ss->dbSNP_sites[k]; stats->dbSNP_var[k] += ss->dbSNP_var[k];
		stats->CpG_ref[k] += ss->CpG_ref[k]; stats->CpG_nonref[k] += ss->CpG_nonref[k];
		for (int m = 0; m < 12; m++) { stats->mut_counts[m][k] += ss->mut_counts[m][k]; stats->dbSNP_mut_counts[m][k] += ss->dbSNP_mut_counts[m][k]; }
		for (int i = 0; i < 101; i++) { stats->CpG_ref_meth[k][i] += ss->CpG_ref_meth[k][i]; stats->CpG_nonref_meth[k][i] += ss->CpG_nonref_meth[k][i]; }
		for (int f = 0; f < 32; f++) stats->filter_counts[k][f] += ss->filter_counts[k][f];
	}
	for (int q = 0; q < 4; q++) for (int i = 0; i < 256; i++) stats->qual[q][i] += ss->qual[q][i];
	add_vec(stats->qd_stats, ss->qd_stats, 256);
	add_vec(stats->mq_stats, ss->mq_stats, 256);
	add_vec(stats->fs_stats, ss->fs_stats, BSGPU_STATS_FS_MAX);
	for (uint32_t c = 0; c < BSGPU_STATS_COV_MAX; c++) {
		const bsgpu_cov_stats * const g = ss->cov + c;
		if (!(g->all | g->var | g->CpG[0] | g->CpG[1] | g->CpG_inf[0] | g->CpG_inf[1])) continue;
		gt_cov_stats *gcov;
		HASH_FIND(hh, stats->cov_stats, &c, sizeof(uint32_t), gcov);
		if (gcov == NULL) {
			gcov = calloc((size_t)1, sizeof(gt_cov_stats));
			gcov->coverage = c;
			HASH_ADD(hh, stats->cov_stats, coverage, sizeof(uint32_t), gcov);
		}
		gcov->all += g->all; gcov->var += g->var;
		for (int k = 0; k < 2; k++) { gcov->CpG[k] += g->CpG[k]; gcov->CpG_inf[k] += g->CpG_inf[k]; }
		for (int i = 0; i < 101; i++) gcov->gc_pcent[i] += g->gc_pcent[i];
	}
	if (ss->fs_overflow | ss->cov_overflow)
		fprintf(stderr, "bsgpu: %" PRIu64 " site(s) with FS >= %d and %" PRIu64 " with a depth >= %d are missing from the report's histograms\n",
				ss->fs_overflow, BSGPU_STATS_FS_MAX, ss->cov_overflow, BSGPU_STATS_COV_MAX);
	for (int k = 0; k < (int)work->n_contigs; k++) {
		ctg_t * const ctg = work->contigs[k];
		if (ctg->ctg_stats == NULL || ctg->vcf_rid < 0 || ctg->vcf_rid >= g_site_ctgs) continue;
		const bsgpu_ctg_site_stats * const c = cs + ctg->vcf_rid;
		for (int j = 0; j < 2; j++) {
			*(hm ? &ctg->ctg_stats->multi[j] : &ctg->ctg_stats->snps[j]) += c->multi[j]; *(om ? &ctg->ctg_stats->multi[j] : &ctg->ctg_stats->snps[j]) += c->snps[j];
			ctg->ctg_stats->dbSNP_sites[j] += c->dbSNP_sites[j];
			ctg->ctg_stats->dbSNP_var[j] += c->dbSNP_var[j]; ctg->ctg_stats->CpG_ref[j] += c->CpG_ref[j]; ctg->ctg_stats->CpG_nonref[j] += c->CpG_nonref[j];
		}
	}
	free(ss); free(cs);
}

void join_calc_threads(sr_param * const param) {
	work_t * const work = &param->work;
	work->calc_end = true;
	if (g_profile && work->stats != NULL) fold_profile(work->stats);
	if (g_site_stats && work->stats != NULL) fold_site_stats(param);
	g_site_stats = 0;
	stamp("join_calc_threads");
	bsgpu_destroy(g_ctx);
	stamp("bsgpu_destroy done");
	g_ctx = NULL;
	pthread_mutex_lock(&work->vcf_mutex);
	pthread_cond_signal(&work->vcf_cond);
	pthread_mutex_unlock(&work->vcf_mutex);
}

void call_genotypes_ML(ctg_t * const ctg, gt_vector * const align_list, const uint32_t x, const uint32_t y, sr_param * const param) {
	gt_fatal_error_msg("bsgpu: call_genotypes_ML reached although the whole chain runs on the device\n");
}

gt_status process_template_vector(gt_vector *align_list, ctg_t * const ctg, uint32_t y, sr_param *param) {
	gt_fatal_error_msg("bsgpu: process_template_vector reached although the whole chain runs on the device\n");
	return GT_STATUS_FAIL;
}

/* ---- the thread that takes the session's results ---- */
typedef struct {
	sr_param *param;
	bsgpu_bam_session *sess;
	uint8_t **codes;             /* per tid: codes of positions 1 .. target_len, or NULL */
	int n_targets;
	int records;                 /* seam D */
	int failed;
} taker_t;

/* seam C: one block of gt_vcf records to the reference's print thread (src/call_genotypes.c:228-258, src/process.c:80-104) */
static void publish_block(taker_t * const tk, const bsgpu_block * const b, const gt_vcf * const vcf) {
	work_t * const work = &tk->param->work;
	const int k = work->tid2id[b->tid];
	assert(k >= 0);
	ctg_t * const ctg = work->contigs[k];
	const uint32_t sz = b->y - b->x + 1;
	/* reference codes of [x, y + 2] into the spare string: N from the contig's last position on (src/get_sequence.c:41-48) */
	gt_string_resize(work->ref1, sz + 3);
	char *rp = work->ref1->buffer;
	const uint32_t len = ctg->end_pos;
	const uint8_t *codes = tk->codes[b->tid];
	for (uint32_t i = 0; i < sz + 2; i++) { const uint64_t pos = (uint64_t)b->x + i; rp[i] = pos < len ? (char)codes[pos - 1] : 0; }
	rp[sz + 2] = 0;
	work->ref1->length = sz + 3;
	pthread_mutex_lock(&work->print_mutex);
	while (work->vcf_n) timed_wait(&work->print_cond2, &work->print_mutex);
	pthread_mutex_unlock(&work->print_mutex);
	work->vcf = (gt_vcf *)vcf;           /* the print thread reads the session's page-locked result in place */
	work->vcf_size = (int)sz;
	work->vcf_x = b->x;
	work->vcf_ctg = ctg;
	gt_string *tp = work->ref;
	work->ref = work->ref1;
	work->ref1 = tp;
	work->vcf_n = sz;
	pthread_mutex_lock(&work->print_mutex);
	pthread_cond_signal(&work->print_cond1);
	pthread_mutex_unlock(&work->print_mutex);
	pthread_mutex_lock(&work->vcf_mutex);
	pthread_cond_signal(&work->vcf_cond);
	pthread_mutex_unlock(&work->vcf_mutex);
}

/* seam D: the records as they lie in a BCF file -> bcf_write(), which is what _print_vcf_entry ends with (src/print_vcf.c:375-380) */
static int write_records(taker_t * const tk, bcf1_t * const bcf, const uint8_t *p, size_t n) {
	work_t * const work = &tk->param->work;
	while (n) {
		uint32_t w[8];
		if (n < 32) return -1;
		memcpy(w, p, 32);
		const size_t l_shared = w[0], l_indiv = w[1];
		if (l_shared < 24 || 8 + l_shared + l_indiv > n) return -1;
		kstring_t sh = bcf->shared, in = bcf->indiv;
		bcf->rid = (int32_t)w[2]; bcf->pos = (int32_t)w[3]; bcf->rlen = (int32_t)w[4];
		memcpy(&bcf->qual, &w[5], 4);
		bcf->n_info = w[6] & 0xffff; bcf->n_allele = w[6] >> 16;
		bcf->n_sample = w[7] & 0xffffff; bcf->n_fmt = w[7] >> 24;
		bcf->shared.s = (char *)(p + 32); bcf->shared.l = l_shared - 24; bcf->shared.m = bcf->shared.l;
		bcf->indiv.s = (char *)(p + 8 + l_shared); bcf->indiv.l = l_indiv; bcf->indiv.m = l_indiv;
		const int rc = bcf_write(work->vcf_file, work->vcf_hdr, bcf);
		bcf->shared = sh; bcf->indiv = in;
		if (rc) return -1;
		p += 8 + l_shared + l_indiv;
		n -= 8 + l_shared + l_indiv;
	}
	return 0;
}

static void *taker_thread(void *arg) {
	taker_t * const tk = arg;
	work_t * const work = &tk->param->work;
	bcf1_t *bcf = tk->records ? bcf_init() : NULL;
	uint64_t held = 0;               /* seam C: the result the print thread is still reading from */
	for (;;) {
		bsgpu_bam_result res;
		if (bsgpu_bam_drain(tk->sess, 1, &res) != BSGPU_OK) { fprintf(stderr, "bsgpu: %s\n", bsgpu_last_error()); tk->failed = 1; break; }
		if (res.id) {
			if (tk->records) {
				if (write_records(tk, bcf, res.data, res.nbytes)) { fprintf(stderr, "bsgpu: malformed record stream or write error\n"); tk->failed = 1; }
				bsgpu_bam_release(tk->sess, res.id);
			} else {
				const gt_vcf *vcf = (const gt_vcf *)res.data;
				for (size_t b = 0; b < res.nblocks; b++) publish_block(tk, res.blocks + b, vcf + res.blocks[b].vcf_off);
				/* the block published last is still with the print thread: its result is released once the next one is out */
				if (held) {
					bsgpu_bam_release(tk->sess, held);
					held = 0;
				}
				if (res.nblocks) held = res.id; else bsgpu_bam_release(tk->sess, res.id);
			}
		}
		if (res.finished) break;
	}
	if (held) {
		pthread_mutex_lock(&work->print_mutex);
		while (work->vcf_n) timed_wait(&work->print_cond2, &work->print_mutex);
		pthread_mutex_unlock(&work->print_mutex);
		bsgpu_bam_release(tk->sess, held);
	}
	if (bcf) bcf_destroy(bcf);
	return NULL;
}

/* codes 0..4 of a whole contig from the reference's own sequence store (src/get_sequence.c:20-55).  The length is the
 * header's / index's (ctg->seq_len, src/process_sam_header.c:188,221-224): end_pos is only set once load_sequence() has run
 * (src/read_reference.c:112).  No previous contig is handed to get_sequence_string(): it would free_sequence() it, which also
 * zeroes that contig's end_pos (src/read_reference.c:35-42) -- and the print thread, blocks behind this reader, still clips
 * every record against end_pos (src/print_vcf.c:157).  The packed copy this call made load_sequence() allocate is released
 * here instead, end_pos / start_pos stay. */
static uint8_t *contig_codes(ctg_t * const ctg, sr_param * const param) {
	gt_string *s = gt_string_new(1024);
	const uint32_t len = ctg->end_pos ? ctg->end_pos : ctg->seq_len;
	if (len == 0) { gt_string_delete(s); return NULL; }
	gt_string_resize(s, (uint64_t)len + 8);
	const bool loaded_here = ctg->seq == NULL;
	if (get_sequence_string(ctg, 1, len, NULL, s, param)) { gt_string_delete(s); return NULL; }
	uint8_t *codes = malloc((size_t)len + 8);
	if (codes != NULL) memcpy(codes, s->buffer, len);
	gt_string_delete(s);
	if (loaded_here && ctg->seq != NULL) { free(ctg->seq); ctg->seq = NULL; }
	return codes;
}

/* seam D with -D: the print thread, which loads a contig's dbSNP entries when its first site arrives (src/print_vcf.c:552-561),
 * never sees a site, so the reader does it when the contig's first record arrives: the reference's own index reader loads the
 * contig (src/dbSNP.c:157-304), every known position is asked for through dbSNP_lookup_name() (:305-350) and the answers -- flags
 * and id bytes -- go to the device writer together with the contig's region (src/print_vcf.c:133,139,154-157,163-167). */
static void contig_annotation(sr_param * const param, ctg_t * const ctg, const int tid) {
	work_t * const work = &param->work;
	const uint32_t r0 = ctg->curr_reg ? ctg->curr_reg->start : 0, r1 = ctg->curr_reg ? ctg->curr_reg->stop : 0;
	dbsnp_ctg_t *dc = NULL;
	if (work->dbSNP_hdr != NULL) HASH_FIND(hh, work->dbSNP_hdr->dbSNP, ctg->name, strlen(ctg->name), dc);
	if (dc != NULL && !load_dbSNP_ctg(work->dbSNP_hdr, dc)) dc = NULL;
	if (dc == NULL) {
		if ((r0 | r1) && bsgpu_set_contig_annotation(g_ctx, tid, r0, r1, NULL) != BSGPU_OK) die("bsgpu_set_contig_annotation");
		return;
	}
	size_t n = 0, cap = 1024, nb = 0, bcap = 16384;
	uint32_t *pos = malloc(cap * sizeof(uint32_t)), *off = malloc((cap + 1) * sizeof(uint32_t));
	uint8_t *flags = malloc(cap), *names = malloc(bcap);
	char rs[512];
	for (int bn = dc->min_bin; bn <= dc->max_bin; bn++) {
		const dbsnp_bin_t * const b = dc->bins + (bn - dc->min_bin);
		for (int ix = 0; ix < 64; ix++) {
			if (!(b->mask >> ix & 1)) continue;
			const uint32_t x = (uint32_t)bn << 6 | (uint32_t)ix;
			size_t len = 0;
			const uint8_t f = dbSNP_lookup_name(work->dbSNP_hdr, dc, rs, &len, x);
			if (!f) continue;
			if (len > BSGPU_DBSNP_MAX_ID) len = BSGPU_DBSNP_MAX_ID;
			if (n == cap) {
				cap *= 2;
				pos = realloc(pos, cap * sizeof(uint32_t)); off = realloc(off, (cap + 1) * sizeof(uint32_t)); flags = realloc(flags, cap);
			}
			if (nb + len > bcap) { bcap = 2 * (nb + len); names = realloc(names, bcap); }
			pos[n] = x; flags[n] = f; off[n] = (uint32_t)nb;
			memcpy(names + nb, rs, len);
			nb += len; n++;
		}
	}
	off[n] = (uint32_t)nb;
	bsgpu_dbsnp db = {(uint32_t)n, pos, flags, off, names};
	if (bsgpu_set_contig_annotation(g_ctx, tid, r0, r1, &db) != BSGPU_OK) die("bsgpu_set_contig_annotation");
	free(pos); free(off); free(flags); free(names);
	unload_dbSNP_ctg(dc);
}

/* A contig begins (src/get_template_vector.c:111-124): its region, its sequence, on seam D its dbSNP entries and GC bins.  Called by
 * the reader when the first record of the contig arrives, or -- when the reader hands over inflated bytes without looking at the
 * records -- by the session's worker before the contig's first results are computed (bsgpu_bam_on_contig). */
static int start_contig(void * const user, const int tid) {
	taker_t * const tk = user;
	sr_param * const param = tk->param;
	work_t * const work = &param->work;
	const int k = tid >= 0 && tid < tk->n_targets ? work->tid2id[tid] : -1;
	if (k < 0) return BSGPU_FAIL;
	ctg_t * const ctg = work->contigs[k];
	fprintf(stderr, "Processing chromosome %s (OK)\n", ctg->name);
	ctg->curr_reg = work->curr_region;
	if (tk->codes[tid] == NULL) tk->codes[tid] = contig_codes(ctg, param);
	if (tk->codes[tid] == NULL) {
		fprintf(stderr, "Problem loading reference sequence for contig '%s'\n", ctg->name);
		return BSGPU_FAIL;
	}
	if (tk->records) contig_annotation(param, ctg, tid);
	if (bsgpu_bam_set_contig(tk->sess, tid, tk->codes[tid]) != BSGPU_OK) die("bsgpu_bam_set_contig");
	if (g_site_stats && ctg->ctg_stats != NULL && ctg->ctg_stats->gc != NULL
			&& bsgpu_set_contig_gc(g_ctx, ctg->vcf_rid, ctg->start_pos, ctg->ctg_stats->gc, (uint32_t)ctg->ctg_stats->nbins) != BSGPU_OK) die("bsgpu_set_contig_gc");
	return BSGPU_OK;
}

gt_status read_input(htsFile *sam_input, gt_vector * align_list, sr_param *param) {
	work_t * const work = &param->work;
	bam_hdr_t * const hdr = work->sam_header;
	const int nt = hdr->n_targets;
	bsgpu_reader_params rp;
	bsgpu_default_reader_params(&rp);
	rp.max_template_len = (uint32_t)param->max_template_len;
	rp.mapq_thresh = param->mapq_thresh;
	rp.keep_unmatched = param->keep_unmatched;
	rp.ignore_duplicates = param->ignore_duplicates;
	rp.keep_duplicates = param->keep_duplicates;
	taker_t tk;
	memset(&tk, 0, sizeof(tk));
	tk.param = param; tk.n_targets = nt;
	tk.records = getenv("BSGPU_SEAM_RECORDS") != NULL && atoi(getenv("BSGPU_SEAM_RECORDS")) != 0;
	tk.codes = calloc((size_t)nt, sizeof(uint8_t *));
	{   /* the options are fixed for a run of bs_call; a host that changed them since init_calc_threads gets a new context */
		bsgpu_params now;
		params_of(param, &now);
		if (memcmp(&now, &g_params, sizeof(now))) {
			bsgpu_destroy(g_ctx);
			g_params = now;
			if (bsgpu_init(&g_params, &g_ctx) != BSGPU_OK) die("bsgpu_init");
			g_profile = 0;
		}
	}
	if (work->stats != NULL) {
		if (bsgpu_profile_enable(g_ctx, 1) != BSGPU_OK) die("bsgpu_profile_enable");
		g_profile = 1;
		if (tk.records && !g_site_stats) {
			/* seam D: no gt_vcf record reaches the reference's writer, so its per-site statistics are gathered where the
			 * records are built (per-contig counters by vcf_rid) */
			g_site_ctgs = 0;
			for (int k = 0; k < (int)work->n_contigs; k++) if (work->contigs[k]->vcf_rid >= g_site_ctgs) g_site_ctgs = work->contigs[k]->vcf_rid + 1;
			if (bsgpu_site_stats_enable(g_ctx, 1, g_site_ctgs) != BSGPU_OK) die("bsgpu_site_stats_enable");
			g_site_stats = 1;
		}
	}
	bsgpu_bcf_params bp;
	int32_t *rid = NULL;
	if (tk.records) {
		bsgpu_default_bcf_params(&bp);
		for (int k = 0; k < 16; k++) bp.ids[k] = work->vcf_ids[k];
		bp.all_positions = param->all_positions;
		rid = calloc((size_t)nt, sizeof(int32_t));
		for (int t = 0; t < nt; t++) { const int k = work->tid2id[t]; rid[t] = k >= 0 ? work->contigs[k]->vcf_rid : 0; }
	}
	const uint8_t **codes0 = calloc((size_t)nt, sizeof(uint8_t *));
	/* Batch size: the library's default (384 MiB) suits a genome; a small input would spend longer page-locking the two
	 * stages and three results of that size than streaming through them, and its first results would only leave when a third
	 * of it had been read.  For a regular file an eighth of its size, between 32 MiB and the default (BGZF level 0 ... 6 inflate
	 * to 1 ... 4 times the file); BSGPU_BATCH_BYTES overrides. */
	size_t batch = 0;
	if (getenv("BSGPU_BATCH_BYTES") == NULL && param->input_file != NULL) {
		struct stat sb;
		if (stat(param->input_file, &sb) == 0 && S_ISREG(sb.st_mode)) {
			batch = (size_t)sb.st_size / 8;
			if (batch < ((size_t)32 << 20)) batch = (size_t)32 << 20;
			if (batch >= ((size_t)384 << 20)) batch = 0;
		}
	}
	stamp("read_input: opening the session");
	if (bsgpu_bam_open(g_ctx, nt, hdr->target_len, codes0, &rp, tk.records ? &bp : NULL, rid, batch, &tk.sess) != BSGPU_OK) die("bsgpu_bam_open");
	stamp("session open");
	free(codes0);
	pthread_t taker;
	pthread_create(&taker, NULL, taker_thread, &tk);

	const int n_reg = work->n_regions;
	int reg_ix = 0;
	hts_itr_t *itr = NULL;
	if (n_reg > 0 && work->sam_idx) {
		region_t * const reg = work->regions + reg_ix++;
		itr = sam_itr_queryi(work->sam_idx, reg->ctg->bam_tid, reg->start - 1, reg->stop);
		fprintf(stderr, "Processing region %s:%u-%u\n", reg->ctg->name, reg->start, reg->stop);
		work->curr_region = reg;
	}
#ifdef BSGPU_SEAM_BULK_IO
	/* Bulk input: when every contig of the header is wanted and no index region restricts the run, the record bytes need not be
	 * looked at on the host at all -- the session frames, filters and decodes them on the device.  bgzf_read() inflates
	 * straight into the session's page-locked stage (bsgpu_bam_reserve / _commit, no copy), and the contig bookkeeping of
	 * src/get_template_vector.c:111-124 happens when the session reaches the contig (bsgpu_bam_on_contig -> start_contig).
	 * sam_read1() + one copy per record was 2.8 of the 5.9 s of a config-1 run.  BSGPU_SEAM_BULK=0 keeps the record loop. */
	{
		const char *e = getenv("BSGPU_SEAM_BULK");
		bool bulk = (e == NULL || atoi(e) != 0) && itr == NULL && n_reg == 0 && sam_input->format.format == bam;
		for (int t = 0; t < nt && bulk; t++) if (work->tid2id[t] < 0) bulk = false;
		if (bulk) {
			if (bsgpu_bam_on_contig(tk.sess, start_contig, &tk) != BSGPU_OK) die("bsgpu_bam_on_contig");
			gt_status st = GT_STATUS_OK;
			bool first_slice = true;
			for (;;) {
				uint8_t *dst = NULL;
				size_t avail = 0;
				if (bsgpu_bam_reserve(tk.sess, &dst, &avail, 1) != BSGPU_OK) { fprintf(stderr, "bsgpu: %s\n", bsgpu_last_error()); st = GT_STATUS_FAIL; break; }
				if (avail > ((size_t)4 << 20)) avail = (size_t)4 << 20;       /* an open reservation holds the stage: a few MB at a time */
				const ssize_t got = bgzf_read(sam_input->fp.bgzf, dst, avail);
				if (first_slice && got >= 8) {
					/* the first contig's sequence now, before the pipeline has anything to wait for (the later ones when the
					 * session reaches them, under the batches already queued) */
					int32_t tid0;
					memcpy(&tid0, dst + 4, 4);
					if (tid0 >= 0 && tid0 < nt && start_contig(&tk, tid0) != BSGPU_OK) { bsgpu_bam_commit(tk.sess, 0); st = GT_STATUS_FAIL; break; }
				}
				first_slice = false;
				if (bsgpu_bam_commit(tk.sess, got > 0 ? (size_t)got : 0) != BSGPU_OK) { fprintf(stderr, "bsgpu: %s\n", bsgpu_last_error()); st = GT_STATUS_FAIL; break; }
				if (got < 0) { st = GT_STATUS_FAIL; break; }
				if (got == 0) break;
				if (tk.failed) { st = GT_STATUS_FAIL; break; }
			}
			stamp("last byte fed (bulk)");
			if (bsgpu_bam_finish(tk.sess) != BSGPU_OK) { fprintf(stderr, "bsgpu: %s\n", bsgpu_last_error()); st = GT_STATUS_FAIL; }
			pthread_join(taker, NULL);
			stamp("last result written");
			if (tk.failed) st = GT_STATUS_FAIL;
			bsgpu_bam_close(tk.sess);
			stamp("session closed");
			for (int t = 0; t < nt; t++) free(tk.codes[t]);
			free(tk.codes);
			free(rid);
			return st;
		}
	}
#endif
	bam1_t *b = bam_init1();
	int curr_tid = -1;
	bool chr_skip = false;
	gt_status st = GT_STATUS_OK;
	uint8_t *dst = NULL;
	size_t avail = 0, used = 0;
	for (;;) {
		int ret = itr == NULL ? sam_read1(sam_input, hdr, b) : sam_itr_next(sam_input, itr, b);
		if (ret < 0) {
			if (ret != -1) { st = GT_STATUS_FAIL; break; }
			if (itr != NULL) hts_itr_destroy(itr);
			itr = NULL;
			if (reg_ix < n_reg && work->sam_idx) {
				region_t * const reg = work->regions + reg_ix++;
				itr = sam_itr_queryi(work->sam_idx, reg->ctg->bam_tid, reg->start - 1, reg->stop);
				fprintf(stderr, "Processing region %s:%u-%u\n", reg->ctg->name, reg->start, reg->stop);
				work->curr_region = reg;
				continue;
			}
			break;                                          /* end of input */
		}
		const bam1_core_t * const c = &b->core;
		if (c->tid >= 0 && c->tid != curr_tid) {
			curr_tid = c->tid;
			const int k = c->tid < nt ? work->tid2id[curr_tid] : -1;
			chr_skip = k < 0;
			if (chr_skip) fprintf(stderr, "Processing chromosome %s (SKIP)\n", hdr->target_name[curr_tid]);
			else if (start_contig(&tk, curr_tid) != BSGPU_OK) { st = GT_STATUS_FAIL; break; }
		}
		if (chr_skip || c->tid < 0) continue;               /* records without a wanted contig never reach a block */
		/* the record as it lies in a BAM file: block_size, the 32 fixed bytes, then qname | cigar | seq | qual | aux */
		const size_t need = 36 + (size_t)b->l_data;
		if (avail - used < need) {
			if (dst != NULL && bsgpu_bam_commit(tk.sess, used) != BSGPU_OK) die("bsgpu_bam_commit");
			dst = NULL; used = 0;
			for (;;) {
				if (bsgpu_bam_reserve(tk.sess, &dst, &avail, 1) != BSGPU_OK) die("bsgpu_bam_reserve");
				if (avail >= need) break;
				/* the tail of a staging buffer: fill it with the head of the record through the copying entry point */
				if (bsgpu_bam_commit(tk.sess, 0) != BSGPU_OK) die("bsgpu_bam_commit");
				dst = NULL;
				break;
			}
		}
		uint8_t hdr36[36];
		const uint32_t bs = 32 + (uint32_t)b->l_data, bin_mq_nl = (uint32_t)c->bin << 16 | (uint32_t)c->qual << 8 | (uint32_t)c->l_qname;
		const uint32_t flag_nc = (uint32_t)c->flag << 16 | (c->n_cigar & 0xffff);
		const int32_t pos = (int32_t)c->pos, mpos = (int32_t)c->mpos, isize = (int32_t)c->isize, lq = c->l_qseq, tid = c->tid, mtid = c->mtid;
		memcpy(hdr36, &bs, 4); memcpy(hdr36 + 4, &tid, 4); memcpy(hdr36 + 8, &pos, 4); memcpy(hdr36 + 12, &bin_mq_nl, 4);
		memcpy(hdr36 + 16, &flag_nc, 4); memcpy(hdr36 + 20, &lq, 4); memcpy(hdr36 + 24, &mtid, 4); memcpy(hdr36 + 28, &mpos, 4);
		memcpy(hdr36 + 32, &isize, 4);
		if (dst != NULL) {
			memcpy(dst + used, hdr36, 36);
			memcpy(dst + used + 36, b->data, (size_t)b->l_data);
			used += need;
			/* an open reservation keeps the session from moving carried records in front of this buffer: commit every few MB */
			if (used >= ((size_t)4 << 20)) {
				if (bsgpu_bam_commit(tk.sess, used) != BSGPU_OK) die("bsgpu_bam_commit");
				dst = NULL; avail = used = 0;
			}
		} else {
			if (bsgpu_bam_feed(tk.sess, hdr36, 36, NULL) != BSGPU_OK || bsgpu_bam_feed(tk.sess, b->data, (size_t)b->l_data, NULL) != BSGPU_OK) die("bsgpu_bam_feed");
			avail = used = 0;
		}
		if (tk.failed) { st = GT_STATUS_FAIL; break; }
	}
	if (dst != NULL && bsgpu_bam_commit(tk.sess, used) != BSGPU_OK) die("bsgpu_bam_commit");
	stamp("last record fed");
	if (bsgpu_bam_finish(tk.sess) != BSGPU_OK) { fprintf(stderr, "bsgpu: %s\n", bsgpu_last_error()); st = GT_STATUS_FAIL; }
	pthread_join(taker, NULL);
	stamp("last result written");
	if (tk.failed) st = GT_STATUS_FAIL;
	bsgpu_bam_close(tk.sess);
	stamp("session closed");
	bam_destroy1(b);
	for (int t = 0; t < nt; t++) free(tk.codes[t]);
	free(tk.codes);
	free(rid);
	return st;
}
