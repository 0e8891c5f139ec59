/*
 * bsgpu_seam_template.c -- seam B: a link-compatible replacement for the reference's src/process_template.c AND
 * src/call_genotypes.c (include/bs_call.h:357-360):
 *
 *     gt_status process_template_vector(gt_vector *align_list, ctg_t *ctg, uint32_t y, sr_param *param);
 *     void init_calc_threads(sr_param *param);
 *     void join_calc_threads(sr_param *param);
 *     void call_genotypes_ML(...)            -- not reached any more (process_template_vector was its only caller)
 *
 * Build bs_call with this file in place of those two (and link libbsgpu.so); read_input, the print thread, the writer and
 * everything else stay the reference's.  The block read_input hands over goes to the device as it is -- raw mates with
 * their CIGAR events -- and trim_read / trim_soft_clips / handle_overlap / indel normalisation (src/read_utils.c:13,
 * src/al_utils.c:122,164, src/process_template.c:36-111), the pileup (src/call_genotypes.c:172-226) and the per-site model
 * (:43-115) all run there through bsgpu_process_block.  Unlike the reference the reads are NOT normalised in place: the
 * align_details go back to read_input's free list untouched, which is all read_input ever does with them.
 *
 * Hand-off to the print thread -- and what is better than in the reference's own schedule:
 *   - results are computed into the spare one of TWO page-locked gt_vcf arrays while the print thread is still writing the
 *     block before (the reference runs its pileup before it waits for the printer, src/call_genotypes.c:180-235, but its
 *     calc threads then write into the single work->vcf); only the publication waits for vcf_n == 0;
 *   - work->ref / work->ref1 are swapped at publication, every vcf[i] carries ready = true;
 *   - with --report-file (work->stats != NULL) the conversion profile and the base / read tallies are gathered on the
 *     device (bsgpu_profile_enable) and folded into work->stats by join_calc_threads; the mprof ring stays empty.
 * This file needs the reference's headers; it contains no arithmetic of the path, only staging and the hand-off.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>

#include "gem_tools.h"
#include "bs_call.h"
#include "bsgpu.h"

static bsgpu_ctx *g_ctx;
static bsgpu_template *g_tmpl;
static size_t g_tmpl_cap;
static bsgpu_misms *g_mm;
static size_t g_mm_cap;
static uint8_t *g_bases;
static size_t g_base_cap;
static gt_vcf *g_vcf[2];        /* page-locked; work->vcf points at the one that was published last */
static size_t g_vcf_cap[2];
static int g_next;              /* the array the next block is computed into */
static int g_profile;

static void die(const char *what) {
	gt_fatal_error_msg("bsgpu: %s: %s\n", what, bsgpu_last_error());
}

static void timed_wait(pthread_cond_t *c, pthread_mutex_t *m) {
	struct timespec ts;
	clock_gettime(CLOCK_REALTIME, &ts);
	ts.tv_sec += 5;
	pthread_cond_timedwait(c, m, &ts);
}

static void *pinned_grow(void *old, size_t *cap, size_t need, size_t elem) {
	if (need <= *cap) return old;
	bsgpu_host_free(old);
	*cap = need + need / 4 + 1024;
	void *p = bsgpu_host_alloc(*cap * elem);
	if (p == NULL) die("bsgpu_host_alloc");
	return p;
}

static bsgpu_params g_params;
static void fold_profile(bs_stats * const stats);

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
	if (bsgpu_init(&g_params, &g_ctx) != BSGPU_OK) die("bsgpu_init");
	g_profile = 0;
	work->calc_end = false;
	work->n_calc_threads = 0;            /* no host calc threads exist */
	work->calc_threads_complete = 0;
	work->calc_threads = NULL;
}

/* the --report-file side channels of the replaced functions, gathered on the device, into the reference's bs_stats */
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
	/* filter_cts / filter_bases: only what process_template_vector and its helpers count (read_input keeps its own) */
	stats->filter_cts[gt_flt_none] += pr->filter_cts[0];
	stats->filter_bases[gt_flt_none] += pr->filter_bases[0];
	free(pr);
}

void join_calc_threads(sr_param * const param) {
	work_t * const work = &param->work;
	work->calc_end = true;
	if (g_profile && work->stats != NULL) fold_profile(work->stats);
	bsgpu_destroy(g_ctx);
	g_ctx = NULL;
	pthread_mutex_lock(&work->vcf_mutex);
	pthread_cond_signal(&work->vcf_cond);
	pthread_mutex_unlock(&work->vcf_mutex);
}

void call_genotypes_ML(ctg_t * const ctg, gt_vector * const align_list, const uint32_t x, const uint32_t y, sr_param * const param) {
	gt_fatal_error_msg("bsgpu: call_genotypes_ML reached although process_template_vector runs on the device\n");
}

gt_status process_template_vector(gt_vector *align_list, ctg_t * const ctg, uint32_t y, sr_param *param) {
	work_t * const work = &param->work;
	const size_t n = gt_vector_get_used(align_list);
	assert(n);
	align_details **al_p = gt_vector_get_mem(align_list, align_details *);
	uint32_t x = (*al_p)->forward_position;
	if (x == 0) x = (*al_p)->reverse_position;
	assert(x > 0 && x <= y);
	x = x > 2 ? x - 2 : 1;
	const uint32_t sz = y - x + 1;
	gt_string_resize(work->ref1, sz + 3);
	if (get_sequence_string(ctg, x, sz + 2, work->vcf_ctg, work->ref1, param)) {
		fprintf(stderr, "Problem loading reference sequence for contig '%s' %" PRIu32 " %" PRIu32 "\n", ctg->name, x, sz);
		return GT_STATUS_FAIL;
	}
	/* -L / -R, -Q and the conversion rates are fixed for a run of bs_call; a host that changes them between blocks (the
	 * test harness does) gets a context for the new values */
	{
		bsgpu_params now;
		params_of(param, &now);
		if (memcmp(&now, &g_params, sizeof(now))) {
			if (g_profile && work->stats != NULL) fold_profile(work->stats);
			bsgpu_destroy(g_ctx);
			g_params = now;
			if (bsgpu_init(&g_params, &g_ctx) != BSGPU_OK) die("bsgpu_init");
			g_profile = 0;
		}
	}
	if (work->stats != NULL && !g_profile) {
		if (bsgpu_profile_enable(g_ctx, 1) != BSGPU_OK) die("bsgpu_profile_enable");
		g_profile = 1;
	}
	/* flat image of the block: templates, packed reads, CIGAR events */
	size_t nb = 0, nm = 0;
	for (size_t i = 0; i < n; i++) for (int k = 0; k < 2; k++) {
		const align_details * const al = al_p[i];
		if (al->read[k] == NULL) continue;
		nb += gt_vector_get_used(al->read[k]);
		nm += gt_vector_get_used(al->mismatches[k]);
	}
	g_tmpl = pinned_grow(g_tmpl, &g_tmpl_cap, n, sizeof(bsgpu_template));
	g_bases = pinned_grow(g_bases, &g_base_cap, nb + 16, 1);
	g_mm = pinned_grow(g_mm, &g_mm_cap, nm + 1, sizeof(bsgpu_misms));
	nb = nm = 0;
	for (size_t i = 0; i < n; i++) {
		const align_details * const al = al_p[i];
		bsgpu_template * const t = g_tmpl + i;
		memset(t, 0, sizeof(*t));
		t->forward_position = al->forward_position;
		t->reverse_position = al->reverse_position;
		t->orientation = (uint8_t)al->orientation;
		t->bs_strand = (uint8_t)al->bs_strand;
		for (int k = 0; k < 2; k++) {
			t->mapq[k] = al->mapq[k];
			t->reference_span[k] = al->reference_span[k];
			if (al->read[k] == NULL) continue;
			const uint32_t rl = gt_vector_get_used(al->read[k]);
			const uint32_t ne = gt_vector_get_used(al->mismatches[k]);
			t->present[k] = 1;
			t->read_off[k] = (uint32_t)nb; t->read_len[k] = rl;
			t->mm_off[k] = (uint32_t)nm; t->mm_n[k] = ne;
			memcpy(g_bases + nb, gt_vector_get_mem(al->read[k], uint8_t), rl);
			nb += rl;
			const gt_misms *mp = gt_vector_get_mem(al->mismatches[k], gt_misms);
			for (uint32_t z = 0; z < ne; z++) { g_mm[nm].type = (uint32_t)mp[z].misms_type; g_mm[nm].position = mp[z].position; g_mm[nm].size = mp[z].size; nm++; }
		}
	}
	/* the whole block on the device, into the spare array: the print thread may still be busy with the other one */
	const int slot = g_next;
	g_vcf[slot] = pinned_grow(g_vcf[slot], &g_vcf_cap[slot], sz, sizeof(gt_vcf));
	uint32_t x_dev = 0;
	const uint8_t *refcodes = (const uint8_t *)gt_string_get_string(work->ref1);
	if (bsgpu_process_block(g_ctx, g_tmpl, n, g_bases, nb, g_mm, nm, refcodes, y, &x_dev, (bsgpu_gt_vcf *)g_vcf[slot]) != BSGPU_OK) {
		fprintf(stderr, "bsgpu: %s\n", bsgpu_last_error());
		return GT_STATUS_FAIL;
	}
	assert(x_dev == x);
	/* publication: the previous block must have left the print thread (src/call_genotypes.c:228-235) */
	pthread_mutex_lock(&work->print_mutex);
	while (work->vcf_n) timed_wait(&work->print_cond2, &work->print_mutex);
	pthread_mutex_unlock(&work->print_mutex);
	work->vcf = g_vcf[slot];
	work->vcf_size = (int)(g_vcf_cap[slot] > 0x7fffffff ? 0x7fffffff : g_vcf_cap[slot]);
	g_next = slot ^ 1;
	work->vcf_x = x;
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
	return GT_STATUS_OK;
}
