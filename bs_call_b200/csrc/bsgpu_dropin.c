/*
 * bsgpu_dropin.c -- link-compatible replacements for the three symbols of the reference's src/call_genotypes.c
 * (declared in include/bs_call.h:358-360):
 *
 *     void init_calc_threads(sr_param *param);
 *     void call_genotypes_ML(ctg_t *ctg, gt_vector *align_list, uint32_t x, uint32_t y, sr_param *param);
 *     void join_calc_threads(sr_param *param);
 *
 * Build bs_call with this file in place of src/call_genotypes.c (and link libbsgpu.so): every other reference file
 * stays as it is.  The pileup loop and the calc threads' per-site body run on the GPU through the C ABI of
 * include/bsgpu.h; the hand-off protocol with the reader, the meth-profile thread and the print thread is the
 * reference's:
 *   - a block is computed into the spare one of two page-locked arrays while the printer drains the block before, and is
 *     published (work->vcf, vcf_x, vcf_n) once the printer has finished (vcf_n == 0);
 *   - work->ref / work->ref1 are swapped only after the meth-profile ring is empty, and before vcf_n is published;
 *   - every vcf[i] carries ready = true when vcf_n is published (the printer consumes strictly in index order, so
 *     publishing a fully computed block is a legal schedule of the reference's per-site signalling);
 *   - join_calc_threads finally signals vcf_cond.
 * This file needs the reference's headers (it is compiled where the reference tree is available); it contains no
 * arithmetic of the path, only staging and the hand-off.
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
static bsgpu_seg *g_segs;
static size_t g_seg_cap;
static uint8_t *g_bases;
static size_t g_base_cap;
static gt_vcf *g_vcf[2];        /* page-locked; work->vcf points at the one that was published last */
static size_t g_vcf_cap[2];
static int g_next;              /* the array the next block is computed into */

static void die(const char *what) {
	gt_fatal_error_msg("bsgpu: %s: %s\n", what, bsgpu_last_error());
}

static void timed_wait(pthread_cond_t *c, pthread_mutex_t *m) {
	struct timespec ts;
	clock_gettime(CLOCK_REALTIME, &ts);
	ts.tv_sec += 5;
	pthread_cond_timedwait(c, m, &ts);
}

void init_calc_threads(sr_param * const param) {
	work_t * const work = &param->work;
	bsgpu_params p;
	bsgpu_default_params(&p);
	p.under_conv = param->under_conv;
	p.over_conv = param->over_conv;
	p.ref_bias = param->ref_bias;
	p.min_qual = param->min_qual;
	for (int i = 0; i < 2; i++) { p.left_trim[i] = param->left_trim[i]; p.right_trim[i] = param->right_trim[i]; }
	const char *dev = getenv("BSGPU_DEVICE");
	p.device = dev ? atoi(dev) : 0;
	if (bsgpu_init(&p, &g_ctx) != BSGPU_OK) die("bsgpu_init");
	work->calc_end = false;
	work->n_calc_threads = 0;            /* no host calc threads exist */
	work->calc_threads_complete = 0;
	work->calc_threads = NULL;
}

void join_calc_threads(sr_param * const param) {
	work_t * const work = &param->work;
	work->calc_end = true;
	bsgpu_destroy(g_ctx);
	g_ctx = NULL;
	pthread_mutex_lock(&work->vcf_mutex);
	pthread_cond_signal(&work->vcf_cond);
	pthread_mutex_unlock(&work->vcf_mutex);
}

/* the mate walk at the top of the reference's pileup loop (src/call_genotypes.c:181-212, 224), emitting segments */
static size_t stage(gt_vector * const align_list, const uint32_t x, const uint32_t y, size_t *nbases_out) {
	const uint32_t nr = gt_vector_get_used(align_list);
	align_details **al_p = gt_vector_get_mem(align_list, align_details *);
	size_t need_b = 0, need_s = 0;
	for (uint32_t ix = 0; ix < nr; ix++) for (int k = 0; k < 2; k++) {
		gt_vector *rd = al_p[ix]->read[k];
		if (rd == NULL) continue;
		const size_t rl = gt_vector_get_used(rd);
		need_b += rl;
		need_s += (rl + BSGPU_MAX_SEG_LEN - 1) / BSGPU_MAX_SEG_LEN;
	}
	if (need_b > g_base_cap) {
		bsgpu_host_free(g_bases);
		g_base_cap = need_b + need_b / 4 + 4096;
		if ((g_bases = bsgpu_host_alloc(g_base_cap)) == NULL) die("bsgpu_host_alloc");
	}
	if (need_s > g_seg_cap) {
		bsgpu_host_free(g_segs);
		g_seg_cap = need_s + need_s / 4 + 256;
		if ((g_segs = bsgpu_host_alloc(g_seg_cap * sizeof(bsgpu_seg))) == NULL) die("bsgpu_host_alloc");
	}
	size_t ns = 0, nb = 0;
	for (uint32_t ix = 0; ix < nr; ix++) {
		const align_details * const al = al_p[ix];
		uint32_t ori = al->orientation;
		assert(ori < 2);
		const uint32_t st = al->bs_strand;
		for (int k = 0; k < 2; k++) {
			if (al->read[k] == NULL) continue;
			const uint32_t rl = gt_vector_get_used(al->read[k]);
			if (rl == 0) continue;
			const uint8_t *sp = gt_vector_get_mem(al->read[k], uint8_t);
			uint32_t first = 0, last = rl;
			while (first < rl) { const uint8_t q = GET_QUAL(sp[first]); if (q > 0 && q != FLT_QUAL) break; first++; }
			if (first == rl) continue;                         /* no usable base: no strand flip either */
			for (;;) { const uint8_t q = GET_QUAL(sp[last - 1]); if (q > 0 && q != FLT_QUAL) break; last--; }
			uint32_t pos = (k ? al->reverse_position : al->forward_position) + first;
			assert(pos >= x);
			uint32_t len = last - first;
			if (pos <= y) {
				if ((uint64_t)pos + len > (uint64_t)y + 1) len = y + 1 - pos;
				memcpy(g_bases + nb, sp + first, len);
				uint32_t off = (uint32_t)nb;
				nb += len;
				while (len) {
					const uint32_t l = len > BSGPU_MAX_SEG_LEN ? BSGPU_MAX_SEG_LEN : len;
					bsgpu_seg *s = g_segs + ns++;
					s->pos = pos; s->off = off; s->len = (uint16_t)l; s->mapq = al->mapq[k]; s->flags = (uint8_t)(ori | (st << 1)); s->pad_ = 0;
					pos += l; off += l; len -= l;
				}
			}
			ori ^= 1;
		}
	}
	*nbases_out = nb;
	return ns;
}

void call_genotypes_ML(ctg_t * const ctg, gt_vector * const align_list, const uint32_t x, const uint32_t y, sr_param * const param) {
	assert(y >= x);
	const uint32_t sz = y - x + 1;
	work_t * const work = &param->work;
	/* host staging of this block (the reader may not reclaim align_list before we return) */
	size_t nbases = 0;
	const size_t nseg = stage(align_list, x, y, &nbases);
	/* Pileup, model and strand test for every site of the block (work->ref1 holds the reference codes of [x, y + 2],
	 * src/process_template.c:29-30), written in the gt_vcf layout with ready = true into the SPARE one of two arrays: the
	 * print thread may still be writing the block before from the other one.  (The reference also runs its pileup before
	 * it waits for the printer, src/call_genotypes.c:180-235; its calc threads then need the single work->vcf.) */
	const int slot = g_next;
	if (sz > g_vcf_cap[slot]) {
		bsgpu_host_free(g_vcf[slot]);
		g_vcf_cap[slot] = (size_t)sz + sz / 4 + 1024;
		if ((g_vcf[slot] = bsgpu_host_alloc(g_vcf_cap[slot] * sizeof(gt_vcf))) == NULL) die("bsgpu_host_alloc");
	}
	const uint8_t *refcodes = (const uint8_t *)gt_string_get_string(work->ref1);
	if (bsgpu_call_block(g_ctx, g_segs, nseg, g_bases, nbases, refcodes, x, sz, (bsgpu_gt_vcf *)g_vcf[slot]) != BSGPU_OK) die("bsgpu_call_block");
	/* publication: the previous block must have left the print thread (src/call_genotypes.c:228-235) */
	pthread_mutex_lock(&work->print_mutex);
	while (work->vcf_n) timed_wait(&work->print_cond2, &work->print_mutex);
	pthread_mutex_unlock(&work->print_mutex);
	work->vcf = g_vcf[slot];
	work->vcf_size = (int)(g_vcf_cap[slot] > 0x7fffffff ? 0x7fffffff : g_vcf_cap[slot]);
	g_next = slot ^ 1;
	work->vcf_x = x;
	work->vcf_ctg = ctg;
	/* meth profiling reads ref1 and the read buffers: let it finish before the buffers change hands (:243-254) */
	pthread_mutex_lock(&work->mprof_mutex);
	while (work->mprof_read_idx != work->mprof_write_idx) timed_wait(&work->mprof_cond2, &work->mprof_mutex);
	pthread_mutex_unlock(&work->mprof_mutex);
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
