/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin driver that links the reference's own, unmodified hot-path sources
 * (compiled in place from /root/reference by oracle/Makefile into oracle/_ref/libbsref.so)
 * and exposes them through a flat C interface that tests/ and bench.py's cpu_baseline /
 * --impl reference legs can call with ctypes.
 *
 * What runs inside each entry point is the reference's code, not a restatement:
 *   bsref_calc_gt_prob   -> calc_gt_prob()            src/genotype_model.c:44
 *   bsref_fisher         -> fisher()                  src/stats_utils.c:25
 *   bsref_call_block     -> call_genotypes_ML()       src/call_genotypes.c:155 (+ call_thread 21)
 *   bsref_process_block  -> process_template_vector() src/process_template.c:18
 *
 * The harness supplies what src/process.c normally supplies around those calls
 * (src/process.c:20-41 meth-profile drainer, 74-110 vcf[] consumer, 158-166 set-up).
 * The pileup[] array is function-static inside call_genotypes_ML; it is observed, without touching
 * the source, by wrapping gt_vector_new at link time (-Wl,--wrap) and remembering the vector whose
 * element size is sizeof(pileup).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>
#include <math.h>

#include "gem_tools.h"
#include "bs_call.h"

/* ---- flat records shared with oracle/bs_oracle.h (same layout, checked in tests) ---- */
typedef struct {
	uint32_t forward_position, reverse_position;
	uint32_t reference_span[2];
	uint32_t read_off[2];   /* offset into bases[] */
	uint32_t read_len[2];
	uint32_t mm_off[2];     /* offset into misms[] */
	uint32_t mm_n[2];
	uint8_t present[2];     /* read[k] != NULL */
	uint8_t mapq[2];
	uint8_t orientation;    /* 0 FORWARD, 1 REVERSE */
	uint8_t bs_strand;      /* 0 NON_CONVERTED, 1 C2T, 2 G2A */
	uint8_t pad[2];
} bsref_template;

typedef struct { uint32_t type, position, size; } bsref_misms;

static sr_param par;
static ctg_t ctg;
static pthread_t drain_thr;
static int inited = 0;
static gt_vector *captured_pileup = NULL;
static gt_vector *align_list = NULL;
static double last_call_seconds = 0.0;    /* wall time of the last call_genotypes_ML + all sites ready (harness set-up excluded) */

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
double bsref_last_call_seconds(void) { return last_call_seconds; }

/* ---- link-time interposition to observe the function-static pileup vector ---- */
gt_vector *__real_gt_vector_new(size_t n, size_t es);
gt_vector *__wrap_gt_vector_new(size_t n, size_t es) {
	gt_vector *v = __real_gt_vector_new(n, es);
	if (es == sizeof(pileup)) captured_pileup = v;
	return v;
}

/* ---- symbols the hot-path objects leave unresolved ---- */
bool load_sequence(ctg_t * const contig, faidx_t *idx, const bool calc_gc) { return contig->seq == NULL; }
void free_sequence(ctg_t * const contig) { }

/* meth-profile ring drainer: same hand-off protocol as src/process.c:20-41 (stats are NULL so
 * meth_profile() itself is a no-op, src/meth_profile.c:49) */
static void *drain_mprof(void *arg) {
	work_t * const w = &par.work;
	pthread_mutex_lock(&w->mprof_mutex);
	for (;;) {
		while (w->mprof_read_idx == w->mprof_write_idx && !w->mprof_end) {
			struct timespec ts;
			clock_gettime(CLOCK_REALTIME, &ts);
			ts.tv_sec += 1;
			pthread_cond_timedwait(&w->mprof_cond1, &w->mprof_mutex, &ts);
		}
		if (w->mprof_read_idx == w->mprof_write_idx) break;
		int ix = w->mprof_read_idx;
		pthread_mutex_unlock(&w->mprof_mutex);
		mprof_thread_t *mp = &w->mprof_thread[ix];
		meth_profile(mp->al, mp->x, mp->orig_pos, mp->max_pos, &par);
		pthread_mutex_lock(&w->mprof_mutex);
		w->mprof_read_idx = (ix + 1) % N_MPROF_BUFFERS;
		pthread_cond_broadcast(&w->mprof_cond2);
	}
	pthread_mutex_unlock(&w->mprof_mutex);
	return NULL;
}

int bsref_sizeof(int what) {
	switch (what) {
	case 0: return (int)sizeof(pileup);
	case 1: return (int)sizeof(gt_meth);
	case 2: return (int)sizeof(gt_vcf);
	case 3: return (int)sizeof(bsref_template);
	case 4: return (int)sizeof(qual_prob);
	}
	return -1;
}

int bsref_init(double under_conv, double over_conv, double ref_bias, int min_qual, int n_extra_calc_threads,
		const uint32_t left_trim[2], const uint32_t right_trim[2]) {
	if (inited) return -1;
	init_param(&par);
	par.under_conv = under_conv;
	par.over_conv = over_conv;
	par.ref_bias = ref_bias;
	par.min_qual = (uint8_t)min_qual;
	par.num_threads[CALC_THREADS] = n_extra_calc_threads;
	for (int i = 0; i < 2; i++) {
		par.left_trim[i] = left_trim ? left_trim[i] : 0;
		par.right_trim[i] = right_trim ? right_trim[i] : 0;
	}
	par.work.ref = gt_string_new(16384);
	par.work.ref1 = gt_string_new(16384);
	for (int i = 0; i < N_MPROF_BUFFERS; i++) {
		par.work.mprof_thread[i].orig_pos[0] = gt_vector_new(256, sizeof(int));
		par.work.mprof_thread[i].orig_pos[1] = gt_vector_new(256, sizeof(int));
	}
	memset(&ctg, 0, sizeof(ctg));
	ctg.name = "ctg";
	fill_base_prob_table();
	init_calc_threads(&par);
	pthread_create(&drain_thr, NULL, drain_mprof, NULL);
	align_list = gt_vector_new(32, sizeof(align_details *));
	inited = 1;
	return 0;
}

/* Parameters can be changed between blocks (all reads of them are per call). */
void bsref_set_params(double under_conv, double over_conv, double ref_bias, int min_qual,
		const uint32_t left_trim[2], const uint32_t right_trim[2]) {
	par.under_conv = under_conv;
	par.over_conv = over_conv;
	par.ref_bias = ref_bias;
	par.min_qual = (uint8_t)min_qual;
	for (int i = 0; i < 2; i++) {
		par.left_trim[i] = left_trim ? left_trim[i] : 0;
		par.right_trim[i] = right_trim ? right_trim[i] : 0;
	}
}

void bsref_shutdown(void) {
	if (!inited) return;
	join_calc_threads(&par);
	pthread_mutex_lock(&par.work.mprof_mutex);
	par.work.mprof_end = true;
	pthread_cond_broadcast(&par.work.mprof_cond1);
	pthread_mutex_unlock(&par.work.mprof_mutex);
	pthread_join(drain_thr, NULL);
	inited = 0;
}

/* Direct calls into the model */
void bsref_calc_gt_prob(const uint64_t counts[8], const int qual[8], int rf, gt_meth *out) {
	memset(out, 0, sizeof(gt_meth));
	for (int i = 0; i < 8; i++) { out->counts[i] = counts[i]; out->qual[i] = qual[i]; }
	calc_gt_prob(out, &par, (char)rf);
}

void bsref_calc_gt_prob_batch(const uint64_t *counts, const int *qual, const uint8_t *rf, size_t n, gt_meth *out) {
	for (size_t i = 0; i < n; i++) bsref_calc_gt_prob(counts + 8 * i, qual + 8 * i, rf[i], out + i);
}

double bsref_fisher(const int c[4]) {
	int t[4] = { c[0], c[1], c[2], c[3] };
	return fisher(t, par.defs.lfact_store);
}

void bsref_lfact_table(double *out) { memcpy(out, par.defs.lfact_store, sizeof(double) * LFACT_STORE_SIZE); }

/* ---- block-level ---- */
static align_details *make_al(const bsref_template *t, const uint8_t *bases, const bsref_misms *mm) {
	align_details *al = malloc(sizeof(align_details));
	memset(al, 0, sizeof(*al));
	al->forward_position = t->forward_position;
	al->reverse_position = t->reverse_position;
	for (int k = 0; k < 2; k++) {
		al->reference_span[k] = t->reference_span[k];
		al->mapq[k] = t->mapq[k];
		al->mismatches[k] = gt_vector_new(8, sizeof(gt_misms));
		if (t->present[k]) {
			al->read[k] = gt_vector_new(t->read_len[k] + 2, sizeof(uint8_t));
			if (t->read_len[k]) memcpy(al->read[k]->memory, bases + t->read_off[k], t->read_len[k]);
			gt_vector_set_used(al->read[k], t->read_len[k]);
		}
		for (uint32_t z = 0; mm && z < t->mm_n[k]; z++) {
			gt_misms m;
			memset(&m, 0, sizeof(m));
			m.misms_type = (gt_misms_t)mm[t->mm_off[k] + z].type;
			m.position = mm[t->mm_off[k] + z].position;
			m.size = mm[t->mm_off[k] + z].size;
			gt_vector_insert(al->mismatches[k], m, gt_misms);
		}
	}
	al->orientation = (gt_strand)t->orientation;
	al->bs_strand = (gt_bs_strand)t->bs_strand;
	return al;
}

static void free_al(align_details *al) {
	for (int k = 0; k < 2; k++) {
		if (al->read[k]) gt_vector_delete(al->read[k]);
		gt_vector_delete(al->mismatches[k]);
	}
	free(al);
}

static void build_list(const bsref_template *t, size_t n, const uint8_t *bases, const bsref_misms *mm) {
	gt_vector_reserve(align_list, n, false);
	align_details **p = gt_vector_get_mem(align_list, align_details *);
	for (size_t i = 0; i < n; i++) p[i] = make_al(t + i, bases, mm);
	gt_vector_set_used(align_list, n);
}

static void drop_list(void) {
	align_details **p = gt_vector_get_mem(align_list, align_details *);
	for (size_t i = 0; i < gt_vector_get_used(align_list); i++) free_al(p[i]);
	gt_vector_clear(align_list);
}

/* Stand-in for print_thread (src/process.c:74-110) without the print_vcf_entry call */
static double call_t0;
static void consume(uint32_t sz, pileup *pile_out, gt_vcf *vcf_out, uint8_t *ref_out) {
	work_t * const w = &par.work;
	for (int i = 0; i < w->vcf_n; i++) {
		while (!w->vcf[i].ready) {
			pthread_mutex_lock(&w->vcf_mutex);
			while (!w->vcf[i].ready) {
				struct timespec ts;
				clock_gettime(CLOCK_REALTIME, &ts);
				ts.tv_nsec += 2000000;
				if (ts.tv_nsec >= 1000000000) { ts.tv_sec++; ts.tv_nsec -= 1000000000; }
				pthread_cond_timedwait(&w->vcf_cond, &w->vcf_mutex, &ts);
			}
			pthread_mutex_unlock(&w->vcf_mutex);
		}
	}
	/* all sites ready; wait for the calc threads to report completion so the next block can start cleanly */
	pthread_mutex_lock(&w->calc_mutex);
	while (w->calc_threads_complete < w->n_calc_threads) {
		struct timespec ts;
		clock_gettime(CLOCK_REALTIME, &ts);
		ts.tv_nsec += 2000000;
		if (ts.tv_nsec >= 1000000000) { ts.tv_sec++; ts.tv_nsec -= 1000000000; }
		pthread_cond_timedwait(&w->calc_cond2, &w->calc_mutex, &ts);
	}
	pthread_mutex_unlock(&w->calc_mutex);
	last_call_seconds = now_s() - call_t0;
	if (vcf_out) {
		for (uint32_t i = 0; i < sz; i++) {
			/* a skipped site's gtm is never written by the reference (src/call_genotypes.c:112): report zeros */
			if (w->vcf[i].skip) { memset(vcf_out + i, 0, sizeof(gt_vcf)); vcf_out[i].skip = true; vcf_out[i].ready = true; }
			else vcf_out[i] = w->vcf[i];
		}
	}
	if (pile_out && captured_pileup) memcpy(pile_out, captured_pileup->memory, sizeof(pileup) * sz);
	if (ref_out) memcpy(ref_out, gt_string_get_string(w->ref), sz);
	pthread_mutex_lock(&w->print_mutex);
	w->vcf_n = 0;
	pthread_cond_signal(&w->print_cond2);
	pthread_mutex_unlock(&w->print_mutex);
}

/* Already-normalised templates -> call_genotypes_ML.  refcodes holds codes 0..4 for positions [x, y+2]. */
int bsref_call_block(const bsref_template *t, size_t n, const uint8_t *bases, const uint8_t *refcodes,
		uint32_t x, uint32_t y, pileup *pile_out, gt_vcf *vcf_out) {
	if (!inited || y < x) return -1;
	uint32_t sz = y - x + 1;
	build_list(t, n, bases, NULL);
	gt_string_resize(par.work.ref1, sz + 3);
	memcpy(par.work.ref1->buffer, refcodes, sz + 2);
	par.work.ref1->buffer[sz + 2] = 0;
	par.work.ref1->length = sz + 3;
	call_t0 = now_s();
	call_genotypes_ML(&ctg, align_list, x, y, &par);
	consume(sz, pile_out, vcf_out, NULL);
	drop_list();
	return 0;
}

/* Raw templates -> process_template_vector (trim, soft clips, overlap, indel normalisation, then call).
 * ctg_codes: codes 0..4 for contig positions 1..ctg_len.  Outputs: *x_out = window start; norm_t/norm_bases
 * receive the templates as they look after normalisation (read_off assigned consecutively). */
int bsref_process_block(const bsref_template *t, size_t n, const uint8_t *bases, const bsref_misms *mm,
		const uint8_t *ctg_codes, uint32_t ctg_len, uint32_t y,
		uint32_t *x_out, pileup *pile_out, gt_vcf *vcf_out, uint8_t *ref_out,
		bsref_template *norm_t, uint8_t *norm_bases, size_t norm_cap) {
	if (!inited || n == 0) return -1;
	/* pack the contig the way load_sequence does (src/read_reference.c:62-110): 5 codes per uint16, first in the top bits */
	size_t nw = ((size_t)ctg_len + 9) / 5;
	uint16_t *seq = calloc(nw + 1, sizeof(uint16_t));
	for (uint32_t i = 0; i < ctg_len; i++) seq[i / 5] |= (uint16_t)(ctg_codes[i] & 7) << (3 * (4 - (i % 5)));
	ctg.seq = seq;
	ctg.start_pos = 1;
	ctg.end_pos = ctg_len;
	ctg.seq_len = ctg_len;
	par.work.vcf_ctg = &ctg;
	build_list(t, n, bases, mm);
	uint32_t x = t[0].forward_position ? t[0].forward_position : t[0].reverse_position;
	x = x > 2 ? x - 2 : 1;
	call_t0 = now_s();
	gt_status st = process_template_vector(align_list, &ctg, y, &par);
	int ret = 0;
	if (st != GT_STATUS_OK) ret = -2;
	else {
		uint32_t sz = y - x + 1;
		consume(sz, pile_out, vcf_out, ref_out);
		if (x_out) *x_out = x;
		if (norm_t) {
			size_t off = 0;
			align_details **p = gt_vector_get_mem(align_list, align_details *);
			for (size_t i = 0; i < n; i++) {
				align_details *al = p[i];
				bsref_template *o = norm_t + i;
				memset(o, 0, sizeof(*o));
				o->forward_position = al->forward_position;
				o->reverse_position = al->reverse_position;
				o->orientation = (uint8_t)al->orientation;
				o->bs_strand = (uint8_t)al->bs_strand;
				for (int k = 0; k < 2; k++) {
					o->reference_span[k] = al->reference_span[k];
					o->mapq[k] = al->mapq[k];
					o->present[k] = al->read[k] != NULL;
					o->read_off[k] = (uint32_t)off;
					uint32_t rl = al->read[k] ? (uint32_t)gt_vector_get_used(al->read[k]) : 0;
					o->read_len[k] = rl;
					if (rl) {
						if (off + rl > norm_cap) { ret = -3; rl = 0; o->read_len[k] = 0; }
						else memcpy(norm_bases + off, al->read[k]->memory, rl);
					}
					off += rl;
				}
			}
		}
	}
	drop_list();
	ctg.seq = NULL;
	free(seq);
	return ret;
}

/* CPU baseline of the likelihood microbench (config 2): pre-built pileup records through the reference's model.
 * The reference has no entry point that takes pileup[] directly (call_thread reads the function-static vector
 * that call_genotypes_ML fills from reads), so each record goes through the reference's own calc_gt_prob() and
 * fisher() objects, with the thin summarise / strand-table glue of src/call_genotypes.c:45-59,62-104 supplied by
 * the restatement (bso_summarise, bso_strand_table in bs_oracle.c).  Sites are strided over `nthreads` threads the
 * way the reference strides them over its calc threads (src/call_genotypes.c:259-270). */
#include "bs_oracle.h"

typedef struct { const bso_pileup *tp; const uint8_t *ref; size_t n, first, step; gt_meth *out; uint8_t *skip; } mt_job;

static void *mt_worker(void *arg) {
	mt_job *j = arg;
	for (size_t i = j->first; i < j->n; i += j->step) {
		gt_meth *tg = j->out + i;
		memset(tg, 0, sizeof(*tg));
		if (!j->tp[i].n) { j->skip[i] = 1; continue; }
		bso_summarise(j->tp + i, (bso_gt_meth *)tg);
		calc_gt_prob(tg, &par, (char)j->ref[i]);
		double fs = 0.0;
		int ftab[4];
		if (par.defs.gt_het[tg->max_gt] && bso_strand_table(j->tp + i, tg->max_gt, ftab)) {
			double z = fisher(ftab, par.defs.lfact_store);
			if (z < 1.0e-20) z = 1.0e-20;
			fs = log(z) / LOG10;
		}
		tg->fisher_strand = fs;
		j->skip[i] = 0;
	}
	return NULL;
}

int bsref_call_sites_mt(const void *pile, const uint8_t *ref, size_t n, gt_meth *out, uint8_t *skip, int nthreads) {
	if (!inited) return -1;
	if (nthreads < 1) nthreads = 1;
	mt_job *jobs = malloc(sizeof(mt_job) * nthreads);
	pthread_t *thr = malloc(sizeof(pthread_t) * nthreads);
	for (int i = 0; i < nthreads; i++) {
		jobs[i] = (mt_job){ (const bso_pileup *)pile, ref, n, (size_t)i, (size_t)nthreads, out, skip };
		if (i) pthread_create(thr + i, NULL, mt_worker, jobs + i);
	}
	mt_worker(jobs);
	for (int i = 1; i < nthreads; i++) pthread_join(thr[i], NULL);
	free(jobs);
	free(thr);
	return 0;
}
