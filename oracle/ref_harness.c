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

/* meth-profile ring drainer: same hand-off protocol as src/process.c:20-41 (meth_profile() is a no-op
 * unless bsref_stats_enable(1) was called, src/meth_profile.c:49) */
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

/* ---- --report-file side channels: give the reference a bs_stats the way init_stats does (src/stats.c:304-312; only the
 * meth_profile vector is touched on this path) and read back what meth_profile(), process_template_vector(),
 * trim_soft_clips(), handle_overlap() and read_input() put there ---- */
static bs_stats ref_stats;
void bsref_stats_enable(int on) {
	if (on && !ref_stats.meth_profile) {
		ref_stats.meth_profile = gt_vector_new(256, sizeof(meth_cts));
		memset(ref_stats.meth_profile->memory, 0, sizeof(meth_cts) * ref_stats.meth_profile->elements_allocated);
	}
	par.work.stats = on ? &ref_stats : NULL;
}
void bsref_stats_reset(void) {
	gt_vector *mp = ref_stats.meth_profile;
	memset(&ref_stats, 0, sizeof(ref_stats));
	ref_stats.meth_profile = mp;
	if (mp) {
		memset(mp->memory, 0, sizeof(meth_cts) * mp->elements_allocated);
		gt_vector_set_used(mp, 0);
	}
}
/* conv: room for cap entries of 4 counters; returns `used` of the profile vector */
uint32_t bsref_stats_read(uint64_t *conv, size_t cap, uint64_t base_filter[5], uint64_t filter_cts[15], uint64_t filter_bases[15]) {
	gt_vector *mp = ref_stats.meth_profile;
	const size_t used = mp ? gt_vector_get_used(mp) : 0;
	for (size_t i = 0; i < used && i < cap; i++) memcpy(conv + 4 * i, gt_vector_get_elm(mp, i, meth_cts)->conv_cts, 4 * sizeof(uint64_t));
	memcpy(base_filter, ref_stats.base_filter, sizeof(ref_stats.base_filter));
	memcpy(filter_cts, ref_stats.filter_cts, sizeof(ref_stats.filter_cts));
	memcpy(filter_bases, ref_stats.filter_bases, sizeof(ref_stats.filter_bases));
	return (uint32_t)used;
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

/* =====================================================================================================
 * Reader side: the reference's own BAM-record decode (src/input_sam.c) and block builder
 * (read_input, src/get_template_vector.c), both compiled unmodified.  htslib is replaced by a memory-backed
 * sam_read1() over a buffer of raw BAM alignment records (the byte stream that follows the header in an
 * uncompressed BAM: int32 block_size, 32 bytes of fixed fields, then qname | cigar | seq | qual | aux).
 * ===================================================================================================== */
#include <htslib/sam.h>

static const uint8_t *mem_bam;
static size_t mem_len, mem_pos;

static int32_t rd_i32(const uint8_t *p) { int32_t v; memcpy(&v, p, 4); return v; }
static uint16_t rd_u16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }

int sam_read1(htsFile *fp, bam_hdr_t *h, bam1_t *b) {
	if (mem_pos == mem_len) return -1;
	if (mem_pos + 4 > mem_len) return -2;
	const int32_t bs = rd_i32(mem_bam + mem_pos);
	if (bs < 32 || mem_pos + 4 + (size_t)bs > mem_len) return -2;
	const uint8_t *p = mem_bam + mem_pos + 4;
	bam1_core_t *c = &b->core;
	c->tid = rd_i32(p);
	c->pos = rd_i32(p + 4);
	c->l_qname = p[8];
	c->qual = p[9];
	c->bin = rd_u16(p + 10);
	c->n_cigar = rd_u16(p + 12);
	c->flag = rd_u16(p + 14);
	c->l_qseq = rd_i32(p + 16);
	c->mtid = rd_i32(p + 20);
	c->mpos = rd_i32(p + 24);
	c->isize = rd_i32(p + 28);
	c->l_extranul = 0;
	b->l_data = bs - 32;
	if ((uint32_t)b->l_data > b->m_data) { b->m_data = b->l_data + 64; b->data = realloc(b->data, b->m_data); }
	memcpy(b->data, p + 32, b->l_data);
	mem_pos += 4 + (size_t)bs;
	return b->l_data;
}
int sam_itr_next(htsFile *fp, hts_itr_t *itr, bam1_t *b) { return -1; }
hts_itr_t *sam_itr_queryi(const hts_idx_t *idx, int tid, hts_pos_t beg, hts_pos_t end) { return NULL; }
void hts_itr_destroy(hts_itr_t *itr) { }
bam1_t *bam_init1(void) { return calloc(1, sizeof(bam1_t)); }
void bam_destroy1(bam1_t *b) { if (b) { free(b->data); free(b); } }

int get_next_align_details(htsFile * const sam_input, bam_hdr_t *hdr, hts_itr_t *itr, bam1_t *b, align_details * const al, const uint64_t thresh,
		const uint64_t max_template_len, bool keep_unmatched, bool ignore_dup, bool *reverse, gt_filter_reason * const filtered,
		uint32_t * const align_length, uint32_t * const alignment_flag);

/* per-record view of what get_next_align_details() produced (src/input_sam.c:222-312) */
typedef struct {
	int32_t ret;                 /* 0 keep, 1 filtered */
	uint32_t filtered;           /* gt_filter_reason */
	uint32_t forward_position, reverse_position;
	uint32_t alignment_flag, align_length;
	uint32_t reference_span;     /* of the mate this record is (index = reverse) */
	uint32_t read_off, read_len; /* into bases_out (only when ret == 0) */
	uint32_t mm_off, mm_n;       /* into misms_out */
	uint8_t reverse, orientation, bs_strand, mapq;
} bsref_record;

int bsref_decode_records(const uint8_t *bam, size_t nbytes, int mapq_thresh, uint32_t max_template_len,
		int keep_unmatched, int ignore_dup, bsref_record *out, size_t cap, size_t *nrec,
		uint8_t *bases_out, size_t bases_cap, size_t *nbases, bsref_misms *misms_out, size_t misms_cap, size_t *nmisms) {
	mem_bam = bam; mem_len = nbytes; mem_pos = 0;
	bam1_t *b = bam_init1();
	gt_vector *fl = gt_vector_new(4, sizeof(align_details *));
	size_t n = 0, nb = 0, nm = 0;
	int rc = 0;
	for (;;) {
		align_details *al = get_new_align_details(fl);
		bool reverse = false;
		gt_filter_reason flt;
		uint32_t alen = 0, aflag = 0;
		int ret = get_next_align_details(NULL, NULL, NULL, b, al, mapq_thresh, max_template_len, keep_unmatched, ignore_dup, &reverse, &flt, &alen, &aflag);
		if (ret < 0) { if (ret != -1) rc = -2; break; }
		if (n >= cap) { rc = -3; break; }
		bsref_record *o = out + n++;
		memset(o, 0, sizeof(*o));
		o->ret = ret; o->filtered = flt;
		o->forward_position = al->forward_position; o->reverse_position = al->reverse_position;
		o->alignment_flag = aflag;
		o->reverse = reverse; o->orientation = (uint8_t)al->orientation;
		const int ix = reverse ? 1 : 0;
		o->mapq = al->mapq[ix];
		if (ret == 0) {
			o->align_length = alen;
			o->bs_strand = (uint8_t)al->bs_strand;
			o->reference_span = al->reference_span[ix];
			o->read_off = (uint32_t)nb; o->read_len = (uint32_t)gt_vector_get_used(al->read[ix]);
			if (nb + o->read_len > bases_cap) { rc = -3; break; }
			memcpy(bases_out + nb, al->read[ix]->memory, o->read_len);
			nb += o->read_len;
			o->mm_off = (uint32_t)nm; o->mm_n = (uint32_t)gt_vector_get_used(al->mismatches[ix]);
			if (nm + o->mm_n > misms_cap) { rc = -3; break; }
			gt_misms *mp = gt_vector_get_mem(al->mismatches[ix], gt_misms);
			for (uint32_t z = 0; z < o->mm_n; z++) { misms_out[nm].type = mp[z].misms_type; misms_out[nm].position = mp[z].position; misms_out[nm].size = mp[z].size; nm++; }
		}
		for (int k = 0; k < 2; k++) { if (al->read[k]) gt_vector_delete(al->read[k]); gt_vector_delete(al->mismatches[k]); }
		free(al);
	}
	bam_destroy1(b);
	gt_vector_delete(fl);
	*nrec = n; *nbases = nb; *nmisms = nm;
	return rc;
}

/* ---- read_input() with a stand-in for process_thread (src/process.c:43-72) that snapshots every block ---- */
typedef struct { uint32_t tid, x, y, first_template, n_templates, pad; uint64_t vcf_off; } bsref_block;

typedef struct {
	bsref_block *blocks; size_t block_cap, nblocks;
	bsref_template *tmpl; size_t tmpl_cap, ntmpl;
	uint8_t *bases; size_t bases_cap, nbases;
	bsref_misms *misms; size_t misms_cap, nmisms;
	gt_vcf *vcf; size_t vcf_cap, nvcf;
	int run_chain;               /* also run process_template_vector + call_genotypes_ML on each block */
	int err;
} reader_out;

static reader_out rdr;

static void snapshot_block(gt_vector *alist, ctg_t *c, uint32_t y) {
	const size_t n = gt_vector_get_used(alist);
	align_details **p = gt_vector_get_mem(alist, align_details *);
	if (rdr.nblocks >= rdr.block_cap || rdr.ntmpl + n > rdr.tmpl_cap) { rdr.err = -3; return; }
	bsref_block *bk = rdr.blocks + rdr.nblocks++;
	memset(bk, 0, sizeof(*bk));
	bk->tid = (uint32_t)c->bam_tid; bk->y = y; bk->first_template = (uint32_t)rdr.ntmpl; bk->n_templates = (uint32_t)n;
	uint32_t x = p[0]->forward_position ? p[0]->forward_position : p[0]->reverse_position;
	bk->x = x > 2 ? x - 2 : 1;
	bk->vcf_off = rdr.nvcf;
	for (size_t i = 0; i < n; i++) {
		align_details *al = p[i];
		bsref_template *o = rdr.tmpl + rdr.ntmpl++;
		memset(o, 0, sizeof(*o));
		o->forward_position = al->forward_position; o->reverse_position = al->reverse_position;
		o->orientation = (uint8_t)al->orientation; o->bs_strand = (uint8_t)al->bs_strand;
		for (int k = 0; k < 2; k++) {
			o->reference_span[k] = al->reference_span[k];
			o->mapq[k] = al->mapq[k];
			o->present[k] = al->read[k] != NULL;
			const uint32_t rl = al->read[k] ? (uint32_t)gt_vector_get_used(al->read[k]) : 0;
			const uint32_t nm = rl ? (uint32_t)gt_vector_get_used(al->mismatches[k]) : 0;
			if (rdr.nbases + rl > rdr.bases_cap || rdr.nmisms + nm > rdr.misms_cap) { rdr.err = -3; return; }
			o->read_off[k] = (uint32_t)rdr.nbases; o->read_len[k] = rl;
			if (rl) memcpy(rdr.bases + rdr.nbases, al->read[k]->memory, rl);
			rdr.nbases += rl;
			o->mm_off[k] = (uint32_t)rdr.nmisms; o->mm_n[k] = nm;
			gt_misms *mp = gt_vector_get_mem(al->mismatches[k], gt_misms);
			for (uint32_t z = 0; z < nm; z++) { rdr.misms[rdr.nmisms].type = mp[z].misms_type; rdr.misms[rdr.nmisms].position = mp[z].position; rdr.misms[rdr.nmisms].size = mp[z].size; rdr.nmisms++; }
		}
	}
	if (rdr.run_chain) {
		const uint32_t sz = y - bk->x + 1;
		if (rdr.nvcf + sz > rdr.vcf_cap) { rdr.err = -3; return; }
		par.work.vcf_ctg = c;
		call_t0 = now_s();
		if (process_template_vector(alist, c, y, &par) != GT_STATUS_OK) { rdr.err = -2; return; }
		consume(sz, NULL, rdr.vcf + rdr.nvcf, NULL);
		rdr.nvcf += sz;
	}
}

static void *reader_process_thread(void *arg) {
	gt_vector *prev_align = NULL;
	work_t * const work = &par.work;
	while (true) {
		pthread_mutex_lock(&work->process_mutex);
		while (!work->align_list_waiting && !work->process_end) {
			struct timespec ts;
			clock_gettime(CLOCK_REALTIME, &ts);
			ts.tv_sec += 1;
			pthread_cond_timedwait(&work->process_cond1, &work->process_mutex, &ts);
		}
		pthread_mutex_unlock(&work->process_mutex);
		gt_vector *alist = work->align_list_waiting;
		if (alist == NULL) break;
		ctg_t *c = work->ctg_waiting;
		const uint32_t pos = work->y_waiting;
		work->ctg_waiting = NULL;
		work->free_list_waiting = prev_align;
		work->align_list_waiting = NULL;
		pthread_mutex_lock(&work->process_mutex);
		pthread_cond_signal(&work->process_cond2);
		pthread_mutex_unlock(&work->process_mutex);
		if (!rdr.err) snapshot_block(alist, c, pos);
		prev_align = alist;
	}
	return NULL;
}

gt_status read_input(htsFile *sam_input, gt_vector * align_list, sr_param *param);

/* ctg_codes[tid] (may be NULL when run_chain == 0): reference codes 0..4 for positions 1..target_len[tid] */
int bsref_read_input(const uint8_t *bam, size_t nbytes, int n_targets, const uint32_t *target_len, const uint8_t *const *ctg_codes,
		int mapq_thresh, uint32_t max_template_len, int keep_unmatched, int ignore_duplicates, int keep_duplicates, int run_chain,
		bsref_block *blocks, size_t block_cap, size_t *nblocks, bsref_template *tmpl, size_t tmpl_cap, size_t *ntmpl,
		uint8_t *bases, size_t bases_cap, size_t *nbases, bsref_misms *misms, size_t misms_cap, size_t *nmisms,
		gt_vcf *vcf, size_t vcf_cap, size_t *nvcf) {
	if (!inited) return -1;
	mem_bam = bam; mem_len = nbytes; mem_pos = 0;
	memset(&rdr, 0, sizeof(rdr));
	rdr.blocks = blocks; rdr.block_cap = block_cap; rdr.tmpl = tmpl; rdr.tmpl_cap = tmpl_cap; rdr.bases = bases; rdr.bases_cap = bases_cap;
	rdr.misms = misms; rdr.misms_cap = misms_cap; rdr.vcf = vcf; rdr.vcf_cap = vcf_cap; rdr.run_chain = run_chain;
	work_t * const w = &par.work;
	bam_hdr_t hdr;
	memset(&hdr, 0, sizeof(hdr));
	hdr.n_targets = n_targets;
	hdr.target_len = (uint32_t *)target_len;
	hdr.target_name = calloc(n_targets, sizeof(char *));
	ctg_t *ctgs = calloc(n_targets, sizeof(ctg_t));
	w->contigs = calloc(n_targets, sizeof(ctg_t *));
	w->tid2id = calloc(n_targets, sizeof(int));
	for (int i = 0; i < n_targets; i++) {
		char nm[32];
		snprintf(nm, sizeof(nm), "ctg%d", i);
		hdr.target_name[i] = strdup(nm);
		ctgs[i].name = hdr.target_name[i];
		ctgs[i].bam_tid = i;
		if (run_chain) {
			const uint32_t len = target_len[i];
			const size_t nw = ((size_t)len + 9) / 5;
			ctgs[i].seq = calloc(nw + 1, sizeof(uint16_t));
			for (uint32_t j = 0; j < len; j++) ctgs[i].seq[j / 5] |= (uint16_t)(ctg_codes[i][j] & 7) << (3 * (4 - (j % 5)));
			ctgs[i].start_pos = 1; ctgs[i].end_pos = len; ctgs[i].seq_len = len;
		}
		w->contigs[i] = ctgs + i;
		w->tid2id[i] = i;
	}
	w->n_contigs = n_targets; w->n_regions = 0; w->sam_idx = NULL; w->sam_header = &hdr; w->curr_region = NULL;
	w->process_end = false; w->align_list_waiting = NULL; w->free_list_waiting = NULL; w->ctg_waiting = NULL;
	par.mapq_thresh = (uint8_t)mapq_thresh; par.max_template_len = max_template_len;
	par.keep_unmatched = keep_unmatched; par.ignore_duplicates = ignore_duplicates; par.keep_duplicates = keep_duplicates;
	pthread_t thr;
	pthread_create(&thr, NULL, reader_process_thread, NULL);
	gt_vector *al_list = gt_vector_new(32, sizeof(align_details *));
	gt_status st = read_input(NULL, al_list, &par);
	w->process_end = true;
	pthread_mutex_lock(&w->process_mutex);
	pthread_cond_broadcast(&w->process_cond1);
	pthread_mutex_unlock(&w->process_mutex);
	pthread_join(thr, NULL);
	*nblocks = rdr.nblocks; *ntmpl = rdr.ntmpl; *nbases = rdr.nbases; *nmisms = rdr.nmisms; *nvcf = rdr.nvcf;
	for (int i = 0; i < n_targets; i++) { free(hdr.target_name[i]); free(ctgs[i].seq); }
	free(hdr.target_name); free(w->contigs); free(w->tid2id); free(ctgs);
	w->contigs = NULL; w->tid2id = NULL; w->sam_header = NULL;
	if (st != GT_STATUS_OK) return -2;
	return rdr.err;
}

#ifdef BSREF_SEAM_READER
/* ---- seams C / D (bs_call_b200/csrc/bsgpu_seam_reader.c linked in place of get_template_vector.c, process_template.c and
 * call_genotypes.c): read_input() is the product's, fed by the memory-backed sam_read1() above.  The harness stands in for
 * what src/process.c runs around it: a print thread that takes every published block off work->vcf (seam C), and the
 * capture buffer behind bcf_write() (seam D, BSGPU_SEAM_RECORDS=1 in the environment). ---- */
static uint8_t *pv_out;
static size_t pv_cap, pv_len, pv_nrec;
static int pv_overflow;
static struct { bsref_block *blocks; size_t cap, n; gt_vcf *vcf; size_t vcf_cap, nvcf; int err; } seam;
static gt_ctg_stats *seam_ctg_sum;       /* sum of the contigs' ctg_stats after the last run with statistics */

static void *seam_print_thread(void *arg) {
	work_t * const w = &par.work;
	for (;;) {
		pthread_mutex_lock(&w->print_mutex);
		while (!w->vcf_n && !w->print_end) {
			struct timespec ts;
			clock_gettime(CLOCK_REALTIME, &ts);
			ts.tv_sec += 1;
			pthread_cond_timedwait(&w->print_cond1, &w->print_mutex, &ts);
		}
		pthread_mutex_unlock(&w->print_mutex);
		if (!w->vcf_n) break;
		const uint32_t sz = (uint32_t)w->vcf_n;
		if (seam.n >= seam.cap || seam.nvcf + sz > seam.vcf_cap) seam.err = -3;
		else {
			bsref_block *bk = seam.blocks + seam.n++;
			memset(bk, 0, sizeof(*bk));
			bk->tid = (uint32_t)w->vcf_ctg->bam_tid; bk->x = w->vcf_x; bk->y = w->vcf_x + sz - 1; bk->vcf_off = seam.nvcf;
			/* the block's reference string must be the codes of [x, y + 2] (what the writer would read) */
			const char *rf = gt_string_get_string(w->ref);
			const ctg_t *c = w->vcf_ctg;
			for (uint32_t i = 0; i < sz + 2 && !seam.err; i++) {
				const uint64_t pos = (uint64_t)w->vcf_x + i;
				char want = 0;
				if (pos < c->end_pos) want = (char)((c->seq[(pos - 1) / 5] >> (3 * (4 - ((pos - 1) % 5)))) & 7);
				if (rf[i] != want) seam.err = -7;
			}
			for (uint32_t i = 0; i < sz; i++) {
				if (!w->vcf[i].ready) seam.err = -6;
				if (w->vcf[i].skip) { memset(seam.vcf + seam.nvcf + i, 0, sizeof(gt_vcf)); seam.vcf[seam.nvcf + i].skip = true; seam.vcf[seam.nvcf + i].ready = true; }
				else seam.vcf[seam.nvcf + i] = w->vcf[i];
			}
			seam.nvcf += sz;
		}
		pthread_mutex_lock(&w->print_mutex);
		w->vcf_n = 0;
		pthread_cond_signal(&w->print_cond2);
		pthread_mutex_unlock(&w->print_mutex);
	}
	return NULL;
}

int bsref_seam_read_input(const uint8_t *bam, size_t nbytes, int n_targets, const uint32_t *target_len, const uint8_t *const *ctg_codes,
		int mapq_thresh, uint32_t max_template_len, int keep_unmatched, int ignore_duplicates, int keep_duplicates,
		const int *vcf_ids, int all_positions,
		bsref_block *blocks, size_t block_cap, size_t *nblocks, gt_vcf *vcf, size_t vcf_cap, size_t *nvcf,
		uint8_t *bcf_out, size_t bcf_cap, size_t *bcf_bytes, size_t *bcf_nrec) {
	if (!inited) return -1;
	mem_bam = bam; mem_len = nbytes; mem_pos = 0;
	memset(&seam, 0, sizeof(seam));
	seam.blocks = blocks; seam.cap = block_cap; seam.vcf = vcf; seam.vcf_cap = vcf_cap;
	pv_out = bcf_out; pv_cap = bcf_cap; pv_len = 0; pv_nrec = 0; pv_overflow = 0;
	work_t * const w = &par.work;
	bam_hdr_t hdr;
	memset(&hdr, 0, sizeof(hdr));
	hdr.n_targets = n_targets;
	hdr.target_len = (uint32_t *)target_len;
	hdr.target_name = calloc(n_targets, sizeof(char *));
	ctg_t *ctgs = calloc(n_targets, sizeof(ctg_t));
	w->contigs = calloc(n_targets, sizeof(ctg_t *));
	w->tid2id = calloc(n_targets, sizeof(int));
	for (int i = 0; i < n_targets; i++) {
		char nm[32];
		snprintf(nm, sizeof(nm), "ctg%d", i);
		hdr.target_name[i] = strdup(nm);
		ctgs[i].name = hdr.target_name[i];
		ctgs[i].bam_tid = i;
		ctgs[i].vcf_rid = i;
		const uint32_t len = target_len[i];
		const size_t nw = ((size_t)len + 9) / 5;
		ctgs[i].seq = calloc(nw + 1, sizeof(uint16_t));
		for (uint32_t j = 0; j < len; j++) ctgs[i].seq[j / 5] |= (uint16_t)(ctg_codes[i][j] & 7) << (3 * (4 - (j % 5)));
		ctgs[i].start_pos = 1; ctgs[i].end_pos = len; ctgs[i].seq_len = len;
		w->contigs[i] = ctgs + i;
		w->tid2id[i] = i;
		if (w->stats != NULL) {
			/* --report-file: the contig's statistics block with GC bins (src/read_reference.c:120-123 computes them while it loads
			 * the sequence; here: percentage of C / G among the known bases of every 100-base bin) */
			gt_ctg_stats *cs = calloc(1, sizeof(gt_ctg_stats));
			cs->nbins = (int)((len + 99) / 100);
			cs->gc = malloc((size_t)cs->nbins + 1);
			for (int b = 0; b < cs->nbins; b++) {
				uint32_t known = 0, gcn = 0;
				for (uint32_t j = (uint32_t)b * 100; j < len && j < (uint32_t)b * 100 + 100; j++) { const int c = ctg_codes[i][j] & 7; known += c != 0; gcn += c == 2 || c == 3; }
				cs->gc[b] = known ? (uint8_t)(100 * gcn / known) : 255;
			}
			ctgs[i].ctg_stats = cs;
		}
	}
	if (w->stats != NULL && w->stats->qd_stats == NULL) {          /* what init_stats() sets up (src/stats.c:304-318) */
		w->stats->qd_stats = gt_vector_new(256, sizeof(fstats_cts));
		w->stats->fs_stats = gt_vector_new(256, sizeof(fstats_cts));
		w->stats->mq_stats = gt_vector_new(256, sizeof(fstats_cts));
		memset(w->stats->qd_stats->memory, 0, sizeof(fstats_cts) * w->stats->qd_stats->elements_allocated);
		memset(w->stats->fs_stats->memory, 0, sizeof(fstats_cts) * w->stats->fs_stats->elements_allocated);
		memset(w->stats->mq_stats->memory, 0, sizeof(fstats_cts) * w->stats->mq_stats->elements_allocated);
	}
	w->n_contigs = n_targets; w->n_regions = 0; w->sam_idx = NULL; w->sam_header = &hdr; w->curr_region = NULL;
	w->process_end = false; w->print_end = false; w->vcf_n = 0; w->vcf_ctg = NULL;
	gt_vcf * const saved_vcf = w->vcf;
	const int saved_size = w->vcf_size;
	par.mapq_thresh = (uint8_t)mapq_thresh; par.max_template_len = max_template_len;
	par.keep_unmatched = keep_unmatched; par.ignore_duplicates = ignore_duplicates; par.keep_duplicates = keep_duplicates;
	par.all_positions = all_positions;
	for (int i = 0; i < 16; i++) w->vcf_ids[i] = vcf_ids[i];
	pthread_t thr;
	pthread_create(&thr, NULL, seam_print_thread, NULL);
	gt_vector *al_list = gt_vector_new(32, sizeof(align_details *));
	gt_status st = read_input(NULL, al_list, &par);
	pthread_mutex_lock(&w->print_mutex);
	w->print_end = true;
	pthread_cond_signal(&w->print_cond1);
	pthread_mutex_unlock(&w->print_mutex);
	pthread_join(thr, NULL);
	*nblocks = seam.n; *nvcf = seam.nvcf; *bcf_bytes = pv_len; *bcf_nrec = pv_nrec;
	w->vcf = saved_vcf; w->vcf_size = saved_size;          /* work->vcf pointed into the session's results: gone now */
	if (w->stats != NULL) {
		/* end of the run as far as the side channels go: the seam folds what the device gathered into bs_stats and the contigs'
		 * ctg_stats (main() calls join_calc_threads before it writes the report); keep the last contig's counters readable */
		static gt_ctg_stats seam_last;
		join_calc_threads(&par);
		init_calc_threads(&par);
		memset(&seam_last, 0, sizeof(seam_last));
		for (int i = 0; i < n_targets; i++) {
			const gt_ctg_stats *c = ctgs[i].ctg_stats;
			for (int k = 0; k < 2; k++) {
				seam_last.snps[k] += c->snps[k]; seam_last.multi[k] += c->multi[k]; seam_last.dbSNP_sites[k] += c->dbSNP_sites[k];
				seam_last.dbSNP_var[k] += c->dbSNP_var[k]; seam_last.CpG_ref[k] += c->CpG_ref[k]; seam_last.CpG_nonref[k] += c->CpG_nonref[k];
			}
		}
		seam_ctg_sum = &seam_last;
		for (int i = 0; i < n_targets; i++) { if (ctgs[i].ctg_stats->gc) free(ctgs[i].ctg_stats->gc); free(ctgs[i].ctg_stats); }
	}
	for (int i = 0; i < n_targets; i++) { free(hdr.target_name[i]); free(ctgs[i].seq); }
	free(hdr.target_name); free(w->contigs); free(w->tid2id); free(ctgs);
	w->contigs = NULL; w->tid2id = NULL; w->sam_header = NULL;
	if (st != GT_STATUS_OK) return -2;
	if (pv_overflow) return -3;
	return seam.err;
}
#endif

/* ------------------------------------------------------------------------------------------------------------------
 * Writer side: the reference's print_vcf_entry() / flush_vcf_entries() / _print_vcf_entry() (src/print_vcf.c, compiled
 * unmodified) driven over a block of gt_vcf records the way print_thread drives them (src/process.c:89-104), with the
 * htslib functions that file calls supplied here.
 *
 * htslib is a system dependency of bs_call (configure.ac:9-14; any 1.x release with the bcf_enc_* API), not vendored
 * under /root/reference and absent from this image.  What follows restates the part of its PUBLISHED behaviour the
 * writer relies on -- the BCF2 typed-value encoding and record layout of the VCF/BCF specification v4.3, section 6.3
 * -- under the API names print_vcf.c calls.  bcf_write() appends the record, laid out as in a BCF file, to a buffer.
 * ------------------------------------------------------------------------------------------------------------------ */
#include <htslib/vcf.h>
#include <htslib/khash.h>
#include "dbSNP.h"

void bcf_enc_size(kstring_t *s, int size, int type) {
	if (size >= 15) {
		kputc(15 << 4 | type, s);
		if (size >= 128) {
			if (size >= 32768) { int32_t x = size; kputc(1 << 4 | BCF_BT_INT32, s); kputsn((char *)&x, 4, s); }
			else { int16_t x = (int16_t)size; kputc(1 << 4 | BCF_BT_INT16, s); kputsn((char *)&x, 2, s); }
		} else { kputc(1 << 4 | BCF_BT_INT8, s); kputc(size, s); }
	} else kputc(size << 4 | type, s);
}
void bcf_enc_int1(kstring_t *s, int32_t x) {
	if (x == bcf_int32_vector_end) { bcf_enc_size(s, 1, BCF_BT_INT8); kputc(bcf_int8_vector_end, s); }
	else if (x == bcf_int32_missing) { bcf_enc_size(s, 1, BCF_BT_INT8); kputc(bcf_int8_missing, s); }
	else if (x <= BCF_MAX_BT_INT8 && x >= BCF_MIN_BT_INT8) { bcf_enc_size(s, 1, BCF_BT_INT8); kputc(x, s); }
	else if (x <= BCF_MAX_BT_INT16 && x >= BCF_MIN_BT_INT16) { int16_t z = (int16_t)x; bcf_enc_size(s, 1, BCF_BT_INT16); kputsn((char *)&z, 2, s); }
	else { int32_t z = x; bcf_enc_size(s, 1, BCF_BT_INT32); kputsn((char *)&z, 4, s); }
}
void bcf_enc_vint(kstring_t *s, int n, int32_t *a, int wsize) {
	if (n <= 0) { bcf_enc_size(s, 0, BCF_BT_NULL); return; }
	if (n == 1) { bcf_enc_int1(s, a[0]); return; }
	int32_t max = INT32_MIN + 1, min = INT32_MAX;
	if (wsize <= 0) wsize = n;
	for (int i = 0; i < n; i++) {
		if (a[i] == bcf_int32_missing || a[i] == bcf_int32_vector_end) continue;
		if (max < a[i]) max = a[i];
		if (min > a[i]) min = a[i];
	}
	if (max <= BCF_MAX_BT_INT8 && min >= BCF_MIN_BT_INT8) {
		bcf_enc_size(s, wsize, BCF_BT_INT8);
		for (int i = 0; i < n; i++) kputc(a[i] == bcf_int32_vector_end ? bcf_int8_vector_end : a[i] == bcf_int32_missing ? bcf_int8_missing : a[i], s);
	} else if (max <= BCF_MAX_BT_INT16 && min >= BCF_MIN_BT_INT16) {
		bcf_enc_size(s, wsize, BCF_BT_INT16);
		for (int i = 0; i < n; i++) {
			int16_t z = a[i] == bcf_int32_vector_end ? bcf_int16_vector_end : a[i] == bcf_int32_missing ? bcf_int16_missing : (int16_t)a[i];
			kputsn((char *)&z, 2, s);
		}
	} else {
		bcf_enc_size(s, wsize, BCF_BT_INT32);
		for (int i = 0; i < n; i++) { int32_t z = a[i]; kputsn((char *)&z, 4, s); }
	}
}
void bcf_enc_vfloat(kstring_t *s, int n, float *a) {
	bcf_enc_size(s, n, BCF_BT_FLOAT);
	kputsn((char *)a, (size_t)n << 2, s);          /* little-endian host */
}
void bcf_enc_vchar(kstring_t *s, int l, const char *a) {
	bcf_enc_size(s, l, BCF_BT_CHAR);
	kputsn(a, l, s);
}
bcf1_t *bcf_init(void) { return calloc(1, sizeof(bcf1_t)); }
void bcf_destroy(bcf1_t *v) { if (v) { free(v->shared.s); free(v->indiv.s); free(v); } }
void bcf_clear(bcf1_t *v) {
	v->rid = 0; v->pos = 0; v->rlen = 0;
	{ uint32_t miss = 0x7F800001u; memcpy(&v->qual, &miss, 4); }          /* bcf_float_missing */
	v->n_info = v->n_allele = v->n_fmt = v->n_sample = 0;
	v->shared.l = v->indiv.l = 0;
}

/* capture buffer of bcf_write(): records as they lie in a BCF file (two length words, six fixed words, shared, indiv) */
#ifndef BSREF_SEAM_READER
static uint8_t *pv_out;
static size_t pv_cap, pv_len, pv_nrec;
static int pv_overflow;
#endif
int bcf_write(htsFile *fp, bcf_hdr_t *h, bcf1_t *v) {
	uint32_t x[8];
	x[0] = (uint32_t)v->shared.l + 24;
	x[1] = (uint32_t)v->indiv.l;
	x[2] = (uint32_t)v->rid;
	x[3] = (uint32_t)v->pos;
	x[4] = (uint32_t)v->rlen;
	memcpy(x + 5, &v->qual, 4);
	x[6] = (uint32_t)v->n_allele << 16 | v->n_info;
	x[7] = (uint32_t)v->n_fmt << 24 | v->n_sample;
	const size_t need = 32 + v->shared.l + v->indiv.l;
	if (pv_len + need > pv_cap) { pv_overflow = 1; return 0; }
	memcpy(pv_out + pv_len, x, 32);
	memcpy(pv_out + pv_len + 32, v->shared.s, v->shared.l);
	memcpy(pv_out + pv_len + 32 + v->shared.l, v->indiv.s, v->indiv.l);
	pv_len += need;
	pv_nrec++;
	return 0;
}

/* header side and dbSNP: never reached from the harness (print_vcf_header is not called, par.work.dbSNP_hdr is NULL) */
static void pv_unreachable(const char *what) { fprintf(stderr, "ref_harness: %s called\n", what); abort(); }
bcf_hdr_t *bcf_hdr_init(const char *mode) { pv_unreachable("bcf_hdr_init"); return NULL; }
int bcf_hdr_append(bcf_hdr_t *h, const char *line) { pv_unreachable("bcf_hdr_append"); return -1; }
int bcf_hdr_printf(bcf_hdr_t *h, const char *format, ...) { pv_unreachable("bcf_hdr_printf"); return -1; }
const char *bcf_hdr_get_version(const bcf_hdr_t *hdr) { pv_unreachable("bcf_hdr_get_version"); return NULL; }
int bcf_hdr_add_sample(bcf_hdr_t *hdr, const char *sample) { pv_unreachable("bcf_hdr_add_sample"); return -1; }
int bcf_hdr_write(htsFile *fp, bcf_hdr_t *h) { pv_unreachable("bcf_hdr_write"); return -1; }
htsFile *hts_open(const char *fn, const char *mode) { pv_unreachable("hts_open"); return NULL; }
int hts_set_threads(htsFile *fp, int n) { pv_unreachable("hts_set_threads"); return -1; }
int bam_name2id(bam_hdr_t *h, const char *ref) { pv_unreachable("bam_name2id"); return -1; }
khint_t bsstub_kh_get(const void *h, const char *key) { pv_unreachable("kh_get"); return 0; }
khint_t bsstub_kh_end(const void *h) { pv_unreachable("kh_end"); return 0; }
/* dbSNP: the index file and its reader (src/dbSNP.c) are host I/O outside the path; what the path's writer sees of them is
 * the answer of dbSNP_lookup_name() per position -- flags (1 known, 3 known and always written) and the ID bytes -- which
 * the harness serves from a table the test supplies (bsref_print_block_ann).  print_vcf.c's use of the answer
 * (src/print_vcf.c:133, 139, 163-167) is the compiled reference's. */
static struct { uint32_t n; const uint32_t *pos; const uint8_t *flags; const uint32_t *name_off; const uint8_t *names; } db_tab;
uint8_t dbSNP_lookup_name(const dbsnp_header_t *const hdr, const dbsnp_ctg_t *c, char *const rs, size_t *const rs_len, const uint32_t x) {
	rs[0] = 0;
	uint32_t lo = 0, hi = db_tab.n;
	while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (db_tab.pos[mid] < x) lo = mid + 1; else hi = mid; }
	if (lo == db_tab.n || db_tab.pos[lo] != x) return 0;
	const uint32_t l = db_tab.name_off[lo + 1] - db_tab.name_off[lo];
	memcpy(rs, db_tab.names + db_tab.name_off[lo], l);
	rs[l] = 0;
	if (rs_len) *rs_len = l;
	return db_tab.flags[lo];
}
bool load_dbSNP_ctg(const dbsnp_header_t *const hdr, dbsnp_ctg_t *const c) { return true; }
void unload_dbSNP_ctg(dbsnp_ctg_t *const c) { }

/* ---- the writer's --report-file statistics (src/print_vcf.c:382-526): bsref_writer_stats(1, gc, nbins, start_pos) makes the
 * blocks printed from then on count into the harness's bs_stats (with the GC bins given as the contig's);
 * bsref_writer_stats_read() flattens what they left there into a bso_site_stats (vectors and the coverage hash by value) ---- */
static int pv_stats_on, pv_nbins;
static const uint8_t *pv_gc;
static uint32_t pv_start = 1;
void bsref_writer_stats(int on, const uint8_t *gc, int nbins, uint32_t start_pos) {
	pv_stats_on = on; pv_gc = gc; pv_nbins = gc ? nbins : 0; pv_start = start_pos ? start_pos : 1;
}
void bsref_writer_stats_reset(void) {
	gt_cov_stats *g, *tmp;
	HASH_ITER(hh, ref_stats.cov_stats, g, tmp) { HASH_DEL(ref_stats.cov_stats, g); free(g); }
	gt_vector *keep[4] = { ref_stats.meth_profile, ref_stats.qd_stats, ref_stats.fs_stats, ref_stats.mq_stats };
	for (int k = 1; k < 4; k++) if (keep[k]) { memset(keep[k]->memory, 0, sizeof(fstats_cts) * keep[k]->elements_allocated); keep[k]->used = 0; }
	const size_t head = offsetof(bs_stats, qual) + sizeof(ref_stats.qual);      /* snps .. qual; the read-level tallies after them stay */
	memset(&ref_stats, 0, head);
	memset(ref_stats.filter_counts, 0, sizeof(ref_stats.filter_counts));
	memset(ref_stats.CpG_ref_meth, 0, sizeof(ref_stats.CpG_ref_meth));
	memset(ref_stats.CpG_nonref_meth, 0, sizeof(ref_stats.CpG_nonref_meth));
}
static void pv_flat_vec(const gt_vector *v, uint64_t (*out)[2], int n, uint64_t *overflow) {
	if (!v) return;
	for (uint64_t i = 0; i < v->used; i++) {
		const fstats_cts *c = gt_vector_get_elm(v, i, fstats_cts);
		if (i < (uint64_t)n) { out[i][0] = c->cts[0]; out[i][1] = c->cts[1]; }
		else if (overflow) *overflow += c->cts[0] + c->cts[1];
	}
}
void bsref_writer_stats_read(bso_site_stats *o) {
	memset(o, 0, sizeof(*o));
	memcpy(o->snps, ref_stats.snps, sizeof(o->snps)); memcpy(o->multi, ref_stats.multi, sizeof(o->multi));
	memcpy(o->dbSNP_sites, ref_stats.dbSNP_sites, sizeof(o->dbSNP_sites)); memcpy(o->dbSNP_var, ref_stats.dbSNP_var, sizeof(o->dbSNP_var));
	memcpy(o->CpG_ref, ref_stats.CpG_ref, sizeof(o->CpG_ref)); memcpy(o->CpG_nonref, ref_stats.CpG_nonref, sizeof(o->CpG_nonref));
	memcpy(o->mut_counts, ref_stats.mut_counts, sizeof(o->mut_counts)); memcpy(o->dbSNP_mut_counts, ref_stats.dbSNP_mut_counts, sizeof(o->dbSNP_mut_counts));
	memcpy(o->qual, ref_stats.qual, sizeof(o->qual));
	memcpy(o->filter_counts, ref_stats.filter_counts, sizeof(o->filter_counts));
	memcpy(o->CpG_ref_meth, ref_stats.CpG_ref_meth, sizeof(o->CpG_ref_meth)); memcpy(o->CpG_nonref_meth, ref_stats.CpG_nonref_meth, sizeof(o->CpG_nonref_meth));
	pv_flat_vec(ref_stats.qd_stats, o->qd_stats, 256, NULL);
	pv_flat_vec(ref_stats.mq_stats, o->mq_stats, 256, NULL);
	pv_flat_vec(ref_stats.fs_stats, o->fs_stats, BSO_STATS_FS_MAX, &o->fs_overflow);
	gt_cov_stats *g, *tmp;
	HASH_ITER(hh, ref_stats.cov_stats, g, tmp) {
		if (g->coverage < BSO_STATS_COV_MAX) {
			bso_cov_stats *d = o->cov + g->coverage;
			d->var = g->var; d->CpG[0] = g->CpG[0]; d->CpG[1] = g->CpG[1]; d->CpG_inf[0] = g->CpG_inf[0]; d->CpG_inf[1] = g->CpG_inf[1];
			d->all = g->all; memcpy(d->gc_pcent, g->gc_pcent, sizeof(d->gc_pcent));
		} else o->cov_overflow += g->all + g->CpG_inf[0] + g->CpG_inf[1];
	}
}
/* per-contig counters of the block printed last (ctg_stats, include/bs_call.h:75-85): snps, multi, dbSNP_sites, dbSNP_var, CpG_ref, CpG_nonref */
static gt_ctg_stats *pv_last_cstats;
void bsref_writer_ctg_stats_read(uint64_t out[12]) {
	memset(out, 0, 12 * sizeof(uint64_t));
#ifdef BSREF_SEAM_READER
	if (seam_ctg_sum) pv_last_cstats = seam_ctg_sum;
#endif
	if (!pv_last_cstats) return;
	const gt_ctg_stats *c = pv_last_cstats;
	const uint64_t *src[6] = { c->snps, c->multi, c->dbSNP_sites, c->dbSNP_var, c->CpG_ref, c->CpG_nonref };
	for (int k = 0; k < 6; k++) { out[2 * k] = src[k][0]; out[2 * k + 1] = src[k][1]; }
}

void print_vcf_entry(bcf1_t *bcf, ctg_t * const ctg, gt_meth *gtm, const char *rf, const uint32_t x, const uint32_t xstart, bool skip, sr_param * const par);
void flush_vcf_entries(bcf1_t *bcf, const sr_param * const par);

/* One block through the reference's writer, as print_thread runs it (src/process.c:89-104).
 * refcodes: sz + 2 codes (positions x .. x + sz + 1), the string get_sequence_string() leaves in work.ref.
 * vcf_ids: the 16 dictionary ids print_vcf_header() looks up (include/bs_call.h:192-208).
 * out receives the records bcf_write() was handed, in BCF layout; returns 0, or -3 when out is too small. */
int bsref_print_block_ann(const gt_vcf *vcf, uint32_t sz, const uint8_t *refcodes, uint32_t x, int rid, uint32_t ctg_end,
		const int *vcf_ids, int all_positions, uint32_t reg_start, uint32_t reg_stop,
		uint32_t db_n, const uint32_t *db_pos, const uint8_t *db_flags, const uint32_t *db_name_off, const uint8_t *db_names,
		uint8_t *out, size_t cap, size_t *nbytes, size_t *nrec);

int bsref_print_block(const gt_vcf *vcf, uint32_t sz, const uint8_t *refcodes, uint32_t x, int rid, uint32_t ctg_end,
		const int *vcf_ids, int all_positions, uint8_t *out, size_t cap, size_t *nbytes, size_t *nrec) {
	return bsref_print_block_ann(vcf, sz, refcodes, x, rid, ctg_end, vcf_ids, all_positions, 0, 0, 0, NULL, NULL, NULL, NULL, out, cap, nbytes, nrec);
}

/* the same with ctg->curr_reg = [reg_start, reg_stop] (0, 0: none; what -C / a region list sets, src/get_template_vector.c:123)
 * and a dbSNP index that knows the db_n positions given (-D) */
int bsref_print_block_ann(const gt_vcf *vcf, uint32_t sz, const uint8_t *refcodes, uint32_t x, int rid, uint32_t ctg_end,
		const int *vcf_ids, int all_positions, uint32_t reg_start, uint32_t reg_stop,
		uint32_t db_n, const uint32_t *db_pos, const uint8_t *db_flags, const uint32_t *db_name_off, const uint8_t *db_names,
		uint8_t *out, size_t cap, size_t *nbytes, size_t *nrec) {
	static ctg_t pv_ctg[2];
	static gt_ctg_stats pv_cstats[2];
	static region_t pv_reg;
	static int flip = 0;
	static bcf1_t *bcf = NULL;
	static dbsnp_header_t db_hdr;
	static dbsnp_ctg_t db_ctg;
	if (!inited) return -1;
	if (!bcf) bcf = bcf_init();
	/* a different ctg_t object every call: _print_vcf_entry then forgets the last position it printed (:117-127) */
	ctg_t *c = pv_ctg + (flip ^= 1);
	memset(c, 0, sizeof(*c));
	c->name = "ctg"; c->vcf_rid = rid; c->start_pos = 1; c->end_pos = ctg_end; c->curr_reg = NULL;
	if (reg_start || reg_stop) { pv_reg.ctg = c; pv_reg.start = reg_start; pv_reg.stop = reg_stop; c->curr_reg = &pv_reg; }
	bs_stats *saved = par.work.stats;
	par.work.stats = NULL;                     /* the writer's own statistics: only when bsref_writer_stats(1, ..) asked for them */
	if (pv_stats_on) {
		/* what init_stats() sets up (src/stats.c:304-318) and the contig's GC bins (src/read_reference.c:120-123); the bins are
		 * a fresh malloc() per call because _print_vcf_entry() frees those of the contig it saw before (:113-119) */
		if (!ref_stats.qd_stats) {
			ref_stats.qd_stats = gt_vector_new(256, sizeof(fstats_cts));
			ref_stats.fs_stats = gt_vector_new(256, sizeof(fstats_cts));
			ref_stats.mq_stats = gt_vector_new(256, sizeof(fstats_cts));
			memset(ref_stats.qd_stats->memory, 0, sizeof(fstats_cts) * ref_stats.qd_stats->elements_allocated);
			memset(ref_stats.fs_stats->memory, 0, sizeof(fstats_cts) * ref_stats.fs_stats->elements_allocated);
			memset(ref_stats.mq_stats->memory, 0, sizeof(fstats_cts) * ref_stats.mq_stats->elements_allocated);
		}
		gt_ctg_stats *cs = pv_cstats + flip;
		if (cs->gc) free(cs->gc);
		memset(cs, 0, sizeof(*cs));
		if (pv_nbins > 0) { cs->gc = malloc((size_t)pv_nbins); memcpy(cs->gc, pv_gc, (size_t)pv_nbins); cs->nbins = pv_nbins; }
		c->ctg_stats = cs;
		pv_last_cstats = cs;
		c->start_pos = pv_start;
		par.work.stats = &ref_stats;
	}
	par.work.vcf_ctg = c;
	par.work.dbSNP_hdr = NULL;
	db_tab.n = 0;
	if (db_n) {
		db_tab.n = db_n; db_tab.pos = db_pos; db_tab.flags = db_flags; db_tab.name_off = db_name_off; db_tab.names = db_names;
		if (db_hdr.dbSNP == NULL) { db_ctg.name = "ctg"; HASH_ADD_KEYPTR(hh, db_hdr.dbSNP, db_ctg.name, strlen(db_ctg.name), &db_ctg); }
		par.work.dbSNP_hdr = &db_hdr;
	}
	par.all_positions = all_positions;
	for (int i = 0; i < 16; i++) par.work.vcf_ids[i] = vcf_ids[i];
	pv_out = out; pv_cap = cap; pv_len = 0; pv_nrec = 0; pv_overflow = 0;
	char *rf = malloc((size_t)sz + 3);
	memcpy(rf, refcodes, (size_t)sz + 2);
	rf[sz + 2] = 0;
	for (uint32_t i = 0; i < sz; i++) {
		gt_vcf v = vcf[i];
		print_vcf_entry(bcf, c, &v.gtm, rf, x + i, x, v.skip, &par);
	}
	flush_vcf_entries(bcf, &par);
	free(rf);
	par.work.stats = saved;
	par.work.dbSNP_hdr = NULL;
	*nbytes = pv_len; *nrec = pv_nrec;
	return pv_overflow ? -3 : 0;
}
